// extern "C" surface of libramdsir.so (see include/ramdsir.h for the contract and the reference
// interfaces each entry point replaces).
#include <atomic>
#include <cstdint>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "ram_internal.h"

namespace ramdsir {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return RAMDSIR_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int check_shape(int n, int C, int H, int W, int b) {
  if (n < 0 || C < 1 || H < 1 || W < 1 || b < 0) {
    set_error("invalid shape n=%d C=%d H=%d W=%d b=%d", n, C, H, W, b);
    return RAMDSIR_ERR_INVALID;
  }
  if (2 * b + 1 > (H < W ? H : W)) {
    set_error("band (2*%d+1) exceeds the spectrum of a %dx%d image", b, H, W);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  return RAMDSIR_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool ranges_overlap(const void* a, size_t na, const void* b, size_t nb) {
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(a), b0 = reinterpret_cast<uintptr_t>(b);
  return a0 < b0 + nb && b0 < a0 + na;
}

// Kernel family for a shape.  RAMDSIR_FLAG_PATH_GENERIC (a per-call flag, no global state) selects the generic
// kernels explicitly so that two independent CUDA implementations can be cross-checked.
static int path_of(int H, int W, int b, unsigned flags = 0) {
  if (flags & RAMDSIR_FLAG_PATH_GENERIC) return generic_supported(H, W, b) ? RAMDSIR_PATH_GENERIC : RAMDSIR_PATH_NONE;
  if (fused_supported(H, W, b)) return RAMDSIR_PATH_FUSED;
  if (multipass_supported(H, W, b)) return RAMDSIR_PATH_MULTIPASS;
  if (generic_supported(H, W, b)) return RAMDSIR_PATH_GENERIC;
  return RAMDSIR_PATH_NONE;
}

}  // namespace ramdsir

using namespace ramdsir;

extern "C" {

int ramdsir_version(void) { return 100; }  // 0.1.0

const char* ramdsir_last_error(void) { return g_err; }

int ramdsir_band_halfwidth(int H, int W, double L) {
  const double m = (double)(H < W ? H : W);
  return (int)std::floor(m * L);  // fundus.py:26, float64
}

int ramdsir_path(int C, int H, int W, int b) {
  if (check_shape(0, C, H, W, b)) return RAMDSIR_PATH_NONE;
  return path_of(H, W, b);
}

int ramdsir_prepare(int C, int H, int W, int b) {
  int rc = check_shape(0, C, H, W, b);
  if (rc) return rc;
  switch (path_of(H, W, b)) {
    case RAMDSIR_PATH_FUSED: return fused_prepare(H, W, b);
    case RAMDSIR_PATH_MULTIPASS:
    case RAMDSIR_PATH_GENERIC: return RAMDSIR_OK;     // these build nothing lazily that blocks
  }
  set_error("no CUDA path for H=%d W=%d b=%d (and there is no CPU fallback)", H, W, b);
  return RAMDSIR_ERR_UNSUPPORTED;
}

size_t ramdsir_workspace_bytes(int n_images, int C, int H, int W, int b) {
  return ramdsir_workspace_bytes_ex(n_images, C, H, W, b, 0u);
}

size_t ramdsir_workspace_bytes_ex(int n_images, int C, int H, int W, int b, unsigned flags) {
  if (check_shape(n_images, C, H, W, b)) return 0;
  switch (path_of(H, W, b, flags)) {
    case RAMDSIR_PATH_GENERIC: return generic_workspace_bytes(n_images, C, H, W, b);
    case RAMDSIR_PATH_MULTIPASS: return multipass_workspace_bytes(n_images, C, H, W, b);
  }
  return 0;
}

int ramdsir_bank_build_f32(const float* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b,
                           void* workspace_dev, size_t workspace_bytes, void* stream) {
  return ramdsir_bank_build_f32_ex(imgs_dev, bank_dev, N, C, H, W, b, workspace_dev, workspace_bytes, stream, 0u);
}

int ramdsir_bank_build_f32_ex(const float* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b,
                              void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags) {
  int rc = check_shape(N, C, H, W, b);
  if (rc) return rc;
  if (N == 0) return RAMDSIR_OK;
  if (!imgs_dev || !bank_dev) {
    set_error("bank_build: null pointer");
    return RAMDSIR_ERR_INVALID;
  }
  const int path = path_of(H, W, b, flags);
  if (path != RAMDSIR_PATH_GENERIC && !aligned16(imgs_dev)) {   // the float4 paths need 16-byte alignment
    set_error("bank_build: imgs must be 16-byte aligned for this shape");
    return RAMDSIR_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (path) {
    case RAMDSIR_PATH_FUSED: return fused_bank(imgs_dev, bank_dev, N, C, H, W, b, flags, st);
    case RAMDSIR_PATH_MULTIPASS: fused_trail_reset(st); return multipass_bank(imgs_dev, bank_dev, N, C, H, W, b, workspace_dev, workspace_bytes, st);
    case RAMDSIR_PATH_GENERIC: fused_trail_reset(st); return generic_bank(imgs_dev, bank_dev, N, C, H, W, b, workspace_dev, workspace_bytes, st);
  }
  set_error("no CUDA path for H=%d W=%d b=%d (and there is no CPU fallback)", H, W, b);
  return RAMDSIR_ERR_UNSUPPORTED;
}

int ramdsir_mix_f32(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                    const int32_t* ratio_k_dev, float* out_dev, int B, int C, int H, int W, int b,
                    int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                    void* workspace_dev, size_t workspace_bytes, void* stream) {
  return ramdsir_mix_f32_ex(src_dev, bank_dev, partner_idx_dev, ratio_k_dev, out_dev, B, C, H, W, b, n_bank, ratio_den,
                            clip_lo, clip_hi, scale, shift, workspace_dev, workspace_bytes, stream, 0u);
}

int ramdsir_mix_f32_ex(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                       const int32_t* ratio_k_dev, float* out_dev, int B, int C, int H, int W, int b,
                       int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                       void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags) {
  return ramdsir_mix_dual_f32(src_dev, bank_dev, partner_idx_dev, ratio_k_dev, out_dev, nullptr, B, C, H, W, b, n_bank,
                              ratio_den, clip_lo, clip_hi, scale, shift, workspace_dev, workspace_bytes, stream, flags);
}

int ramdsir_mix_dual_f32(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                         const int32_t* ratio_k_dev, float* out_dev, float* clean_out_dev, int B, int C, int H, int W,
                         int b, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                         void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags) {
  int rc = check_shape(B, C, H, W, b);
  if (rc) return rc;
  if (B == 0) return RAMDSIR_OK;
  if (!src_dev || !bank_dev || !partner_idx_dev || !ratio_k_dev || !out_dev) {
    set_error("mix: null pointer");
    return RAMDSIR_ERR_INVALID;
  }
  const int path = path_of(H, W, b, flags);
  if (path != RAMDSIR_PATH_GENERIC && (!aligned16(src_dev) || !aligned16(out_dev))) {
    set_error("mix: src and out must be 16-byte aligned for this shape (float4 kernels)");
    return RAMDSIR_ERR_INVALID;
  }
  if (n_bank < 1 || ratio_den < 1) {
    set_error("mix: n_bank=%d ratio_den=%d must be positive", n_bank, ratio_den);
    return RAMDSIR_ERR_INVALID;
  }
  const size_t batch_bytes = (size_t)B * C * H * W * sizeof(float);
  if (ranges_overlap(src_dev, batch_bytes, out_dev, batch_bytes)) {
    set_error("mix: out must not overlap src (the source is read twice)");
    return RAMDSIR_ERR_INVALID;
  }
  MixArgs a{src_dev, bank_dev, partner_idx_dev, ratio_k_dev, out_dev, B, C, H, W, b, n_bank, ratio_den,
            Epilogue{clip_lo, clip_hi, scale, shift}};
  a.flags = flags;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (clean_out_dev) {
    if ((path != RAMDSIR_PATH_GENERIC && !aligned16(clean_out_dev)) || ranges_overlap(clean_out_dev, batch_bytes, out_dev, batch_bytes) ||
        ranges_overlap(clean_out_dev, batch_bytes, src_dev, batch_bytes)) {
      set_error("mix_dual: clean_out must be 16-byte aligned and overlap neither src nor out");
      return RAMDSIR_ERR_INVALID;
    }
    if (path == RAMDSIR_PATH_FUSED && fused_dual_supported(H, W, b, flags)) {
      a.clean = clean_out_dev;      // written by the RAM kernel itself, from the same read of src
    } else {
      fused_trail_reset(st);
      rc = affine_f32(src_dev, clean_out_dev, (size_t)B * C * H * W, scale, shift, st);
      if (rc) return rc;
    }
  }
  switch (path) {
    case RAMDSIR_PATH_FUSED: return fused_mix(a, st);
    case RAMDSIR_PATH_MULTIPASS: fused_trail_reset(st); return multipass_mix(a, workspace_dev, workspace_bytes, st);
    case RAMDSIR_PATH_GENERIC: fused_trail_reset(st); return generic_mix(a, workspace_dev, workspace_bytes, st);
  }
  set_error("no CUDA path for H=%d W=%d b=%d (and there is no CPU fallback)", H, W, b);
  return RAMDSIR_ERR_UNSUPPORTED;
}

int ramdsir_u8_supported(int C, int H, int W, int b) {
  if (check_shape(0, C, H, W, b)) return 0;
  return (path_of(H, W, b) == RAMDSIR_PATH_FUSED && fused_u8_supported(H, W, b)) ? 1 : 0;
}

int ramdsir_bank_build_u8(const uint8_t* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b, void* stream) {
  int rc = check_shape(N, C, H, W, b);
  if (rc) return rc;
  if (N == 0) return RAMDSIR_OK;
  if (!imgs_dev || !bank_dev || (reinterpret_cast<uintptr_t>(imgs_dev) & 3u)) {
    set_error("bank_build_u8: null or misaligned pointer (imgs must be 4-byte aligned)");
    return RAMDSIR_ERR_INVALID;
  }
  if (!ramdsir_u8_supported(C, H, W, b)) {
    set_error("bank_build_u8: uint8 sources are read directly only at 256x256/b=25 and 384x384/b=38 (got H=%d W=%d b=%d)", H, W, b);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  return fused_bank_u8(imgs_dev, bank_dev, N, C, H, W, b, reinterpret_cast<cudaStream_t>(stream));
}

int ramdsir_mix_dual_u8(const uint8_t* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                        const int32_t* ratio_k_dev, float* out_dev, float* clean_out_dev, int B, int C, int H, int W,
                        int b, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                        void* stream, unsigned flags) {
  int rc = check_shape(B, C, H, W, b);
  if (rc) return rc;
  if (B == 0) return RAMDSIR_OK;
  if (!src_dev || !bank_dev || !partner_idx_dev || !ratio_k_dev || !out_dev) {
    set_error("mix_u8: null pointer");
    return RAMDSIR_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(src_dev) & 3u) || !aligned16(out_dev) || (clean_out_dev && !aligned16(clean_out_dev))) {
    set_error("mix_u8: src must be 4-byte aligned, out and clean_out 16-byte aligned");
    return RAMDSIR_ERR_INVALID;
  }
  if (n_bank < 1 || ratio_den < 1) {
    set_error("mix_u8: n_bank=%d ratio_den=%d must be positive", n_bank, ratio_den);
    return RAMDSIR_ERR_INVALID;
  }
  if (clean_out_dev && clean_out_dev == out_dev) {
    set_error("mix_u8: clean_out must be distinct from out");
    return RAMDSIR_ERR_INVALID;
  }
  if (!ramdsir_u8_supported(C, H, W, b)) {
    set_error("mix_u8: uint8 sources are read directly only at 256x256/b=25 and 384x384/b=38 (got H=%d W=%d b=%d)", H, W, b);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  const size_t out_bytes = (size_t)B * C * H * W * sizeof(float);
  if (ranges_overlap(src_dev, out_bytes / 4, out_dev, out_bytes) ||
      (clean_out_dev && (ranges_overlap(clean_out_dev, out_bytes, out_dev, out_bytes) ||
                         ranges_overlap(clean_out_dev, out_bytes, src_dev, out_bytes / 4)))) {
    set_error("mix_u8: src, out and clean_out must not overlap");
    return RAMDSIR_ERR_INVALID;
  }
  MixArgs a{nullptr, bank_dev, partner_idx_dev, ratio_k_dev, out_dev, B, C, H, W, b, n_bank, ratio_den,
            Epilogue{clip_lo, clip_hi, scale, shift}};
  a.flags = flags;
  a.clean = clean_out_dev;
  return fused_mix_u8(a, src_dev, reinterpret_cast<cudaStream_t>(stream));
}

size_t ramdsir_amp_spectrum_workspace_bytes(int N, int C, int H, int W) {
  if (N < 0 || C < 1 || H < 1 || W < 1) return 0;
  return full_amp_workspace_bytes(N, C, H, W);
}

int ramdsir_amp_spectrum_f32(const float* imgs_dev, float* amp_dev, int N, int C, int H, int W,
                             void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (N < 0 || C < 1 || H < 1 || W < 1 || !imgs_dev || !amp_dev) {
    set_error("amp_spectrum: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  return full_amp(imgs_dev, amp_dev, N, C, H, W, workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int ramdsir_band_blend_f32(const float* amp_src_dev, const float* amp_trg_dev, float* out_dev, int C, int H, int W,
                           int b, int ratio_k, int ratio_den, void* stream) {
  if (C < 1 || H < 1 || W < 1 || b < 0 || ratio_den < 1 || !amp_src_dev || !amp_trg_dev || !out_dev) {
    set_error("band_blend: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  return band_blend(amp_src_dev, amp_trg_dev, out_dev, C, H, W, b, ratio_k, ratio_den, reinterpret_cast<cudaStream_t>(stream));
}

uint64_t ramdsir_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// Host-buffer pipeline: NSLOT device slots, one stream each (H2D -> kernel -> D2H in stream order);
// consecutive chunks go to different slots so the copy engines and the SMs overlap.
struct ramdsir_pipe {
  static constexpr int NSLOT = 3;
  int chunk, C, H, W, b;
  size_t img_elems, ws_bytes;
  cudaStream_t st[NSLOT];
  float* d_src[NSLOT];
  unsigned char* d_u8[NSLOT];   // staging for uint8 ingestion (allocated on first use)
  float* d_out[NSLOT];
  int64_t* d_pidx[NSLOT];
  int32_t* d_rk[NSLOT];
  void* d_ws[NSLOT];
};

int ramdsir_pipe_create(ramdsir_pipe_t** pipe, int max_chunk_images, int C, int H, int W, int b) {
  if (!pipe || max_chunk_images < 1) {
    set_error("pipe_create: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  int rc = check_shape(max_chunk_images, C, H, W, b);
  if (rc) return rc;
  if (path_of(H, W, b) == RAMDSIR_PATH_NONE) {
    set_error("no CUDA path for H=%d W=%d b=%d", H, W, b);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  ramdsir_pipe* p = new ramdsir_pipe();
  memset(p, 0, sizeof(*p));
  p->chunk = max_chunk_images; p->C = C; p->H = H; p->W = W; p->b = b;
  p->img_elems = (size_t)C * H * W;
  p->ws_bytes = ramdsir_workspace_bytes(max_chunk_images, C, H, W, b);
  for (int i = 0; i < ramdsir_pipe::NSLOT; ++i) {
    cudaError_t e = cudaStreamCreateWithFlags(&p->st[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_src[i], p->img_elems * sizeof(float) * p->chunk);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_out[i], p->img_elems * sizeof(float) * p->chunk);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_pidx[i], sizeof(int64_t) * p->chunk);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_rk[i], sizeof(int32_t) * p->chunk);
    if (e == cudaSuccess && p->ws_bytes) e = cudaMalloc(&p->d_ws[i], p->ws_bytes);
    if (e != cudaSuccess) {
      ramdsir_pipe_destroy(p);
      return cuda_fail(e, "pipe_create allocation");
    }
  }
  *pipe = p;
  return RAMDSIR_OK;
}

void ramdsir_pipe_destroy(ramdsir_pipe_t* p) {
  if (!p) return;
  for (int i = 0; i < ramdsir_pipe::NSLOT; ++i) {
    if (p->st[i]) { cudaStreamSynchronize(p->st[i]); cudaStreamDestroy(p->st[i]); }
    cudaFree(p->d_src[i]); cudaFree(p->d_u8[i]); cudaFree(p->d_out[i]); cudaFree(p->d_pidx[i]); cudaFree(p->d_rk[i]); cudaFree(p->d_ws[i]);
  }
  delete p;
}

static int mix_host_impl(ramdsir_pipe_t* p, const float* src_host, const unsigned char* src_u8_host, const float* bank_dev,
                         const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host, int B,
                         int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift);
static int mix_host_chunks(ramdsir_pipe_t* p, const float* src_host, const unsigned char* src_u8_host, const float* bank_dev,
                           const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host, int B,
                           int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift);

int ramdsir_mix_host_f32(ramdsir_pipe_t* p, const float* src_host, const float* bank_dev,
                         const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host,
                         int B, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale,
                         float shift) {
  return mix_host_impl(p, src_host, nullptr, bank_dev, partner_idx_host, ratio_k_host, out_host, B, n_bank, ratio_den,
                       clip_lo, clip_hi, scale, shift);
}

int ramdsir_mix_host_u8(ramdsir_pipe_t* p, const unsigned char* src_host, const float* bank_dev,
                        const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host,
                        int B, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift) {
  return mix_host_impl(p, nullptr, src_host, bank_dev, partner_idx_host, ratio_k_host, out_host, B, n_bank, ratio_den,
                       clip_lo, clip_hi, scale, shift);
}

static int mix_host_impl(ramdsir_pipe_t* p, const float* src_host, const unsigned char* src_u8_host, const float* bank_dev,
                         const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host, int B,
                         int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift) {
  if (!p || (!src_host && !src_u8_host) || !bank_dev || !partner_idx_host || !ratio_k_host || !out_host || B < 0) {
    set_error("mix_host: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  for (int i = 0; i < B; ++i) {
    if (partner_idx_host[i] < 0 || partner_idx_host[i] >= n_bank) {
      set_error("mix_host: partner_idx[%d]=%lld outside the bank of %d images", i, (long long)partner_idx_host[i], n_bank);
      return RAMDSIR_ERR_INVALID;
    }
    if (ratio_k_host[i] < 0 || ratio_k_host[i] > ratio_den) {
      set_error("mix_host: ratio_k[%d]=%d outside [0, %d]", i, (int)ratio_k_host[i], ratio_den);
      return RAMDSIR_ERR_INVALID;
    }
  }
  const int rc_all = mix_host_chunks(p, src_host, src_u8_host, bank_dev, partner_idx_host, ratio_k_host, out_host, B, n_bank,
                                     ratio_den, clip_lo, clip_hi, scale, shift);
  // Success or failure, no copy may still be in flight on the caller's buffers when this returns.
  int rc_sync = RAMDSIR_OK;
  for (int i = 0; i < ramdsir_pipe::NSLOT; ++i) {
    const cudaError_t e = cudaStreamSynchronize(p->st[i]);
    if (e != cudaSuccess && rc_sync == RAMDSIR_OK && rc_all == RAMDSIR_OK) rc_sync = cuda_fail(e, "cudaStreamSynchronize(pipe stream)");
  }
  return rc_all ? rc_all : rc_sync;
}

static int mix_host_chunks(ramdsir_pipe_t* p, const float* src_host, const unsigned char* src_u8_host, const float* bank_dev,
                           const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host, int B,
                           int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift) {
  int slot = 0;
  for (int i0 = 0; i0 < B; i0 += p->chunk, slot = (slot + 1) % ramdsir_pipe::NSLOT) {
    const int n = (B - i0 < p->chunk) ? B - i0 : p->chunk;
    cudaStream_t st = p->st[slot];
    const size_t off = (size_t)i0 * p->img_elems, bytes = (size_t)n * p->img_elems * sizeof(float);
    RD_CUDA(cudaMemcpyAsync(p->d_pidx[slot], partner_idx_host + i0, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    RD_CUDA(cudaMemcpyAsync(p->d_rk[slot], ratio_k_host + i0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    int rc;
    if (src_u8_host) {   // uint8 ingestion: a quarter of the H2D bytes
      if (!p->d_u8[slot]) RD_CUDA(cudaMalloc(&p->d_u8[slot], p->img_elems * (size_t)p->chunk));
      RD_CUDA(cudaMemcpyAsync(p->d_u8[slot], src_u8_host + off, (size_t)n * p->img_elems, cudaMemcpyHostToDevice, st));
      if (ramdsir_u8_supported(p->C, p->H, p->W, p->b)) {   // the RAM kernel reads the bytes itself
        rc = ramdsir_mix_dual_u8(p->d_u8[slot], bank_dev, p->d_pidx[slot], p->d_rk[slot], p->d_out[slot], nullptr, n, p->C,
                                 p->H, p->W, p->b, n_bank, ratio_den, clip_lo, clip_hi, scale, shift, st, 0u);
      } else {                                              // widened on the device first
        rc = u8_to_f32(p->d_u8[slot], p->d_src[slot], (size_t)n * p->img_elems, st);
        if (!rc) rc = ramdsir_mix_f32(p->d_src[slot], bank_dev, p->d_pidx[slot], p->d_rk[slot], p->d_out[slot], n, p->C, p->H,
                                      p->W, p->b, n_bank, ratio_den, clip_lo, clip_hi, scale, shift, p->d_ws[slot], p->ws_bytes, st);
      }
    } else {
      RD_CUDA(cudaMemcpyAsync(p->d_src[slot], src_host + off, bytes, cudaMemcpyHostToDevice, st));
      rc = ramdsir_mix_f32(p->d_src[slot], bank_dev, p->d_pidx[slot], p->d_rk[slot], p->d_out[slot], n, p->C, p->H,
                           p->W, p->b, n_bank, ratio_den, clip_lo, clip_hi, scale, shift, p->d_ws[slot], p->ws_bytes, st);
    }
    if (rc) return rc;
    RD_CUDA(cudaMemcpyAsync(out_host + off, p->d_out[slot], bytes, cudaMemcpyDeviceToHost, st));
  }
  return RAMDSIR_OK;
}

}  // extern "C"
