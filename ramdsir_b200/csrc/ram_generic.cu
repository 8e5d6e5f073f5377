// Any-shape RAM kernels: direct band-pruned DFTs, four small passes through a global workspace.
// Serves shapes the fused path cannot (odd sizes, H % 16 != 0, W % 64 != 0, planes whose band
// buffers exceed shared memory).  Same algorithm as ram_fused_core.cuh:
//   T[h][kw]  = sum_w x[h][w] e^{-2 pi i kw w / W}             kw in [0, b]      (row pass)
//   F[kh][kw] = sum_h T[h][kw] e^{-2 pi i kh h / H}            kh in [-b, b]     (column pass)
//   dF        = F (1-ratio) (A_t/|F| - 1) / (HW)                                  (fundus.py:37,49,55)
//   G[h][kw]  = sum_kh dF[kh][kw] e^{+2 pi i kh h / H}
//   y[h][w]   = x + sum_kw c_kw Re(G[h][kw] e^{+2 pi i kw w / W}),  c_0 = 1, c_kw = 2 (Hermitian half)
// Twiddles come from exact tables e^{-2 pi i n / N} indexed by (k * n) mod N in integers.
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <cmath>
#include "ram_internal.h"

namespace ramdsir {
namespace {

struct GShape {
  int H, W, b, nb;
  float invHW;
};

// one block per (row h, plane): T[plane][h][kw], thread -> kw
__global__ void g_row_fwd(GShape S, int nkw, const float* __restrict__ src, const float2* __restrict__ twW, float2* __restrict__ T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw);
  float* row = reinterpret_cast<float*>(tw + S.W);
  const int h = blockIdx.x;
  const size_t plane = blockIdx.y;
  const float* x = src + (plane * S.H + h) * (size_t)S.W;
  for (int w = threadIdx.x; w < S.W; w += blockDim.x) {
    tw[w] = twW[w];
    row[w] = x[w];
  }
  __syncthreads();
  // Degenerate rows.  A constant row transforms to exactly W c at kw == 0 and exactly 0 elsewhere under a butterfly
  // FFT (pocketfft behind the reference's np.fft.fft2, fundus.py:16,46); a direct sum of roots of unity leaves rounding
  // noise instead, whose phase the blend would then keep (np.angle(0) == 0, fundus.py:49).  Detect it exactly.
  bool eq = true;
  for (int w = threadIdx.x; w < S.W; w += blockDim.x) eq = eq && row[w] == row[0];
  const bool constant = __syncthreads_and(eq);
  for (int kw = threadIdx.x; kw < nkw; kw += blockDim.x) {
    float re = 0.f, im = 0.f;
    if (constant) {
      re = (kw == 0) ? row[0] * (float)S.W : 0.f;
    } else {
      int idx = 0;
      for (int w = 0; w < S.W; ++w) {
        const float2 t = tw[idx];
        re = fmaf(row[w], t.x, re);
        im = fmaf(row[w], t.y, im);
        idx += kw;
        if (idx >= S.W) idx -= S.W;
      }
    }
    T[(plane * S.H + h) * (size_t)nkw + kw] = make_float2(re, im);
  }
}

// FULL amplitude spectrum (drop-in extract_amp_spectrum, fundus.py:13-19): one block per (kw, plane),
// kw in [0, W/2]; thread -> kh in [0, H).  Writes amp[kh][kw] and its Hermitian mirror.
__global__ void g_col_full_amp(GShape S, int nkw, const float2* __restrict__ T, const float2* __restrict__ twH, float* __restrict__ amp) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw);
  float2* col = tw + S.H;
  const int kw = blockIdx.x;
  const size_t plane = blockIdx.y;
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) {
    tw[h] = twH[h];
    col[h] = T[(plane * S.H + h) * (size_t)nkw + kw];
  }
  __syncthreads();
  bool eq = true;                       // degenerate columns, as in g_row_fwd
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) eq = eq && col[h].x == col[0].x && col[h].y == col[0].y;
  const bool constant = __syncthreads_and(eq);
  float* a = amp + plane * (size_t)S.H * S.W;
  for (int kh = threadIdx.x; kh < S.H; kh += blockDim.x) {
    float re = 0.f, im = 0.f;
    if (constant) {
      if (kh == 0) { re = col[0].x * (float)S.H; im = col[0].y * (float)S.H; }
    } else {
      int idx = 0;
      for (int h = 0; h < S.H; ++h) {
        const float2 t = tw[idx], v = col[h];
        re = fmaf(v.x, t.x, re);
        re = fmaf(-v.y, t.y, re);
        im = fmaf(v.x, t.y, im);
        im = fmaf(v.y, t.x, im);
        idx += kh;
        if (idx >= S.H) idx -= S.H;
      }
    }
    const float m = sqrtf(re * re + im * im);
    a[(size_t)kh * S.W + kw] = m;
    a[(size_t)((S.H - kh) % S.H) * S.W + ((S.W - kw) % S.W)] = m;
  }
}

// low_freq_mutate_np on un-shifted spectra (fundus.py:21-39): out = amp_src, and inside the centred
// band out = amp_src * ratio + amp_trg * (1 - ratio).  One thread per element.
__global__ void g_band_blend(int C, int H, int W, int b, float ratio, float omr, const float* __restrict__ a_src,
                             const float* __restrict__ a_trg, float* __restrict__ out) {
  const size_t n = (size_t)C * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    // position after fftshift is (k + N/2) mod N; the band is [N/2 - b, N/2 + b] there (fundus.py:27-33)
    const int sh = (h + H / 2) % H - H / 2, sw = (w + W / 2) % W - W / 2;
    float v = a_src[i];
    if (sh >= -b && sh <= b && sw >= -b && sw <= b) v = v * ratio + a_trg[i] * omr;
    out[i] = v;
  }
}

// one block per (kw, plane): F[kh][kw] for all kh, then the band update.  thread -> kh
template <bool BANK>
__global__ void g_col_fwd_band(GShape S, const float2* __restrict__ T, const float2* __restrict__ twH,
                               const float* __restrict__ bank, const int64_t* __restrict__ pidx,
                               const int32_t* __restrict__ rk, int ratio_den, int n_bank, int C,
                               float2* __restrict__ D, float* __restrict__ bank_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw);
  float2* col = tw + S.H;
  const int kw = blockIdx.x;
  const size_t plane = blockIdx.y;
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) {
    tw[h] = twH[h];
    col[h] = T[(plane * S.H + h) * (size_t)(S.b + 1) + kw];
  }
  __syncthreads();
  bool eq = true;                       // degenerate columns, as in g_row_fwd
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) eq = eq && col[h].x == col[0].x && col[h].y == col[0].y;
  const bool constant = __syncthreads_and(eq);
  const size_t patch_px = (size_t)S.nb * S.nb;
  for (int khi = threadIdx.x; khi < S.nb; khi += blockDim.x) {
    const int kh = khi - S.b;
    const int step = ((kh % S.H) + S.H) % S.H;
    float re = 0.f, im = 0.f;
    if (constant) {
      if (kh == 0) { re = col[0].x * (float)S.H; im = col[0].y * (float)S.H; }
    } else {
      int idx = 0;
      for (int h = 0; h < S.H; ++h) {
        const float2 t = tw[idx], v = col[h];
        re = fmaf(v.x, t.x, re);
        re = fmaf(-v.y, t.y, re);
        im = fmaf(v.x, t.y, im);
        im = fmaf(v.y, t.x, im);
        idx += step;
        if (idx >= S.H) idx -= S.H;
      }
    }
    const float a2 = re * re + im * im;
    if (BANK) {
      const float a = sqrtf(a2);
      float* patch = bank_out + plane * patch_px;
      patch[(size_t)khi * S.nb + (kw + S.b)] = a;
      patch[(size_t)(S.nb - 1 - khi) * S.nb + (S.b - kw)] = a;
    } else {
      const int img = (int)(plane / C), ch = (int)(plane - (size_t)img * C);
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)n_bank);
      const float omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den : __int_as_float(0x7fc00000);
      const float* patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      const float at = 0.5f * (patch[(size_t)khi * S.nb + (kw + S.b)] + patch[(size_t)(S.nb - 1 - khi) * S.nb + (S.b - kw)]);
      float2 d;
      if (a2 > 0.f) {
        const float s = omr * (at / sqrtf(a2) - 1.0f) * S.invHW;
        d = make_float2(re * s, im * s);
      } else {
        d = make_float2(omr * at * S.invHW, 0.f);
      }
      D[(plane * S.nb + khi) * (size_t)(S.b + 1) + kw] = d;
    }
  }
}

// one block per (kw, plane): G[h][kw] = sum_kh D[kh][kw] e^{+2 pi i kh h / H}; thread -> h.  Writes over T.
__global__ void g_col_inv(GShape S, const float2* __restrict__ D, const float2* __restrict__ twH, float2* __restrict__ G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw);
  float2* dcol = tw + S.H;
  const int kw = blockIdx.x;
  const size_t plane = blockIdx.y;
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) tw[h] = twH[h];
  for (int khi = threadIdx.x; khi < S.nb; khi += blockDim.x) dcol[khi] = D[(plane * S.nb + khi) * (size_t)(S.b + 1) + kw];
  __syncthreads();
  for (int h = threadIdx.x; h < S.H; h += blockDim.x) {
    float re = 0.f, im = 0.f;
    // idx = (kh * h) mod H, starting at kh = -b
    int idx = (int)((((long long)(-S.b) * h) % S.H + S.H) % S.H);
    for (int khi = 0; khi < S.nb; ++khi) {
      const float2 t = tw[idx], v = dcol[khi];  // multiply by conj(t)
      re = fmaf(v.x, t.x, re);
      re = fmaf(v.y, t.y, re);
      im = fmaf(v.y, t.x, im);
      im = fmaf(-v.x, t.y, im);
      idx += h;
      if (idx >= S.H) idx -= S.H;
    }
    G[(plane * S.H + h) * (size_t)(S.b + 1) + kw] = make_float2(re, im);
  }
}

// one block per (row h, plane): y = clamp(x + delta) * scale + shift; thread -> w
__global__ void g_row_inv(GShape S, const float* __restrict__ src, const float2* __restrict__ G,
                          const float2* __restrict__ twW, float* __restrict__ out, Epilogue ep,
                          const int64_t* __restrict__ pidx, int n_bank, int C) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw);
  float2* g = tw + S.W;
  const int h = blockIdx.x;
  const size_t plane = blockIdx.y;
  for (int w = threadIdx.x; w < S.W; w += blockDim.x) tw[w] = twW[w];
  for (int kw = threadIdx.x; kw <= S.b; kw += blockDim.x) {
    float2 v = G[(plane * S.H + h) * (size_t)(S.b + 1) + kw];
    if (kw > 0) { v.x *= 2.f; v.y *= 2.f; }
    g[kw] = v;
  }
  __syncthreads();
  const size_t base = (plane * S.H + h) * (size_t)S.W;
  {
    const long long p = pidx[plane / C];
    if (p < 0 || p >= (long long)n_bank) ep.scale = __int_as_float(0x7fc00000);  // poison (see ramdsir.h)
  }
  for (int w = threadIdx.x; w < S.W; w += blockDim.x) {
    float acc = 0.f;
    int idx = 0;
    for (int kw = 0; kw <= S.b; ++kw) {
      const float2 t = tw[idx], v = g[kw];  // Re(v * conj(t))
      acc = fmaf(v.x, t.x, acc);
      acc = fmaf(v.y, t.y, acc);
      idx += w;
      if (idx >= S.W) idx -= S.W;
    }
    out[base + w] = fmaf(fminf(fmaxf(src[base + w] + acc, ep.lo), ep.hi), ep.scale, ep.shift);
  }
}

struct GPlan {
  float2* twH = nullptr;
  float2* twW = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, int>, float2*> g_tw;  // (device, N) -> e^{-2 pi i n / N}

int get_tw(int N, float2** out) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, N);
  auto it = g_tw.find(key);
  if (it == g_tw.end()) {
    std::vector<float2> host(N);
    const double tau = 6.283185307179586476925286766559;
    for (int n = 0; n < N; ++n) {
      const double a = -tau * (double)n / N;
      host[n] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    float2* d = nullptr;
    RD_CUDA(cudaMalloc(&d, sizeof(float2) * N));
    RD_CUDA(cudaMemcpy(d, host.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    RD_CUDA(cudaDeviceSynchronize());
    it = g_tw.emplace(key, d).first;
  }
  *out = it->second;
  return RAMDSIR_OK;
}

size_t per_plane_bytes(int H, int b) { return ((size_t)H * (b + 1) + (size_t)(2 * b + 1) * (b + 1)) * sizeof(float2); }

// images per internal chunk: grid.y <= 65535 planes, and at most ~256 MB of workspace
int chunk_images(int C, int H, int b) {
  size_t n = 65535 / (size_t)C;
  const size_t by_ws = ((size_t)256 << 20) / ((size_t)C * per_plane_bytes(H, b));
  if (by_ws < n) n = by_ws;
  return n < 1 ? 1 : (int)n;
}

int run_chunk(bool bank_mode, const MixArgs& a, float* bank_out, int nimg, void* ws, size_t ws_bytes, cudaStream_t st);

int run(bool bank_mode, const MixArgs& a, float* bank_out, int nimg, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int per = chunk_images(a.C, a.H, a.b);
  const size_t img_px = (size_t)a.C * a.H * a.W, patch_px = (size_t)a.C * (2 * a.b + 1) * (2 * a.b + 1);
  for (int i0 = 0; i0 < nimg; i0 += per) {
    const int n = nimg - i0 < per ? nimg - i0 : per;
    MixArgs c = a;
    c.src = a.src + (size_t)i0 * img_px;
    if (!bank_mode) {
      c.out = a.out + (size_t)i0 * img_px;
      c.pidx = a.pidx + i0;
      c.rk = a.rk + i0;
    }
    const int rc = run_chunk(bank_mode, c, bank_mode ? bank_out + (size_t)i0 * patch_px : nullptr, n, ws, ws_bytes, st);
    if (rc) return rc;
  }
  return RAMDSIR_OK;
}

int run_chunk(bool bank_mode, const MixArgs& a, float* bank_out, int nimg, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int H = a.H, W = a.W, b = a.b, C = a.C;
  const size_t planes = (size_t)nimg * C;
  if (planes == 0) return RAMDSIR_OK;
  const size_t need = planes * per_plane_bytes(H, b);
  if (ws == nullptr || ws_bytes < need) {
    set_error("generic path needs %zu workspace bytes, got %zu", need, ws_bytes);
    return RAMDSIR_ERR_WORKSPACE;
  }
  float2 *twH, *twW;
  int rc = get_tw(H, &twH);
  if (rc) return rc;
  rc = get_tw(W, &twW);
  if (rc) return rc;
  GShape S{H, W, b, 2 * b + 1, (float)(1.0 / ((double)H * W))};
  float2* T = reinterpret_cast<float2*>(ws);
  float2* D = T + planes * (size_t)H * (b + 1);
  const size_t smRow = (size_t)W * (sizeof(float2) + sizeof(float2));
  const size_t smCol = (size_t)H * 2 * sizeof(float2) + (size_t)S.nb * sizeof(float2);
  if (smRow > 200 * 1024 || smCol > 200 * 1024) {
    set_error("generic path: H=%d W=%d too large for the staging buffers", H, W);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  if (smRow > 48 * 1024) {
    RD_CUDA(cudaFuncSetAttribute(g_row_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smRow));
    RD_CUDA(cudaFuncSetAttribute(g_row_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smRow));
  }
  if (smCol > 48 * 1024) {
    RD_CUDA(cudaFuncSetAttribute(g_col_fwd_band<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smCol));
    RD_CUDA(cudaFuncSetAttribute(g_col_fwd_band<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smCol));
    RD_CUDA(cudaFuncSetAttribute(g_col_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smCol));
  }
  const dim3 gRow(H, (unsigned)planes), gCol(b + 1, (unsigned)planes);
  g_row_fwd<<<gRow, 64, smRow, st>>>(S, b + 1, a.src, twW, T);
  if (bank_mode) {
    g_col_fwd_band<true><<<gCol, 64, smCol, st>>>(S, T, twH, nullptr, nullptr, nullptr, 10, 0, C, D, bank_out);
    count_launch(2);
  } else {
    g_col_fwd_band<false><<<gCol, 64, smCol, st>>>(S, T, twH, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, C, D, nullptr);
    g_col_inv<<<gCol, 128, smCol, st>>>(S, D, twH, T);
    g_row_inv<<<gRow, 128, smRow, st>>>(S, a.src, T, twW, a.out, a.ep, a.pidx, a.n_bank, C);
    count_launch(4);
  }
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

}  // namespace

// y = x * scale + shift (the clean-image normalisation, fundus.py:217-218) for the paths without the fused dual output
__global__ void g_affine(const float4* __restrict__ x, float4* __restrict__ y, size_t n4, float scale, float shift) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_float4(fmaf(v.x, scale, shift), fmaf(v.y, scale, shift), fmaf(v.z, scale, shift), fmaf(v.w, scale, shift));
  }
}

__global__ void g_affine_scalar(const float* __restrict__ x, float* __restrict__ y, size_t n, float scale, float shift) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = fmaf(x[i], scale, shift);
}

int affine_f32(const float* x, float* y, size_t n, float scale, float shift, cudaStream_t st) {
  if (n == 0) return RAMDSIR_OK;
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
  if (vec) g_affine<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n / 4, scale, shift);
  else g_affine_scalar<<<148 * 8, 256, 0, st>>>(x, y, n, scale, shift);
  count_launch(1);
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

// uint8 -> float32 widening for the host pipeline's uint8 ingestion (fundus loaders hold uint8 images, fundus.py:212)
__global__ void g_u8_to_f32(const uchar4* __restrict__ x, float4* __restrict__ y, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 v = x[i];
    y[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  }
}

int u8_to_f32(const unsigned char* x, float* y, size_t n, cudaStream_t st) {
  if (n == 0) return RAMDSIR_OK;
  if (n % 4 || (reinterpret_cast<uintptr_t>(x) & 3u) || (reinterpret_cast<uintptr_t>(y) & 15u)) {
    set_error("u8_to_f32: size must be a multiple of 4 and the buffers aligned");
    return RAMDSIR_ERR_INVALID;
  }
  g_u8_to_f32<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const uchar4*>(x), reinterpret_cast<float4*>(y), n / 4);
  count_launch(1);
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

size_t full_amp_workspace_bytes(int N, int C, int H, int W) { return (size_t)N * C * H * (W / 2 + 1) * sizeof(float2); }

int full_amp(const float* imgs, float* amp, int N, int C, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t planes = (size_t)N * C;
  if (planes == 0) return RAMDSIR_OK;
  const int nkw = W / 2 + 1;
  if (ws == nullptr || ws_bytes < full_amp_workspace_bytes(N, C, H, W)) {
    set_error("amp_spectrum needs %zu workspace bytes, got %zu", full_amp_workspace_bytes(N, C, H, W), ws_bytes);
    return RAMDSIR_ERR_WORKSPACE;
  }
  const size_t smRow = (size_t)W * 16, smCol = (size_t)H * 16;
  if (planes > 65535 || smRow > 200 * 1024 || smCol > 200 * 1024) {
    set_error("amp_spectrum: shape too large (planes=%zu H=%d W=%d)", planes, H, W);
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  float2 *twH, *twW;
  int rc = get_tw(H, &twH);
  if (rc) return rc;
  rc = get_tw(W, &twW);
  if (rc) return rc;
  if (smRow > 48 * 1024) RD_CUDA(cudaFuncSetAttribute(g_row_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smRow));
  if (smCol > 48 * 1024) RD_CUDA(cudaFuncSetAttribute(g_col_full_amp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smCol));
  GShape S{H, W, 0, 1, 1.f};
  float2* T = reinterpret_cast<float2*>(ws);
  g_row_fwd<<<dim3(H, (unsigned)planes), 128, smRow, st>>>(S, nkw, imgs, twW, T);
  g_col_full_amp<<<dim3(nkw, (unsigned)planes), 128, smCol, st>>>(S, nkw, T, twH, amp);
  count_launch(2);
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

int band_blend(const float* a_src, const float* a_trg, float* out, int C, int H, int W, int b, int ratio_k, int ratio_den, cudaStream_t st) {
  const size_t n = (size_t)C * H * W;
  if (n == 0) return RAMDSIR_OK;
  // same float64 arithmetic as fundus.py:35,37 (ratio = k/10; 1 - ratio), rounded once to float32
  const double ratio = (double)ratio_k / (double)ratio_den;
  const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
  g_band_blend<<<grid, 256, 0, st>>>(C, H, W, b, (float)ratio, (float)(1.0 - ratio), a_src, a_trg, out);
  count_launch(1);
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

bool generic_supported(int H, int W, int b) {
  if (H < 1 || W < 1 || b < 0) return false;
  const int m = H < W ? H : W;
  return 2 * b + 1 <= m && (size_t)(H > W ? H : W) * 16 + (size_t)(2 * b + 1) * 8 <= 200 * 1024;
}

size_t generic_workspace_bytes(int n_images, int C, int H, int W, int b) {
  (void)W;
  const int per = chunk_images(C, H, b);
  return (size_t)(n_images < per ? n_images : per) * C * per_plane_bytes(H, b);
}

int generic_mix(const MixArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) { return run(false, a, nullptr, a.B, ws, ws_bytes, st); }

int generic_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, void* ws, size_t ws_bytes, cudaStream_t st) {
  MixArgs a{};
  a.src = imgs; a.C = C; a.H = H; a.W = W; a.b = b;
  return run(true, a, bank, N, ws, ws_bytes, st);
}

}  // namespace ramdsir
