// Fused RAM kernels for sm_100a: one persistent CTA per image plane, band buffers in shared memory.
// See ram_fused_core.cuh for the algorithm and the reference lines it follows
// (/root/reference/code/dataset/fundus.py:13-61).
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include "ram_fused_host.h"
#include "ram_internal.h"

namespace ramdsir {

template <int MMAX, bool BANK>
__global__ void __launch_bounds__(512, 1)
ram_fused_kernel(const FusedShape S, const float* __restrict__ src, const float* __restrict__ bank,
                 const int64_t* __restrict__ pidx, const int32_t* __restrict__ rk, int ratio_den, int n_bank,
                 int C, float* __restrict__ out, const Epilogue ep, const float* __restrict__ tables, int nplanes) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, NT = S.nthreads;
  {
    float* tab = reinterpret_cast<float*>(sm + S.offTab);
    for (int i = tid; i < S.tabBytes / 4; i += NT) tab[i] = tables[i];
  }
  __syncthreads();
  const size_t plane_px = (size_t)S.H * S.W;
  const size_t patch_px = (size_t)S.nb * S.nb;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    const float4* x4 = reinterpret_cast<const float4*>(src + plane * plane_px);
    auto ld = [&](size_t off) { return __ldg(x4 + off); };
    // ---- forward columns: radix-16 in registers, band rows by direct sums through E ----
    for (int strip = 0; strip < S.nstrips; ++strip) {
      cf va[16], vb[16];
      f1_load_fft(S, ld, strip, tid, va, vb);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        f1_store_half(S, sm, tid, half, va, vb);
        __syncthreads();
        for (int u = tid; u < 256; u += NT) f1_direct<MMAX>(S, sm, strip, half, u);
        __syncthreads();
      }
    }
    // ---- forward rows on the band rows only ----
    for (int r = tid; r < S.nb * S.N1w; r += NT) row_fwd(S, sm, r);
    __syncthreads();
    if (BANK) {
      float* patch = out + plane * patch_px;
      for (int q = tid; q < S.nb * (S.b + 1); q += NT) band_amp(S, sm, patch, q);
      __syncthreads();
      continue;
    }
    // ---- band: |F|, blend with the partner amplitude, dF ----
    Epilogue epp = ep;
    {
      const int img = plane / C, ch = plane - img * C;
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)n_bank);
      if (!ok) epp.scale = __int_as_float(0x7fc00000);  // poison the plane: fminf/fmaxf would swallow a NaN delta
      const float omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den : __int_as_float(0x7fc00000);
      const float* patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      for (int q = tid; q < S.nb * (S.b + 1); q += NT) band_mix(S, sm, patch, omr, q);
    }
    __syncthreads();
    // ---- inverse rows, then inverse columns + x + epilogue ----
    for (int r = tid; r < S.nb * S.N1w; r += NT) row_inv<MMAX>(S, sm, r);
    __syncthreads();
    {
      float4* y4 = reinterpret_cast<float4*>(out + plane * plane_px);
      auto st = [&](size_t off, float4 v) { y4[off] = v; };
      for (int strip = 0; strip < S.nstrips; ++strip) col_inv_store<MMAX>(S, sm, ld, st, epp, strip, tid);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
namespace {
struct FusedPlan {
  FusedShape S;
  float* tables = nullptr;  // device
  int sm_count = 0;
};
std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, FusedPlan> g_plans;  // (device, H, W, b)

int get_plan(int H, int W, int b, FusedPlan** out) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, H, W, b);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    FusedPlan p;
    if (!fused_shape_init(p.S, H, W, b)) {
      set_error("fused path does not support H=%d W=%d b=%d", H, W, b);
      return RAMDSIR_ERR_UNSUPPORTED;
    }
    std::vector<float> host(p.S.tabBytes / 4);
    fused_tables_fill(p.S, host.data());
    RD_CUDA(cudaMalloc(&p.tables, p.S.tabBytes));
    RD_CUDA(cudaMemcpy(p.tables, host.data(), p.S.tabBytes, cudaMemcpyHostToDevice));
    RD_CUDA(cudaDeviceSynchronize());  // once per shape: tables are visible to every stream afterwards
    RD_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
    it = g_plans.emplace(key, p).first;
  }
  *out = &it->second;
  return RAMDSIR_OK;
}

template <int MMAX, bool BANK>
int launch(const FusedPlan& p, const float* src, const float* bank, const int64_t* pidx, const int32_t* rk,
           int ratio_den, int n_bank, int C, float* out, Epilogue ep, int nplanes, cudaStream_t st) {
  auto kern = ram_fused_kernel<MMAX, BANK>;
  static thread_local int configured_smem = -1;  // per instantiation
  if (configured_smem < p.S.smemBytes) {
    RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.S.smemBytes));
    configured_smem = p.S.smemBytes;
  }
  int per_sm = 1;
  RD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p.S.nthreads, p.S.smemBytes));
  if (per_sm < 1) per_sm = 1;
  int grid = p.sm_count * per_sm;
  if (grid > nplanes) grid = nplanes;
  kern<<<grid, p.S.nthreads, p.S.smemBytes, st>>>(p.S, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, p.tables, nplanes);
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

template <bool BANK>
int dispatch(const FusedPlan& p, const float* src, const float* bank, const int64_t* pidx, const int32_t* rk,
             int ratio_den, int n_bank, int C, float* out, Epilogue ep, int nplanes, cudaStream_t st) {
  switch (p.S.MM) {
    case 0:
    case 1: return launch<1, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
    case 2: return launch<2, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
    case 3: return launch<3, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
    case 4: return launch<4, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  }
  set_error("band half-width %d too large for the fused path", p.S.b);
  return RAMDSIR_ERR_UNSUPPORTED;
}
}  // namespace

bool fused_supported(int H, int W, int b) {
  FusedShape S;
  return fused_shape_init(S, H, W, b);
}

int fused_mix(const MixArgs& a, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(a.H, a.W, a.b, &p);
  if (rc) return rc;
  return dispatch<false>(*p, a.src, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, a.C, a.out, a.ep, a.B * a.C, st);
}

int fused_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(H, W, b, &p);
  if (rc) return rc;
  return dispatch<true>(*p, imgs, nullptr, nullptr, nullptr, 10, 0, C, bank, Epilogue{0, 0, 1, 0}, N * C, st);
}

}  // namespace ramdsir
