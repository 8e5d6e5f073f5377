// Fused RAM kernels for sm_100a: one persistent CTA per image plane, band buffers in shared memory.
// See ram_fused_core.cuh for the algorithm and the reference lines it follows
// (/root/reference/code/dataset/fundus.py:13-61).
#include <atomic>
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
// Static-shape kernel (ram_fused_static.cuh): the BASELINE shapes.  Two CTAs per SM for 256^2.
//
// Memory-system choreography.  A plane is read twice (forward columns, then y = x + delta), and the
// second read must come from L2, not HBM, for DRAM traffic to stay at the algorithmic 2 * plane bytes:
//   * first read:  ld.global.nc with L2::evict_last  -> keep x resident while the band is processed
//   * second read: L2::evict_first                   -> the line is dead after this use
//   * output:      st.global with L2::evict_first    -> streaming stores must not displace x
//   * the second pass walks the column strips in REVERSE order (most recently read strip first)
//   * while a strip is being processed, the next strip's lines are prefetched into L2.
__device__ __forceinline__ float4 ldg_policy(const float4* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// uint8 sources (images straight from the decoder, fundus.py:205-214 hands RAM uint8-valued pixels): one 32-bit load
// is four adjacent pixels.  Byte -> float through the 2^23 trick (PRMT + FADD, exact; no quarter-rate I2F).
__device__ __forceinline__ float4 quad_of_bytes(unsigned v) {
  const float m = 8388608.f;
  return make_float4(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7541)) - m,
                     __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7542)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7543)) - m);
}
__device__ __forceinline__ float4 ldg_policy(const unsigned* p, unsigned long long pol) {
  unsigned v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return quad_of_bytes(v);
}
__device__ __forceinline__ float4 ldg_plain(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldg_plain(const unsigned* p) { return quad_of_bytes(__ldg(p)); }
template <class T> struct QuadOf;                                   // the 4-pixel load unit of a source type
template <> struct QuadOf<float> { using type = float4; };
template <> struct QuadOf<unsigned char> { using type = unsigned; };
__device__ __forceinline__ void stg_policy(float4* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Each thread prefetches 2 of the 128-byte lines its warp will load for column strip `strip`:
// a warp covers rows (h1w + (lane >> 4)) + N1h * (lane & 15), 256 bytes (2 lines) each.
template <class CFG>
__device__ __forceinline__ void prefetch_strip(const float4* x4, int strip, int tid) {
  const int lane = tid & 31, h1w = (tid >> 5) * 2;
  const int row = h1w + (lane >> 4) + CFG::N1h * (lane & 15);
  const float4* p = x4 + (size_t)row * CFG::M + strip * 16;
  prefetch_l2(p);
  prefetch_l2(p + 8);
}

// uint8 planes: a strip is 64 bytes per row, so two neighbouring strips share each 128-byte line.  Prefetch one line
// per row when the walk enters a new line (ascending: even strips; descending: odd strips).
template <class CFG>
__device__ __forceinline__ void prefetch_strip(const unsigned* x4, int strip, int tid, bool descending = false) {
  if (((strip & 1) != 0) != descending) return;
  for (int row = tid; row < CFG::H; row += CFG::nthreads)
    prefetch_l2(reinterpret_cast<const char*>(x4 + (size_t)row * CFG::M) + (strip >> 1) * 128);
}
template <class CFG>
__device__ __forceinline__ void prefetch_strip(const float4* x4, int strip, int tid, bool) { prefetch_strip<CFG>(x4, strip, tid); }

template <class CFG, bool BANK, bool DUAL, class TIn = float>
__global__ void __launch_bounds__(CFG::nthreads, (CFG::smemBytes <= 113 * 1024) ? 2 : 1)
ram_fused_static_kernel(const TIn* __restrict__ src, const float* __restrict__ bank,
                        const int64_t* __restrict__ pidx, const int32_t* __restrict__ rk, int ratio_den, int n_bank,
                        int C, float* __restrict__ out, const Epilogue ep, const float* __restrict__ tables, int nplanes,
                        float* __restrict__ clean) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int NT = CFG::nthreads;
  const int tid = threadIdx.x;
  // let a launch that was declared independent (RAMDSIR_FLAG_INDEPENDENT) take over SM slots as ours drain
  asm volatile("griddepcontrol.launch_dependents;");
  unsigned long long pol_keep, pol_stream;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  {
    float4* tab = reinterpret_cast<float4*>(sm + CFG::offTab);
    const float4* g = reinterpret_cast<const float4*>(tables);
    for (int i = tid; i < CFG::tabBytes / 16; i += NT) tab[i] = g[i];
  }
  __syncthreads();
  constexpr size_t plane_px = (size_t)CFG::H * CFG::W;
  constexpr size_t patch_px = (size_t)CFG::nb * CFG::nb;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    using Quad = typename QuadOf<TIn>::type;
    const Quad* x4 = reinterpret_cast<const Quad*>(src + plane * plane_px);
    auto ld_first = [&](size_t off) { return BANK ? ldg_plain(x4 + off) : ldg_policy(x4 + off, pol_keep); };
    Epilogue epp = ep;
    float omr = 0.f;
    const float* patch = nullptr;
    if (!BANK) {
      const int img = plane / C, ch = plane - img * C;
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)n_bank);
      if (!ok) epp.scale = __int_as_float(0x7fc00000);   // poison the plane (fminf/fmaxf swallow a NaN delta)
      omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den : __int_as_float(0x7fc00000);
      patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      // pull the partner's patch (10-24 KB) towards L2 now; it is consumed after the forward columns
      const char* pl = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(patch) & ~(uintptr_t)127);
      if (tid * 128 < (int)(patch_px * 4) + 128) prefetch_l2(pl + tid * 128);
    }
    // ---- forward columns ----
    if constexpr (CFG::smemBytes > 113 * 1024) {
      // one CTA per SM (384^2): registers are not capped at 128, so rows j = 0..7 of the next strip are requested
      // one stage early (32 registers in flight across the last exchange stage): the 64 KB load burst of a strip
      // is split in two (+1 % measured).  At two CTAs per SM the same trick spills and loses 4 %.
      float4 lo[8];
      s_f1_load_rows<CFG, 0>(ld_first, 0, tid, lo);
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        cf va[16], vb[16];
        {
          float4 hi[8];
          s_f1_load_rows<CFG, 8>(ld_first, strip, tid, hi);
          if (strip + 1 < CFG::nstrips) prefetch_strip<CFG>(x4, strip + 1, tid);
          s_f1_fft_tw_raw<CFG>(sm, tid, lo, hi, va, vb);
        }
        s_f1_store_half<CFG>(sm, tid, 0, va, vb);
        __syncthreads();
        for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, 0, u);
        __syncthreads();
        s_f1_store_half<CFG>(sm, tid, 1, va, vb);
        if (strip + 1 < CFG::nstrips) s_f1_load_rows<CFG, 0>(ld_first, strip + 1, tid, lo);
        __syncthreads();
        for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, 1, u);
        __syncthreads();
      }
    } else {
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        cf va[16], vb[16];
        s_f1_load_fft_tw<CFG>(ld_first, sm, strip, tid, va, vb);
        if (strip + 1 < CFG::nstrips) prefetch_strip<CFG>(x4, strip + 1, tid);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          s_f1_store_half<CFG>(sm, tid, half, va, vb);
          __syncthreads();
          for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, half, u);
          __syncthreads();
        }
      }
    }
    // ---- forward rows on the band rows.  The partner patch is staged around them: its loads are issued first,
    //      the row transforms run while they are in flight, the symmetrised values are stored afterwards ----
    constexpr int PPT = ((int)patch_px + NT - 1) / NT;      // patch elements per thread
    float pv[PPT], pm[PPT];
    if (!BANK) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int i = tid + k * NT;
        if (i < (int)patch_px) {
          pv[k] = __ldg(patch + i);
          pm[k] = __ldg(patch + ((int)patch_px - 1 - i));
        }
      }
    }
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) s_row_fwd<CFG>(sm, r);
    if (!BANK) {
      float* P = reinterpret_cast<float*>(sm + CFG::offP);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int i = tid + k * NT;
        if (i < (int)patch_px) P[i] = 0.5f * (pv[k] + pm[k]);    // Re(ifft2) only sees the +-k average
      }
    }
    __syncthreads();
    if (BANK) {
      float* patch = out + plane * patch_px;
      for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) s_band_amp<CFG>(sm, patch, q);
      __syncthreads();
      continue;
    }
    for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) s_band_mix<CFG>(sm, omr, q);
    __syncthreads();
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) s_row_inv<CFG>(sm, r);
    __syncthreads();
    // ---- inverse columns + x + epilogue, strips in reverse order ----
    {
      float4* y4 = reinterpret_cast<float4*>(out + plane * plane_px);
      auto ld_again = [&](size_t off) { return ldg_policy(x4 + off, pol_stream); };
      // fused dual output: the normalised CLEAN image x * scale + shift (fundus.py:217-218) is
      // written from the same read of x that feeds y = x + delta
      float4* c4 = DUAL ? reinterpret_cast<float4*>(clean + plane * plane_px) : nullptr;
      auto st = [&](size_t off, float4 v, float4 x) {
        stg_policy(y4 + off, v, pol_stream);
        if (DUAL) stg_policy(c4 + off, make_float4(fmaf(x.x, ep.scale, ep.shift), fmaf(x.y, ep.scale, ep.shift),
                                                 fmaf(x.z, ep.scale, ep.shift), fmaf(x.w, ep.scale, ep.shift)), pol_stream);
      };
      const int next = plane + gridDim.x;
      for (int strip = CFG::nstrips - 1; strip >= 0; --strip) {
        cf eva[8], evb[8], oda[8], odb[8];
        if (strip > 0) prefetch_strip<CFG>(x4, strip - 1, tid, true);
        else if (next < nplanes) prefetch_strip<CFG>(reinterpret_cast<const Quad*>(src + next * plane_px), 0, tid);
        for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 0, u);
        __syncthreads();
        s_inv_fft8<CFG>(sm, tid, 0, eva, evb);
        __syncthreads();
        if constexpr (CFG::smemBytes > 113 * 1024) {
          // one CTA per SM: request half of the strip's x quads before the second-parity stage
          float4 xpre[8];
          s_inv_preload_x<CFG>(ld_again, strip, tid, xpre);
          for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 1, u);
          __syncthreads();
          s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
          __syncthreads();
          s_inv_combine_store_pre<CFG>(ld_again, st, epp, strip, tid, xpre, eva, evb, oda, odb);
        } else {
          for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 1, u);
          __syncthreads();
          s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
          __syncthreads();
          s_inv_combine_store<CFG>(ld_again, st, epp, strip, tid, eva, evb, oda, odb);
        }
      }
    }
    // next plane: its first E write happens after its loads and Z is only rewritten two barriers later
  }
}

// ---------------------------------------------------------------------------------------------
// x2 kernel (ram_fused_x2.cuh): the static float4 pipeline with packed f32x2 arithmetic (FADD2 / FMUL2 / FFMA2).
template <class CFG, bool BANK, bool DUAL>
__global__ void __launch_bounds__(CFG::nthreads, (CFG::smemBytes <= 113 * 1024) ? 2 : 1)
ram_fused_x2_kernel(const float* __restrict__ src, const float* __restrict__ bank, const int64_t* __restrict__ pidx,
                    const int32_t* __restrict__ rk, int ratio_den, int n_bank, int C, float* __restrict__ out,
                    const Epilogue ep, const float* __restrict__ tables, int nplanes, float* __restrict__ clean) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int NT = CFG::nthreads;
  const int tid = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;");
  unsigned long long pol_keep, pol_stream;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  {
    float4* tab = reinterpret_cast<float4*>(sm + CFG::offTab);
    const float4* g = reinterpret_cast<const float4*>(tables);
    for (int i = tid; i < CFG::tabBytes / 16; i += NT) tab[i] = g[i];
  }
  __syncthreads();
  constexpr size_t plane_px = (size_t)CFG::H * CFG::W;
  constexpr size_t patch_px = (size_t)CFG::nb * CFG::nb;
  const bool worker = tid < 2 * CFG::ntask;      // exchange-stage tasks: 2 parities x 128
  const int par2 = tid >> 7, u = tid & 127;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    const float4* x4 = reinterpret_cast<const float4*>(src + plane * plane_px);
    auto ld_first = [&](size_t off) { return BANK ? __ldg(x4 + off) : ldg_policy(x4 + off, pol_keep); };
    for (int strip = 0; strip < CFG::nstrips; ++strip) {
      cf2 v[16];
      x_f1_load_fft_tw<CFG>(ld_first, sm, strip, tid, v);
      if (strip + 1 < CFG::nstrips) prefetch_strip<CFG>(x4, strip + 1, tid);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        x_f1_store_half<CFG>(sm, tid, half, v);
        __syncthreads();
        if (worker) {
          if (par2 == 0) x_f1_pruned<CFG, 0>(sm, strip, half, u);
          else x_f1_pruned<CFG, 1>(sm, strip, half, u);
        }
        __syncthreads();
      }
    }
    Epilogue epp = ep;
    float omr = 0.f;
    if (!BANK) {
      const int img = plane / C, ch = plane - img * C;
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)n_bank);
      if (!ok) epp.scale = __int_as_float(0x7fc00000);
      omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den : __int_as_float(0x7fc00000);
      const float* patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      for (int i = tid; i < (int)patch_px; i += NT) s_stage_patch<CFG>(sm, patch, i);
    }
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) x_row_fwd<CFG>(sm, r);
    __syncthreads();
    if (BANK) {
      float* po = out + plane * patch_px;
      for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) x_band_amp<CFG>(sm, po, q);
      __syncthreads();
      continue;
    }
    for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) x_band_mix<CFG>(sm, omr, q);
    __syncthreads();
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) x_row_inv<CFG>(sm, r);
    __syncthreads();
    {
      float4* y4 = reinterpret_cast<float4*>(out + plane * plane_px);
      auto ld_again = [&](size_t off) { return ldg_policy(x4 + off, pol_stream); };
      float4* c4 = DUAL ? reinterpret_cast<float4*>(clean + plane * plane_px) : nullptr;
      auto st = [&](size_t off, float4 v, float4 x) {
        stg_policy(y4 + off, v, pol_stream);
        if (DUAL) stg_policy(c4 + off, make_float4(fmaf(x.x, ep.scale, ep.shift), fmaf(x.y, ep.scale, ep.shift),
                                                   fmaf(x.z, ep.scale, ep.shift), fmaf(x.w, ep.scale, ep.shift)), pol_stream);
      };
      const int next = plane + gridDim.x;
      for (int strip = CFG::nstrips - 1; strip >= 0; --strip) {
        cf2 ev[8], od[8];
        if (strip > 0) prefetch_strip<CFG>(x4, strip - 1, tid);
        else if (next < nplanes) prefetch_strip<CFG>(reinterpret_cast<const float4*>(src + next * plane_px), 0, tid);
        if (worker) {
          if (par2 == 0) x_inv_pruned<CFG, 0>(sm, strip, 0, u);
          else x_inv_pruned<CFG, 1>(sm, strip, 0, u);
        }
        __syncthreads();
        x_inv_fft8<CFG>(sm, tid, 0, ev);
        __syncthreads();
        if (worker) {
          if (par2 == 0) x_inv_pruned<CFG, 0>(sm, strip, 1, u);
          else x_inv_pruned<CFG, 1>(sm, strip, 1, u);
        }
        __syncthreads();
        x_inv_fft8<CFG>(sm, tid, 1, od);
        __syncthreads();
        x_inv_combine_store<CFG>(ld_again, st, epp, strip, tid, ev, od);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// P2 kernel (ram_fused_p2.cuh): one packed complex column per thread, 64 registers, 1024 threads/SM.
__device__ __forceinline__ float2 ldg_policy2(const float2* p, unsigned long long pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_policy2(float2* p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
// a warp (one h1, 32 columns) loads 16 rows x 256 bytes per strip = 32 lines: one prefetch per lane
template <class CFG>
__device__ __forceinline__ void prefetch_strip_p2(const float2* x2, int strip, int tid) {
  const int lane = tid & 31, h1 = tid >> 5;
  prefetch_l2(x2 + (size_t)(h1 + CFG::N1h * (lane >> 1)) * CFG::M2 + strip * 32 + (lane & 1) * 16);
}

template <class CFG, bool BANK>
__global__ void __launch_bounds__(CFG::nthreads, (CFG::smemBytes <= 113 * 1024) ? 2 : 1)
ram_fused_p2_kernel(const float* __restrict__ src, const float* __restrict__ bank, const int64_t* __restrict__ pidx,
                    const int32_t* __restrict__ rk, int ratio_den, int n_bank, int C, float* __restrict__ out,
                    const Epilogue ep, const float* __restrict__ tables, int nplanes) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int NT = CFG::nthreads;
  const int tid = threadIdx.x;
  // let a launch that was declared independent (RAMDSIR_FLAG_INDEPENDENT) take over SM slots as ours drain
  asm volatile("griddepcontrol.launch_dependents;");
  unsigned long long pol_keep, pol_stream;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  {
    float4* tab = reinterpret_cast<float4*>(sm + CFG::offTab);
    const float4* g = reinterpret_cast<const float4*>(tables);
    for (int i = tid; i < CFG::tabBytes / 16; i += NT) tab[i] = g[i];
  }
  __syncthreads();
  constexpr size_t plane_px = (size_t)CFG::H * CFG::W;
  constexpr size_t patch_px = (size_t)CFG::nb * CFG::nb;
  const bool worker = tid < CFG::npruned;          // pruned-stage tasks: 2 parities x 256
  const int par2 = tid >> 8, u = tid & 255;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    const float2* x2 = reinterpret_cast<const float2*>(src + plane * plane_px);
    auto ld_first = [&](size_t off) { return BANK ? __ldg(x2 + off) : ldg_policy2(x2 + off, pol_keep); };
    for (int strip = 0; strip < CFG::nstrips; ++strip) {
      cf v[16];
      p_f1_load_fft_tw<CFG>(ld_first, sm, strip, tid, v);
      if (strip + 1 < CFG::nstrips) prefetch_strip_p2<CFG>(x2, strip + 1, tid);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        p_f1_store_half<CFG>(sm, tid, half, v);
        __syncthreads();
        if (worker) {
          if (par2 == 0) p_f1_pruned<CFG, 0>(sm, strip, half, u);
          else p_f1_pruned<CFG, 1>(sm, strip, half, u);
        }
        __syncthreads();
      }
    }
    Epilogue epp = ep;
    float omr = 0.f;
    if (!BANK) {
      const int img = plane / C, ch = plane - img * C;
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)n_bank);
      if (!ok) epp.scale = __int_as_float(0x7fc00000);
      omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den : __int_as_float(0x7fc00000);
      const float* patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      for (int i = tid; i < (int)patch_px; i += NT) s_stage_patch<CFG>(sm, patch, i);
    }
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) p_row_fwd<CFG>(sm, r);
    __syncthreads();
    if (BANK) {
      float* patch = out + plane * patch_px;
      for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) p_band_amp<CFG>(sm, patch, q);
      __syncthreads();
      continue;
    }
    for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) p_band_mix<CFG>(sm, omr, q);
    __syncthreads();
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) p_row_inv<CFG>(sm, r);
    __syncthreads();
    {
      float2* y2 = reinterpret_cast<float2*>(out + plane * plane_px);
      auto ld_again = [&](size_t off) { return ldg_policy2(x2 + off, pol_stream); };
      auto st = [&](size_t off, float2 v) { stg_policy2(y2 + off, v, pol_stream); };
      const int next = plane + gridDim.x;
      for (int strip = CFG::nstrips - 1; strip >= 0; --strip) {
        cf ev[8], od[8];
        if (strip > 0) prefetch_strip_p2<CFG>(x2, strip - 1, tid);
        else if (next < nplanes) prefetch_strip_p2<CFG>(reinterpret_cast<const float2*>(src + next * plane_px), 0, tid);
        if (worker) {
          if (par2 == 0) p_inv_pruned<CFG, 0>(sm, strip, 0, u);
          else p_inv_pruned<CFG, 1>(sm, strip, 0, u);
        }
        __syncthreads();
        p_inv_fft8<CFG>(sm, tid, 0, ev);
        __syncthreads();
        if (worker) {
          if (par2 == 0) p_inv_pruned<CFG, 0>(sm, strip, 1, u);
          else p_inv_pruned<CFG, 1>(sm, strip, 1, u);
        }
        __syncthreads();
        p_inv_fft8<CFG>(sm, tid, 1, od);
        __syncthreads();
        p_inv_combine_store<CFG>(ld_again, st, epp, strip, tid, ev, od);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
namespace {
struct FusedPlan {
  FusedShape S;
  float* tables = nullptr;  // device
  int sm_count = 0;
  int static_id = 0;        // 0: runtime-shape kernel; 1/2: SCfg<256..>/<384..> (float4); 3/4: PCfg (float2, default)
};
using Cfg256 = SCfg<256, 256, 25>;
using Cfg384 = SCfg<384, 384, 38>;
using P256 = PCfg<256, 256, 25>;
using P384 = PCfg<384, 384, 38>;
using X256 = XCfg<256, 256, 25>;
using X384 = XCfg<384, 384, 38>;
thread_local float* t_clean = nullptr;   // fused dual output target of the calling thread's current call (or null)
thread_local int t_overlap = 0;   // per-call flag of the calling thread (set by fused_mix, read by launch_static)
std::atomic<int> g_variant{0};   // RAMDSIR_VARIANT_*: 0 auto (= float4 static kernels), 1 P2, 2 runtime-shape
int static_id_of(int H, int W, int b, int variant) {
  if (variant == 2) return 0;
  if (variant == 4) {
    if (H == 256 && W == 256 && b == 25) return 7;
    if (H == 384 && W == 384 && b == 38) return 8;
    return 0;
  }
  const int base = (variant == 1) ? 2 : 0;
  if (H == 256 && W == 256 && b == 25) return base + 1;
  if (H == 384 && W == 384 && b == 38) return base + 2;
  return 0;
}
std::mutex g_mu;
std::map<std::tuple<int, int, int, int, int>, FusedPlan> g_plans;  // (device, H, W, b, variant)

int get_plan(int H, int W, int b, FusedPlan** out) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  const int variant = g_variant.load(std::memory_order_relaxed);
  auto key = std::make_tuple(dev, H, W, b, variant);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    FusedPlan p;
    if (!fused_shape_init(p.S, H, W, b)) {
      set_error("fused path does not support H=%d W=%d b=%d", H, W, b);
      return RAMDSIR_ERR_UNSUPPORTED;
    }
    p.static_id = static_id_of(H, W, b, variant);
    const int tab_bytes = p.static_id == 1 ? Cfg256::tabBytes : p.static_id == 2 ? Cfg384::tabBytes
                        : p.static_id == 3 ? P256::tabBytes : p.static_id == 4 ? P384::tabBytes
                        : p.static_id == 7 ? X256::tabBytes : p.static_id == 8 ? X384::tabBytes : p.S.tabBytes;
    std::vector<float> host(tab_bytes / 4);
    if (p.static_id == 1) fused_tables_fill_static<Cfg256>(host.data());
    else if (p.static_id == 2) fused_tables_fill_static<Cfg384>(host.data());
    else if (p.static_id == 3) fused_tables_fill_p2<P256>(host.data());
    else if (p.static_id == 4) fused_tables_fill_p2<P384>(host.data());
    else if (p.static_id == 7) fused_tables_fill_x2<X256>(host.data());
    else if (p.static_id == 8) fused_tables_fill_x2<X384>(host.data());
    else fused_tables_fill(p.S, host.data());
    RD_CUDA(cudaMalloc(&p.tables, tab_bytes));
    RD_CUDA(cudaMemcpy(p.tables, host.data(), tab_bytes, cudaMemcpyHostToDevice));
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
  static thread_local int configured_smem_dev[64];  // per (instantiation, device); zero-initialised
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  int& configured_smem = configured_smem_dev[dev & 63];
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

template <class CFG, bool BANK, class TIn = float>
struct KernelOf {
  static constexpr bool dual = true;    // takes the extra `clean` pointer; a second instantiation writes it
  static auto get() { return ram_fused_static_kernel<CFG, BANK, false, TIn>; }
  static auto get_dual() { return ram_fused_static_kernel<CFG, BANK, !BANK, TIn>; }
};
template <int H, int W, int B, bool BANK>
struct KernelOf<PCfg<H, W, B>, BANK, float> {
  static constexpr bool dual = false;
  static auto get() { return ram_fused_p2_kernel<PCfg<H, W, B>, BANK>; }
};

template <int H, int W, int B, bool BANK>
struct KernelOf<XCfg<H, W, B>, BANK, float> {
  static constexpr bool dual = true;
  static auto get() { return ram_fused_x2_kernel<XCfg<H, W, B>, BANK, false>; }
  static auto get_dual() { return ram_fused_x2_kernel<XCfg<H, W, B>, BANK, !BANK>; }
};

template <class CFG, bool BANK, class TIn = float>
int launch_static(const FusedPlan& p, const TIn* src, const float* bank, const int64_t* pidx, const int32_t* rk,
                  int ratio_den, int n_bank, int C, float* out, Epilogue ep, int nplanes, cudaStream_t st) {
  using KO = KernelOf<CFG, BANK, TIn>;
  auto kern = KO::get();
  if constexpr (KO::dual) {
    if (t_clean && !BANK) kern = KO::get_dual();
  }
  // cudaFuncSetAttribute and occupancy are per device: cache per (instantiation, device)
  static thread_local int per_sm_dev[64] = {0};
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  int& per_sm = per_sm_dev[dev & 63];
  if (per_sm == 0) {
    RD_CUDA(cudaFuncSetAttribute(KO::get(), cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    if constexpr (KO::dual)   // whichever of the two instantiations this first call uses, configure both
      RD_CUDA(cudaFuncSetAttribute(KO::get_dual(), cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    RD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CFG::nthreads, CFG::smemBytes));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = p.sm_count * per_sm;
  if (grid > nplanes) grid = nplanes;
  if (t_overlap && !BANK) {
    // programmatic dependent launch: do not wait for the previous kernel in the stream to finish
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(CFG::nthreads);
    cfg.dynamicSmemBytes = CFG::smemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if constexpr (KO::dual)
      RD_CUDA(cudaLaunchKernelEx(&cfg, kern, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, (const float*)p.tables, nplanes, t_clean));
    else
      RD_CUDA(cudaLaunchKernelEx(&cfg, kern, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, (const float*)p.tables, nplanes));
  } else {
    if constexpr (KO::dual)
      kern<<<grid, CFG::nthreads, CFG::smemBytes, st>>>(src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, p.tables, nplanes, BANK ? nullptr : t_clean);
    else
      kern<<<grid, CFG::nthreads, CFG::smemBytes, st>>>(src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, p.tables, nplanes);
  }
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

template <bool BANK>
int dispatch(const FusedPlan& p, const float* src, const float* bank, const int64_t* pidx, const int32_t* rk,
             int ratio_den, int n_bank, int C, float* out, Epilogue ep, int nplanes, cudaStream_t st) {
  if (p.static_id == 1) return launch_static<Cfg256, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 2) return launch_static<Cfg384, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 3) return launch_static<P256, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 4) return launch_static<P384, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 7) return launch_static<X256, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 8) return launch_static<X384, BANK>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
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

template <bool BANK>
int dispatch_u8(const FusedPlan& p, const unsigned char* src, const float* bank, const int64_t* pidx, const int32_t* rk,
                int ratio_den, int n_bank, int C, float* out, Epilogue ep, int nplanes, cudaStream_t st) {
  if (p.static_id == 1) return launch_static<Cfg256, BANK, unsigned char>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  if (p.static_id == 2) return launch_static<Cfg384, BANK, unsigned char>(p, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, nplanes, st);
  set_error("uint8 sources are read directly only by the static fused kernels (256x256/b=25, 384x384/b=38)");
  return RAMDSIR_ERR_UNSUPPORTED;
}
}  // namespace

bool fused_u8_supported(int H, int W, int b) {
  const int id = static_id_of(H, W, b, g_variant.load(std::memory_order_relaxed));
  return id == 1 || id == 2;
}

int fused_mix_u8(const MixArgs& a, const unsigned char* src, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(a.H, a.W, a.b, &p);
  if (rc) return rc;
  t_overlap = a.independent ? 1 : 0;
  t_clean = a.clean;
  return dispatch_u8<false>(*p, src, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, a.C, a.out, a.ep, a.B * a.C, st);
}

int fused_bank_u8(const unsigned char* imgs, float* bank, int N, int C, int H, int W, int b, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(H, W, b, &p);
  if (rc) return rc;
  return dispatch_u8<true>(*p, imgs, nullptr, nullptr, nullptr, 10, 0, C, bank, Epilogue{0, 0, 1, 0}, N * C, st);
}

bool fused_dual_supported(int H, int W, int b) {
  const int id = static_id_of(H, W, b, g_variant.load(std::memory_order_relaxed));
  return id == 1 || id == 2 || id == 7 || id == 8;
}

void fused_set_variant(int v) { g_variant.store(v, std::memory_order_relaxed); }

bool fused_supported(int H, int W, int b) {
  FusedShape S;
  return fused_shape_init(S, H, W, b);
}

int fused_mix(const MixArgs& a, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(a.H, a.W, a.b, &p);
  if (rc) return rc;
  t_overlap = a.independent ? 1 : 0;
  t_clean = (p->static_id == 1 || p->static_id == 2 || p->static_id == 7 || p->static_id == 8) ? a.clean : nullptr;
  return dispatch<false>(*p, a.src, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, a.C, a.out, a.ep, a.B * a.C, st);
}

int fused_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(H, W, b, &p);
  if (rc) return rc;
  return dispatch<true>(*p, imgs, nullptr, nullptr, nullptr, 10, 0, C, bank, Epilogue{0, 0, 1, 0}, N * C, st);
}

}  // namespace ramdsir
