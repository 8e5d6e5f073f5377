// Fused RAM kernels for sm_100a: one persistent CTA per image plane, band buffers in shared memory.
// See ram_fused_core.cuh for the algorithm and the reference lines it follows
// (/root/reference/code/dataset/fundus.py:13-61).
#include <cstdlib>
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
// The twiddle tables (once per CTA) and the partner's amplitude patch (once per plane, issued before the
// forward columns, consumed after them) arrive by bulk-async copies (cp.async.bulk, the non-tensor TMA
// path) that signal an mbarrier: no thread holds patch values in registers across the column passes.
__device__ __forceinline__ float4 ldg_policy(const float4* p, unsigned long long pol) {
  float4 v;
#ifdef RD_EXP_NOALLOC
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
#else
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
#endif
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

// ---- mbarrier + bulk-async copy (cp.async.bulk global -> shared, completion counted in bytes) ---------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" :: "r"(bar), "r"(parity) : "memory");
}

// ---- tensor memory as a stash for x (no MMA involved) ---------------------------------------------------------------
// TMEM is 128 lanes x 512 columns x 32 bit per SM; a warp reaches the 32 lanes of quadrant (warp % 4) and each thread
// its own lane, so `tcgen05.st / ld .32x32b.xN` moves N registers of every thread to / from N columns of its lane: a
// private, explicitly addressed scratch that never touches the LSU / L1TEX path.  With two CTAs per SM each CTA owns
// 256 columns = 128 per thread (warps w and w + 4 share a quadrant) = the x quads of TWO of the plane's four strips:
// the inverse pass takes those from TMEM instead of reading them a second time through L2.
__device__ __forceinline__ void tmem_alloc(unsigned smem_slot, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// 8 quads (32 registers) -> 32 columns
__device__ __forceinline__ void tmem_st_quads8(unsigned taddr, float4 q0, float4 q1, float4 q2, float4 q3, float4 q4, float4 q5,
                                               float4 q6, float4 q7) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "f"(q0.x), "f"(q0.y), "f"(q0.z), "f"(q0.w), "f"(q1.x), "f"(q1.y), "f"(q1.z), "f"(q1.w), "f"(q2.x), "f"(q2.y),
         "f"(q2.z), "f"(q2.w), "f"(q3.x), "f"(q3.y), "f"(q3.z), "f"(q3.w), "f"(q4.x), "f"(q4.y), "f"(q4.z), "f"(q4.w), "f"(q5.x),
         "f"(q5.y), "f"(q5.z), "f"(q5.w), "f"(q6.x), "f"(q6.y), "f"(q6.z), "f"(q6.w), "f"(q7.x), "f"(q7.y), "f"(q7.z), "f"(q7.w)
      : "memory");
}
// 2 quads (8 columns) -> registers; the caller waits with tmem_wait_ld() before using them
__device__ __forceinline__ void tmem_ld_quads2(unsigned taddr, float4& a, float4& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

__device__ __forceinline__ EpiClamp poisoned(EpiClamp e) { e.scale = __int_as_float(0x7fc00000); return e; }
__device__ __forceinline__ EpiSat poisoned(EpiSat e) { e.m = __int_as_float(0x7fc00000); return e; }   // sat() would swallow a NaN

// Per-launch scalars that are not part of the epilogue policy.
struct MixScalars {
  int ratio_den, n_bank, C, nplanes;
  float dscale;    // EpiSat: 1 / (hi - lo), folded into dF;  EpiClamp: 1
  float dc_add;    // EpiSat: -lo / (hi - lo), added to the DC bin of dF / (H W);  EpiClamp: 0
  float cscale, cshift;   // fused dual output: clean = x * cscale + cshift
  // Completion watermark (see "automatic overlap" below).  A CTA of launch number `seq` that has passed the
  // griddepcontrol.wait at its end knows that launch seq - 1 is complete and flushed, and raises *complete_upto to seq
  // ("every launch before seq is complete").  A launch that may overlap its predecessors first waits until the
  // watermark has reached wait_upto = seq - 2, i.e. until the launch three back is complete: at most the launches one
  // and two back can be alive beside it.
  unsigned long long* complete_upto;
  unsigned long long seq, wait_upto;     // wait_upto == 0: nothing to wait for
};

// which static kernels stash x in tensor memory (-DRD_NO_TMEM_STASH: none, for A/B runs)
template <class CFG, bool BANK, class TIn>
struct StashX {
#ifdef RD_NO_TMEM_STASH
  static constexpr bool value = false;
#else
#ifdef RD_NO_TMEM_STASH_1CTA
  static constexpr bool value = CFG::kTwoCtas && !BANK && std::is_same<TIn, float>::value;
#else
  static constexpr bool value = !BANK && std::is_same<TIn, float>::value;
#endif
#endif
};

// Combine (ram_fused_static.cuh) with the x quads of rows J and J + 8 coming from the thread's TMEM columns
// [8 J, 8 J + 8): the load of pair J + 1 is in flight while pair J is combined and stored.
template <class CFG, int J>
struct CombineT {
  template <class StoreF4, class EPI, class CX>
  static __device__ __forceinline__ void run(unsigned taddr, StoreF4 st, EPI ep, size_t base, float4 x0, float4 x1, const CX (&eva)[8],
                                             const CX (&evb)[8], const CX (&oda)[8], const CX (&odb)[8]) {
    tmem_wait_ld();                                   // pair J has landed in (x0, x1)
    float4 n0 = x0, n1 = x1;
    if (J < 7) tmem_ld_quads2(taddr + 8u * (J + 1), n0, n1);
    const CX wa = twidN<16, J, 1>(oda[pos8(J)]), wb = twidN<16, J, 1>(odb[pos8(J)]);
    const CX ea = eva[pos8(J)], eb = evb[pos8(J)];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = base + (size_t)(CFG::N1h * (J + 8 * hi)) * CFG::M;
      const CX a = hi ? ea - wa : ea + wa, bq = hi ? eb - wb : eb + wb;
      const float4 x = hi ? x1 : x0;
      const float4 y = make_float4(ep(x.x, cx_re(a)), ep(x.y, cx_re(bq)), ep(x.z, cx_im(a)), ep(x.w, cx_im(bq)));
      st(off, y, x);
    }
    CombineT<CFG, J + 1>::run(taddr, st, ep, base, n0, n1, eva, evb, oda, odb);
  }
};
template <class CFG>
struct CombineT<CFG, 8> {
  template <class StoreF4, class EPI, class CX>
  static __device__ __forceinline__ void run(unsigned, StoreF4, EPI, size_t, float4, float4, const CX (&)[8], const CX (&)[8],
                                             const CX (&)[8], const CX (&)[8]) {}
};

template <class CFG, bool BANK, bool DUAL, class TIn, class EPI>
__global__ void __launch_bounds__(CFG::nthreads, (CFG::smemBytes <= 113 * 1024) ? 2 : 1)
ram_fused_static_kernel(const TIn* __restrict__ src, const float* __restrict__ bank,
                        const int64_t* __restrict__ pidx, const int32_t* __restrict__ rk,
                        float* __restrict__ out, const EPI ep, const MixScalars ms, const float* __restrict__ tables,
                        float* __restrict__ clean) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int NT = CFG::nthreads;
  const int tid = threadIdx.x;
  const int nplanes = ms.nplanes, C = ms.C;
  // A following launch that was found (or declared) independent may take over SM slots as ours drain; it orders
  // itself after this grid with the griddepcontrol.wait at its own end.
  asm volatile("griddepcontrol.launch_dependents;");
  // pull the first strip of this CTA's first plane towards L2 while the prologue (barriers, tables, TMEM) runs
  if (!BANK && (int)blockIdx.x < nplanes) {
    const auto* xf = reinterpret_cast<const typename QuadOf<TIn>::type*>(src + (size_t)blockIdx.x * CFG::H * CFG::W);
    prefetch_strip<CFG>(xf, 0, tid);      // strip 0 only: +0.55 %; strips 0 and 1: +0.1 %
  }
  unsigned long long pol_keep, pol_stream;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  const unsigned bar_tab = smem_u32(sm + CFG::offBar), bar_patch = bar_tab + 8;
  if (tid == 0) {
    mbar_init(bar_tab, 1);
    mbar_init(bar_patch, 1);
    mbar_init_fence();
    mbar_expect_tx(bar_tab, CFG::tabBytes);
    bulk_g2s(smem_u32(sm + CFG::offTab), tables, CFG::tabBytes, bar_tab);
    if (ms.wait_upto) {            // the launch three back must be complete (it is, unless a CTA of it straggles)
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ms.complete_upto) : "memory");
      } while (seen < ms.wait_upto);
    }
  }
  // x stash in tensor memory: two-CTAs-per-SM mix kernels with float32 sources (see tmem_* above)
  constexpr bool kStash = StashX<CFG, BANK, TIn>::value;
  constexpr int kStashStrips = 2;                            // strips 0 and 1: read first, re-read last
  constexpr unsigned kTmemCols = CFG::kTwoCtas ? 256 : 512;   // per CTA: half of the SM's tensor memory, or all of it
  // (warp 1 allocates, so that the allocation runs beside thread 0's barrier set-up and table copy)
  if (kStash && (tid >> 5) == 1) tmem_alloc(smem_u32(sm + CFG::offBar + 16), kTmemCols);
  if (kStash) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  unsigned tmem_mine = 0, tmem_base = 0;
  if (kStash) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = *reinterpret_cast<const volatile unsigned*>(sm + CFG::offBar + 16);
    // this thread's 128 columns: lane quadrant = warp % 4 (address bits 31:16 = lane), column half = warp / 4
    tmem_mine = tmem_base + ((unsigned)((tid >> 5) & 3) << 21) + (unsigned)(tid >> 7) * 128u;
  }
  // The tables are first needed by the twiddles of the first strip: the kernels that stash x issue that strip's loads
  // before waiting for them (a CTA lives for 2-3 planes, so its prologue is worth hiding).
  bool tables_pending = kStash;
  if (!tables_pending) mbar_wait(bar_tab, 0);
  unsigned patch_parity = 0;
  constexpr size_t plane_px = (size_t)CFG::H * CFG::W;
  constexpr int patch_px = CFG::nb * CFG::nb;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    using Quad = typename QuadOf<TIn>::type;
    const Quad* x4 = reinterpret_cast<const Quad*>(src + plane * plane_px);
    auto ld_first = [&](size_t off) { return BANK ? ldg_plain(x4 + off) : ldg_policy(x4 + off, pol_keep); };
    EPI epp = ep;
    float omr = 0.f;
    const float* patch = nullptr; // the partner's patch in the bank
    if (!BANK) {
      const int img = plane / C, ch = plane - img * C;
      const long long p = pidx[img];
      const bool ok = (p >= 0 && p < (long long)ms.n_bank);
      if (!ok) epp = poisoned(ep);                      // out-of-range partner: the whole plane becomes NaN
      omr = ok ? (float)(ms.ratio_den - rk[img]) / (float)ms.ratio_den * ms.dscale : __int_as_float(0x7fc00000);
      patch = bank + ((size_t)(ok ? p : 0) * C + ch) * patch_px;
      // pull the partner's patch (10-24 KB) towards L2 now; the bulk copy after the forward columns finds it there
      const char* pl = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(patch) & ~(uintptr_t)127);
      if (tid * 128 < patch_px * 4 + 128) prefetch_l2(pl + tid * 128);
    }
    // ---- forward columns ----
    if constexpr (CFG::smemBytes > 113 * 1024) {
      // one CTA per SM (384^2): registers are not capped at 128, so rows j = 0..7 of the next strip are requested
      // one stage early (32 registers in flight across the last exchange stage): the 64 KB load burst of a strip
      // is split in two (+1 % measured).  At two CTAs per SM the same trick spills and loses 4 %.
      // strips whose x goes to the tensor-memory stash are not read again: no reason to keep them in L2
      auto ld_strip = [&](int strip) {
        const unsigned long long pol = (kStash && strip < kStashStrips) ? pol_stream : pol_keep;
        return [=](size_t off) { return BANK ? ldg_plain(x4 + off) : ldg_policy(x4 + off, pol); };
      };
      float4 lo[8];
      s_f1_load_rows<CFG, 0>(ld_strip(0), 0, tid, lo);
      if (tables_pending) {
        mbar_wait(bar_tab, 0);
        tables_pending = false;
      }
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        cf2 v[16];
        {
          float4 hi[8];
          s_f1_load_rows<CFG, 8>(ld_strip(strip), strip, tid, hi);
          if (strip + 1 < CFG::nstrips) prefetch_strip<CFG>(x4, strip + 1, tid);
          if (kStash && strip < kStashStrips) {     // quads in the order the combine asks for them: (row J, row J + 8)
            tmem_st_quads8(tmem_mine + 64u * strip, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], lo[3], hi[3]);
            tmem_st_quads8(tmem_mine + 64u * strip + 32u, lo[4], hi[4], lo[5], hi[5], lo[6], hi[6], lo[7], hi[7]);
          }
          s_f1_fft_tw_raw_x2<CFG>(sm, tid, lo, hi, v);
          if (kStash && strip < kStashStrips) tmem_wait_st();
        }
        s_f1_store_half_x2<CFG>(sm, tid, 0, v);
        __syncthreads();
        for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, 0, u);
        __syncthreads();
        s_f1_store_half_x2<CFG>(sm, tid, 1, v);
        if (strip + 1 < CFG::nstrips) s_f1_load_rows<CFG, 0>(ld_strip(strip + 1), strip + 1, tid, lo);
        __syncthreads();
        for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, 1, u);
        __syncthreads();
      }
    } else {
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        cf2 v[16];
        const bool stash_this = kStash && strip < kStashStrips;
        if (stash_this) {
          // these strips are not read again: no reason to keep them in L2; columns [64 strip, 64 strip + 64) of this
          // thread's lane take the 16 quads in the order the combine asks for them: (row J, row J + 8), J = 0..7
          auto ld_once = [&](size_t off) { return ldg_policy(x4 + off, pol_stream); };
          const int h1 = tid >> 4, c4 = strip * 16 + (tid & 15);
          float4 q[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) q[j] = ld_once((size_t)(h1 + CFG::N1h * j) * CFG::M + c4);
          if (tables_pending) {
            mbar_wait(bar_tab, 0);
            tables_pending = false;
          }
          tmem_st_quads8(tmem_mine + 64u * strip, q[0], q[8], q[1], q[9], q[2], q[10], q[3], q[11]);
          tmem_st_quads8(tmem_mine + 64u * strip + 32u, q[4], q[12], q[5], q[13], q[6], q[14], q[7], q[15]);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = cf2_from(q[j]);
          s_f1_fft_tw_x2<CFG>(sm, tid, v);       // the stores to tensor memory drain behind the transform
          tmem_wait_st();
        } else {
          s_f1_load_fft_tw_x2<CFG>(ld_first, sm, strip, tid, v);
        }
        // (issued before this strip's own loads, the same prefetch delays them: 0.0915 vs 0.0896 ms)
        if (strip + 1 < CFG::nstrips) prefetch_strip<CFG>(x4, strip + 1, tid);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          s_f1_store_half_x2<CFG>(sm, tid, half, v);
          __syncthreads();
          for (int u = tid; u < 256; u += NT) s_f1_pruned<CFG>(sm, strip, half, u);
          __syncthreads();
        }
      }
    }
    // ---- forward rows on the band rows.  E is idle from here to the inverse columns: the partner's patch is
    //      bulk-copied into it (behind D) while the row transforms run ----
    const float* P = nullptr;     // the partner's patch in shared memory (element 0)
    if (!BANK) {
      // the 16-byte-aligned superset of the patch (a patch starts on any 4-byte boundary of the bank)
      const uintptr_t a0 = reinterpret_cast<uintptr_t>(patch) & ~(uintptr_t)15;
      const int lead = (int)(reinterpret_cast<uintptr_t>(patch) - a0);                 // 0, 4, 8 or 12
      int bytes = (lead + patch_px * 4 + 15) & ~15;
      // never read past the end of the bank: the last patch may be short of a full 16-byte tail
      const uintptr_t bank_end = reinterpret_cast<uintptr_t>(bank + (size_t)ms.n_bank * C * patch_px);
      if (a0 + bytes > bank_end) bytes = (int)((bank_end - a0) & ~(uintptr_t)15);
      P = reinterpret_cast<const float*>(sm + CFG::offPraw + lead);
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // E was last touched through the generic proxy
        mbar_expect_tx(bar_patch, bytes);
        bulk_g2s(smem_u32(sm + CFG::offPraw), reinterpret_cast<const void*>(a0), bytes, bar_patch);
      }
      const int covered = (bytes - lead) >> 2;           // patch elements the bulk copy delivers
      if (covered < patch_px && tid < patch_px - covered)   // <= 3 tail elements, last patch of the bank only
        reinterpret_cast<float*>(sm + CFG::offPraw + lead)[covered + tid] = __ldg(patch + covered + tid);
    }
    // bank builder: no inverse pass to hide the next plane's first loads behind -- start pulling them into L2 here
    // (-0.7 %; its second strip as well, or two strips ahead inside a plane: +1.5 ... +3 %)
    if (BANK && plane + (int)gridDim.x < nplanes)
      prefetch_strip<CFG>(reinterpret_cast<const Quad*>(src + (plane + gridDim.x) * plane_px), 0, tid);
    for (int r = tid; r < CFG::nb * CFG::N1w; r += NT) s_row_fwd<CFG>(sm, r);
    __syncthreads();
    if (BANK) {
      float* po = out + (size_t)plane * patch_px;
      for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) s_band_amp<CFG>(sm, po, q);
      __syncthreads();
      continue;
    }
    mbar_wait(bar_patch, patch_parity);     // the patch has landed (requested before the row transforms)
    patch_parity ^= 1u;
    for (int q = tid; q < CFG::nb * (CFG::b + 1); q += NT) s_band_mix<CFG>(sm, P, omr, ms.dc_add, q);
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
        if (DUAL) stg_policy(c4 + off, make_float4(fmaf(x.x, ms.cscale, ms.cshift), fmaf(x.y, ms.cscale, ms.cshift),
                                                 fmaf(x.z, ms.cscale, ms.cshift), fmaf(x.w, ms.cscale, ms.cshift)), pol_stream);
      };
      const int next = plane + gridDim.x;
      for (int strip = CFG::nstrips - 1; strip >= 0; --strip) {
        typename CFG::cx_fft8 eva[8], evb[8], oda[8], odb[8];
        if (strip > 0 && !(kStash && strip - 1 < kStashStrips)) prefetch_strip<CFG>(x4, strip - 1, tid, true);
        else if (next < nplanes) prefetch_strip<CFG>(reinterpret_cast<const Quad*>(src + next * plane_px), 0, tid);
        for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 0, u);
        __syncthreads();
        s_inv_fft8<CFG>(sm, tid, 0, eva, evb);
        __syncthreads();
        if constexpr (CFG::smemBytes > 113 * 1024) {
          if (kStash && strip < kStashStrips) {     // x of this strip waits in tensor memory
            for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 1, u);
            __syncthreads();
            s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
            __syncthreads();
            const unsigned taddr = tmem_mine + 64u * strip;
            float4 x0, x1;
            tmem_ld_quads2(taddr, x0, x1);
            CombineT<CFG, 0>::run(taddr, st, epp, (size_t)(tid >> 4) * CFG::M + strip * 16 + (tid & 15), x0, x1, eva, evb, oda, odb);
          } else {
            // one CTA per SM: request half of the strip's x quads before the second-parity stage
            float4 xpre[8];
            s_inv_preload_x<CFG>(ld_again, strip, tid, xpre);
            for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 1, u);
            __syncthreads();
            s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
            __syncthreads();
            s_inv_combine_store_pre<CFG>(ld_again, st, epp, strip, tid, xpre, eva, evb, oda, odb);
          }
        } else {
          for (int u = tid; u < 256; u += NT) s_inv_pruned<CFG>(sm, strip, 1, u);
          __syncthreads();
          s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
          __syncthreads();
          if (kStash && strip < kStashStrips) {
            const unsigned taddr = tmem_mine + 64u * strip;
            float4 x0, x1;
            tmem_ld_quads2(taddr, x0, x1);
            CombineT<CFG, 0>::run(taddr, st, epp, (size_t)(tid >> 4) * CFG::M + strip * 16 + (tid & 15), x0, x1, eva, evb, oda, odb);
          } else {
            s_inv_combine_store<CFG>(ld_again, st, epp, strip, tid, eva, evb, oda, odb);
          }
        }
      }
    }
    // next plane: its first E write happens after its loads and Z is only rewritten two barriers later
  }
  if (kStash) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((tid >> 5) == 1) tmem_dealloc(tmem_base, kTmemCols);
  }
  // Programmatic dependent launch: this grid may have started before the previous kernel in the stream finished.
  // Do not let it COMPLETE before that kernel has (completion and memory flush), so that whatever the stream runs
  // next (a copy of the previous batch, a consumer kernel) is ordered after both.  No-op for a plain launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the previous launch in the stream (number seq - 1) is now complete and flushed: raise the watermark
  if (tid == 0) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" :: "l"(ms.complete_upto), "l"(ms.seq) : "memory");
}

// ---------------------------------------------------------------------------------------------
namespace {
struct FusedPlan {
  FusedShape S;
  float* tables = nullptr;  // device
  int sm_count = 0;
  int static_id = 0;        // 0: runtime-shape kernel; 1: SCfg<256,256,25>; 2: SCfg<384,384,38>
};
using Cfg256 = SCfg<256, 256, 25>;
using Cfg384 = SCfg<384, 384, 38>;

int static_id_of(int H, int W, int b) {
  if (H == 256 && W == 256 && b == 25) return 1;
  if (H == 384 && W == 384 && b == 38) return 2;
  return 0;
}
std::mutex g_mu;
std::map<std::tuple<int, int, int, int, int>, FusedPlan> g_plans;  // (device, H, W, b, runtime-shape?)
std::map<int, cudaStream_t> g_upload_stream;                       // per device, for table uploads

// Tables are uploaded on a private stream and that stream alone is synchronised: first use of a shape blocks the
// calling host thread for the few microseconds of a 7 KB copy, but never the device or anybody else's stream.
int get_plan(int H, int W, int b, bool runtime_shape, FusedPlan** out) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, H, W, b, runtime_shape ? 1 : 0);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    FusedPlan p;
    if (!fused_shape_init(p.S, H, W, b)) {
      set_error("fused path does not support H=%d W=%d b=%d", H, W, b);
      return RAMDSIR_ERR_UNSUPPORTED;
    }
    p.static_id = runtime_shape ? 0 : static_id_of(H, W, b);
    const int tab_bytes = p.static_id == 1 ? Cfg256::tabBytes : p.static_id == 2 ? Cfg384::tabBytes : p.S.tabBytes;
    std::vector<float> host(tab_bytes / 4);
    if (p.static_id == 1) fused_tables_fill_static<Cfg256>(host.data());
    else if (p.static_id == 2) fused_tables_fill_static<Cfg384>(host.data());
    else fused_tables_fill(p.S, host.data());
    cudaStream_t& up = g_upload_stream[dev];
    if (!up) RD_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    RD_CUDA(cudaMalloc(&p.tables, tab_bytes));
    RD_CUDA(cudaMemcpyAsync(p.tables, host.data(), tab_bytes, cudaMemcpyHostToDevice, up));
    RD_CUDA(cudaStreamSynchronize(up));   // `host` dies at scope exit; the tables are complete for every stream
    RD_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
    it = g_plans.emplace(key, p).first;
  }
  *out = &it->second;
  return RAMDSIR_OK;
}

// ---- automatic overlap of consecutive launches -------------------------------------------------------------
// A static fused launch may start while the previous fused launches in the same stream drain their last CTAs
// (programmatic dependent launch) when they touch disjoint memory: nothing this launch reads or writes was written by
// them, and nothing it writes was read by them.  The library knows every buffer of every launch, so it decides this
// itself; the caller's RAMDSIR_FLAG_INDEPENDENT only skips the check.  A dependent launch is dispatched once its
// predecessor has dispatched all its CTAs, not once older launches have retired theirs, so CTAs of the launches one
// AND two back can still be running: the last kTrailDepth = 2 launches are compared, and the completion watermark
// (Tickets, below) makes every launch wait until the launch three back is complete, so that nothing older can be
// alive.  Anything else the library enqueues on the stream clears the trail (the next launch is then a plain,
// fully ordered one).
struct Range { uintptr_t lo, hi; };
struct LaunchFootprint { Range rd[4]; int nrd = 0; Range wr[2]; int nwr = 0; };
constexpr int kTrailDepth = 2;
struct Trail { LaunchFootprint last[kTrailDepth]; int n = 0; };
std::mutex g_trail_mu;
std::map<std::pair<int, cudaStream_t>, Trail> g_trails;

// Completion watermark: one monotonic device word per (device, stream), "every launch numbered below this is complete"
// (numbers start at 1).  Never freed or reset (8 bytes per stream).
struct Tickets {
  unsigned long long* complete_upto = nullptr;   // device
  unsigned long long seq = 1;                     // number of the next launch
  unsigned long long plain_seq = 1;               // number of the most recent launch WITHOUT overlap permission (it
                                                  // starts after everything before it has completed)
};
std::map<std::pair<int, cudaStream_t>, Tickets> g_tickets;   // guarded by g_trail_mu

bool overlaps(const Range& a, const Range& b) { return a.lo < b.hi && b.lo < a.hi; }
bool conflicts(const LaunchFootprint& cur, const LaunchFootprint& prev) {
  for (int i = 0; i < prev.nwr; ++i) {
    for (int j = 0; j < cur.nrd; ++j) if (overlaps(prev.wr[i], cur.rd[j])) return true;   // read after write
    for (int j = 0; j < cur.nwr; ++j) if (overlaps(prev.wr[i], cur.wr[j])) return true;   // write after write
  }
  for (int i = 0; i < prev.nrd; ++i)
    for (int j = 0; j < cur.nwr; ++j) if (overlaps(prev.rd[i], cur.wr[j])) return true;   // write after read
  return false;
}
const bool g_auto_overlap = [] { const char* e = getenv("RAMDSIR_NO_AUTO_OVERLAP"); return !(e && e[0] == '1'); }();

// Returns whether `cur` may be launched with programmatic stream serialisation, and records it as the newest launch.
bool trail_push(int dev, cudaStream_t st, const LaunchFootprint& cur, bool assume_independent) {
  std::lock_guard<std::mutex> lk(g_trail_mu);
  Trail& t = g_trails[{dev, st}];
  bool ok = t.n > 0;                               // nothing of ours precedes: a plain launch
  if (ok && !assume_independent) {
    ok = g_auto_overlap;
    for (int i = 0; ok && i < t.n; ++i) ok = !conflicts(cur, t.last[i]);
  }
  for (int i = kTrailDepth - 1; i > 0; --i) t.last[i] = t.last[i - 1];
  t.last[0] = cur;
  if (t.n < kTrailDepth) ++t.n;
  return ok;
}
// Numbers the launch and, when it may overlap its predecessors (pdl), tells it which watermark to wait for.
int ticket_take(int dev, cudaStream_t st, bool pdl, MixScalars& ms) {
  std::lock_guard<std::mutex> lk(g_trail_mu);
  Tickets& t = g_tickets[{dev, st}];
  if (!t.complete_upto) {   // once per (device, stream): zeroed on the private upload stream, which alone is synchronised
    cudaStream_t up = nullptr;
    {
      std::lock_guard<std::mutex> lk2(g_mu);
      cudaStream_t& u = g_upload_stream[dev];
      if (!u) RD_CUDA(cudaStreamCreateWithFlags(&u, cudaStreamNonBlocking));
      up = u;
    }
    RD_CUDA(cudaMalloc(&t.complete_upto, sizeof(unsigned long long)));
    RD_CUDA(cudaMemsetAsync(t.complete_upto, 0, sizeof(unsigned long long), up));
    RD_CUDA(cudaStreamSynchronize(up));
  }
  const unsigned long long q = t.seq++;
  if (!pdl) t.plain_seq = q;
  ms.complete_upto = t.complete_upto;
  ms.seq = q;
  // launches q - 2, q - 1 and q all overlap their predecessors: q - 3 (and everything older) must be complete, which a
  // CTA of launch q - 2 announces (watermark >= q - 2) when it has passed its final griddepcontrol.wait
  ms.wait_upto = (pdl && t.plain_seq + 3 <= q) ? q - 2 : 0;
  return RAMDSIR_OK;
}
}  // namespace

void fused_trail_reset(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lk(g_trail_mu);
  g_trails.erase({dev, st});
}

namespace {
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
  fused_trail_reset(st);
  kern<<<grid, p.S.nthreads, p.S.smemBytes, st>>>(p.S, src, bank, pidx, rk, ratio_den, n_bank, C, out, ep, p.tables, nplanes);
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

struct StaticArgs {
  const void* src; size_t src_bytes;      // float or uint8 planes
  const float* bank; const int64_t* pidx; const int32_t* rk;
  float* out; float* clean;
  int ratio_den, n_bank, C, nplanes, nimg;
  Epilogue ep;
  unsigned flags;
};

template <class CFG, bool BANK, bool DUAL, class TIn, class EPI>
int launch_static_one(const FusedPlan& p, const StaticArgs& a, const EPI& epi, const MixScalars& ms, bool pdl, cudaStream_t st) {
  auto kern = ram_fused_static_kernel<CFG, BANK, DUAL, TIn, EPI>;
  // cudaFuncSetAttribute and occupancy are per device: cache per (instantiation, device)
  static thread_local int per_sm_dev[64] = {0};
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  int& per_sm = per_sm_dev[dev & 63];
  if (per_sm == 0) {
    RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    RD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CFG::nthreads, CFG::smemBytes));
    if (per_sm < 1) per_sm = 1;
    // The occupancy API answers 1 for any kernel that contains tcgen05.alloc (cubin attribute TCGEN05_1CTA_USED); two
    // CTAs that allocate 256 of the 512 columns each do co-reside, and registers and shared memory admit exactly two.
    if (StashX<CFG, BANK, TIn>::value && CFG::kTwoCtas) per_sm = 2;
  }
  int grid = p.sm_count * per_sm;
  if (grid > a.nplanes) grid = a.nplanes;
  // A launch that may overlap its predecessor does not need to fill every CTA slot itself: the neighbouring launches
  // take the rest.  Give every CTA the same number of planes instead (768 planes: 256 CTAs x 3 rather than 176 x 3 +
  // 120 x 2) -- fewer CTA prologues and no short CTAs; +1.0 % on the two-CTAs-per-SM kernels (measured: -0.75 % on the
  // 384^2 kernel and worse for grids far below the slot count, hence the conditions).
  if (pdl && CFG::kTwoCtas) {
    const int rounds = (a.nplanes + grid - 1) / grid;
    const int balanced = (a.nplanes + rounds - 1) / rounds;
    if (balanced * 5 >= grid * 4) grid = balanced;
  }
  MixScalars msl = ms;
  {
    const int rc = ticket_take(dev, st, pdl, msl);
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CFG::nthreads);
  cfg.dynamicSmemBytes = CFG::smemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  RD_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<const TIn*>(a.src), a.bank, a.pidx, a.rk, a.out, epi, msl,
                             (const float*)p.tables, a.clean));
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

template <class CFG, bool BANK, class TIn>
int launch_static(const FusedPlan& p, const StaticArgs& a, cudaStream_t st) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  constexpr size_t patch_bytes = (size_t)CFG::nb * CFG::nb * 4;
  auto rng = [](const void* q, size_t n) { return Range{reinterpret_cast<uintptr_t>(q), reinterpret_cast<uintptr_t>(q) + n}; };
  LaunchFootprint fp;
  fp.rd[fp.nrd++] = rng(a.src, a.src_bytes);
  MixScalars ms{a.ratio_den, a.n_bank, a.C, a.nplanes, 1.f, 0.f, a.ep.scale, a.ep.shift, nullptr, 0ull, 0ull};
  if (BANK) {
    fp.wr[fp.nwr++] = rng(a.out, (size_t)a.nplanes * patch_bytes);
    trail_push(dev, st, fp, false);       // recorded (a following mix reads this bank); itself a plain launch
    return launch_static_one<CFG, true, false, TIn, EpiClamp>(p, a, EpiClamp{0, 0, 1, 0}, ms, false, st);
  } else {
    fp.rd[fp.nrd++] = rng(a.bank, (size_t)a.n_bank * a.C * patch_bytes);
    fp.rd[fp.nrd++] = rng(a.pidx, (size_t)a.nimg * sizeof(int64_t));
    fp.rd[fp.nrd++] = rng(a.rk, (size_t)a.nimg * sizeof(int32_t));
    const size_t out_bytes = (size_t)a.nplanes * CFG::H * CFG::W * 4;
    fp.wr[fp.nwr++] = rng(a.out, out_bytes);
    if (a.clean) fp.wr[fp.nwr++] = rng(a.clean, out_bytes);
    const bool pdl = trail_push(dev, st, fp, (a.flags & RAMDSIR_FLAG_INDEPENDENT) != 0) && !(a.flags & RAMDSIR_FLAG_SERIAL);
    EpiSat es;
#ifdef RD_EXP_NOSAT
    if (false) {
#else
    if (epilogue_sat(a.ep, es, ms.dscale, ms.dc_add)) {
#endif
      if (a.clean) return launch_static_one<CFG, false, true, TIn, EpiSat>(p, a, es, ms, pdl, st);
      return launch_static_one<CFG, false, false, TIn, EpiSat>(p, a, es, ms, pdl, st);
    }
    ms.dscale = 1.f;
    ms.dc_add = 0.f;
    const EpiClamp ec{a.ep.lo, a.ep.hi, a.ep.scale, a.ep.shift};
    if (a.clean) return launch_static_one<CFG, false, true, TIn, EpiClamp>(p, a, ec, ms, pdl, st);
    return launch_static_one<CFG, false, false, TIn, EpiClamp>(p, a, ec, ms, pdl, st);
  }
}

template <bool BANK, class TIn>
int dispatch_static(const FusedPlan& p, const StaticArgs& a, cudaStream_t st) {
  if (p.static_id == 1) return launch_static<Cfg256, BANK, TIn>(p, a, st);
  return launch_static<Cfg384, BANK, TIn>(p, a, st);
}

template <bool BANK>
int dispatch_runtime(const FusedPlan& p, const float* src, const float* bank, const int64_t* pidx, const int32_t* rk,
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

bool fused_static_shape(int H, int W, int b) { return static_id_of(H, W, b) != 0; }
bool fused_u8_supported(int H, int W, int b) { return fused_static_shape(H, W, b); }
bool fused_dual_supported(int H, int W, int b, unsigned flags) {
  return fused_static_shape(H, W, b) && !(flags & RAMDSIR_FLAG_RUNTIME_SHAPE);
}

bool fused_supported(int H, int W, int b) {
  FusedShape S;
  return fused_shape_init(S, H, W, b);
}

int fused_prepare(int H, int W, int b) {
  FusedPlan* p = nullptr;
  return get_plan(H, W, b, false, &p);
}

int fused_mix_u8(const MixArgs& a, const unsigned char* src, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(a.H, a.W, a.b, false, &p);
  if (rc) return rc;
  if (!p->static_id) {
    set_error("uint8 sources are read directly only by the static fused kernels (256x256/b=25, 384x384/b=38)");
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  const StaticArgs s{src, (size_t)a.B * a.C * a.H * a.W, a.bank, a.pidx, a.rk, a.out, a.clean,
                     a.ratio_den, a.n_bank, a.C, a.B * a.C, a.B, a.ep, a.flags};
  return dispatch_static<false, unsigned char>(*p, s, st);
}

int fused_bank_u8(const unsigned char* imgs, float* bank, int N, int C, int H, int W, int b, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(H, W, b, false, &p);
  if (rc) return rc;
  if (!p->static_id) {
    set_error("uint8 sources are read directly only by the static fused kernels (256x256/b=25, 384x384/b=38)");
    return RAMDSIR_ERR_UNSUPPORTED;
  }
  const StaticArgs s{imgs, (size_t)N * C * H * W, nullptr, nullptr, nullptr, bank, nullptr, 10, 0, C, N * C, N,
                     Epilogue{0, 0, 1, 0}, 0u};
  return dispatch_static<true, unsigned char>(*p, s, st);
}

int fused_mix(const MixArgs& a, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(a.H, a.W, a.b, (a.flags & RAMDSIR_FLAG_RUNTIME_SHAPE) != 0, &p);
  if (rc) return rc;
  if (p->static_id) {
    const StaticArgs s{a.src, (size_t)a.B * a.C * a.H * a.W * 4, a.bank, a.pidx, a.rk, a.out, a.clean,
                       a.ratio_den, a.n_bank, a.C, a.B * a.C, a.B, a.ep, a.flags};
    return dispatch_static<false, float>(*p, s, st);
  }
  return dispatch_runtime<false>(*p, a.src, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, a.C, a.out, a.ep, a.B * a.C, st);
}

int fused_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, unsigned flags, cudaStream_t st) {
  FusedPlan* p = nullptr;
  int rc = get_plan(H, W, b, (flags & RAMDSIR_FLAG_RUNTIME_SHAPE) != 0, &p);
  if (rc) return rc;
  if (p->static_id) {
    const StaticArgs s{imgs, (size_t)N * C * H * W * 4, nullptr, nullptr, nullptr, bank, nullptr, 10, 0, C, N * C, N,
                       Epilogue{0, 0, 1, 0}, 0u};
    return dispatch_static<true, float>(*p, s, st);
  }
  return dispatch_runtime<true>(*p, imgs, nullptr, nullptr, nullptr, 10, 0, C, bank, Epilogue{0, 0, 1, 0}, N * C, st);
}

}  // namespace ramdsir
