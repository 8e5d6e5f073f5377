// Multi-pass RAM kernels (K2) for sm_100a: 512^2 (b = 51) and 1024^2 (b = 102).  See ram_multipass.cuh.
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>
#include "ram_internal.h"
#include "ram_fused_host.h"
#include "ram_multipass_host.h"

namespace ramdsir {

__device__ __forceinline__ float4 mp_ldg_policy(const float4* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// 16-byte global -> shared copies that need no registers (tables): issue early, wait in front of the barrier before first use
__device__ __forceinline__ void mp_copy_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mp_copy_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <bool ASYNC, int NT>
__device__ __forceinline__ void mp_copy_table(void* smem_dst, const float* gsrc, int bytes, int tid) {
  float4* t = reinterpret_cast<float4*>(smem_dst);
  const float4* g = reinterpret_cast<const float4*>(gsrc);
  for (int i = tid; i < bytes / 16; i += NT) {
    if (ASYNC) mp_copy_async16(t + i, g + i);
    else t[i] = g[i];
  }
}
__device__ __forceinline__ void mp_stg_policy(float4* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// ---- latency hiding, tuned per shape on the device (same-box A/B, tools/kbench.py) --------------
//  pf_c       pass C pulls polyphase sub-column p + 1 of x into L2 while it transforms sub-column p (and sub-column 0
//             while it loads its band rows): 1024^2 -4.1 %
//  pf_a       the same in pass A, issued behind the loads of sub-column p: 1024^2 -1.5 %, 512^2 +0.3 %
//  pf_a_next  pass A, in its last sub-column, pulls sub-column 0 of the block that will follow it on its slot
//             (blocks are dealt in index order, kMpSlots are resident): 1024^2 -1.2 %, 512^2 +-0
//  pdl        passes B and C are launched as programmatic dependents of the pass before them: their CTAs take the slots
//             its last wave vacates, load their tables (pass C: start on x) and only then wait for its results.
//             512^2 -2.2 %, 1024^2 +3.7 % (slower)
//  async_tables  the per-CTA table copy as cp.async (no registers, no wait until the barrier before first use) instead
//             of loads + stores: 1024^2 -0.8 %, 512^2 +1.8 % (slower, beside pdl)
// Deeper prefetch (two sub-columns ahead +4 %, the whole strip at entry +11 %) and the next-block prefetch in pass C
// (+2 %) were slower; the next block's band rows in pass B gave -0.3 % and were not kept; band-row stores with an
// L2 evict_last policy (to keep the workspace resident between the passes) cost +5 %.
template <class CFG>
struct MpTune {
  static constexpr bool pf_c = true;
  static constexpr bool pf_a = CFG::H > 512;
  static constexpr bool pf_a_next = CFG::H > 512;
  static constexpr bool pdl = CFG::H <= 512;
  static constexpr bool async_tables = CFG::H > 512;
};
constexpr int kMpSlots = 296;   // CTAs of pass A resident at once (2 per SM)
template <class CFG> __device__ __forceinline__ void mp_release_dependents() { if (MpTune<CFG>::pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class CFG> __device__ __forceinline__ void mp_wait_predecessor() { if (MpTune<CFG>::pdl) asm volatile("griddepcontrol.wait;" ::: "memory"); }
// One polyphase sub-column (256 rows x 256 B of this strip) into L2: two 128-byte lines a thread.
template <class CFG>
__device__ __forceinline__ void mp_prefetch_phase(const float4* x4, int strip, int p, int tid) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int line = tid + 256 * k, r = line >> 1;
    asm volatile("prefetch.global.L2 [%0];" :: "l"(x4 + (size_t)(r * CFG::HP + p) * CFG::M + strip * 16 + (line & 1) * 8));
  }
}
template <class CFG>
__device__ __forceinline__ void mp_prefetch_next_block(const float* src_chunk, int tid) {
  const unsigned nb = blockIdx.x + kMpSlots;
  if (nb < gridDim.x)
    mp_prefetch_phase<CFG>(reinterpret_cast<const float4*>(src_chunk + (size_t)(nb / CFG::nstrips) * CFG::H * CFG::W), nb % CFG::nstrips, 0, tid);
}

// ===== task bodies (one CTA each).  `zg` is the plane's band rows in the workspace: [nb][M] float4, read with
// ld.global.cg (they are produced by other CTAs and have no reuse in L1). ============================================

// ---- A: forward columns of one (plane, strip) -> zg[kh][strip columns] ------------------------------------------
template <class CFG>
__device__ __forceinline__ void mp_task_a(unsigned char* sm, const float* __restrict__ src, float4* __restrict__ zgp,
                                          int plane, int strip, unsigned long long pol_keep, const float* __restrict__ tab_ac) {
  const int tid = threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(src + (size_t)plane * CFG::H * CFG::W);
  auto ld = [&](size_t off) { return mp_ldg_policy(x4 + off, pol_keep); };
  // one polyphase sub-column; FIRST (p = 0) also brings the tables in and stores Z where the others accumulate
  auto sub_column = [&](auto first, int p) {
    constexpr bool FIRST = decltype(first)::value;
    cf2 v[16];
    if (MpTune<CFG>::pf_a_next && p + 1 == CFG::HP) mp_prefetch_next_block<CFG>(src, tid);
    m_f1_load_rows_x2<CFG>(ld, strip, p, tid, v);
    if (FIRST) {   // the tables travel behind the first sub-column's loads
      mp_copy_table<MpTune<CFG>::async_tables, 256>(sm + CFG::offTab, tab_ac, CFG::tabBytes, tid);
      mp_copy_async_wait();
      __syncthreads();
    }
    s_f1_fft_tw_x2<CFG>(sm, tid, v);
    if (MpTune<CFG>::pf_a && p + 1 < CFG::HP) mp_prefetch_phase<CFG>(x4, strip, p + 1, tid);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      s_f1_store_half_x2<CFG>(sm, tid, half, v);
      __syncthreads();
      m_f1_pruned<CFG, FIRST>(sm, p, half, tid);
      __syncthreads();
    }
  };
  sub_column(std::true_type{}, 0);
  for (int p = 1; p < CFG::HP; ++p) sub_column(std::false_type{}, p);
  // element i = tid + 256 k of the strip's band rows is row (tid >> 4) + 16 k, column tid & 15
  const float4* Zs = reinterpret_cast<const float4*>(sm + CFG::offZ) + (tid >> 4) * CFG::pitchZ + (tid & 15);
  float4* zg = zgp + strip * 16 + (size_t)(tid >> 4) * CFG::M + (tid & 15);
  constexpr int NZ = (CFG::nb + 15) / 16;
  float4 q[NZ];
#pragma unroll
  for (int k = 0; k < NZ; ++k)
    if ((tid >> 4) + 16 * k < CFG::nb) q[k] = Zs[16 * k * CFG::pitchZ];
#pragma unroll
  for (int k = 0; k < NZ; ++k)
    if ((tid >> 4) + 16 * k < CFG::nb) zg[(size_t)16 * k * CFG::M] = q[k];
}

// ---- B: KG pairs of rows (+kh, -kh) of one plane: row FFTs, band update, inverse row FFTs, in place on zg ----------
// `sm` is the CTA's work area (R, D, P at CFG::bOff*), `tab` the pass-B tables wherever the caller keeps them.
template <class CFG, bool BANK, int NT>
__device__ __forceinline__ void mp_task_b(unsigned char* sm, const float* tab, float4* __restrict__ zplane,
                                          const float* __restrict__ bank, const int64_t* __restrict__ pidx,
                                          const int32_t* __restrict__ rk, int ratio_den, int n_bank, int C,
                                          float* __restrict__ bank_out, int plane, int gplane, int kg,
                                          float dscale = 1.f, float dc_add = 0.f) {
  constexpr int KG = CFG::KG, M = CFG::M, RP = 2 * CFG::rowPitch, DP = 2 * (CFG::b + 1);
  const int tid = threadIdx.x, kh0 = kg * KG;
  const int ng = (CFG::b + 1 - kh0 < KG) ? CFG::b + 1 - kh0 : KG;     // kh values handled here
  float4* R = reinterpret_cast<float4*>(sm + CFG::bOffR);
  float2* D = reinterpret_cast<float2*>(sm + CFG::bOffD);
  // row slot r = 2 g + s  <->  band row  b + kh (s = 0)  /  b - kh (s = 1),  kh = kh0 + g  (kh = 0 uses s = 0 only)
  auto zrow = [&](int r) { const int kh = kh0 + (r >> 1); return zplane + (size_t)(CFG::b + ((r & 1) ? -kh : kh)) * M; };
  auto live = [&](int r) { return !((r & 1) && kh0 + (r >> 1) == 0); };
  // column m = tid + NT k of a row lives at slot (m % N1w) + SP (m / N1w): a per-thread base plus a constant step.
  // All loads of the task are issued before the first of them is stored (16 float4 in flight per thread at 1024^2).
  static_assert(M % NT == 0 && NT % CFG::N1w == 0, "row <-> thread mapping");
  constexpr int CPT = M / NT, slotK = CFG::SP * (NT / CFG::N1w);
  const int slot0 = (tid % CFG::N1w) + CFG::SP * (tid / CFG::N1w);
  {
    float4 q[2 * KG][CPT];
#pragma unroll
    for (int r = 0; r < 2 * KG; ++r)
      if (r < 2 * ng && live(r)) {
        const float4* z = zrow(r) + tid;
#pragma unroll
        for (int k = 0; k < CPT; ++k) q[r][k] = __ldcg(z + NT * k);
      }
#pragma unroll
    for (int r = 0; r < 2 * KG; ++r)
      if (r < 2 * ng && live(r)) {
        float4* Rr = R + (size_t)(r >> 1) * RP + (r & 1) * CFG::rowPitch + slot0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) Rr[slotK * k] = q[r][k];
      }
  }
  // stage the partner-patch rows b + kh and b - kh of every kh handled here (coalesced; consumed by the band tasks)
  float omr = 0.f;
  if (!BANK) {
    const int img = gplane / C, ch = gplane - img * C;
    const long long p = pidx[img];
    const bool ok = (p >= 0 && p < (long long)n_bank);
    omr = ok ? (float)(ratio_den - rk[img]) / (float)ratio_den * dscale : __int_as_float(0x7fc00000);
    const float* patch = bank + ((size_t)(ok ? p : 0) * C + ch) * ((size_t)CFG::nb * CFG::nb);
    float* P = reinterpret_cast<float*>(sm + CFG::bOffP);
#pragma unroll
    for (int r = 0; r < 2 * KG; ++r)
      if (r < 2 * ng) {
        const int kh = kh0 + (r >> 1);
        const float* prow = patch + (size_t)(CFG::b + ((r & 1) ? -kh : kh)) * CFG::nb;
#pragma unroll
        for (int k = 0; k < (CFG::nb + NT - 1) / NT; ++k)
          if (tid + NT * k < CFG::nb) P[r * CFG::nb + tid + NT * k] = prow[tid + NT * k];
      }
  }
  mp_copy_async_wait();      // the caller's table copy
  __syncthreads();
  for (int t = tid; t < ng * 2 * CFG::N1w; t += NT) {
    const int g = t / (2 * CFG::N1w), tt = t % (2 * CFG::N1w);
    if (live(2 * g + tt / CFG::N1w)) m_row_stage1<CFG, -1>(R + (size_t)g * RP, tab, tt);
  }
  __syncthreads();
  for (int t = tid; t < ng * 32; t += NT) {
    const int g = t / 32, tt = t % 32;
    if (live(2 * g + tt / 16)) m_row_stage2<CFG, -1>(R + (size_t)g * RP, tt);
  }
  __syncthreads();
  const size_t patch_px = (size_t)CFG::nb * CFG::nb;
  // band bins row by row (two rows at a time when a row's b + 1 bins fill at most half of the CTA): no divisions
  constexpr int P2 = (CFG::b + 1 <= NT / 2) ? NT / 2 : NT, RPI = NT / P2;
  static_assert(CFG::b + 1 <= P2, "one pass of the CTA covers a row of band bins");
  const int kwb = tid % P2, sub = tid / P2;
  if (BANK) {
    float* patch = bank_out + (size_t)plane * patch_px;
#pragma unroll
    for (int r0 = 0; r0 < 2 * KG; r0 += RPI) {
      const int r = r0 + sub, g = r >> 1;
      if (r < 2 * ng && kwb <= CFG::b && live(r)) m_band<CFG, true>(R + (size_t)g * RP, D + g * DP, tab, nullptr, patch, 0.f, 0.f, kh0 + g, r & 1, kwb);
    }
  } else {
    {
      const float* P = reinterpret_cast<const float*>(sm + CFG::bOffP);
#pragma unroll
      for (int r0 = 0; r0 < 2 * KG; r0 += RPI) {
        const int r = r0 + sub, g = r >> 1;
        if (r < 2 * ng && kwb <= CFG::b && live(r)) m_band<CFG, false>(R + (size_t)g * RP, D + g * DP, tab, P + g * 2 * CFG::nb, nullptr, omr, dc_add, kh0 + g, r & 1, kwb);
      }
    }
    __syncthreads();
    for (int i = tid; i < ng * RP; i += NT) R[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // everything off the band is zero
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2 * KG; ++r)       // row by row: no division, the row's offsets are compile-time constants
      if (r < 2 * ng && live(r)) {
#pragma unroll
        for (int k = 0; k < (CFG::nb + NT - 1) / NT; ++k)
          if (tid + NT * k < CFG::nb) m_put_v<CFG>(R + (size_t)(r >> 1) * RP, D + (r >> 1) * DP, tab, kh0 + (r >> 1), r & 1, tid + NT * k);
      }
    __syncthreads();
    for (int t = tid; t < ng * 32; t += NT) {
      const int g = t / 32, tt = t % 32;
      if (live(2 * g + tt / 16)) m_row_stage2<CFG, 1>(R + (size_t)g * RP, tt);
    }
    __syncthreads();
    for (int t = tid; t < ng * 2 * CFG::N1w; t += NT) {
      const int g = t / (2 * CFG::N1w), tt = t % (2 * CFG::N1w);
      if (live(2 * g + tt / CFG::N1w)) m_row_stage1<CFG, 1>(R + (size_t)g * RP, tab, tt);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2 * KG; ++r)
      if (r < 2 * ng && live(r)) {
        const float4* Rr = R + (size_t)(r >> 1) * RP + (r & 1) * CFG::rowPitch + slot0;
        float4* z = zrow(r) + tid;
#pragma unroll
        for (int k = 0; k < CPT; ++k) z[NT * k] = Rr[slotK * k];
      }
  }
}

// ---- C: inverse columns of one (plane, strip), + x, epilogue ---------------------------------------------------
__device__ __forceinline__ EpiClamp mp_poisoned(EpiClamp e) { e.scale = __int_as_float(0x7fc00000); return e; }
__device__ __forceinline__ EpiSat mp_poisoned(EpiSat e) { e.m = __int_as_float(0x7fc00000); return e; }   // sat() would swallow a NaN

template <class CFG, class EPI>
__device__ __forceinline__ void mp_task_c(unsigned char* sm, const float* __restrict__ src, const float4* __restrict__ zgp,
                                          const int64_t* __restrict__ pidx, int n_bank, int C, float* __restrict__ out,
                                          EPI ep, int plane, int gplane, int strip, unsigned long long pol_stream) {
  const int tid = threadIdx.x;
  const size_t plane_px = (size_t)CFG::H * CFG::W;
  const float4* x4 = reinterpret_cast<const float4*>(src + plane * plane_px);
  if (MpTune<CFG>::pf_c) mp_prefetch_phase<CFG>(x4, strip, 0, tid);
  mp_wait_predecessor<CFG>();
  {
    // the strip's band rows: element i = tid + 256 k is row (tid >> 4) + 16 k, column tid & 15.  Every load is issued
    // before the first one is stored (13 float4 in flight per thread at 1024^2) -- a rolled load/store loop pays the
    // DRAM latency once per iteration and was 40 % of this pass's stall samples.
    float4* Zs = reinterpret_cast<float4*>(sm + CFG::offZ) + (tid >> 4) * CFG::pitchZ + (tid & 15);
    const float4* zg = zgp + strip * 16 + (size_t)(tid >> 4) * CFG::M + (tid & 15);
    constexpr int NZ = (CFG::nb + 15) / 16;
    float4 q[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k)
      if ((tid >> 4) + 16 * k < CFG::nb) q[k] = __ldcg(zg + (size_t)16 * k * CFG::M);
#pragma unroll
    for (int k = 0; k < NZ; ++k)
      if ((tid >> 4) + 16 * k < CFG::nb) Zs[16 * k * CFG::pitchZ] = q[k];
  }
  {
    const long long p = pidx[gplane / C];
    if (p < 0 || p >= (long long)n_bank) ep = mp_poisoned(ep);
  }
  mp_copy_async_wait();      // the caller's table copy
  __syncthreads();
  float4* y4 = reinterpret_cast<float4*>(out + plane * plane_px);
  auto ld = [&](size_t off) { return mp_ldg_policy(x4 + off, pol_stream); };
  auto st = [&](size_t off, float4 v) { mp_stg_policy(y4 + off, v, pol_stream); };
  const int h1 = tid >> 4, c4 = strip * 16 + (tid & 15);
  for (int p = 0; p < CFG::HP; ++p) {
    cp eva[8], evb[8], oda[8], odb[8];
    if (MpTune<CFG>::pf_c && p + 1 < CFG::HP) mp_prefetch_phase<CFG>(x4, strip, p + 1, tid);
    m_inv_pruned<CFG>(sm, p, 0, tid);
    __syncthreads();
    s_inv_fft8<CFG>(sm, tid, 0, eva, evb);
    __syncthreads();
    m_inv_pruned<CFG>(sm, p, 1, tid);
    __syncthreads();
    s_inv_fft8<CFG>(sm, tid, 1, oda, odb);
    __syncthreads();
    MCombine<CFG, 0>::run(ld, st, ep, h1, p, c4, eva, evb, oda, odb);
  }
}

// ===== the three passes as separate launches over all planes (same task bodies) ====================================
template <class CFG>
__global__ void __launch_bounds__(256, 2) ram_mp_pass_a(const float* __restrict__ src, float4* __restrict__ Zg, const float* __restrict__ tab_ac) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int strip = blockIdx.x % CFG::nstrips, plane = blockIdx.x / CFG::nstrips;
  mp_release_dependents<CFG>();
  // x of a whole-batch pass is far larger than L2 and is not met again before it has been evicted: stream it, and leave
  // L2 to the band rows that pass B is about to read
  unsigned long long pol_x;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_x));
  mp_task_a<CFG>(sm, src, Zg + (size_t)plane * CFG::nb * CFG::M, plane, strip, pol_x, tab_ac);
}
template <class CFG, bool BANK>
__global__ void __launch_bounds__(CFG::bThreads) ram_mp_pass_b(float4* __restrict__ Zg, const float* __restrict__ bank,
    const int64_t* __restrict__ pidx, const int32_t* __restrict__ rk, int ratio_den, int n_bank, int C, float* __restrict__ bank_out,
    const float* __restrict__ tab_b, int plane0, float dscale, float dc_add) {
  extern __shared__ __align__(16) unsigned char sm[];
  // Planes in REVERSE launch order: pass A wrote the band rows of the last planes most recently, so pass B starts where
  // they are still in L2 -- and finishes with the first planes, which is where pass C (forward order) starts reading.
  const int tid = threadIdx.x, kg = blockIdx.x % CFG::nkg, plane = (int)(gridDim.x / CFG::nkg) - 1 - (int)(blockIdx.x / CFG::nkg);
  mp_release_dependents<CFG>();
  mp_copy_table<MpTune<CFG>::async_tables, CFG::bThreads>(sm + CFG::bOffTab, tab_b, CFG::bTabBytes, tid);
  mp_wait_predecessor<CFG>();
  // no barrier here: the task reads the tables only behind its own first barrier, so their loads overlap its row loads
  // `plane` indexes this chunk's workspace and (bank mode) output; plane0 + plane is the batch-global plane whose
  // image selects the partner and the ratio
  mp_task_b<CFG, BANK, CFG::bThreads>(sm, reinterpret_cast<const float*>(sm + CFG::bOffTab), Zg + (size_t)plane * CFG::nb * CFG::M,
                                      bank, pidx, rk, ratio_den, n_bank, C, bank_out, plane, plane0 + plane, kg, dscale, dc_add);
}
template <class CFG, class EPI>
__global__ void __launch_bounds__(256, 2) ram_mp_pass_c(const float* __restrict__ src, const float4* __restrict__ Zg,
    const int64_t* __restrict__ pidx, int n_bank, int C, float* __restrict__ out, EPI ep, const float* __restrict__ tab_ac, int plane0) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, strip = blockIdx.x % CFG::nstrips, plane = blockIdx.x / CFG::nstrips;
  mp_copy_table<MpTune<CFG>::async_tables, 256>(sm + CFG::offTab, tab_ac, CFG::tabBytes, tid);
  unsigned long long pol_stream;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  // no barrier here: the task reads the tables only behind its own first barrier, so their loads overlap its band-row loads
  mp_task_c<CFG, EPI>(sm, src, Zg + (size_t)plane * CFG::nb * CFG::M, pidx, n_bank, C, out, ep, plane, plane0 + plane, strip, pol_stream);
}


// ---------------------------------------------------------------------------------------------
namespace {
using K512 = KCfg<512, 512, 51>;
using K1024 = KCfg<1024, 1024, 102>;

struct MpPlan {
  float* tab_ac = nullptr;
  float* tab_b = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, int>, MpPlan> g_plans;  // (device, H)
std::map<int, cudaStream_t> g_upload;

template <class CFG>
int get_plan(MpPlan** out) {
  int dev = 0;
  RD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, CFG::H);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    MpPlan p;
    std::vector<float> a(CFG::tabBytes / 4), b(CFG::bTabBytes / 4);
    multipass_tables_ac<CFG>(a.data());
    multipass_tables_b<CFG>(b.data());
    cudaStream_t& up = g_upload[dev];
    if (!up) RD_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    RD_CUDA(cudaMalloc(&p.tab_ac, CFG::tabBytes));
    RD_CUDA(cudaMalloc(&p.tab_b, CFG::bTabBytes));
    RD_CUDA(cudaMemcpyAsync(p.tab_ac, a.data(), CFG::tabBytes, cudaMemcpyHostToDevice, up));
    RD_CUDA(cudaMemcpyAsync(p.tab_b, b.data(), CFG::bTabBytes, cudaMemcpyHostToDevice, up));
    RD_CUDA(cudaStreamSynchronize(up));      // a private stream: nobody else's work is synchronised
    RD_CUDA(cudaFuncSetAttribute(ram_mp_pass_a<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    RD_CUDA(cudaFuncSetAttribute(ram_mp_pass_c<CFG, EpiClamp>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    RD_CUDA(cudaFuncSetAttribute(ram_mp_pass_c<CFG, EpiSat>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::smemBytes));
    RD_CUDA(cudaFuncSetAttribute(ram_mp_pass_b<CFG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::bSmemBytes));
    RD_CUDA(cudaFuncSetAttribute(ram_mp_pass_b<CFG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::bSmemBytes));
    it = g_plans.emplace(key, p).first;
  }
  *out = &it->second;
  return RAMDSIR_OK;
}

// Band rows of the planes in flight: one pass over (up to) kChunkPlanes planes at a time, so the workspace stays below
// 1 GiB however large the batch is.  (Smaller chunks were measured slower: 9-plane chunks 42 k img/s, 18: 65 k, 36: 67 k,
// whole 192-plane batch 84 k -- the kernels are latency-bound, not DRAM-bound, and fewer, larger launches win.)
template <class CFG>
constexpr int chunk_planes() { return (int)(((size_t)1 << 30) / ((size_t)CFG::nb * CFG::M * 16)); }
template <class CFG>
size_t ws_bytes_for(int planes) {
  const int n = planes < chunk_planes<CFG>() ? planes : chunk_planes<CFG>();
  return (size_t)n * CFG::nb * CFG::M * 16;
}

template <class CFG>
int run(bool bank_mode, const MixArgs& a, float* bank_out, int nimg, void* ws, size_t ws_bytes, cudaStream_t st) {
  MpPlan* pl = nullptr;
  int rc = get_plan<CFG>(&pl);
  if (rc) return rc;
  const int planes = nimg * a.C;
  if (planes == 0) return RAMDSIR_OK;
  const size_t need = ws_bytes_for<CFG>(planes);
  if (ws == nullptr || ws_bytes < need) {
    set_error("multi-pass path needs %zu workspace bytes, got %zu", need, ws_bytes);
    return RAMDSIR_ERR_WORKSPACE;
  }
  float4* Zg = reinterpret_cast<float4*>(ws);
  const size_t plane_px = (size_t)CFG::H * CFG::W, patch_px = (size_t)CFG::nb * CFG::nb;
  // finite clip range: the two-instruction epilogue of the fused kernels (1 / (hi - lo) folded into dF in pass B)
  EpiSat es{};
  float dscale = 1.f, dc_add = 0.f;
  const bool sat = !bank_mode && epilogue_sat(a.ep, es, dscale, dc_add);
  if (!sat) { dscale = 1.f; dc_add = 0.f; }
  for (int p0 = 0; p0 < planes; p0 += chunk_planes<CFG>()) {
    const int n = (planes - p0 < chunk_planes<CFG>()) ? planes - p0 : chunk_planes<CFG>();
    // chunks start on image boundaries only if chunk_planes is a multiple of C; pass plane offsets through pointers
    const float* src = a.src + (size_t)p0 * plane_px;
    // pass A is an ordinary launch (ordered after whatever precedes it in the stream); B and C may chain on it (MpTune::pdl)
    cudaLaunchAttribute chain[1];
    chain[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    chain[0].val.programmaticStreamSerializationAllowed = 1;
    auto cfg_of = [&](int grid, int threads, size_t smem) {
      cudaLaunchConfig_t c{};
      c.gridDim = dim3(grid); c.blockDim = dim3(threads); c.dynamicSmemBytes = smem; c.stream = st;
      c.attrs = chain; c.numAttrs = MpTune<CFG>::pdl ? 1 : 0;
      return c;
    };
    ram_mp_pass_a<CFG><<<n * CFG::nstrips, 256, CFG::smemBytes, st>>>(src, Zg, pl->tab_ac);
    if (bank_mode) {
      const cudaLaunchConfig_t cb = cfg_of(n * CFG::nkg, CFG::bThreads, CFG::bSmemBytes);
      RD_CUDA(cudaLaunchKernelEx(&cb, ram_mp_pass_b<CFG, true>, Zg, (const float*)nullptr, (const int64_t*)nullptr, (const int32_t*)nullptr, 10, 0, a.C,
                                 bank_out + (size_t)p0 * patch_px, (const float*)pl->tab_b, 0, 1.f, 0.f));
      count_launch(2);
    } else {
      const cudaLaunchConfig_t cb = cfg_of(n * CFG::nkg, CFG::bThreads, CFG::bSmemBytes), cc = cfg_of(n * CFG::nstrips, 256, CFG::smemBytes);
      RD_CUDA(cudaLaunchKernelEx(&cb, ram_mp_pass_b<CFG, false>, Zg, a.bank, a.pidx, a.rk, a.ratio_den, a.n_bank, a.C, (float*)nullptr,
                                 (const float*)pl->tab_b, p0, dscale, dc_add));
      float* outp = a.out + (size_t)p0 * plane_px;
      if (sat)
        RD_CUDA(cudaLaunchKernelEx(&cc, ram_mp_pass_c<CFG, EpiSat>, src, (const float4*)Zg, a.pidx, a.n_bank, a.C, outp, es, (const float*)pl->tab_ac, p0));
      else
        RD_CUDA(cudaLaunchKernelEx(&cc, ram_mp_pass_c<CFG, EpiClamp>, src, (const float4*)Zg, a.pidx, a.n_bank, a.C, outp,
                                   EpiClamp{a.ep.lo, a.ep.hi, a.ep.scale, a.ep.shift}, (const float*)pl->tab_ac, p0));
      count_launch(3);
    }
  }
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}
}  // namespace

bool multipass_supported(int H, int W, int b) {
  return (H == 512 && W == 512 && b == 51) || (H == 1024 && W == 1024 && b == 102);
}

size_t multipass_workspace_bytes(int n_images, int C, int H, int W, int b) {
  (void)W; (void)b;
  const int planes = n_images * C;
  if (planes <= 0) return 0;
  return H == 512 ? ws_bytes_for<K512>(planes) : ws_bytes_for<K1024>(planes);
}

int multipass_mix(const MixArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (a.H == 512) return run<K512>(false, a, nullptr, a.B, ws, ws_bytes, st);
  return run<K1024>(false, a, nullptr, a.B, ws, ws_bytes, st);
}

int multipass_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, void* ws, size_t ws_bytes, cudaStream_t st) {
  MixArgs a{};
  a.src = imgs; a.C = C; a.H = H; a.W = W; a.b = b;
  if (H == 512) return run<K512>(true, a, bank, N, ws, ws_bytes, st);
  return run<K1024>(true, a, bank, N, ws, ws_bytes, st);
}

}  // namespace ramdsir
