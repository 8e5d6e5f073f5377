// Fused RAM pipeline, compile-time shape variant (the BASELINE shapes: 256x256 b=25, 384x384 b=38).
//
// Same algorithm and data layout as ram_fused_core.cuh (which stays as the any-shape fallback), with
// two changes that matter for issue-slot count:
//   * H, W and the band half-width b are template constants, so every loop bound, band-validity test
//     and shared-memory offset folds to an immediate;
//   * the length-N1h stage of the column transforms is a real pruned FFT (radix-4 over h1 = a + Q b'
//     followed by Q-term twiddled sums with compile-time twiddles) instead of N1h-term direct sums:
//     ~4x fewer multiply-adds, no twiddle loads, and bit-exact zeros for constant planes, like the
//     butterflies of the reference's pocketfft.
// Column index algebra (forward; the inverse conjugates every factor), H = 16 N1h, N1h = 4 Q:
//   h = h1 + N1h j,  kh = k2 + 16 k1,  h1 = a + Q b'
//   W_H^(h kh) = W_16^(j k2) * W_H^(h1 k2) * W_N1h^(a k1) * W_4^(b' k1)
//   fft16 over j (registers) -> twiddle table TWH[h1][k2] -> exchange through E -> fft4 over b' ->
//   sum over a for the k1 in [-MM, MM-1] whose kh lies in the band.
#pragma once
#include <type_traits>
#include "ram_fused_core.cuh"
#include "ram_math_x2.cuh"

namespace ramdsir {

constexpr int s_al16(int x) { return (x + 15) & ~15; }
constexpr int s_max(int a, int b) { return a > b ? a : b; }

template <int H_, int W_, int B_>
struct SCfg {
  static constexpr int H = H_, W = W_, b = B_, nb = 2 * B_ + 1;
  static constexpr int N1h = H_ / 16, M = W_ / 4, N1w = M / 16;
  static constexpr int Q = N1h / 4;
  static constexpr int pitchZ = M + 1;             // = 1 (mod 8): row/band tasks map lanes to consecutive kh rows
  static constexpr int nthreads = 16 * N1h, nstrips = M / 16;
  static constexpr int MM = B_ / 16 + 1;          // k1 candidates [-MM, MM-1] cover [-16 MM, 16 MM - 1] >= band
  static constexpr int MMcore = (B_ + 15) / 16;   // what the shared row_inv<MMAX> expects
  static constexpr int twhPitch = 18;             // float2 per h1 row of TWH (16 used; 18 avoids a 2-way bank conflict)
  // shared memory (bytes)
  static constexpr int offZ = 0;
  static constexpr int zBytes = s_al16(nb * pitchZ * 16);
  static constexpr int eBytes = 8 * N1h * 16 * 16;
  // D[nb][dPitch] float2: an odd pitch keeps the lanes <-> consecutive kh mapping of the band / row tasks free of bank
  // conflicts (b = 25: pitch 26 float2 = 52 words put lanes l and l + 8 on the same banks)
  static constexpr int dPitch = (B_ + 1) | 1;
  static constexpr int dBytes = nb * dPitch * 8;
  static constexpr int pBytes = nb * nb * 4;      // the partner's raw patch
  static constexpr int offE = zBytes, offD = zBytes;   // D aliases E (never live together)
  static constexpr int offTab = zBytes + s_al16(s_max(eBytes, dBytes));
  // tables (float offsets inside the table block)
  static constexpr int tPT = 0;
  static constexpr int tTW2 = tPT + N1w * nb * 4;
  static constexpr int tTW3 = tTW2 + N1w * (B_ + 1) * 2;
  static constexpr int tTWH = ((tTW3 + 3 * (B_ + 1) * 2) + 3) & ~3;   // 16-byte aligned for LDS.128
  static constexpr int tTWM = tTWH + N1h * twhPitch * 2;              // [N1w][16] float2: W_M^(c1 k2)
  static constexpr int tabBytes = s_al16((tTWM + N1w * 16 * 2) * 4);
  // Landing zone of the bulk-async (TMA) copy of the partner's patch: the 16-byte-aligned superset of the patch (a
  // patch starts at any 4-byte boundary of the bank).  It lives in E behind D: E is idle between the forward and
  // the inverse column passes, which is exactly when the patch is in flight and in use.  A dedicated 10 KB region
  // would push two CTAs over the 196 KB shared-memory carve-out step and leave the L1 too small for the column
  // loads in flight (measured: -6 % serial, -10 % pipelined).
  static constexpr int prawBytes = s_al16(pBytes + 16);
  static constexpr int offPraw = offD + s_al16(dBytes);
  static_assert(s_al16(dBytes) + prawBytes <= eBytes, "D and the patch landing zone must fit inside E");
  static constexpr int offBar = offTab + tabBytes;      // two mbarriers (tables, patch) + the TMEM base address slot
  static constexpr int smemBytes = offBar + 32;
  static constexpr float invHW = (float)(1.0 / ((double)H_ * (double)W_));
  // Complex arithmetic of each phase: cf = scalar FADD / FMUL / FFMA, cp = (re, im) register pairs on FADD2 / FMUL2 /
  // FFMA2 (ram_math_x2.cuh): half the issue slots for the same FP-pipe time -- and scalar butterflies can use both FP
  // pipes while packed ones cannot (tools/microbench/pipes.cu: 321 vs 243 lane-ops/ns/SM), so packing pays only where
  // a phase is short of issue slots.  Measured phase by phase on B200 (kernel time, 256^2 with two CTAs per SM / 384^2
  // with one): second stage of the forward columns +-0 / -0.8 %, first stage of the inverse columns +-0 / -5.7 %,
  // second stage of the inverse columns + combine +1.35 % / -6.5 %, row transforms -0.3 % / -2.9 %, band bins -0.2 % /
  // +0.3 %.  (-DRD_CX_<PHASE>=cp|cf overrides a phase for such A/B runs.)
  static constexpr bool kTwoCtas = (H_ * W_ <= 256 * 256);
#ifdef RD_CX_F1P
  using cx_f1p = RD_CX_F1P;
#else
  using cx_f1p = cf;             // forward columns, second stage
#endif
#ifdef RD_CX_INVP
  using cx_invp = RD_CX_INVP;
#else
  using cx_invp = cf;            // inverse columns, first stage
#endif
#ifdef RD_CX_FFT8
  using cx_fft8 = RD_CX_FFT8;
#else
  using cx_fft8 = typename std::conditional<kTwoCtas, cp, cf>::type;   // inverse columns, second stage + combine
#endif
#ifdef RD_CX_ROW
  using cx_row = RD_CX_ROW;
#else
  using cx_row = cf;             // row transforms
#endif
#ifdef RD_CX_BAND
  using cx_band = RD_CX_BAND;
#else
  using cx_band = cf;            // band bins (one output of an N1w-point butterfly)
#endif
  static_assert(H_ % 64 == 0 && W_ % 64 == 0, "static fused path needs H, W multiples of 64");
  static_assert(2 * B_ + 1 <= M && 2 * B_ + 1 <= H_, "band must fit the packed spectrum");
  static_assert(nthreads >= 256 && nthreads <= 512, "block size out of range");
};

// ---- F1a: load 16 strided rows (h1 + N1h j), radix-16 over j, multiply by TWH[h1][k2] = W_H^(h1 k2) ------------
// Half-ahead form: rows j = 0..7 of a strip can be requested one stage early (s_f1_load_rows<0>), rows j = 8..15 at the
// start of the strip (s_f1_load_rows<8>); s_f1_fft_tw_raw_x2 then transforms and twiddles.
template <class CFG, int JBASE, class LoadF4>
RD_HD void s_f1_load_rows(LoadF4 ld, int strip, int t, float4 (&raw)[8]) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
#pragma unroll
  for (int j = 0; j < 8; ++j) raw[j] = ld((size_t)(h1 + CFG::N1h * (JBASE + j)) * CFG::M + c4);
}

// ---- F1a / F1b on column pairs (cf2, ram_math_x2.cuh): the two packed columns of a thread travel as register pairs
// through the radix-16 and the TWH twiddles at half the issue slots -- this phase is issue-bound once its loads have
// landed.  E keeps its (a.re, a.im, b.re, b.im) layout at the price of two or three MOVs per published float4.
// Measured on B200: +1.2 % (256^2), +3.2 % (384^2) of kernel time.
template <class CFG>
RD_HD void s_f1_fft_tw_x2(const unsigned char* sm, int t, cf2 (&v)[16]) {
  const int h1 = t >> 4;
  fft16x<-1>(v);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH) +
                     h1 * (CFG::twhPitch / 2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = tw[i];
    if (i > 0) v[pos16(2 * i)] = cmul2(v[pos16(2 * i)], f2_bcast(w.x), f2_bcast(w.y));
    v[pos16(2 * i + 1)] = cmul2(v[pos16(2 * i + 1)], f2_bcast(w.z), f2_bcast(w.w));
  }
}
template <class CFG, class LoadF4>
RD_HD void s_f1_load_fft_tw_x2(LoadF4 ld, const unsigned char* sm, int strip, int t, cf2 (&v)[16]) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = cf2_from(ld((size_t)(h1 + CFG::N1h * j) * CFG::M + c4));
  s_f1_fft_tw_x2<CFG>(sm, t, v);
}
template <class CFG>
RD_HD void s_f1_fft_tw_raw_x2(const unsigned char* sm, int t, const float4 (&lo)[8], const float4 (&hi)[8], cf2 (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = cf2_from(lo[j]);
    v[j + 8] = cf2_from(hi[j]);
  }
  s_f1_fft_tw_x2<CFG>(sm, t, v);
}
template <class CFG>
RD_HD void s_f1_store_half_x2(unsigned char* sm, int t, int half, const cf2 (&v)[16]) {
  float4* E = reinterpret_cast<float4*>(sm + CFG::offE) + (t >> 4) * 16 + (t & 15);
#pragma unroll
  for (int k2l = 0; k2l < 8; ++k2l) {
    const cf2 a = half ? v[pos16(8 + k2l)] : v[pos16(k2l)];
    E[k2l * CFG::N1h * 16] = make_float4(f2_lo(a.re), f2_lo(a.im), f2_hi(a.re), f2_hi(a.im));
  }
}

// second stage of the pruned N1h-point transform: out = sum_a W_N1h^(DIR a k1) y[a]
template <class CFG, int K1, int DIR, int A>
struct SumA {
  template <class CX>
  static RD_HD CX run(const CX (&y)[CFG::Q][4]) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    return SumA<CFG, K1, DIR, A - 1>::run(y) + twidN<CFG::N1h, A * K1, DIR>(y[A][r]);
  }
};
template <class CFG, int K1, int DIR>
struct SumA<CFG, K1, DIR, 0> {
  template <class CX>
  static RD_HD CX run(const CX (&y)[CFG::Q][4]) { return y[0][((K1 % 4) + 4) % 4]; }
};

// kh = k2 + 16 k1, k2 in [0, 15]: is the row inside the band whatever k2 is?  (Then the test folds away at compile
// time: no compare / branch / reconvergence point around the emit or the gather of that residue.)
template <class CFG>
RD_HD constexpr bool band_always(int k1) { return 16 * k1 >= -CFG::b && 16 * k1 + 15 <= CFG::b; }

template <class CFG, int K1>
struct F1Out {   // emits band rows kh = k2 + 16 k1 for k1 = K1 .. MM-1
  template <class CX>
  static RD_HD void run(const CX (&y)[CFG::Q][4], float2* zcol, int k2) {
    const int kh = k2 + 16 * K1;
    if (band_always<CFG>(K1) || (kh >= -CFG::b && kh <= CFG::b)) {
      const CX o = SumA<CFG, K1, -1, CFG::Q - 1>::run(y);
      zcol[(size_t)(kh + CFG::b) * CFG::pitchZ * 2] = make_float2(cx_re(o), cx_im(o));
    }
    F1Out<CFG, K1 + 1>::run(y, zcol, k2);
  }
};
template <class CFG>
struct F1Out<CFG, CFG::MM> {
  template <class CX>
  static RD_HD void run(const CX (&)[CFG::Q][4], float2*, int) {}
};

// ---- F1c: task u = (col = u & 1, c = (u >> 1) & 15, k2l = u >> 5): pruned N1h-point DFT over h1 -----
template <class CFG, class CX = typename CFG::cx_f1p>
RD_HD void s_f1_pruned(unsigned char* sm, int strip, int half, int u) {
  const int col = u & 1, c = (u >> 1) & 15, k2l = u >> 5, k2 = 8 * half + k2l;
  const float2* e = reinterpret_cast<const float2*>(sm + CFG::offE) + ((size_t)k2l * CFG::N1h * 16 + c) * 2 + col;
  CX y[CFG::Q][4];
#pragma unroll
  for (int a = 0; a < CFG::Q; ++a) {
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
      const float2 v = e[(a + CFG::Q * bq) * 32];
      y[a][bq] = cx_make<CX>(v.x, v.y);
    }
    fft4<-1>(y[a][0], y[a][1], y[a][2], y[a][3]);
  }
  float2* zcol = reinterpret_cast<float2*>(sm + CFG::offZ) + (strip * 16 + c) * 2 + col;
  F1Out<CFG, -CFG::MM>::run(y, zcol, k2);
}

// ---- inverse, first stage: task u = (col, c, k2l), k2 = 2 k2l + par: pruned inverse N1h-point DFT ----
// E'[k2l][h1][c].col = sum_k1 z'[k2 + 16 k1][c].col W_N1h^(-h1 k1)
template <class CFG, int K1>
struct InvGather {   // t[r] += in[k1] * W_N1h^(+a k1), for k1 = K1 .. MM-1
  template <int A, class CX>
  static RD_HD void run(const CX (&in)[2 * CFG::MM], CX (&t)[4]) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    t[r] = t[r] + twidN<CFG::N1h, A * K1, 1>(in[K1 + CFG::MM]);
    InvGather<CFG, K1 + 1>::template run<A>(in, t);
  }
};
template <class CFG>
struct InvGather<CFG, CFG::MM> {
  template <int A, class CX>
  static RD_HD void run(const CX (&)[2 * CFG::MM], CX (&)[4]) {}
};

template <class CFG, int A>
struct InvA {   // for a = A .. Q-1
  template <class CX>
  static RD_HD void run(const CX (&in)[2 * CFG::MM], float2* e) {
    CX t[4] = {cx_make<CX>(0.f, 0.f), cx_make<CX>(0.f, 0.f), cx_make<CX>(0.f, 0.f), cx_make<CX>(0.f, 0.f)};
    InvGather<CFG, -CFG::MM>::template run<A>(in, t);
    fft4<1>(t[0], t[1], t[2], t[3]);
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) e[(A + CFG::Q * bq) * 32] = make_float2(cx_re(t[bq]), cx_im(t[bq]));
    InvA<CFG, A + 1>::run(in, e);
  }
};
template <class CFG>
struct InvA<CFG, CFG::Q> {
  template <class CX>
  static RD_HD void run(const CX (&)[2 * CFG::MM], float2*) {}
};

template <class CFG, class CX = typename CFG::cx_invp>
RD_HD void s_inv_pruned(unsigned char* sm, int strip, int par, int u) {
  const int col = u & 1, c = (u >> 1) & 15, k2l = u >> 5, k2 = 2 * k2l + par;
  const float2* zcol = reinterpret_cast<const float2*>(sm + CFG::offZ) + (strip * 16 + c) * 2 + col;
  CX in[2 * CFG::MM];
#pragma unroll
  for (int i = 0; i < 2 * CFG::MM; ++i) {
    const int kh = k2 + 16 * (i - CFG::MM);
    CX v = cx_make<CX>(0.f, 0.f);
    if (band_always<CFG>(i - CFG::MM) || (kh >= -CFG::b && kh <= CFG::b)) {
      const float2 z = zcol[(size_t)(kh + CFG::b) * CFG::pitchZ * 2];
      v = cx_make<CX>(z.x, z.y);
    }
    in[i] = v;
  }
  float2* e = reinterpret_cast<float2*>(sm + CFG::offE) + ((size_t)k2l * CFG::N1h * 16 + c) * 2 + col;
  InvA<CFG, 0>::run(in, e);
}

// ---- inverse, second stage part 1: thread (h1, c) takes the 8 k2 of one parity from E', applies
// conj(TWH[h1][k2]) and a radix-8 inverse; results stay in registers until both parities are done ----
template <class CFG, class CX>
RD_HD void s_inv_fft8(const unsigned char* sm, int t, int par, CX (&pa)[8], CX (&pb)[8]) {
  const int h1 = t >> 4;
  const float4* E = reinterpret_cast<const float4*>(sm + CFG::offE) + h1 * 16 + (t & 15);
  const float2* tw = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH) +
                     h1 * CFG::twhPitch + par;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 v = E[m * CFG::N1h * 16];
    const float2 w = tw[2 * m];
    pa[m] = cmulc(cx_make<CX>(v.x, v.y), cf{w.x, w.y});
    pb[m] = cmulc(cx_make<CX>(v.z, v.w), cf{w.x, w.y});
  }
  fft8<1>(pa);
  fft8<1>(pb);
}

// ---- epilogue policies ------------------------------------------------------------------------------
// EpiClamp: y = clamp(x + d, lo, hi) * scale + shift, the caller epilogue as written (fundus.py:215,222-223,
//           prostate.py:188); used when a clip bound is infinite.
// EpiSat:   the same value for finite lo < hi in TWO instructions per pixel instead of four.  The band stage
//           delivers d' = (d - lo... see fused_epilogue_sat()) so that  w = sat(x * a + d')  (one FFMA.SAT)
//           is (clamp(x + d, lo, hi) - lo) / (hi - lo), and y = w * m + t with m = (hi - lo) * scale,
//           t = lo * scale + shift.  Rounding differs from EpiClamp in the last bit only.
RD_HD float rd_sat(float v) {
#if defined(__CUDA_ARCH__)
  return __saturatef(v);
#else
  return fminf(fmaxf(v, 0.f), 1.f);
#endif
}
// 1 / sqrt(v) for a normal positive v: MUFU.RSQ (2 ulp) refined by one Newton step
RD_HD float rd_rsqrt(float v) {
#if defined(__CUDA_ARCH__)
  const float r = rsqrtf(v);
#else
  const float r = 1.0f / sqrtf(v);
#endif
  return r * fmaf(-0.5f * v, r * r, 1.5f);
}
struct EpiClamp {
  float lo, hi, scale, shift;
  RD_HD float operator()(float x, float d) const { return fmaf(fminf(fmaxf(x + d, lo), hi), scale, shift); }
};
struct EpiSat {
  float a, m, t;
  RD_HD float operator()(float x, float d) const { return fmaf(rd_sat(fmaf(x, a, d)), m, t); }
};

// ---- inverse, second stage part 2: radix-2 combine of the two parities, add x, epilogue, store -----
template <class CFG, int J>
struct Combine {
  // `st(off, y, x)`: y is the augmented pixel quad, x the source quad just read (for the fused dual output)
  template <class LoadF4, class StoreF4, class EPI, class CX>
  static RD_HD void run(LoadF4 ld, StoreF4 st, EPI ep, size_t base, const CX (&eva)[8], const CX (&evb)[8],
                        const CX (&oda)[8], const CX (&odb)[8]) {
    const CX wa = twidN<16, J, 1>(oda[pos8(J)]), wb = twidN<16, J, 1>(odb[pos8(J)]);
    const CX ea = eva[pos8(J)], eb = evb[pos8(J)];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = base + (size_t)(CFG::N1h * (J + 8 * hi)) * CFG::M;
      const CX a = hi ? ea - wa : ea + wa, bq = hi ? eb - wb : eb + wb;
      const float4 x = ld(off);
      const float4 y = make_float4(ep(x.x, cx_re(a)), ep(x.y, cx_re(bq)), ep(x.z, cx_im(a)), ep(x.w, cx_im(bq)));
      st(off, y, x);
    }
    Combine<CFG, J + 1>::run(ld, st, ep, base, eva, evb, oda, odb);
  }
};
template <class CFG>
struct Combine<CFG, 8> {
  template <class LoadF4, class StoreF4, class EPI, class CX>
  static RD_HD void run(LoadF4, StoreF4, EPI, size_t, const CX (&)[8], const CX (&)[8], const CX (&)[8], const CX (&)[8]) {}
};

// Variant with the x quads of rows J = 0..3 (both halves) preloaded one stage early (xpre[2 J + hi]).
template <class CFG, int J>
struct CombineP {
  template <class LoadF4, class StoreF4, class EPI, class CX>
  static RD_HD void run(LoadF4 ld, StoreF4 st, EPI ep, size_t base, const float4 (&xpre)[8], const CX (&eva)[8],
                        const CX (&evb)[8], const CX (&oda)[8], const CX (&odb)[8]) {
    const CX wa = twidN<16, J, 1>(oda[pos8(J)]), wb = twidN<16, J, 1>(odb[pos8(J)]);
    const CX ea = eva[pos8(J)], eb = evb[pos8(J)];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = base + (size_t)(CFG::N1h * (J + 8 * hi)) * CFG::M;
      const CX a = hi ? ea - wa : ea + wa, bq = hi ? eb - wb : eb + wb;
      const float4 x = (J < 4) ? xpre[2 * (J & 3) + hi] : ld(off);
      const float4 y = make_float4(ep(x.x, cx_re(a)), ep(x.y, cx_re(bq)), ep(x.z, cx_im(a)), ep(x.w, cx_im(bq)));
      st(off, y, x);
    }
    CombineP<CFG, J + 1>::run(ld, st, ep, base, xpre, eva, evb, oda, odb);
  }
};
template <class CFG>
struct CombineP<CFG, 8> {
  template <class LoadF4, class StoreF4, class EPI, class CX>
  static RD_HD void run(LoadF4, StoreF4, EPI, size_t, const float4 (&)[8], const CX (&)[8], const CX (&)[8],
                        const CX (&)[8], const CX (&)[8]) {}
};

template <class CFG, class LoadF4>
RD_HD void s_inv_preload_x(LoadF4 ld, int strip, int t, float4 (&xpre)[8]) {
  const size_t base = (size_t)(t >> 4) * CFG::M + strip * 16 + (t & 15);
#pragma unroll
  for (int i = 0; i < 8; ++i) xpre[i] = ld(base + (size_t)(CFG::N1h * ((i >> 1) + 8 * (i & 1))) * CFG::M);
}

template <class CFG, class LoadF4, class StoreF4, class EPI, class CX>
RD_HD void s_inv_combine_store_pre(LoadF4 ld, StoreF4 st, EPI ep, int strip, int t, const float4 (&xpre)[8],
                                   const CX (&eva)[8], const CX (&evb)[8], const CX (&oda)[8], const CX (&odb)[8]) {
  const size_t base = (size_t)(t >> 4) * CFG::M + strip * 16 + (t & 15);
  CombineP<CFG, 0>::run(ld, st, ep, base, xpre, eva, evb, oda, odb);
}

template <class CFG, class LoadF4, class StoreF4, class EPI, class CX>
RD_HD void s_inv_combine_store(LoadF4 ld, StoreF4 st, EPI ep, int strip, int t, const CX (&eva)[8],
                               const CX (&evb)[8], const CX (&oda)[8], const CX (&odb)[8]) {
  const size_t base = (size_t)(t >> 4) * CFG::M + strip * 16 + (t & 15);
  Combine<CFG, 0>::run(ld, st, ep, base, eva, evb, oda, odb);
}

// =============================================================================================
// Row passes and band update, static-shape versions.  Lane mapping: consecutive lanes take consecutive
// band rows kh (row pitch = 1 float4 mod 8 -> conflict-free LDS.128/STS.128).
//
// Row index algebra (M = 16 N1w):  c4 = c1 + N1w j,  kw = k2 + 16 q
//   W_M^(c4 kw) = W_16^(j k2) * W_M^(c1 k2) * W_N1w^(c1 q)
// row_fwd does the radix-16 over j and the twiddle TWM[c1][k2]; the band task finishes with ONE
// output of an N1w-point butterfly (exact +-1, +-i for N1w = 4; radix-2 x radix-3 for N1w = 6).

template <class CFG, class CX = typename CFG::cx_row>
RD_HD void s_row_fwd(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  float4* Z = reinterpret_cast<float4*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
  CX va[16], vb[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 p = Z[CFG::N1w * j];
    va[j] = cx_make<CX>(p.x, p.y);
    vb[j] = cx_make<CX>(p.z, p.w);
  }
  fft16<-1>(va);
  fft16<-1>(vb);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWM) + c1 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = tw[i];
    CX a0 = va[pos16(2 * i)], b0 = vb[pos16(2 * i)];
    if (i > 0) {
      a0 = cmul(a0, cf{w.x, w.y});
      b0 = cmul(b0, cf{w.x, w.y});
    }
    const CX a1 = cmul(va[pos16(2 * i + 1)], cf{w.z, w.w}), b1 = cmul(vb[pos16(2 * i + 1)], cf{w.z, w.w});
    Z[CFG::N1w * (2 * i)] = make_float4(cx_re(a0), cx_im(a0), cx_re(b0), cx_im(b0));
    Z[CFG::N1w * (2 * i + 1)] = make_float4(cx_re(a1), cx_im(a1), cx_re(b1), cx_im(b1));
  }
}

// one output X[q] = sum_c a[c] W_N^(DIR c q) of an N-point DFT, butterfly-structured
template <int N, int DIR>
struct DftOut;
template <int DIR>
struct DftOut<4, DIR> {
  template <class CX>
  static RD_HD CX run(const CX (&a)[4], int q) {
    const CX t02p = a[0] + a[2], t02m = a[0] - a[2], t13p = a[1] + a[3], t13m = a[1] - a[3];
    const CX jt = cx_rot<DIR>(t13m);  // (DIR i) * t13m
    const CX e = (q & 1) ? t02m : t02p, o = (q & 1) ? jt : t13p;
    return (q & 2) ? e - o : e + o;
  }
};
// The same output with the dependence on q moved into three per-task factors: X[q] = (a0 + s2 a2) + (c -+ i s)(a1 + s2 a3),
// s2 = (-1)^q, (c, s) = (cos, sin)(pi q / 2) -- eight FFMA and no selects per vector (the band task evaluates four
// vectors with two values of q).  The factors are 0 and +-1, so every product is exact and a constant row still gives
// exact zeros off q = 0.
struct Rot4 {
  float s2, c, s;
};
RD_HD Rot4 rot4_of(int q) {
  Rot4 r;
  const float sgn = (q & 2) ? -1.f : 1.f;
  r.s2 = (q & 1) ? -1.f : 1.f;
  r.c = (q & 1) ? 0.f : sgn;
  r.s = (q & 1) ? sgn : 0.f;
  return r;
}
template <int DIR>
RD_HD cf dft4_out(const cf (&a)[4], Rot4 r) {
  const cf t1{fmaf(r.s2, a[2].re, a[0].re), fmaf(r.s2, a[2].im, a[0].im)};
  const cf t2{fmaf(r.s2, a[3].re, a[1].re), fmaf(r.s2, a[3].im, a[1].im)};
  const float sg = (DIR < 0) ? r.s : -r.s;     // forward: (c - i s) t2;  inverse: (c + i s) t2
  return cf{fmaf(sg, t2.im, fmaf(r.c, t2.re, t1.re)), fmaf(-sg, t2.re, fmaf(r.c, t2.im, t1.im))};
}
template <int DIR>
struct DftOut<6, DIR> {
  template <class CX>
  static RD_HD CX run(const CX (&a)[6], int q) {
    CX e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) e[c] = (q & 1) ? a[c] - a[c + 3] : a[c] + a[c + 3];   // W_6^(3q) = (-1)^q
    // remaining: sum_c e[c] W_6^(c q); for odd q the factor W_6^(c q) = W_3^(c q') * (W_6^c)^(odd)...
    // handle by reducing to a 3-point DFT on twiddled inputs: W_6^(c q) = W_3^(c (q >> 1)) * W_6^(c (q & 1))
    if (q & 1) {
      e[1] = twidN<6, 1, DIR>(e[1]);
      e[2] = twidN<6, 2, DIR>(e[2]);
    }
    const int r3 = (q >> 1) % 3;
    const CX s = e[1] + e[2], d = e[1] - e[2];
    if (r3 == 0) return e[0] + s;
    constexpr float kS = 0.86602540378443864676f;  // sin(2 pi / 3)
    const CX m = cx_fma_s(s, -0.5f, e[0]);
    // W_3^(DIR) = -1/2 + DIR i kS ; r3 = 1: e0 + W e1 + W^2 e2 = m + (DIR i kS) d ; r3 = 2: m - (DIR i kS) d
    const float sg = ((r3 == 1) == (DIR > 0)) ? kS : -kS;
    return cx_fma_s(cx_rot<1>(d), sg, m);     // (m.re - sg d.im, m.im + sg d.re)
  }
};

// F[kh, kw] for kw >= 0 from the row-transformed, twiddled band buffer
template <class CFG, class CX = typename CFG::cx_band>
RD_HD cf s_band_spectrum(const unsigned char* sm, int kh, int kw) {
  const float4* Z = reinterpret_cast<const float4*>(sm + CFG::offZ);
  const float* tab = reinterpret_cast<const float*>(sm + CFG::offTab);
  const int k2p = kw & 15, qp = kw >> 4;
  const int k2m = (16 - k2p) & 15, qm = (((-kw - k2m) / 16) % CFG::N1w + CFG::N1w) % CFG::N1w;
  const float4* zp = Z + (size_t)(kh + CFG::b) * CFG::pitchZ + CFG::N1w * k2p;
  const float4* zm = Z + (size_t)(CFG::b - kh) * CFG::pitchZ + CFG::N1w * k2m;
  CX pa[CFG::N1w], pb[CFG::N1w], ma[CFG::N1w], mb[CFG::N1w];
#pragma unroll
  for (int c1 = 0; c1 < CFG::N1w; ++c1) {
    const float4 p = zp[c1], q = zm[c1];
    pa[c1] = cx_make<CX>(p.x, p.y);
    pb[c1] = cx_make<CX>(p.z, p.w);
    ma[c1] = cx_make<CX>(q.x, q.y);
    mb[c1] = cx_make<CX>(q.z, q.w);
  }
  cf upa, upb, uma, umb;
  if constexpr (CFG::N1w == 4 && std::is_same<CX, cf>::value) {
    const Rot4 rp = rot4_of(qp), rm = rot4_of(qm);
    upa = dft4_out<-1>(pa, rp);
    upb = dft4_out<-1>(pb, rp);
    uma = dft4_out<-1>(ma, rm);
    umb = dft4_out<-1>(mb, rm);
  } else {
    const CX Upa = DftOut<CFG::N1w, -1>::run(pa, qp), Upb = DftOut<CFG::N1w, -1>::run(pb, qp);
    const CX Uma = DftOut<CFG::N1w, -1>::run(ma, qm), Umb = DftOut<CFG::N1w, -1>::run(mb, qm);
    upa = cf{cx_re(Upa), cx_im(Upa)};
    upb = cf{cx_re(Upb), cx_im(Upb)};
    uma = cf{cx_re(Uma), cx_im(Uma)};
    umb = cf{cx_re(Umb), cx_im(Umb)};
  }
  const cf g0{0.5f * (upa.re + uma.re), 0.5f * (upa.im - uma.im)};
  const cf g2{0.5f * (upa.im + uma.im), -0.5f * (upa.re - uma.re)};
  const cf g1{0.5f * (upb.re + umb.re), 0.5f * (upb.im - umb.im)};
  const cf g3{0.5f * (upb.im + umb.im), -0.5f * (upb.re - umb.re)};
  const float* tw3 = tab + CFG::tTW3 + 2 * kw;
  cf f = g0;
  cmac(f, g1, ldcf(tw3));
  cmac(f, g2, ldcf(tw3 + 2 * (CFG::b + 1)));
  cmac(f, g3, ldcf(tw3 + 4 * (CFG::b + 1)));
  return f;
}

// Band task of the mix kernel.  `patch` is the partner's raw [nb][nb] amplitude patch (shared memory on the device,
// landed there by the bulk-async copy); Re(ifft2) only sees the +-k average, taken here on the fly.
// `dc_add` is added to the DC bin of dF / (H W), i.e. to every pixel of delta (EpiSat's -lo / (hi - lo)).
template <class CFG>
RD_HD void s_band_mix(unsigned char* sm, const float* patch, float one_minus_ratio, float dc_add, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = s_band_spectrum<CFG>(sm, kh, kw);
  constexpr int n = CFG::nb * CFG::nb;
  const int i = khi * CFG::nb + kw + CFG::b;
  const float at = 0.5f * (patch[i] + patch[n - 1 - i]);
  const float a2 = f.re * f.re + f.im * f.im;
  cf d;
  // |F|^2 a normal float: 1 / |F| by reciprocal square root + one Newton step (~1 ulp; +0.4 % over sqrt + divide).
  // Below that the bin counts as zero amplitude and takes phase 0 like an exact zero (fundus.py:49, np.angle(0)).
  if (a2 >= 1.17549435e-38f) {
    const float s = one_minus_ratio * (at * rd_rsqrt(a2) - 1.0f) * CFG::invHW;
    d = cf{f.re * s, f.im * s};
  } else {
    d = cf{one_minus_ratio * at * CFG::invHW, 0.f};
  }
  float2* D = reinterpret_cast<float2*>(sm + CFG::offD);
  if (kw == 0) {
    if (kh == 0) {
      d.im = 0.f;
      d.re += dc_add;
    }
    D[(size_t)(CFG::nb - 1 - khi) * CFG::dPitch] = make_float2(d.re, -d.im);
  }
  D[(size_t)khi * CFG::dPitch + kw] = make_float2(d.re, d.im);
}

template <class CFG>
RD_HD void s_band_amp(const unsigned char* sm, float* patch, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = s_band_spectrum<CFG>(sm, kh, kw);
  const float a = sqrtf(f.re * f.re + f.im * f.im);
  patch[(size_t)khi * CFG::nb + (kw + CFG::b)] = a;
  patch[(size_t)(CFG::nb - 1 - khi) * CFG::nb + (CFG::b - kw)] = a;
}

// inverse rows (direct band sums into the 16 radix-16 inputs, then radix-16 inverse), lanes <-> kh
template <class CFG, class CX = typename CFG::cx_row>
RD_HD void s_row_inv(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  const float2* D = reinterpret_cast<const float2*>(sm + CFG::offD);
  const float2* Dp = D + (size_t)khi * CFG::dPitch;
  const float2* Dm = D + (size_t)(CFG::nb - 1 - khi) * CFG::dPitch;
  const float4* PT = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tPT) +
                     (size_t)c1 * CFG::nb + CFG::b;
  CX va[16], vb[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    CX aa = cx_make<CX>(0.f, 0.f), ab = cx_make<CX>(0.f, 0.f);
#pragma unroll
    for (int mi = -CFG::MM; mi < CFG::MM; ++mi) {
      const int kw = k2 + 16 * mi;               // compile-time after unrolling
      if (kw >= -CFG::b && kw <= CFG::b) {
        CX d;
        if (kw >= 0) { const float2 t = Dp[kw]; d = cx_make<CX>(t.x, t.y); }
        else { const float2 t = Dm[-kw]; d = cx_make<CX>(t.x, -t.y); }
        const float4 p = PT[kw];
        cmac(aa, d, cf{p.x, p.y});
        cmac(ab, d, cf{p.z, p.w});
      }
    }
    va[k2] = aa;
    vb[k2] = ab;
  }
  fft16<1>(va);
  fft16<1>(vb);
  float4* Z = reinterpret_cast<float4*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const CX a = va[pos16(j)], bq = vb[pos16(j)];
    Z[CFG::N1w * j] = make_float4(cx_re(a), cx_im(a), cx_re(bq), cx_im(bq));
  }
}

}  // namespace ramdsir
