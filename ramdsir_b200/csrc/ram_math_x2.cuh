// Packed-pair arithmetic for the in-register FFT stages: two independent fp32 lanes (the two complex columns a, b
// that a thread owns) travel in one 64-bit register pair and are processed by Blackwell's FADD2 / FMUL2 / FFMA2
// (PTX add/sub/mul/fma .f32x2, sm_100+).  One issue slot does the work of two scalar instructions; the FP pipe
// time is unchanged (measured in tools/microbench/pipes.cu), so this frees issue slots in exactly the phases that
// are issue-bound.  ptxas folds scalar broadcasts (`R.F32`), immediates and negations into the packed instruction's
// operand modifiers, so twiddles stay scalar (no duplicated tables, no MOVs).  A loaded pixel quad (x0, x1, x2, x3)
// IS the packed layout: re = (x0, x1), im = (x2, x3).
// The host definitions compute the same IEEE results lane by lane, so tools/emulate can check the kernels.
#pragma once
#include "ram_math.cuh"

namespace ramdsir {

#if defined(__CUDA_ARCH__)
struct f2 {
  unsigned long long v;
};
RD_HD f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
RD_HD float f2_lo(f2 a) { float lo; asm("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(lo) : "l"(a.v)); return lo; }
RD_HD float f2_hi(f2 a) { float hi; asm("{ .reg .b32 t; mov.b64 {t, %0}, %1; }" : "=f"(hi) : "l"(a.v)); return hi; }
RD_HD f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
RD_HD f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
RD_HD f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
RD_HD f2 xfma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct alignas(8) f2 {
  float lo, hi;
};
RD_HD f2 f2_make(float lo, float hi) { return f2{lo, hi}; }
RD_HD float f2_lo(f2 a) { return a.lo; }
RD_HD float f2_hi(f2 a) { return a.hi; }
RD_HD f2 operator+(f2 a, f2 b) { return f2{a.lo + b.lo, a.hi + b.hi}; }
RD_HD f2 operator-(f2 a, f2 b) { return f2{a.lo - b.lo, a.hi - b.hi}; }
RD_HD f2 operator*(f2 a, f2 b) { return f2{a.lo * b.lo, a.hi * b.hi}; }
RD_HD f2 xfma(f2 a, f2 b, f2 c) { return f2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
#endif
RD_HD f2 f2_bcast(float c) { return f2_make(c, c); }
RD_HD f2 f2_neg(f2 a) { return f2_bcast(0.f) - a; }

// complex number whose components are packed pairs: element k of (re, im) belongs to column k (0 = a, 1 = b)
struct cf2 {
  f2 re, im;
};
RD_HD cf2 operator+(cf2 a, cf2 b) { return cf2{a.re + b.re, a.im + b.im}; }
RD_HD cf2 operator-(cf2 a, cf2 b) { return cf2{a.re - b.re, a.im - b.im}; }
// a * w, w given as packed (broadcast) twiddle
RD_HD cf2 cmul2(cf2 a, f2 wr, f2 wi) { return cf2{xfma(a.re, wr, f2_neg(a.im * wi)), xfma(a.re, wi, a.im * wr)}; }
// memory layout of a cf2: float4 (re.a, re.b, im.a, im.b)  ("RRII")
RD_HD cf2 cf2_from(float4 q) { return cf2{f2_make(q.x, q.y), f2_make(q.z, q.w)}; }
RD_HD float4 cf2_to(cf2 a) { return make_float4(f2_lo(a.re), f2_hi(a.re), f2_lo(a.im), f2_hi(a.im)); }

template <int DIR>
RD_HD void fft4x(cf2& a0, cf2& a1, cf2& a2, cf2& a3) {
  const cf2 s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
  a0 = s02 + s13;
  a2 = s02 - s13;
  if (DIR < 0) {   // a1 = d02 - i d13, a3 = d02 + i d13
    a1 = cf2{d02.re + d13.im, d02.im - d13.re};
    a3 = cf2{d02.re - d13.im, d02.im + d13.re};
  } else {
    a1 = cf2{d02.re - d13.im, d02.im + d13.re};
    a3 = cf2{d02.re + d13.im, d02.im - d13.re};
  }
}

// multiply by exp(DIR 2 pi i M / N), compile-time constants become FFMA2 / FMUL2 immediates
template <int N, int M, int DIR>
RD_HD cf2 twidNx(cf2 a) {
  constexpr int m = ((M % N) + N) % N;
  if (m == 0) return a;
  if (4 * m == N) return (DIR < 0) ? cf2{a.im, f2_neg(a.re)} : cf2{f2_neg(a.im), a.re};
  if (2 * m == N) return cf2{f2_neg(a.re), f2_neg(a.im)};
  if (4 * m == 3 * N) return (DIR < 0) ? cf2{f2_neg(a.im), a.re} : cf2{a.im, f2_neg(a.re)};
  constexpr float c = (float)ct::cos2pi(m, N);
  constexpr float s = (float)((DIR < 0) ? -ct::sin2pi(m, N) : ct::sin2pi(m, N));
  // (re c - im s, re s + im c)
  return cf2{xfma(a.im, f2_bcast(-s), a.re * f2_bcast(c)), xfma(a.im, f2_bcast(c), a.re * f2_bcast(s))};
}

template <int DIR>
RD_HD void fft16x(cf2 (&v)[16]) {
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fft4x<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  v[1 + 4 * 1] = twidNx<16, 1, DIR>(v[1 + 4 * 1]);
  v[1 + 4 * 2] = twidNx<16, 2, DIR>(v[1 + 4 * 2]);
  v[1 + 4 * 3] = twidNx<16, 3, DIR>(v[1 + 4 * 3]);
  v[2 + 4 * 1] = twidNx<16, 2, DIR>(v[2 + 4 * 1]);
  v[2 + 4 * 2] = twidNx<16, 4, DIR>(v[2 + 4 * 2]);
  v[2 + 4 * 3] = twidNx<16, 6, DIR>(v[2 + 4 * 3]);
  v[3 + 4 * 1] = twidNx<16, 3, DIR>(v[3 + 4 * 1]);
  v[3 + 4 * 2] = twidNx<16, 6, DIR>(v[3 + 4 * 2]);
  v[3 + 4 * 3] = twidNx<16, 9, DIR>(v[3 + 4 * 3]);
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) fft4x<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// =================================================================================================================
// Complex numbers as ONE register pair (re, im).  ptxas folds the half swap (.LO_HI), per-half negation (.NP / .PN),
// scalar broadcasts (R.F32) and immediates of the expressions below into the operand modifiers of FADD2 / FMUL2 /
// FFMA2, so that  a + b,  a +- i b,  a * w  (w scalar pair or compile-time constant) cost 1, 1 and 2 issue slots
// instead of 2, 2 and 4 -- with the natural (re, im) adjacency of every float2 / float4 the kernels keep in shared
// memory, hence without any packing moves.  Host definitions: the same IEEE operations lane by lane.
struct cp {
  f2 v;
};
RD_HD cp cp_make(float re, float im) { return cp{f2_make(re, im)}; }
RD_HD float cp_re(cp a) { return f2_lo(a.v); }
RD_HD float cp_im(cp a) { return f2_hi(a.v); }
RD_HD cp operator+(cp a, cp b) { return cp{a.v + b.v}; }
RD_HD cp operator-(cp a, cp b) { return cp{a.v - b.v}; }
RD_HD cp cp_conj(cp a) { return cp_make(cp_re(a), -cp_im(a)); }
RD_HD cp cp_neg(cp a) { return cp_make(-cp_re(a), -cp_im(a)); }
// (DIR i) * b as a value: DIR < 0: (b.im, -b.re);  DIR > 0: (-b.im, b.re)
template <int DIR>
RD_HD cp cp_rot(cp b) { return (DIR < 0) ? cp_make(cp_im(b), -cp_re(b)) : cp_make(-cp_im(b), cp_re(b)); }
// a * w and a * conj(w), w = (wr, wi) scalars
RD_HD cp cmul(cp a, cf w) { return cp{xfma(cp_rot<1>(a).v, f2_bcast(w.im), a.v * f2_bcast(w.re))}; }
RD_HD cp cmulc(cp a, cf w) { return cp{xfma(cp_rot<-1>(a).v, f2_bcast(w.im), a.v * f2_bcast(w.re))}; }
// acc += a * w
RD_HD void cmac(cp& acc, cp a, cf w) {
  acc.v = xfma(a.v, f2_bcast(w.re), acc.v);
  acc.v = xfma(cp_rot<1>(a).v, f2_bcast(w.im), acc.v);
}

template <> struct CxMake<cp> { static RD_HD cp make(float re, float im) { return cp_make(re, im); } };
RD_HD float cx_re(cp a) { return cp_re(a); }
RD_HD float cx_im(cp a) { return cp_im(a); }
template <int DIR> RD_HD cp cx_rot(cp b) { return cp_rot<DIR>(b); }
RD_HD cp cx_fma_s(cp a, float s, cp c) { return cp{xfma(a.v, f2_bcast(s), c.v)}; }

template <int DIR>
RD_HD void fft4(cp& a0, cp& a1, cp& a2, cp& a3) {
  const cp s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
  const cp jd = cp_rot<DIR>(d13);
  a0 = s02 + s13;
  a2 = s02 - s13;
  a1 = d02 + jd;
  a3 = d02 - jd;
}

// multiply by exp(DIR 2 pi i M / N), compile-time constants become FMUL2 / FFMA2 immediates
template <int N, int M, int DIR>
RD_HD cp twidN(cp a) {
  constexpr int m = ((M % N) + N) % N;
  if (m == 0) return a;
  if (4 * m == N) return cp_rot<DIR>(a);
  if (2 * m == N) return cp_neg(a);
  if (4 * m == 3 * N) return cp_rot<-DIR>(a);
  constexpr float c = (float)ct::cos2pi(m, N);
  constexpr float s = (float)((DIR < 0) ? -ct::sin2pi(m, N) : ct::sin2pi(m, N));
  return cp{xfma(cp_rot<1>(a).v, f2_bcast(s), a.v * f2_bcast(c))};    // (re c - im s, im c + re s)
}

template <int DIR>
RD_HD void fft16(cp (&v)[16]) {
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fft4<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  v[1 + 4 * 1] = twidN<16, 1, DIR>(v[1 + 4 * 1]);
  v[1 + 4 * 2] = twidN<16, 2, DIR>(v[1 + 4 * 2]);
  v[1 + 4 * 3] = twidN<16, 3, DIR>(v[1 + 4 * 3]);
  v[2 + 4 * 1] = twidN<16, 2, DIR>(v[2 + 4 * 1]);
  v[2 + 4 * 2] = twidN<16, 4, DIR>(v[2 + 4 * 2]);
  v[2 + 4 * 3] = twidN<16, 6, DIR>(v[2 + 4 * 3]);
  v[3 + 4 * 1] = twidN<16, 3, DIR>(v[3 + 4 * 1]);
  v[3 + 4 * 2] = twidN<16, 6, DIR>(v[3 + 4 * 2]);
  v[3 + 4 * 3] = twidN<16, 9, DIR>(v[3 + 4 * 3]);
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) fft4<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

template <int DIR>
RD_HD void fft8(cp (&v)[8]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const cp a = v[n] + v[n + 4], d = v[n] - v[n + 4];
    v[n] = a;
    v[n + 4] = d;
  }
  v[5] = twidN<8, 1, DIR>(v[5]);
  v[6] = twidN<8, 2, DIR>(v[6]);
  v[7] = twidN<8, 3, DIR>(v[7]);
  fft4<DIR>(v[0], v[1], v[2], v[3]);
  fft4<DIR>(v[4], v[5], v[6], v[7]);
}

}  // namespace ramdsir
