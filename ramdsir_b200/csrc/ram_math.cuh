// Small complex / radix-16 helpers shared by the RAM kernels (sm_100a) and by the host
// emulation harness in tools/emulate (same code compiled for the CPU, so the index
// algebra can be checked without a GPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RD_HD __host__ __device__ __forceinline__
#define RD_CX __host__ __device__ constexpr
#else
#define RD_HD inline
#define RD_CX constexpr
#endif

namespace ramdsir {

struct cf {
  float re, im;
};

RD_HD cf operator+(cf a, cf b) { return cf{a.re + b.re, a.im + b.im}; }
RD_HD cf operator-(cf a, cf b) { return cf{a.re - b.re, a.im - b.im}; }
RD_HD cf cmul(cf a, cf w) { return cf{a.re * w.re - a.im * w.im, a.re * w.im + a.im * w.re}; }
// a * conj(w)
RD_HD cf cmulc(cf a, cf w) { return cf{a.re * w.re + a.im * w.im, a.im * w.re - a.re * w.im}; }
RD_HD cf conjf2(cf a) { return cf{a.re, -a.im}; }
// acc += a * w
RD_HD void cmac(cf& acc, cf a, cf w) {
  acc.re = fmaf(a.re, w.re, acc.re);
  acc.re = fmaf(-a.im, w.im, acc.re);
  acc.im = fmaf(a.re, w.im, acc.im);
  acc.im = fmaf(a.im, w.re, acc.im);
}
// acc += a * conj(w)
RD_HD void cmacc(cf& acc, cf a, cf w) {
  acc.re = fmaf(a.re, w.re, acc.re);
  acc.re = fmaf(a.im, w.im, acc.re);
  acc.im = fmaf(a.im, w.re, acc.im);
  acc.im = fmaf(-a.re, w.im, acc.im);
}

// Generic accessors shared by the scalar (cf) and the register-pair (cp, ram_math_x2.cuh) complex types, so that a
// phase body can be instantiated with either.
template <class CX> struct CxMake;
template <> struct CxMake<cf> { static RD_HD cf make(float re, float im) { return cf{re, im}; } };
template <class CX> RD_HD CX cx_make(float re, float im) { return CxMake<CX>::make(re, im); }
RD_HD float cx_re(cf a) { return a.re; }
RD_HD float cx_im(cf a) { return a.im; }
template <int DIR> RD_HD cf cx_rot(cf b) { return (DIR < 0) ? cf{b.im, -b.re} : cf{-b.im, b.re}; }   // (DIR i) * b
RD_HD cf cx_fma_s(cf a, float s, cf c) { return cf{c.re + s * a.re, c.im + s * a.im}; }               // a * s + c

// 4-point DFT in place. DIR = -1: forward kernel exp(-2 pi i nk/4); DIR = +1: inverse (unnormalised).
template <int DIR>
RD_HD void fft4(cf& a0, cf& a1, cf& a2, cf& a3) {
  cf s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
  cf jd = (DIR < 0) ? cf{d13.im, -d13.re} : cf{-d13.im, d13.re};  // (DIR * i) * d13
  a0 = s02 + s13;
  a2 = s02 - s13;
  a1 = d02 + jd;
  a3 = d02 - jd;
}

// multiply by exp(DIR * 2 pi i * M / 16) with compile-time M: immediates only, no table
template <int M, int DIR>
RD_HD cf twid16(cf a) {
  constexpr int m = ((M % 16) + 16) % 16;
  constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
  constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
  constexpr float R2 = 0.70710678118654752f;
  constexpr float cs[16] = {1.f, C1, R2, S1, 0.f, -S1, -R2, -C1, -1.f, -C1, -R2, -S1, 0.f, S1, R2, C1};
  constexpr float sn[16] = {0.f, S1, R2, C1, 1.f, C1, R2, S1, 0.f, -S1, -R2, -C1, -1.f, -C1, -R2, -S1};
  if (m == 0) return a;
  if (m == 4) return (DIR < 0) ? cf{a.im, -a.re} : cf{-a.im, a.re};
  if (m == 8) return cf{-a.re, -a.im};
  if (m == 12) return (DIR < 0) ? cf{-a.im, a.re} : cf{a.im, -a.re};
  const float c = cs[m], s = (DIR < 0) ? -sn[m] : sn[m];
  return cf{a.re * c - a.im * s, a.re * s + a.im * c};
}

// 16-point DFT in place on v[0..15] (input index n). On return X[k] sits at v[pos16(k)].
RD_HD constexpr int pos16(int k) { return (k >> 2) + 4 * (k & 3); }

template <int DIR>
RD_HD void fft16(cf (&v)[16]) {
  // n = n1 + 4 n2, k = 4 k1 + k2
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fft4<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // v[n1 + 4 k2] *= W16^(n1 k2)
  v[1 + 4 * 1] = twid16<1, DIR>(v[1 + 4 * 1]);
  v[1 + 4 * 2] = twid16<2, DIR>(v[1 + 4 * 2]);
  v[1 + 4 * 3] = twid16<3, DIR>(v[1 + 4 * 3]);
  v[2 + 4 * 1] = twid16<2, DIR>(v[2 + 4 * 1]);
  v[2 + 4 * 2] = twid16<4, DIR>(v[2 + 4 * 2]);
  v[2 + 4 * 3] = twid16<6, DIR>(v[2 + 4 * 3]);
  v[3 + 4 * 1] = twid16<3, DIR>(v[3 + 4 * 1]);
  v[3 + 4 * 2] = twid16<6, DIR>(v[3 + 4 * 2]);
  v[3 + 4 * 3] = twid16<9, DIR>(v[3 + 4 * 3]);
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) fft4<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// ---------------------------------------------------------------------------------------------
// Compile-time twiddles exp(DIR * 2 pi i * M / N) for arbitrary N (used by the pruned N1 stages).
// Taylor series in double on an argument reduced to |x| <= pi/4: exact to double rounding.
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;
RD_CX double sin_small(double x) {  // |x| <= pi/4
  double term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
RD_CX double cos_small(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
// cos / sin of 2 pi m / n with exact octant reduction on the integers
RD_CX double cos2pi(long m, long n) {
  m = ((m % n) + n) % n;
  if (2 * m > n) m = n - m;                 // cos is even about pi
  if (4 * m > n) return -cos2pi(n - 2 * m, 2 * n) ;  // cos(x) = -cos(pi - x); pi - x = 2 pi (n - 2m) / (2n)
  if (8 * m > n) return sin_small(2 * kPi * (double)(n - 4 * m) / (double)(4 * n));  // cos(x) = sin(pi/2 - x)
  return cos_small(2 * kPi * (double)m / (double)n);
}
RD_CX double sin2pi(long m, long n) { return cos2pi(4 * m - n, 4 * n); }  // sin(x) = cos(x - pi/2)
}  // namespace ct

template <int N, int M, int DIR>
RD_HD cf twidN(cf a) {
  constexpr int m = ((M % N) + N) % N;
  if (m == 0) return a;
  if (4 * m == N) return (DIR < 0) ? cf{a.im, -a.re} : cf{-a.im, a.re};
  if (2 * m == N) return cf{-a.re, -a.im};
  if (4 * m == 3 * N) return (DIR < 0) ? cf{-a.im, a.re} : cf{a.im, -a.re};
  constexpr float c = (float)ct::cos2pi(m, N);
  constexpr float s = (float)((DIR < 0) ? -ct::sin2pi(m, N) : ct::sin2pi(m, N));
  return cf{a.re * c - a.im * s, a.re * s + a.im * c};
}

// 8-point DFT in place; X[k] ends at v[pos8(k)].
RD_HD constexpr int pos8(int k) { return (k >> 1) + 4 * (k & 1); }

template <int DIR>
RD_HD void fft8(cf (&v)[8]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const cf a = v[n] + v[n + 4], d = v[n] - v[n + 4];
    v[n] = a;
    v[n + 4] = d;
  }
  v[5] = twidN<8, 1, DIR>(v[5]);
  v[6] = twidN<8, 2, DIR>(v[6]);
  v[7] = twidN<8, 3, DIR>(v[7]);
  fft4<DIR>(v[0], v[1], v[2], v[3]);  // X[2k]
  fft4<DIR>(v[4], v[5], v[6], v[7]);  // X[2k+1]
}

}  // namespace ramdsir
