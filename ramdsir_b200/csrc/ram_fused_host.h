// Host-side geometry + twiddle tables for the fused RAM pipeline (pure C++, no CUDA types),
// shared by the library (ram_fused.cu) and the CPU emulation harness (tools/emulate).
#pragma once
#include <cmath>
#include <vector>
#include "ram_fused_core.cuh"
#include "ram_fused_static.cuh"

namespace ramdsir {

constexpr int kFusedMaxMM = 4;             // band half-width b <= 64
constexpr int kFusedMaxSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100a

// Returns false when the shape is outside what the fused single-CTA pipeline supports
// (callers then use the multi-pass or the generic kernels -- never a CPU path).
inline bool fused_shape_init(FusedShape& S, int H, int W, int b) {
  if (H < 32 || W < 64 || (H % 16) || (W % 64) || b < 0) return false;
  S.H = H; S.W = W; S.b = b; S.nb = 2 * b + 1;
  S.N1h = H / 16; S.M = W / 4; S.N1w = S.M / 16;
  if (S.nb > S.M || S.nb > H) return false;          // band indices must stay distinct mod M and mod H
  S.MM = (b + 15) / 16;
  if (S.MM > kFusedMaxMM) return false;
  S.nthreads = 16 * S.N1h;
  if (S.nthreads > 512) return false;
  S.nstrips = S.M / 16;
  S.pitchZ = S.M + (S.N1w % 8);                       // bank-conflict-free row_fwd/row_inv access
  S.invHW = (float)(1.0 / ((double)H * (double)W));
  auto al16 = [](int x) { return (x + 15) & ~15; };
  int off = 0;
  S.offZ = off; off += al16(S.nb * S.pitchZ * 16);
  const int eBytes = 8 * S.N1h * 16 * 16;
  const int dBytes = S.nb * (b + 1) * 8;
  S.offE = off; S.offD = off;                          // D aliases E (never live together)
  off += al16(eBytes > dBytes ? eBytes : dBytes);
  S.offTab = off;
  int t = 0;                                           // in floats
  S.tPT = t;  t += S.N1w * S.nb * 4;
  S.tTW1 = t; t += S.N1h * S.nb * 2;
  S.tTW2 = t; t += S.N1w * (b + 1) * 2;
  S.tTW3 = t; t += 3 * (b + 1) * 2;
  S.tabBytes = al16(t * 4);
  off += S.tabBytes;
  S.smemBytes = off;
  return S.smemBytes <= kFusedMaxSmem;
}

// PT / TW2 / TW3: the row-pass and band tables shared by the runtime-shape and static-shape kernels
template <class SH>
inline void fused_tables_fill_common(const SH& S, float* tab) {
  auto put = [&](int idx, double re, double im) { tab[idx] = (float)re; tab[idx + 1] = (float)im; };
  for (int i = 0; i < S.tabBytes / 4; ++i) tab[i] = 0.f;
  for (int c1 = 0; c1 < S.N1w; ++c1)
    for (int kw = 0; kw <= S.b; ++kw) {
      const long n = ((long)c1 * kw) % S.M;
      put(S.tTW2 + 2 * (c1 * (S.b + 1) + kw), ct::cos2pi(n, S.M), -ct::sin2pi(n, S.M));
    }
  for (int g = 1; g <= 3; ++g)
    for (int kw = 0; kw <= S.b; ++kw) {
      const long n = ((long)g * kw) % S.W;
      put(S.tTW3 + 2 * ((g - 1) * (S.b + 1) + kw), ct::cos2pi(n, S.W), -ct::sin2pi(n, S.W));
    }
  // PT[c1][kw] = (Qa, Qb)[kw] * W_M^(-c1 kw);  Qa = 1 + i W_W^(-2kw),  Qb = W_W^(-kw) + i W_W^(-3kw)
  auto eW = [&](long n, double& re, double& im) {  // exp(+2 pi i n / W), n reduced exactly
    n = ((n % S.W) + S.W) % S.W;
    re = ct::cos2pi(n, S.W); im = ct::sin2pi(n, S.W);
  };
  for (int c1 = 0; c1 < S.N1w; ++c1)
    for (int kwi = 0; kwi < S.nb; ++kwi) {
      const int kw = kwi - S.b;
      double r1, i1, r2, i2, r3, i3, rm, im;
      eW(kw, r1, i1); eW(2L * kw, r2, i2); eW(3L * kw, r3, i3);
      eW(4L * c1 * kw, rm, im);                       // W_M^(-c1 kw) = exp(+2 pi i 4 c1 kw / W)
      const double qar = 1.0 - i2, qai = r2;          // 1 + i (r2 + i i2)
      const double qbr = r1 - i3, qbi = i1 + r3;      // (r1 + i i1) + i (r3 + i i3)
      float* p = tab + S.tPT + 4 * (c1 * S.nb + kwi);
      p[0] = (float)(qar * rm - qai * im); p[1] = (float)(qar * im + qai * rm);
      p[2] = (float)(qbr * rm - qbi * im); p[3] = (float)(qbr * im + qbi * rm);
    }
}

inline void fused_tables_fill(const FusedShape& S, float* tab) {
  fused_tables_fill_common(S, tab);
  // exact integer reduction of the phase index keeps the double-precision argument small
  for (int h1 = 0; h1 < S.N1h; ++h1)
    for (int khi = 0; khi < S.nb; ++khi) {
      const long n = (((long)h1 * (khi - S.b)) % S.H + S.H) % S.H;
      tab[S.tTW1 + 2 * (h1 * S.nb + khi)] = (float)ct::cos2pi(n, S.H);
      tab[S.tTW1 + 2 * (h1 * S.nb + khi) + 1] = (float)-ct::sin2pi(n, S.H);
    }
}

// static-shape kernels: common tables + TWH[h1][k2] = W_H^(h1 k2)
template <class CFG>
inline void fused_tables_fill_static(float* tab) {
  CFG S;
  fused_tables_fill_common(S, tab);
  for (int h1 = 0; h1 < CFG::N1h; ++h1)
    for (int k2 = 0; k2 < 16; ++k2) {
      const long n = ((long)h1 * k2) % CFG::H;
      tab[CFG::tTWH + 2 * (h1 * CFG::twhPitch + k2)] = (float)ct::cos2pi(n, CFG::H);
      tab[CFG::tTWH + 2 * (h1 * CFG::twhPitch + k2) + 1] = (float)-ct::sin2pi(n, CFG::H);
    }
  for (int c1 = 0; c1 < CFG::N1w; ++c1)
    for (int k2 = 0; k2 < 16; ++k2) {
      const long n = ((long)c1 * k2) % CFG::M;
      tab[CFG::tTWM + 2 * (c1 * 16 + k2)] = (float)ct::cos2pi(n, CFG::M);
      tab[CFG::tTWM + 2 * (c1 * 16 + k2) + 1] = (float)-ct::sin2pi(n, CFG::M);
    }
}

// EpiSat parameters of a finite clip range (see ram_fused_static.cuh): computed in double, rounded once.
inline bool epilogue_sat(const Epilogue& e, EpiSat& s, float& dscale, float& dc_add) {
  if (!(std::isfinite(e.lo) && std::isfinite(e.hi) && e.hi > e.lo)) return false;
  const double span = (double)e.hi - (double)e.lo;
  s.a = (float)(1.0 / span);
  s.m = (float)(span * (double)e.scale);
  s.t = (float)((double)e.lo * (double)e.scale + (double)e.shift);
  dscale = s.a;
  dc_add = (float)(-(double)e.lo / span);
  return std::isfinite(s.a) && std::isfinite(s.m) && std::isfinite(s.t);
}

}  // namespace ramdsir
