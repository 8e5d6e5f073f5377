// Host-side tables for the multi-pass pipeline (shared by ram_multipass.cu and tools/emulate).
#pragma once
#include "ram_multipass.cuh"

namespace ramdsir {

template <class CFG>
inline void multipass_tables_ac(float* tab) {   // pass A / C: TWH (256-row sub-transform), TWP (polyphase)
  for (int i = 0; i < CFG::tabBytes / 4; ++i) tab[i] = 0.f;
  for (int h1 = 0; h1 < 16; ++h1)
    for (int k2 = 0; k2 < 16; ++k2) {
      const long n = ((long)h1 * k2) % 256;
      tab[CFG::tTWH + 2 * (h1 * CFG::twhPitch + k2)] = (float)ct::cos2pi(n, 256);
      tab[CFG::tTWH + 2 * (h1 * CFG::twhPitch + k2) + 1] = (float)-ct::sin2pi(n, 256);
    }
  for (int p = 0; p < CFG::HP; ++p)
    for (int khi = 0; khi < CFG::nb; ++khi) {
      const long n = (((long)p * (khi - CFG::b)) % CFG::H + CFG::H) % CFG::H;
      tab[CFG::tTWP + 2 * (p * CFG::nb + khi)] = (float)ct::cos2pi(n, CFG::H);
      tab[CFG::tTWP + 2 * (p * CFG::nb + khi) + 1] = (float)-ct::sin2pi(n, CFG::H);
    }
}

template <class CFG>
inline void multipass_tables_b(float* tab) {    // pass B: TWM, TW3, Q
  for (int i = 0; i < CFG::bTabBytes / 4; ++i) tab[i] = 0.f;
  for (int c1 = 0; c1 < CFG::N1w; ++c1)
    for (int k2 = 0; k2 < 16; ++k2) {
      const long n = ((long)c1 * k2) % CFG::M;
      tab[CFG::tTWM + 2 * (k2 * CFG::N1w + c1)] = (float)ct::cos2pi(n, CFG::M);
      tab[CFG::tTWM + 2 * (k2 * CFG::N1w + c1) + 1] = (float)-ct::sin2pi(n, CFG::M);
    }
  for (int g = 1; g <= 3; ++g)
    for (int kw = 0; kw <= CFG::b; ++kw) {
      tab[CFG::tTW3 + 2 * ((g - 1) * (CFG::b + 1) + kw)] = (float)ct::cos2pi((long)g * kw, CFG::W);
      tab[CFG::tTW3 + 2 * ((g - 1) * (CFG::b + 1) + kw) + 1] = (float)-ct::sin2pi((long)g * kw, CFG::W);
    }
  for (int kwi = 0; kwi < CFG::nb; ++kwi) {
    const long kw = kwi - CFG::b;
    const double r1 = ct::cos2pi(kw, CFG::W), i1 = ct::sin2pi(kw, CFG::W);
    const double r2 = ct::cos2pi(2 * kw, CFG::W), i2 = ct::sin2pi(2 * kw, CFG::W);
    const double r3 = ct::cos2pi(3 * kw, CFG::W), i3 = ct::sin2pi(3 * kw, CFG::W);
    float* q = tab + CFG::tQ + 4 * kwi;
    q[0] = (float)(1.0 - i2); q[1] = (float)r2;          // Qa = 1 + i W_W^(-2kw)
    q[2] = (float)(r1 - i3);  q[3] = (float)(i1 + r3);   // Qb = W_W^(-kw) + i W_W^(-3kw)
  }
}

}  // namespace ramdsir
