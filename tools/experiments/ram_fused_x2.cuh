// Fused RAM pipeline, "x2" variant: the static float4 pipeline of ram_fused_static.cuh with PACKED arithmetic.
// A thread still owns the two complex columns za = x[4c] + i x[4c+2], zb = x[4c+1] + i x[4c+3]; they now live
// in f32x2 register pairs (cf2: re = (za.re, zb.re), im = (za.im, zb.im)) and every butterfly / twiddle is
// one FADD2 / FMUL2 / FFMA2 instead of two scalar instructions.  The loaded float4 (x0, x1, x2, x3) IS the
// packed layout (re, im) = ((x0, x1), (x2, x3)), and so is the stored delta quad.
// Shared-memory buffers keep their sizes; their element layout changes from (a.re, a.im, b.re, b.im) to
// (a.re, b.re, a.im, b.im) ("RRII").  The exchange stages that used to be one-column-per-task become
// both-columns-per-task, split across two threads by radix-4 residue parity (forward) / by `a` half (inverse),
// so all threads stay busy.  Same factorisation and table contents as ram_fused_static.cuh.
#pragma once
#include "ram_fused_static.cuh"
#include "ram_math_x2.cuh"

namespace ramdsir {

template <int H_, int W_, int B_>
struct XCfg : SCfg<H_, W_, B_> {
  using S = SCfg<H_, W_, B_>;
  static constexpr int tw4Pitch = 17;                                  // float4 per h1 row of TWH2
  static constexpr int tPT = 0;                                        // [N1w][nb] float4 RRII: (Pa.re, Pb.re, Pa.im, Pb.im)
  static constexpr int tTW3 = tPT + S::N1w * S::nb * 4;                // [3][b+1] float2
  static constexpr int tTWH2 = ((tTW3 + 3 * (B_ + 1) * 2) + 3) & ~3;   // [N1h][17] float4: (w.re, w.re, w.im, w.im)
  static constexpr int tTWM2 = tTWH2 + S::N1h * tw4Pitch * 4;          // [N1w][16] float4
  static constexpr int tabBytes = s_al16((tTWM2 + S::N1w * 16 * 4) * 4);
  static constexpr int smemBytes = S::offTab + tabBytes;
  static constexpr int ntask = 128;                                    // exchange-stage tasks per parity: 8 k2 x 16 c
};

// ---- F1a ------------------------------------------------------------------------------------------
template <class CFG, class LoadF4>
RD_HD void x_f1_load_fft_tw(LoadF4 ld, const unsigned char* sm, int strip, int t, cf2 (&v)[16]) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = cf2_from(ld((size_t)(h1 + CFG::N1h * j) * CFG::M + c4));
  fft16x<-1>(v);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH2) +
                     h1 * CFG::tw4Pitch;
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float4 w = tw[k2];
    v[pos16(k2)] = cmul2(v[pos16(k2)], f2_make(w.x, w.y), f2_make(w.z, w.w));
  }
}

template <class CFG>
RD_HD void x_f1_store_half(unsigned char* sm, int t, int half, const cf2 (&v)[16]) {
  float4* E = reinterpret_cast<float4*>(sm + CFG::offE) + (t >> 4) * 16 + (t & 15);
#pragma unroll
  for (int k2l = 0; k2l < 8; ++k2l) E[k2l * CFG::N1h * 16] = cf2_to(half ? v[pos16(8 + k2l)] : v[pos16(k2l)]);
}

template <class CFG, int K1, int DIR, int SLOT, int A>
struct XSumA {
  static RD_HD cf2 run(const cf2 (&y)[CFG::Q][2]) {
    return XSumA<CFG, K1, DIR, SLOT, A - 1>::run(y) + twidNx<CFG::N1h, A * K1, DIR>(y[A][SLOT]);
  }
};
template <class CFG, int K1, int DIR, int SLOT>
struct XSumA<CFG, K1, DIR, SLOT, 0> {
  static RD_HD cf2 run(const cf2 (&y)[CFG::Q][2]) { return y[0][SLOT]; }
};

template <class CFG, int RPAR, int K1>
struct XF1Out {
  static RD_HD void run(const cf2 (&y)[CFG::Q][2], float4* zcol, int k2) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    if (r == RPAR || r == RPAR + 2) {
      const int kh = k2 + 16 * K1;
      if (kh >= -CFG::b && kh <= CFG::b)
        zcol[(size_t)(kh + CFG::b) * CFG::pitchZ] = cf2_to(XSumA<CFG, K1, -1, (r == RPAR) ? 0 : 1, CFG::Q - 1>::run(y));
    }
    XF1Out<CFG, RPAR, K1 + 1>::run(y, zcol, k2);
  }
};
template <class CFG, int RPAR>
struct XF1Out<CFG, RPAR, CFG::MM> {
  static RD_HD void run(const cf2 (&)[CFG::Q][2], float4*, int) {}
};

// F1c: task u in [0, 128) of residue parity RPAR: c = u & 15, k2l = u >> 4; both columns, residues {RPAR, RPAR+2}
template <class CFG, int RPAR>
RD_HD void x_f1_pruned(unsigned char* sm, int strip, int half, int u) {
  const int c = u & 15, k2l = u >> 4, k2 = 8 * half + k2l;
  const float4* e = reinterpret_cast<const float4*>(sm + CFG::offE) + (size_t)k2l * CFG::N1h * 16 + c;
  cf2 y[CFG::Q][2];
#pragma unroll
  for (int a = 0; a < CFG::Q; ++a) {
    const cf2 i0 = cf2_from(e[(a + CFG::Q * 0) * 16]), i1 = cf2_from(e[(a + CFG::Q * 1) * 16]);
    const cf2 i2 = cf2_from(e[(a + CFG::Q * 2) * 16]), i3 = cf2_from(e[(a + CFG::Q * 3) * 16]);
    if (RPAR == 0) {
      const cf2 t02 = i0 + i2, t13 = i1 + i3;
      y[a][0] = t02 + t13;                                   // r = 0
      y[a][1] = t02 - t13;                                   // r = 2
    } else {
      const cf2 t02 = i0 - i2, t13 = i1 - i3;                // r = 1: t02 - i t13 ; r = 3: t02 + i t13
      y[a][0] = cf2{t02.re + t13.im, t02.im - t13.re};
      y[a][1] = cf2{t02.re - t13.im, t02.im + t13.re};
    }
  }
  float4* zcol = reinterpret_cast<float4*>(sm + CFG::offZ) + strip * 16 + c;
  XF1Out<CFG, RPAR, -CFG::MM>::run(y, zcol, k2);
}

// ---- rows and band ----------------------------------------------------------------------------------
template <class CFG>
RD_HD void x_row_fwd(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  float4* Z = reinterpret_cast<float4*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
  cf2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = cf2_from(Z[CFG::N1w * j]);
  fft16x<-1>(v);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWM2) + c1 * 16;
  Z[0] = cf2_to(v[pos16(0)]);
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float4 w = tw[k2];
    Z[CFG::N1w * k2] = cf2_to(cmul2(v[pos16(k2)], f2_make(w.x, w.y), f2_make(w.z, w.w)));
  }
}

template <int N, int DIR>
struct XDftOut;
template <int DIR>
struct XDftOut<4, DIR> {
  static RD_HD cf2 run(const cf2 (&a)[4], int q) {
    const cf2 t02p = a[0] + a[2], t02m = a[0] - a[2], t13p = a[1] + a[3], t13m = a[1] - a[3];
    const cf2 jt = (DIR < 0) ? cf2{t13m.im, f2_neg(t13m.re)} : cf2{f2_neg(t13m.im), t13m.re};
    const cf2 e = (q & 1) ? t02m : t02p, o = (q & 1) ? jt : t13p;
    return (q & 2) ? e - o : e + o;
  }
};
template <int DIR>
struct XDftOut<6, DIR> {
  static RD_HD cf2 run(const cf2 (&a)[6], int q) {
    cf2 e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) e[c] = (q & 1) ? a[c] - a[c + 3] : a[c] + a[c + 3];
    if (q & 1) {
      e[1] = twidNx<6, 1, DIR>(e[1]);
      e[2] = twidNx<6, 2, DIR>(e[2]);
    }
    const int r3 = (q >> 1) % 3;
    const cf2 s = e[1] + e[2], d = e[1] - e[2];
    if (r3 == 0) return e[0] + s;
    constexpr float kS = 0.86602540378443864676f;
    const f2 mh = f2_bcast(-0.5f);
    const cf2 m{xfma(s.re, mh, e[0].re), xfma(s.im, mh, e[0].im)};
    const f2 sg = f2_bcast(((r3 == 1) == (DIR > 0)) ? kS : -kS);
    return cf2{xfma(d.im, f2_neg(sg), m.re), xfma(d.re, sg, m.im)};
  }
};

template <class CFG>
RD_HD cf x_band_spectrum(const unsigned char* sm, int kh, int kw) {
  const float4* Z = reinterpret_cast<const float4*>(sm + CFG::offZ);
  const float* tab = reinterpret_cast<const float*>(sm + CFG::offTab);
  const int k2p = kw & 15, qp = kw >> 4;
  const int k2m = (16 - k2p) & 15, qm = (((-kw - k2m) / 16) % CFG::N1w + CFG::N1w) % CFG::N1w;
  const float4* zp = Z + (size_t)(kh + CFG::b) * CFG::pitchZ + CFG::N1w * k2p;
  const float4* zm = Z + (size_t)(CFG::b - kh) * CFG::pitchZ + CFG::N1w * k2m;
  cf2 p[CFG::N1w], m[CFG::N1w];
#pragma unroll
  for (int c1 = 0; c1 < CFG::N1w; ++c1) {
    p[c1] = cf2_from(zp[c1]);
    m[c1] = cf2_from(zm[c1]);
  }
  const cf2 up = XDftOut<CFG::N1w, -1>::run(p, qp), um = XDftOut<CFG::N1w, -1>::run(m, qm);
  const cf upa{f2_lo(up.re), f2_lo(up.im)}, upb{f2_hi(up.re), f2_hi(up.im)};
  const cf uma{f2_lo(um.re), f2_lo(um.im)}, umb{f2_hi(um.re), f2_hi(um.im)};
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

template <class CFG>
RD_HD void x_band_mix(unsigned char* sm, float one_minus_ratio, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = x_band_spectrum<CFG>(sm, kh, kw);
  const float at = reinterpret_cast<const float*>(sm + CFG::offP)[khi * CFG::nb + kw + CFG::b];
  const float a2 = f.re * f.re + f.im * f.im;
  cf d;
  if (a2 > 0.f) {
    const float s = one_minus_ratio * (at / sqrtf(a2) - 1.0f) * CFG::invHW;
    d = cf{f.re * s, f.im * s};
  } else {
    d = cf{one_minus_ratio * at * CFG::invHW, 0.f};
  }
  float2* D = reinterpret_cast<float2*>(sm + CFG::offD);
  if (kw == 0) {
    if (kh == 0) d.im = 0.f;
    D[(size_t)(CFG::nb - 1 - khi) * (CFG::b + 1)] = make_float2(d.re, -d.im);
  }
  D[(size_t)khi * (CFG::b + 1) + kw] = make_float2(d.re, d.im);
}

template <class CFG>
RD_HD void x_band_amp(const unsigned char* sm, float* patch, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = x_band_spectrum<CFG>(sm, kh, kw);
  const float a = sqrtf(f.re * f.re + f.im * f.im);
  patch[(size_t)khi * CFG::nb + (kw + CFG::b)] = a;
  patch[(size_t)(CFG::nb - 1 - khi) * CFG::nb + (CFG::b - kw)] = a;
}

template <class CFG>
RD_HD void x_row_inv(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  const float2* D = reinterpret_cast<const float2*>(sm + CFG::offD);
  const float2* Dp = D + (size_t)khi * (CFG::b + 1);
  const float2* Dm = D + (size_t)(CFG::nb - 1 - khi) * (CFG::b + 1);
  const float4* PT = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tPT) +
                     (size_t)c1 * CFG::nb + CFG::b;
  cf2 v[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    cf2 acc{f2_bcast(0.f), f2_bcast(0.f)};
#pragma unroll
    for (int mi = -CFG::MM; mi < CFG::MM; ++mi) {
      const int kw = k2 + 16 * mi;
      if (kw >= -CFG::b && kw <= CFG::b) {
        float dre, dim;
        if (kw >= 0) { const float2 t = Dp[kw]; dre = t.x; dim = t.y; }
        else { const float2 t = Dm[-kw]; dre = t.x; dim = -t.y; }
        const cf2 p = cf2_from(PT[kw]);
        const f2 dr = f2_bcast(dre), di = f2_bcast(dim);
        acc.re = xfma(dr, p.re, acc.re);
        acc.re = xfma(f2_neg(di), p.im, acc.re);
        acc.im = xfma(dr, p.im, acc.im);
        acc.im = xfma(di, p.re, acc.im);
      }
    }
    v[k2] = acc;
  }
  fft16x<1>(v);
  float4* Z = reinterpret_cast<float4*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
#pragma unroll
  for (int j = 0; j < 16; ++j) Z[CFG::N1w * j] = cf2_to(v[pos16(j)]);
}

// ---- inverse columns ------------------------------------------------------------------------------
template <class CFG, int K1>
struct XInvGather {
  template <int A>
  static RD_HD void run(const cf2 (&in)[2 * CFG::MM], cf2 (&t)[4]) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    t[r] = t[r] + twidNx<CFG::N1h, A * K1, 1>(in[K1 + CFG::MM]);
    XInvGather<CFG, K1 + 1>::template run<A>(in, t);
  }
};
template <class CFG>
struct XInvGather<CFG, CFG::MM> {
  template <int A>
  static RD_HD void run(const cf2 (&)[2 * CFG::MM], cf2 (&)[4]) {}
};
template <class CFG, int A, int AEND>
struct XInvA {
  static RD_HD void run(const cf2 (&in)[2 * CFG::MM], float4* e) {
    const cf2 z{f2_bcast(0.f), f2_bcast(0.f)};
    cf2 t[4] = {z, z, z, z};
    XInvGather<CFG, -CFG::MM>::template run<A>(in, t);
    fft4x<1>(t[0], t[1], t[2], t[3]);
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) e[(A + CFG::Q * bq) * 16] = cf2_to(t[bq]);
    XInvA<CFG, A + 1, AEND>::run(in, e);
  }
};
template <class CFG, int AEND>
struct XInvA<CFG, AEND, AEND> {
  static RD_HD void run(const cf2 (&)[2 * CFG::MM], float4*) {}
};

// task u in [0, 128) of half APAR (a in [APAR Q/2, (APAR+1) Q/2)): c = u & 15, k2l = u >> 4, k2 = 2 k2l + par
template <class CFG, int APAR>
RD_HD void x_inv_pruned(unsigned char* sm, int strip, int par, int u) {
  const int c = u & 15, k2l = u >> 4, k2 = 2 * k2l + par;
  const float4* zcol = reinterpret_cast<const float4*>(sm + CFG::offZ) + strip * 16 + c;
  cf2 in[2 * CFG::MM];
#pragma unroll
  for (int i = 0; i < 2 * CFG::MM; ++i) {
    const int kh = k2 + 16 * (i - CFG::MM);
    cf2 v{f2_bcast(0.f), f2_bcast(0.f)};
    if (kh >= -CFG::b && kh <= CFG::b) v = cf2_from(zcol[(size_t)(kh + CFG::b) * CFG::pitchZ]);
    in[i] = v;
  }
  float4* e = reinterpret_cast<float4*>(sm + CFG::offE) + (size_t)k2l * CFG::N1h * 16 + c;
  XInvA<CFG, APAR * (CFG::Q / 2), (APAR + 1) * (CFG::Q / 2)>::run(in, e);
}

template <class CFG>
RD_HD void x_inv_fft8(const unsigned char* sm, int t, int par, cf2 (&p)[8]) {
  const int h1 = t >> 4;
  const float4* E = reinterpret_cast<const float4*>(sm + CFG::offE) + h1 * 16 + (t & 15);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH2) +
                     h1 * CFG::tw4Pitch + par;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 w = tw[2 * m];
    p[m] = cmulc2(cf2_from(E[m * CFG::N1h * 16]), f2_make(w.x, w.y), f2_make(w.z, w.w));
  }
  fft8x<1>(p);
}

template <class CFG, int J>
struct XCombine {
  template <class LoadF4, class StoreF4>
  static RD_HD void run(LoadF4 ld, StoreF4 st, Epilogue ep, size_t base, const cf2 (&ev)[8], const cf2 (&od)[8]) {
    const cf2 w = twidNx<16, J, 1>(od[pos8(J)]);
    const cf2 e = ev[pos8(J)];
    const f2 sc = f2_bcast(ep.scale), sh = f2_bcast(ep.shift);
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = base + (size_t)(CFG::N1h * (J + 8 * hi)) * CFG::M;
      const cf2 a = hi ? e - w : e + w;
      const float4 x = ld(off);
      const f2 vr = f2_make(x.x, x.y) + a.re, vi = f2_make(x.z, x.w) + a.im;   // x + delta for (4c, 4c+1), (4c+2, 4c+3)
      const f2 cr = f2_make(fminf(fmaxf(f2_lo(vr), ep.lo), ep.hi), fminf(fmaxf(f2_hi(vr), ep.lo), ep.hi));
      const f2 ci = f2_make(fminf(fmaxf(f2_lo(vi), ep.lo), ep.hi), fminf(fmaxf(f2_hi(vi), ep.lo), ep.hi));
      const f2 yr = xfma(cr, sc, sh), yi = xfma(ci, sc, sh);
      st(off, make_float4(f2_lo(yr), f2_hi(yr), f2_lo(yi), f2_hi(yi)), x);
    }
    XCombine<CFG, J + 1>::run(ld, st, ep, base, ev, od);
  }
};
template <class CFG>
struct XCombine<CFG, 8> {
  template <class LoadF4, class StoreF4>
  static RD_HD void run(LoadF4, StoreF4, Epilogue, size_t, const cf2 (&)[8], const cf2 (&)[8]) {}
};

template <class CFG, class LoadF4, class StoreF4>
RD_HD void x_inv_combine_store(LoadF4 ld, StoreF4 st, Epilogue ep, int strip, int t, const cf2 (&ev)[8], const cf2 (&od)[8]) {
  const size_t base = (size_t)(t >> 4) * CFG::M + strip * 16 + (t & 15);
  XCombine<CFG, 0>::run(ld, st, ep, base, ev, od);
}

}  // namespace ramdsir
