// Fused RAM pipeline, "P2" variant: ONE packed complex column per thread.
//
// Same maths as ram_fused_static.cuh, different packing and thread mapping, chosen for latency hiding:
//   * the real plane is read as float2 = 2 adjacent pixels = one complex column  z[h][c] = x[h][2c] + i x[h][2c+1]
//     (c < M2 = W/2).  With U = DFT2(z):  Ev = (U[k] + conj(U[-k]))/2,  Od = (U[k] - conj(U[-k]))/(2i),
//     F[kh,kw] = Ev + W_W^kw Od   (the classic real-FFT-by-half-length-complex-FFT split, applied along w);
//     the inverse feeds  V[kh,kw] = dF[kh,kw] (1 + i W_W^-kw)  through the inverse DFT2 and reads delta off
//     the real / imaginary parts.
//   * a thread holds 16 complex values (32 registers) instead of 2 x 16, so the kernel fits 64 registers
//     per thread and 1024 threads per SM (two CTAs of 16 N1h... = 512 threads for 256^2): twice the warps of
//     the float4 variant to cover global-load latency and the per-strip barriers.
// Column factorisation is unchanged: h = h1 + N1h j, kh = k2 + 16 k1, h1 = a + Q b' (see ram_fused_static.cuh).
// Rows: M2 = 16 N1w (N1w = 8 for W = 256, 12 for W = 384); one output of an N1w-point butterfly per band bin.
#pragma once
#include "ram_fused_static.cuh"

namespace ramdsir {

template <int H_, int W_, int B_>
struct PCfg {
  static constexpr int H = H_, W = W_, b = B_, nb = 2 * B_ + 1;
  static constexpr int N1h = H_ / 16, M2 = W_ / 2, N1w = M2 / 16;
  static constexpr int Q = N1h / 4;
  static constexpr int pitchZ = M2 + 1;            // float2 units; odd -> conflict-free with lanes <-> kh rows
  static constexpr int nthreads = 32 * N1h, nstrips = M2 / 32;
  static constexpr int MM = B_ / 16 + 1;
  static constexpr int twhPitch = 18;
  static constexpr int npruned = 512;              // pruned-stage tasks per pass: 2 (r / a parity) x 8 k2 x 32 c
  // shared memory (bytes)
  static constexpr int offZ = 0;
  static constexpr int zBytes = s_al16(nb * pitchZ * 8);
  static constexpr int eBytes = 8 * N1h * 32 * 8;
  static constexpr int dBytes = nb * (B_ + 1) * 8;
  static constexpr int pBytes = nb * nb * 4;
  static constexpr int offE = zBytes, offD = zBytes;
  static constexpr int offP = offD + s_al16(dBytes);
  static constexpr int offTab = zBytes + s_al16(s_max(eBytes, s_al16(dBytes) + pBytes));
  // tables (float offsets)
  static constexpr int tPT = 0;                                   // [N1w][nb] float2: (1 + i W_W^-kw) W_M2^(-c1 kw)
  static constexpr int tTW3 = tPT + N1w * nb * 2;                 // [b+1] float2: W_W^kw
  static constexpr int tTWH = ((tTW3 + (B_ + 1) * 2) + 3) & ~3;   // [N1h][18] float2: W_H^(h1 k2)
  static constexpr int tTWM = tTWH + N1h * twhPitch * 2;          // [N1w][16] float2: W_M2^(c1 k2)
  static constexpr int tabBytes = s_al16((tTWM + N1w * 16 * 2) * 4);
  static constexpr int smemBytes = offTab + tabBytes;
  static constexpr float invHW = (float)(1.0 / ((double)H_ * (double)W_));
  static_assert(H_ % 64 == 0 && W_ % 64 == 0, "P2 fused path needs H, W multiples of 64");
  static_assert(2 * B_ + 1 <= M2 && 2 * B_ + 1 <= H_, "band must fit the packed spectrum");
  static_assert(N1w == 8 || N1w == 12, "row butterfly implemented for N1w = 8, 12");
  static_assert(nthreads >= 512 && nthreads <= 1024, "block size out of range");
};

// ---- F1a: 16 strided rows of one complex column -> radix-16 -> x TWH[h1][k2] ------------------------
template <class CFG, class LoadF2>
RD_HD void p_f1_load_fft_tw(LoadF2 ld, const unsigned char* sm, int strip, int t, cf (&v)[16]) {
  const int h1 = t >> 5, c = strip * 32 + (t & 31);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 p = ld((size_t)(h1 + CFG::N1h * j) * CFG::M2 + c);
    v[j] = cf{p.x, p.y};
  }
  fft16<-1>(v);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH) +
                     h1 * (CFG::twhPitch / 2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = tw[i];
    if (i > 0) v[pos16(2 * i)] = cmul(v[pos16(2 * i)], cf{w.x, w.y});
    v[pos16(2 * i + 1)] = cmul(v[pos16(2 * i + 1)], cf{w.z, w.w});
  }
}

template <class CFG>
RD_HD void p_f1_store_half(unsigned char* sm, int t, int half, const cf (&v)[16]) {
  float2* E = reinterpret_cast<float2*>(sm + CFG::offE) + (t >> 5) * 32 + (t & 31);
#pragma unroll
  for (int k2l = 0; k2l < 8; ++k2l) {
    const cf a = half ? v[pos16(8 + k2l)] : v[pos16(k2l)];
    E[k2l * CFG::N1h * 32] = make_float2(a.re, a.im);
  }
}

// sum over a of W_N1h^(DIR a k1) y[a][SLOT]
template <class CFG, int K1, int DIR, int SLOT, int A>
struct PSumA {
  static RD_HD cf run(const cf (&y)[CFG::Q][2]) {
    return PSumA<CFG, K1, DIR, SLOT, A - 1>::run(y) + twidN<CFG::N1h, A * K1, DIR>(y[A][SLOT]);
  }
};
template <class CFG, int K1, int DIR, int SLOT>
struct PSumA<CFG, K1, DIR, SLOT, 0> {
  static RD_HD cf run(const cf (&y)[CFG::Q][2]) { return y[0][SLOT]; }
};

template <class CFG, int RPAR, int K1>
struct PF1Out {
  static RD_HD void run(const cf (&y)[CFG::Q][2], float2* zcol, int k2) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    if (r == RPAR || r == RPAR + 2) {
      const int kh = k2 + 16 * K1;
      if (kh >= -CFG::b && kh <= CFG::b) {
        const cf o = PSumA<CFG, K1, -1, (r == RPAR) ? 0 : 1, CFG::Q - 1>::run(y);
        zcol[(size_t)(kh + CFG::b) * CFG::pitchZ] = make_float2(o.re, o.im);
      }
    }
    PF1Out<CFG, RPAR, K1 + 1>::run(y, zcol, k2);
  }
};
template <class CFG, int RPAR>
struct PF1Out<CFG, RPAR, CFG::MM> {
  static RD_HD void run(const cf (&)[CFG::Q][2], float2*, int) {}
};

// ---- F1c: task u in [0, 256) of parity RPAR: c = u & 31, k2l = u >> 5.  Half of a pruned N1h-point DFT:
// the residues r = k1 mod 4 in {RPAR, RPAR + 2} of the radix-4 first stage and the band rows that use them.
template <class CFG, int RPAR>
RD_HD void p_f1_pruned(unsigned char* sm, int strip, int half, int u) {
  const int c = u & 31, k2l = u >> 5, k2 = 8 * half + k2l;
  const float2* e = reinterpret_cast<const float2*>(sm + CFG::offE) + (size_t)k2l * CFG::N1h * 32 + c;
  cf y[CFG::Q][2];
#pragma unroll
  for (int a = 0; a < CFG::Q; ++a) {
    cf in[4];
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
      const float2 v = e[(a + CFG::Q * bq) * 32];
      in[bq] = cf{v.x, v.y};
    }
    if (RPAR == 0) {
      const cf t02 = in[0] + in[2], t13 = in[1] + in[3];
      y[a][0] = t02 + t13;   // r = 0
      y[a][1] = t02 - t13;   // r = 2
    } else {
      const cf t02 = in[0] - in[2], t13 = in[1] - in[3];
      const cf jt{t13.im, -t13.re};  // -i * t13
      y[a][0] = t02 + jt;    // r = 1
      y[a][1] = t02 - jt;    // r = 3
    }
  }
  float2* zcol = reinterpret_cast<float2*>(sm + CFG::offZ) + strip * 32 + c;
  PF1Out<CFG, RPAR, -CFG::MM>::run(y, zcol, k2);
}

// ---- row forward: radix-16 over j + twiddle TWM[c1][k2], in place; lanes <-> band rows ---------------
template <class CFG>
RD_HD void p_row_fwd(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  float2* Z = reinterpret_cast<float2*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
  cf v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 p = Z[CFG::N1w * j];
    v[j] = cf{p.x, p.y};
  }
  fft16<-1>(v);
  const float4* tw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWM) + c1 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = tw[i];
    cf a0 = v[pos16(2 * i)];
    if (i > 0) a0 = cmul(a0, cf{w.x, w.y});
    const cf a1 = cmul(v[pos16(2 * i + 1)], cf{w.z, w.w});
    Z[CFG::N1w * (2 * i)] = make_float2(a0.re, a0.im);
    Z[CFG::N1w * (2 * i + 1)] = make_float2(a1.re, a1.im);
  }
}

// one output of an 8- / 12-point DFT, butterfly structured (exact zeros for constant input)
template <int DIR>
struct DftOut<8, DIR> {
  static RD_HD cf run(const cf (&a)[8], int q) {
    const cf ev4[4] = {a[0], a[2], a[4], a[6]}, od4[4] = {a[1], a[3], a[5], a[7]};
    const cf ev = DftOut<4, DIR>::run(ev4, q & 3), od = DftOut<4, DIR>::run(od4, q & 3);
    switch (q & 7) {   // + W_8^(DIR q) * od
      case 0: return ev + od;
      case 1: return ev + twidN<8, 1, DIR>(od);
      case 2: return ev + twidN<8, 2, DIR>(od);
      case 3: return ev + twidN<8, 3, DIR>(od);
      case 4: return ev - od;
      case 5: return ev - twidN<8, 1, DIR>(od);
      case 6: return ev - twidN<8, 2, DIR>(od);
      default: return ev - twidN<8, 3, DIR>(od);
    }
  }
};
template <int DIR>
struct DftOut<12, DIR> {
  static RD_HD cf run(const cf (&a)[12], int q) {
    const int qa = q & 3, qb = q >> 2;   // q = qa + 4 qb;  c = 3 m + n
    cf in[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const cf g[4] = {a[n], a[n + 3], a[n + 6], a[n + 9]};
      in[n] = DftOut<4, DIR>::run(g, qa);
    }
    switch (qa) {   // in[n] *= W_12^(DIR n qa)
      case 1: in[1] = twidN<12, 1, DIR>(in[1]); in[2] = twidN<12, 2, DIR>(in[2]); break;
      case 2: in[1] = twidN<12, 2, DIR>(in[1]); in[2] = twidN<12, 4, DIR>(in[2]); break;
      case 3: in[1] = twidN<12, 3, DIR>(in[1]); in[2] = twidN<12, 6, DIR>(in[2]); break;
      default: break;
    }
    const cf s = in[1] + in[2], d = in[1] - in[2];
    if (qb == 0) return in[0] + s;
    constexpr float kS = 0.86602540378443864676f;
    const cf m{in[0].re - 0.5f * s.re, in[0].im - 0.5f * s.im};
    const float sg = ((qb == 1) == (DIR > 0)) ? kS : -kS;
    return cf{m.re - sg * d.im, m.im + sg * d.re};
  }
};

// F[kh, kw], kw >= 0
template <class CFG>
RD_HD cf p_band_spectrum(const unsigned char* sm, int kh, int kw) {
  const float2* Z = reinterpret_cast<const float2*>(sm + CFG::offZ);
  const int k2p = kw & 15, qp = kw >> 4;
  const int k2m = (16 - k2p) & 15, qm = (((-kw - k2m) / 16) % CFG::N1w + CFG::N1w) % CFG::N1w;
  const float2* zp = Z + (size_t)(kh + CFG::b) * CFG::pitchZ + CFG::N1w * k2p;
  const float2* zm = Z + (size_t)(CFG::b - kh) * CFG::pitchZ + CFG::N1w * k2m;
  cf p[CFG::N1w], m[CFG::N1w];
#pragma unroll
  for (int c1 = 0; c1 < CFG::N1w; ++c1) {
    const float2 a = zp[c1], bq = zm[c1];
    p[c1] = cf{a.x, a.y};
    m[c1] = cf{bq.x, bq.y};
  }
  const cf up = DftOut<CFG::N1w, -1>::run(p, qp), um = DftOut<CFG::N1w, -1>::run(m, qm);
  const cf ev{0.5f * (up.re + um.re), 0.5f * (up.im - um.im)};
  const cf od{0.5f * (up.im + um.im), -0.5f * (up.re - um.re)};
  cf f = ev;
  cmac(f, od, ldcf(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTW3 + 2 * kw));
  return f;
}

template <class CFG>
RD_HD void p_band_mix(unsigned char* sm, float one_minus_ratio, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = p_band_spectrum<CFG>(sm, kh, kw);
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
RD_HD void p_band_amp(const unsigned char* sm, float* patch, int q) {
  const int khi = q % CFG::nb, kw = q / CFG::nb, kh = khi - CFG::b;
  if (kw == 0 && kh < 0) return;
  const cf f = p_band_spectrum<CFG>(sm, kh, kw);
  const float a = sqrtf(f.re * f.re + f.im * f.im);
  patch[(size_t)khi * CFG::nb + (kw + CFG::b)] = a;
  patch[(size_t)(CFG::nb - 1 - khi) * CFG::nb + (CFG::b - kw)] = a;
}

// ---- inverse rows: band sums into the 16 radix-16 inputs, radix-16 inverse ---------------------------
template <class CFG>
RD_HD void p_row_inv(unsigned char* sm, int r) {
  const int khi = r % CFG::nb, c1 = r / CFG::nb;
  const float2* D = reinterpret_cast<const float2*>(sm + CFG::offD);
  const float2* Dp = D + (size_t)khi * (CFG::b + 1);
  const float2* Dm = D + (size_t)(CFG::nb - 1 - khi) * (CFG::b + 1);
  const float2* PT = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tPT) +
                     (size_t)c1 * CFG::nb + CFG::b;
  cf v[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    cf acc{0, 0};
#pragma unroll
    for (int mi = -CFG::MM; mi < CFG::MM; ++mi) {
      const int kw = k2 + 16 * mi;
      if (kw >= -CFG::b && kw <= CFG::b) {
        cf d;
        if (kw >= 0) { const float2 t = Dp[kw]; d = cf{t.x, t.y}; }
        else { const float2 t = Dm[-kw]; d = cf{t.x, -t.y}; }
        const float2 p = PT[kw];
        cmac(acc, d, cf{p.x, p.y});
      }
    }
    v[k2] = acc;
  }
  fft16<1>(v);
  float2* Z = reinterpret_cast<float2*>(sm + CFG::offZ) + (size_t)khi * CFG::pitchZ + c1;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const cf a = v[pos16(j)];
    Z[CFG::N1w * j] = make_float2(a.re, a.im);
  }
}

// ---- inverse columns, first stage: task u in [0, 256) of parity APAR handles a in [APAR Q/2, (APAR+1) Q/2) --
template <class CFG, int K1>
struct PInvGather {
  template <int A>
  static RD_HD void run(const cf (&in)[2 * CFG::MM], cf (&t)[4]) {
    constexpr int r = ((K1 % 4) + 4) % 4;
    t[r] = t[r] + twidN<CFG::N1h, A * K1, 1>(in[K1 + CFG::MM]);
    PInvGather<CFG, K1 + 1>::template run<A>(in, t);
  }
};
template <class CFG>
struct PInvGather<CFG, CFG::MM> {
  template <int A>
  static RD_HD void run(const cf (&)[2 * CFG::MM], cf (&)[4]) {}
};
template <class CFG, int A, int AEND>
struct PInvA {
  static RD_HD void run(const cf (&in)[2 * CFG::MM], float2* e) {
    cf t[4] = {cf{0.f, 0.f}, cf{0.f, 0.f}, cf{0.f, 0.f}, cf{0.f, 0.f}};
    PInvGather<CFG, -CFG::MM>::template run<A>(in, t);
    fft4<1>(t[0], t[1], t[2], t[3]);
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) e[(A + CFG::Q * bq) * 32] = make_float2(t[bq].re, t[bq].im);
    PInvA<CFG, A + 1, AEND>::run(in, e);
  }
};
template <class CFG, int AEND>
struct PInvA<CFG, AEND, AEND> {
  static RD_HD void run(const cf (&)[2 * CFG::MM], float2*) {}
};

template <class CFG, int APAR>
RD_HD void p_inv_pruned(unsigned char* sm, int strip, int par, int u) {
  const int c = u & 31, k2l = u >> 5, k2 = 2 * k2l + par;
  const float2* zcol = reinterpret_cast<const float2*>(sm + CFG::offZ) + strip * 32 + c;
  cf in[2 * CFG::MM];
#pragma unroll
  for (int i = 0; i < 2 * CFG::MM; ++i) {
    const int kh = k2 + 16 * (i - CFG::MM);
    cf v{0.f, 0.f};
    if (kh >= -CFG::b && kh <= CFG::b) {
      const float2 z = zcol[(size_t)(kh + CFG::b) * CFG::pitchZ];
      v = cf{z.x, z.y};
    }
    in[i] = v;
  }
  float2* e = reinterpret_cast<float2*>(sm + CFG::offE) + (size_t)k2l * CFG::N1h * 32 + c;
  PInvA<CFG, APAR * (CFG::Q / 2), (APAR + 1) * (CFG::Q / 2)>::run(in, e);
}

template <class CFG>
RD_HD void p_inv_fft8(const unsigned char* sm, int t, int par, cf (&p)[8]) {
  const int h1 = t >> 5;
  const float2* E = reinterpret_cast<const float2*>(sm + CFG::offE) + h1 * 32 + (t & 31);
  const float2* tw = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWH) +
                     h1 * CFG::twhPitch + par;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float2 v = E[m * CFG::N1h * 32];
    const float2 w = tw[2 * m];
    p[m] = cmulc(cf{v.x, v.y}, cf{w.x, w.y});
  }
  fft8<1>(p);
}

template <class CFG, int J>
struct PCombine {
  template <class LoadF2, class StoreF2>
  static RD_HD void run(LoadF2 ld, StoreF2 st, Epilogue ep, size_t base, const cf (&ev)[8], const cf (&od)[8]) {
    const cf w = twidN<16, J, 1>(od[pos8(J)]);
    const cf e = ev[pos8(J)];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = base + (size_t)(CFG::N1h * (J + 8 * hi)) * CFG::M2;
      const cf a = hi ? e - w : e + w;
      const float2 x = ld(off);
      float2 y;
      y.x = fmaf(fminf(fmaxf(x.x + a.re, ep.lo), ep.hi), ep.scale, ep.shift);
      y.y = fmaf(fminf(fmaxf(x.y + a.im, ep.lo), ep.hi), ep.scale, ep.shift);
      st(off, y);
    }
    PCombine<CFG, J + 1>::run(ld, st, ep, base, ev, od);
  }
};
template <class CFG>
struct PCombine<CFG, 8> {
  template <class LoadF2, class StoreF2>
  static RD_HD void run(LoadF2, StoreF2, Epilogue, size_t, const cf (&)[8], const cf (&)[8]) {}
};

template <class CFG, class LoadF2, class StoreF2>
RD_HD void p_inv_combine_store(LoadF2 ld, StoreF2 st, Epilogue ep, int strip, int t, const cf (&ev)[8], const cf (&od)[8]) {
  const size_t base = (size_t)(t >> 5) * CFG::M2 + strip * 32 + (t & 31);
  PCombine<CFG, 0>::run(ld, st, ep, base, ev, od);
}

}  // namespace ramdsir
