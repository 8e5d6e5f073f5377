// Multi-pass RAM pipeline for planes whose band buffers do not fit shared memory (K2: 512^2, 1024^2).
//
// Same maths and float4 packing as ram_fused_static.cuh.  Differences:
//   * a column of H = 256 HP rows is split into HP polyphase sub-columns x[HP h' + p] of 256 rows:
//       X[kh] = sum_p W_H^(p kh) X_p[kh],  |kh| <= b < 128
//     so every sub-column reuses the 256-row machinery (radix-16 in registers, TWH, exchange through E,
//     pruned radix-4 x Q=4 stage), and the HP results are combined with the table TWP[p][kh] = W_H^(p kh)
//     into a strip-local band buffer Zs[kh][16 columns] in shared memory;
//   * the band rows of a plane (nb x M float4, 840 KB for 1024^2) live in a global workspace that stays in L2:
//       pass A  (CTA per plane x column strip)  forward columns            -> Zg[plane][kh][c4]
//       pass B  (CTA per plane x {+kh, -kh})     row FFTs, band update, inverse row FFTs, in place on Zg
//       pass C  (CTA per plane x column strip)  inverse columns, + x, epilogue -> out
//   * the host runs the three passes per CHUNK of planes sized so that x and Zg of a chunk stay L2-resident.
// Row transforms here are complete M-point FFTs (M = 16 N1w) in two in-register stages, natural bin order.
#pragma once
#include "ram_fused_static.cuh"

namespace ramdsir {

template <int H_, int W_, int B_>
struct KCfg {
  static constexpr int H = H_, W = W_, b = B_, nb = 2 * B_ + 1;
  static constexpr int HP = H_ / 256;             // polyphase factor along h
  static constexpr int N1h = 16, Q = 4;           // sub-column length 256 = 16 * 16
  static constexpr int M = W_ / 4, N1w = M / 16;
  static constexpr int nstrips = M / 16;
  static constexpr int MM = B_ / 16 + 1;
  static constexpr int pitchZ = 17;               // strip-local band buffer: 16 float4 columns + 1 pad
  static constexpr int twhPitch = 18;
  static constexpr int nthreads = 256;
  // pass A / C shared memory
  static constexpr int offZ = 0;
  static constexpr int zBytes = s_al16(nb * pitchZ * 16);
  static constexpr int eBytes = 8 * N1h * 16 * 16;
  static constexpr int offE = zBytes;
  static constexpr int offTab = zBytes + eBytes;
  static constexpr int tTWH = 0;                               // [16][18] float2  W_256^(h1 k2)
  static constexpr int tTWP = tTWH + N1h * twhPitch * 2;       // [HP][nb] float2  W_H^(p kh)
  static constexpr int tabBytes = s_al16((tTWP + HP * nb * 2) * 4);
  static constexpr int smemBytes = offTab + tabBytes;
  // pass B shared memory: KG pairs of rows (+kh, -kh) of M float4 (+1 pad), their D rows, tables
#ifndef RD_MP_KG
#define RD_MP_KG 4
#endif
  static constexpr int KG = RD_MP_KG;                          // kh values per CTA
  static constexpr int nkg = (B_ + 1 + KG - 1) / KG;           // CTAs per plane
  static constexpr int SP = N1w + 1;                           // padded pitch of a k2 / j group inside a row: conflict-free
  static constexpr int rowPitch = 16 * SP;                     // float4 per staged row (element c1 + N1w j lives at c1 + SP j)
  static constexpr int bOffR = 0;                              // [KG][2][rowPitch] float4
  static constexpr int bOffD = s_al16(KG * 2 * rowPitch * 16); // [KG][2][b+1] float2
  static constexpr int bOffP = bOffD + s_al16(KG * 2 * (B_ + 1) * 8);   // [KG][2][nb] float: partner patch rows b+kh, b-kh
  static constexpr int bOffTab = bOffP + s_al16(KG * 2 * nb * 4);
  static constexpr int tTWM = 0;                               // [16][N1w] float2  W_M^(c1 k2), k2-major
  static constexpr int tTW3 = tTWM + N1w * 16 * 2;             // [3][b+1] float2   W_W^(g kw)
  static constexpr int tQ = ((tTW3 + 3 * (B_ + 1) * 2) + 3) & ~3;   // [nb] float4: (Qa, Qb)[kw]
  static constexpr int bTabBytes = s_al16((tQ + nb * 4) * 4);
  static constexpr int bSmemBytes = bOffTab + bTabBytes;
  static constexpr int bThreads = 128;
  static constexpr float invHW = (float)(1.0 / ((double)H_ * (double)W_));
  static_assert(H_ % 256 == 0 && W_ % 256 == 0, "multi-pass path: H, W multiples of 256");
  static_assert(2 * B_ + 1 <= 255 && 2 * B_ + 1 <= M, "band must fit the 256-row sub-transform and the packed rows");
  static_assert(N1w == 8 || N1w == 16, "row FFT implemented for M = 128, 256");
};

template <int N, int DIR>
struct FftN;
template <int DIR>
struct FftN<16, DIR> {
  static RD_HD void run(cf (&v)[16]) { fft16<DIR>(v); }
  static RD_HD void run(cp (&v)[16]) { fft16<DIR>(v); }
  static RD_HD constexpr int pos(int k) { return pos16(k); }
};
template <int DIR>
struct FftN<8, DIR> {
  static RD_HD void run(cf (&v)[8]) { fft8<DIR>(v); }
  static RD_HD void run(cp (&v)[8]) { fft8<DIR>(v); }
  static RD_HD constexpr int pos(int k) { return pos8(k); }
};

// ===== pass A ==================================================================================
// F1a of a polyphase sub-column on packed pairs (see s_f1_fft_tw_x2): rows (h1 + 16 j) HP + p
template <class CFG, class LoadF4>
RD_HD void m_f1_load_rows_x2(LoadF4 ld, int strip, int p, int t, cf2 (&v)[16]) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = cf2_from(ld((size_t)((h1 + 16 * j) * CFG::HP + p) * CFG::M + c4));
}
template <class CFG, class LoadF4>
RD_HD void m_f1_load_fft_tw_x2(LoadF4 ld, const unsigned char* sm, int strip, int p, int t, cf2 (&v)[16]) {
  m_f1_load_rows_x2<CFG>(ld, strip, p, t, v);
  s_f1_fft_tw_x2<CFG>(sm, t, v);
}

// FIRST (sub-column p = 0) stores, the later sub-columns accumulate: a compile-time choice, so that the emit is
// LDS (twiddle) [+ LDS (old)], 2 packed multiplies [+ 1 packed add], STS with no predicated copies around it.
template <class CFG, int K1, bool FIRST>
struct MF1Out {   // Zs[kh] (+)= TWP[p][kh] * sum_a W_16^(a k1) y[a][k1 mod 4]
  static RD_HD void run(const cp (&y)[4][4], float2* zcol, const float2* twp, int k2) {
    const int kh = k2 + 16 * K1;
    if (band_always<CFG>(K1) || (kh >= -CFG::b && kh <= CFG::b)) {
      cp o = SumA<CFG, K1, -1, 3>::run(y);
      const float2 w = twp[kh];
      o = cmul(o, cf{w.x, w.y});
      float2* z = zcol + (size_t)(kh + CFG::b) * CFG::pitchZ * 2;
      if (!FIRST) {
        const float2 old = *z;
        o = o + cp_make(old.x, old.y);
      }
      *z = make_float2(cp_re(o), cp_im(o));
    }
    MF1Out<CFG, K1 + 1, FIRST>::run(y, zcol, twp, k2);
  }
};
template <class CFG, bool FIRST>
struct MF1Out<CFG, CFG::MM, FIRST> {
  static RD_HD void run(const cp (&)[4][4], float2*, const float2*, int) {}
};

// Emit the band rows kh = k2 + 16 K1 from the finished 16-point transform X[k1 mod 16] (at v[pos16(.)]).
template <class CFG, int K1, bool FIRST>
struct MF1OutFull {
  static RD_HD void run(const cp (&v)[16], float2* zcol, const float2* twp, int k2) {
    const int kh = k2 + 16 * K1;
    if (band_always<CFG>(K1) || (kh >= -CFG::b && kh <= CFG::b)) {
      cp o = v[pos16((K1 + 16) & 15)];
      const float2 w = twp[kh];
      o = cmul(o, cf{w.x, w.y});
      float2* z = zcol + (size_t)(kh + CFG::b) * CFG::pitchZ * 2;
      if (!FIRST) {
        const float2 old = *z;
        o = o + cp_make(old.x, old.y);
      }
      *z = make_float2(cp_re(o), cp_im(o));
    }
    MF1OutFull<CFG, K1 + 1, FIRST>::run(v, zcol, twp, k2);
  }
};
template <class CFG, bool FIRST>
struct MF1OutFull<CFG, CFG::MM, FIRST> {
  static RD_HD void run(const cp (&)[16], float2*, const float2*, int) {}
};

// Second stage of the 256-row sub-transform.  With a wide band (1024^2: 13 of the 16 residues k1 of every k2 are in
// the band) the stage is a complete 16-point FFT over h1 -- 64 butterfly additions and 9 twiddles instead of 13
// four-term twiddled sums; with a narrow band (512^2: 6.4 of 16) the pruned radix-4 + sums form is cheaper.
template <class CFG, bool FIRST>
RD_HD void m_f1_pruned(unsigned char* sm, int p, int half, int u) {
  const int col = u & 1, c = (u >> 1) & 15, k2l = u >> 5, k2 = 8 * half + k2l;
  const float2* e = reinterpret_cast<const float2*>(sm + CFG::offE) + ((size_t)k2l * 16 * 16 + c) * 2 + col;
  float2* zcol = reinterpret_cast<float2*>(sm + CFG::offZ) + c * 2 + col;
  const float2* twp = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWP) +
                      p * CFG::nb + CFG::b;
  if constexpr (CFG::MM >= 6) {
    cp v[16];
#pragma unroll
    for (int h1 = 0; h1 < 16; ++h1) {
      const float2 t = e[h1 * 32];
      v[h1] = cp_make(t.x, t.y);
    }
    fft16<-1>(v);
    MF1OutFull<CFG, -CFG::MM, FIRST>::run(v, zcol, twp, k2);
  } else {
    cp y[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int bq = 0; bq < 4; ++bq) {
        const float2 v = e[(a + 4 * bq) * 32];
        y[a][bq] = cp_make(v.x, v.y);
      }
      fft4<-1>(y[a][0], y[a][1], y[a][2], y[a][3]);
    }
    MF1Out<CFG, -CFG::MM, FIRST>::run(y, zcol, twp, k2);
  }
}

// ===== pass C ==================================================================================
template <class CFG>
RD_HD void m_inv_pruned(unsigned char* sm, int p, int par, int u) {
  const int col = u & 1, c = (u >> 1) & 15, k2l = u >> 5, k2 = 2 * k2l + par;
  const float2* zcol = reinterpret_cast<const float2*>(sm + CFG::offZ) + c * 2 + col;
  const float2* twp = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(sm + CFG::offTab) + CFG::tTWP) +
                      p * CFG::nb + CFG::b;
  cp in[2 * CFG::MM];
#pragma unroll
  for (int i = 0; i < 2 * CFG::MM; ++i) {
    const int kh = k2 + 16 * (i - CFG::MM);
    cp v = cp_make(0.f, 0.f);
    if (band_always<CFG>(i - CFG::MM) || (kh >= -CFG::b && kh <= CFG::b)) {
      const float2 z = zcol[(size_t)(kh + CFG::b) * CFG::pitchZ * 2];
      const float2 w = twp[kh];
      v = cmulc(cp_make(z.x, z.y), cf{w.x, w.y});   // x W_H^(-p kh)
    }
    in[i] = v;
  }
  float2* e = reinterpret_cast<float2*>(sm + CFG::offE) + ((size_t)k2l * 16 * 16 + c) * 2 + col;
  if constexpr (CFG::MM >= 6) {
    // wide band: a complete inverse 16-point FFT over the residues k1 (those outside [-MM, MM) are zero)
    cp v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int k1 = k < 8 ? k : k - 16;
      v[k] = (k1 >= -CFG::MM && k1 < CFG::MM) ? in[k1 + CFG::MM] : cp_make(0.f, 0.f);
    }
    fft16<1>(v);
#pragma unroll
    for (int h1 = 0; h1 < 16; ++h1) e[h1 * 32] = make_float2(cp_re(v[pos16(h1)]), cp_im(v[pos16(h1)]));
  } else {
    InvA<CFG, 0>::run(in, e);
  }
}

template <class CFG, int J>
struct MCombine {
  template <class LoadF4, class StoreF4, class EPI>
  static RD_HD void run(LoadF4 ld, StoreF4 st, EPI ep, int h1, int p, int c4, const cp (&eva)[8], const cp (&evb)[8],
                        const cp (&oda)[8], const cp (&odb)[8]) {
    const cp wa = twidN<16, J, 1>(oda[pos8(J)]), wb = twidN<16, J, 1>(odb[pos8(J)]);
    const cp ea = eva[pos8(J)], eb = evb[pos8(J)];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const size_t off = (size_t)((h1 + 16 * (J + 8 * hi)) * CFG::HP + p) * CFG::M + c4;
      const cp a = hi ? ea - wa : ea + wa, bq = hi ? eb - wb : eb + wb;
      const float4 x = ld(off);
      st(off, make_float4(ep(x.x, cp_re(a)), ep(x.y, cp_re(bq)), ep(x.z, cp_im(a)), ep(x.w, cp_im(bq))));
    }
    MCombine<CFG, J + 1>::run(ld, st, ep, h1, p, c4, eva, evb, oda, odb);
  }
};
template <class CFG>
struct MCombine<CFG, 8> {
  template <class LoadF4, class StoreF4, class EPI>
  static RD_HD void run(LoadF4, StoreF4, EPI, int, int, int, const cp (&)[8], const cp (&)[8], const cp (&)[8], const cp (&)[8]) {}
};

// ===== pass B ==================================================================================
// Row FFT stage 1 (c1 fixed): radix-16 over j on slots c1 + N1w j, x TWM[c1][k2], back to slot c1 + N1w k2.
template <class CFG, int DIR>
RD_HD void m_row_stage1(float4* Rbase, const float* tab, int task) {
  const int row = task / CFG::N1w, c1 = task % CFG::N1w;
  float4* R = Rbase + (size_t)row * CFG::rowPitch + c1;
  const float2* tw = reinterpret_cast<const float2*>(tab + CFG::tTWM) + c1;   // TWM is stored [k2][c1]: lanes <-> c1 read contiguously
  cp va[16], vb[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 q = R[CFG::SP * j];
    va[j] = cp_make(q.x, q.y);
    vb[j] = cp_make(q.z, q.w);
    if (DIR > 0) {   // inverse: undo the twiddle first (input index is k2)
      const float2 w = tw[j * CFG::N1w];
      va[j] = cmulc(va[j], cf{w.x, w.y});
      vb[j] = cmulc(vb[j], cf{w.x, w.y});
    }
  }
  fft16<DIR>(va);
  fft16<DIR>(vb);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    cp a = va[pos16(k)], bq = vb[pos16(k)];
    if (DIR < 0) {
      const float2 w = tw[k * CFG::N1w];
      a = cmul(a, cf{w.x, w.y});
      bq = cmul(bq, cf{w.x, w.y});
    }
    R[CFG::SP * k] = make_float4(cp_re(a), cp_im(a), cp_re(bq), cp_im(bq));
  }
}

// Row FFT stage 2 (k2 fixed): N1w-point FFT over c1 on the contiguous slots N1w k2 + c1 -> bin kw = k2 + 16 q at slot N1w k2 + q.
template <class CFG, int DIR>
RD_HD void m_row_stage2(float4* Rbase, int task) {
  const int row = task / 16, k2 = task % 16;
  float4* R = Rbase + (size_t)row * CFG::rowPitch + CFG::SP * k2;
  cp va[CFG::N1w], vb[CFG::N1w];
#pragma unroll
  for (int i = 0; i < CFG::N1w; ++i) {
    const float4 q = R[i];
    va[i] = cp_make(q.x, q.y);
    vb[i] = cp_make(q.z, q.w);
  }
  FftN<CFG::N1w, DIR>::run(va);
  FftN<CFG::N1w, DIR>::run(vb);
#pragma unroll
  for (int i = 0; i < CFG::N1w; ++i) {
    const cp a = va[FftN<CFG::N1w, DIR>::pos(i)], bq = vb[FftN<CFG::N1w, DIR>::pos(i)];
    R[i] = make_float4(cp_re(a), cp_im(a), cp_re(bq), cp_im(bq));
  }
}

// slot of row bin kw (any sign) after the two forward stages
template <class CFG>
RD_HD int m_slot(int kw) {
  const int k = (kw + CFG::M) % CFG::M;
  return CFG::SP * (k & 15) + (k >> 4);
}

// Band task: bin (s kh, kw), kw in [0, b], s = +1 (row 0 = +kh) or -1 (row 1 = -kh).  Writes D[s][kw].
template <class CFG, bool BANK>
RD_HD void m_band(const float4* R, float2* D, const float* tab, const float* patch, float* patch_out, float one_minus_ratio,
                  float dc_add, int kh, int srow, int kw) {
  const int khs = srow ? -kh : kh;
  const float4 up4 = R[(size_t)srow * CFG::rowPitch + m_slot<CFG>(kw)];
  const float4 um4 = R[(size_t)(kh == 0 ? 0 : 1 - srow) * CFG::rowPitch + m_slot<CFG>(-kw)];
  const cf upa{up4.x, up4.y}, upb{up4.z, up4.w}, uma{um4.x, um4.y}, umb{um4.z, um4.w};
  const cf g0{0.5f * (upa.re + uma.re), 0.5f * (upa.im - uma.im)};
  const cf g2{0.5f * (upa.im + uma.im), -0.5f * (upa.re - uma.re)};
  const cf g1{0.5f * (upb.re + umb.re), 0.5f * (upb.im - umb.im)};
  const cf g3{0.5f * (upb.im + umb.im), -0.5f * (upb.re - umb.re)};
  const float* tw3 = tab + CFG::tTW3 + 2 * kw;
  cf f = g0;
  cmac(f, g1, ldcf(tw3));
  cmac(f, g2, ldcf(tw3 + 2 * (CFG::b + 1)));
  cmac(f, g3, ldcf(tw3 + 4 * (CFG::b + 1)));
  const int khi = khs + CFG::b;
  const float a2 = f.re * f.re + f.im * f.im;
  if (BANK) {
    const float a = sqrtf(a2);
    patch_out[(size_t)khi * CFG::nb + (kw + CFG::b)] = a;
    patch_out[(size_t)(CFG::nb - 1 - khi) * CFG::nb + (CFG::b - kw)] = a;
    return;
  }
  // `patch` points at the two staged rows of this kh pair: [0] = row b + kh, [1] = row b - kh (shared memory)
  const float at = 0.5f * (patch[srow * CFG::nb + (kw + CFG::b)] + patch[(kh == 0 ? 0 : 1 - srow) * CFG::nb + (CFG::b - kw)]);
  cf d;
  if (a2 >= 1.17549435e-38f) {       // a normal float (see s_band_mix)
    const float s = one_minus_ratio * (at * rd_rsqrt(a2) - 1.0f) * CFG::invHW;
    d = cf{f.re * s, f.im * s};
  } else {
    d = cf{one_minus_ratio * at * CFG::invHW, 0.f};
  }
  if (kw == 0 && kh == 0) {
    d.im = 0.f;
    d.re += dc_add;          // EpiSat: -lo / (hi - lo) on every pixel of delta (see ram_fused_static.cuh)
  }
  D[srow * (CFG::b + 1) + kw] = make_float2(d.re, d.im);
}

// Write the inverse-side packed spectrum V{a,b}[kw] = dF[s kh][kw] * (Qa, Qb)[kw] of band bin kw = kwi - b into its
// slot of row `srow` (the row has been zero-filled; lanes walk consecutive kw: D and Q reads are contiguous and the
// slot stride SP is conflict-free).
template <class CFG>
RD_HD void m_put_v(float4* R, const float2* D, const float* tab, int kh, int srow, int kwi) {
  const int kw = kwi - CFG::b;
  const int other = (kh == 0) ? 0 : 1 - srow;
  cf d;
  if (kw >= 0) { const float2 t = D[srow * (CFG::b + 1) + kw]; d = cf{t.x, t.y}; }
  else { const float2 t = D[other * (CFG::b + 1) - kw]; d = cf{t.x, -t.y}; }   // dF[-k] = conj dF[k]
  const float4 qq = reinterpret_cast<const float4*>(tab + CFG::tQ)[kwi];
  const cf a = cmul(d, cf{qq.x, qq.y}), bq = cmul(d, cf{qq.z, qq.w});
  R[(size_t)srow * CFG::rowPitch + m_slot<CFG>(kw)] = make_float4(a.re, a.im, bq.re, bq.im);
}

}  // namespace ramdsir
