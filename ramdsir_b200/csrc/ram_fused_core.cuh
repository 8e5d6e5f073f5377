// Fused RAM pipeline for one image plane -- per-task phase bodies.
//
// Reference semantics: source_to_target_freq / low_freq_mutate_np / extract_amp_spectrum,
// /root/reference/code/dataset/fundus.py:13-61 (= prostate.py:10-62).  Only the centred
// (2b+1)^2 low band of the spectrum changes, so
//     y = x + Re(ifft2(dF)),  dF = F * (1-ratio) * (A_partner/|F| - 1) on the band, 0 elsewhere
// and both transforms are pruned to the band.
//
// Data layout.  A real H x W plane is read as float4 = 4 adjacent pixels and treated as TWO
// packed complex columns  za = x[4c] + i x[4c+2],  zb = x[4c+1] + i x[4c+3]  (c in [0, M), M = W/4).
// The 2-D complex DFT of za / zb over (h, c) is   Ua = G0 + i G2,  Ub = G1 + i G3   with G_g the
// DFT of the real sub-image x[:, 4c+g]; Hermitian symmetry of G_g separates them on the band and
// F[kh,kw] = sum_g W_W^(g kw) G_g[kh,kw].  The inverse runs the same packing backwards.
//
// Every length-N transform here (N = H along columns, N = M along rows) is split  N = 16 * N1:
// an in-register radix-16 over the stride-N1 index plus a direct, band-pruned twiddled sum over
// the N1 residues (only 2b+1 of the N outputs / inputs exist).
//
// The phase bodies are __host__ __device__ so tools/emulate can run the identical index algebra
// on the CPU; the kernels in ram_fused.cu call them with tid-strided task loops.
#pragma once
#include <stdint.h>
#include "ram_math.cuh"

namespace ramdsir {

struct FusedShape {
  int H, W, b, nb;   // nb = 2b+1
  int N1h;           // H / 16
  int M;             // W / 4   (float4 per image row)
  int N1w;           // M / 16
  int pitchZ;        // float4 per row of the band buffer Z
  int nthreads;      // 16 * N1h
  int nstrips;       // M / 16
  int MM;            // ceil(b / 16): band index = k2 + 16 m, m in [-MM, MM]
  // shared-memory byte offsets
  int offZ, offE, offD, offTab, tabBytes, smemBytes;
  // table sub-offsets (in floats, relative to the table block)
  int tTW1, tTW2, tTW3, tPT;
  float invHW;
};

struct Epilogue {
  float lo, hi, scale, shift;
};

#ifndef __CUDACC__
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
#endif

template <class SH>
RD_HD float4* smemZ(const SH& S, unsigned char* sm) { return reinterpret_cast<float4*>(sm + S.offZ); }
template <class SH>
RD_HD float4* smemE(const SH& S, unsigned char* sm) { return reinterpret_cast<float4*>(sm + S.offE); }
template <class SH>
RD_HD float2* smemD(const SH& S, unsigned char* sm) { return reinterpret_cast<float2*>(sm + S.offD); }
template <class SH>
RD_HD const float* smemTab(const SH& S, const unsigned char* sm) { return reinterpret_cast<const float*>(sm + S.offTab); }
RD_HD cf ldcf(const float* p) { return cf{p[0], p[1]}; }

// ---------------------------------------------------------------------------------------------
// F1a: task t = (h1 = t / 16, c = t % 16) of column strip `strip`: load rows h1 + N1h * j,
// radix-16 forward over j for both packed columns.
template <class SH, class LoadF4>
RD_HD void f1_load_fft(const SH& S, LoadF4 ld, int strip, int t, cf (&va)[16], cf (&vb)[16]) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 p = ld((size_t)(h1 + S.N1h * j) * S.M + c4);
    va[j] = cf{p.x, p.z};
    vb[j] = cf{p.y, p.w};
  }
  fft16<-1>(va);
  fft16<-1>(vb);
}

// F1b: publish k2 in [8*half, 8*half+8) to E[k2l][h1][c] as (a.re, a.im, b.re, b.im).
template <class SH>
RD_HD void f1_store_half(const SH& S, unsigned char* sm, int t, int half, const cf (&va)[16], const cf (&vb)[16]) {
  float4* E = smemE(S, sm);
  const int h1 = t >> 4, c = t & 15;
#pragma unroll
  for (int k2l = 0; k2l < 8; ++k2l) {
    // static register index: both halves are unrolled and selected by the (uniform) flag
    const cf a = half ? va[pos16(8 + k2l)] : va[pos16(k2l)];
    const cf bq = half ? vb[pos16(8 + k2l)] : vb[pos16(k2l)];
    E[(k2l * S.N1h + h1) * 16 + c] = make_float4(a.re, a.im, bq.re, bq.im);
  }
}

// F1c: task u in [0, 256): col = u & 1, c = (u >> 1) & 15, k2l = u >> 5.
// Z[kh][strip*16 + c].col = sum_h1 W_H^(h1 kh) * E[k2l][h1][c].col for every band row kh = k2 (mod 16).
template <int MMAX, class SH>
RD_HD void f1_direct(const SH& S, unsigned char* sm, int strip, int half, int u) {
  const float2* E2 = reinterpret_cast<const float2*>(smemE(S, sm));
  float2* Z2 = reinterpret_cast<float2*>(smemZ(S, sm));
  const float* tw1 = smemTab(S, sm) + S.tTW1;
  const int col = u & 1, c = (u >> 1) & 15, k2l = u >> 5, k2 = 8 * half + k2l;
  cf acc[2 * MMAX + 1];
  int khi[2 * MMAX + 1];
#pragma unroll
  for (int mi = 0; mi <= 2 * MMAX; ++mi) {
    const int kh = k2 + 16 * (mi - MMAX);
    khi[mi] = (kh >= -S.b && kh <= S.b) ? kh + S.b : -1;
    acc[mi] = cf{0.f, 0.f};
  }
  const float2 e0 = E2[((k2l * S.N1h) * 16 + c) * 2 + col];
  bool alleq = true;
  for (int h1 = 0; h1 < S.N1h; ++h1) {
    const float2 e = E2[((k2l * S.N1h + h1) * 16 + c) * 2 + col];
    const cf ev{e.x, e.y};
    alleq = alleq && e.x == e0.x && e.y == e0.y;
#pragma unroll
    for (int mi = 0; mi <= 2 * MMAX; ++mi)
      if (khi[mi] >= 0) cmac(acc[mi], ev, ldcf(tw1 + 2 * (h1 * S.nb + khi[mi])));
  }
  // Degenerate columns.  For k2 == 0 the rows are kh = 16 k1 and the sum is an N1h-point DFT of the stride sums.  When
  // those are all EQUAL (a constant column) it is exactly N1h e at k1 == 0 and exactly 0 elsewhere -- which is what a
  // butterfly FFT (pocketfft behind the reference's np.fft.fft2, fundus.py:16,46) returns, whereas a direct sum of
  // roots of unity leaves rounding noise whose phase the blend would then keep (np.angle(0) == 0, fundus.py:49).
  if (k2 == 0 && alleq) {
#pragma unroll
    for (int mi = 0; mi <= 2 * MMAX; ++mi)
      acc[mi] = (mi == MMAX) ? cf{e0.x * (float)S.N1h, e0.y * (float)S.N1h} : cf{0.f, 0.f};
  }
#pragma unroll
  for (int mi = 0; mi <= 2 * MMAX; ++mi)
    if (khi[mi] >= 0) Z2[((size_t)khi[mi] * S.pitchZ + strip * 16 + c) * 2 + col] = make_float2(acc[mi].re, acc[mi].im);
}

// ---------------------------------------------------------------------------------------------
// Row forward, in place on Z: task r = (khi = r / N1w, c1 = r % N1w); radix-16 over j for the
// elements c1 + N1w * j; X[k2] is written back to slot c1 + N1w * k2.
template <class SH>
RD_HD void row_fwd(const SH& S, unsigned char* sm, int r) {
  float4* Z = smemZ(S, sm) + (size_t)(r / S.N1w) * S.pitchZ + (r % S.N1w);
  cf va[16], vb[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 p = Z[S.N1w * j];
    va[j] = cf{p.x, p.y};
    vb[j] = cf{p.z, p.w};
  }
  fft16<-1>(va);
  fft16<-1>(vb);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const cf a = va[pos16(k2)], bq = vb[pos16(k2)];
    Z[S.N1w * k2] = make_float4(a.re, a.im, bq.re, bq.im);
  }
}

// Spectrum value of the real plane at (kh, kw >= 0) from the row-transformed band buffer.
// Returns F[kh, kw] (unnormalised numpy fft2 convention).
template <class SH>
RD_HD cf band_spectrum(const SH& S, const unsigned char* sm, int kh, int kw) {
  const float4* Z = reinterpret_cast<const float4*>(sm + S.offZ);
  const float* tab = smemTab(S, sm);
  const float* tw2 = tab + S.tTW2 + 2 * kw;  // [c1][kw], row stride 2 * (b + 1)
  const float4* zp = Z + (size_t)(kh + S.b) * S.pitchZ + S.N1w * (kw & 15);
  const float4* zm = Z + (size_t)(S.b - kh) * S.pitchZ + S.N1w * ((16 - (kw & 15)) & 15);
  cf upa{0, 0}, upb{0, 0}, uma{0, 0}, umb{0, 0};
  const float4 p0 = zp[0], q0 = zm[0];
  bool peq = true, qeq = true;
  for (int c1 = 0; c1 < S.N1w; ++c1) {
    const cf w = ldcf(tw2 + 2 * c1 * (S.b + 1));  // W_M^(c1 kw)
    const float4 p = zp[c1], q = zm[c1];
    peq = peq && p.x == p0.x && p.y == p0.y && p.z == p0.z && p.w == p0.w;
    qeq = qeq && q.x == q0.x && q.y == q0.y && q.z == q0.z && q.w == q0.w;
    cmac(upa, cf{p.x, p.y}, w);
    cmac(upb, cf{p.z, p.w}, w);
    cmacc(uma, cf{q.x, q.y}, w);  // U[-kh][-kw] uses W_M^(-c1 kw)
    cmacc(umb, cf{q.z, q.w}, w);
  }
  // Degenerate rows (see f1_direct): for kw = 16 q the sum over c1 is an N1w-point DFT; equal terms give exactly
  // N1w z at q == 0 and exactly 0 elsewhere, as the butterflies of the reference's FFT do.
  if ((kw & 15) == 0) {
    const float n = (float)S.N1w;
    if (peq) {
      upa = (kw == 0) ? cf{p0.x * n, p0.y * n} : cf{0.f, 0.f};
      upb = (kw == 0) ? cf{p0.z * n, p0.w * n} : cf{0.f, 0.f};
    }
    if (qeq) {
      uma = (kw == 0) ? cf{q0.x * n, q0.y * n} : cf{0.f, 0.f};
      umb = (kw == 0) ? cf{q0.z * n, q0.w * n} : cf{0.f, 0.f};
    }
  }
  // G0 = (Ua+ + conj(Ua-))/2, G2 = (Ua+ - conj(Ua-))/(2i); G1, G3 likewise from Ub
  const cf g0{0.5f * (upa.re + uma.re), 0.5f * (upa.im - uma.im)};
  const cf g2{0.5f * (upa.im + uma.im), -0.5f * (upa.re - uma.re)};
  const cf g1{0.5f * (upb.re + umb.re), 0.5f * (upb.im - umb.im)};
  const cf g3{0.5f * (upb.im + umb.im), -0.5f * (upb.re - umb.re)};
  const float* tw3 = tab + S.tTW3 + 2 * kw;  // [g-1][kw]
  cf f = g0;
  cmac(f, g1, ldcf(tw3));
  cmac(f, g2, ldcf(tw3 + 2 * (S.b + 1)));
  cmac(f, g3, ldcf(tw3 + 4 * (S.b + 1)));
  return f;
}

// Band task q = (khi = q / (b+1), kw = q % (b+1)) for the MIX kernel: dF/(HW) into D[khi][kw].
// `patch` is the partner's [nb][nb] low-band amplitude (fftshift layout, centre = DC);
// one_minus_ratio = (den - k) / den.
template <class SH>
RD_HD void band_mix(const SH& S, unsigned char* sm, const float* patch, float one_minus_ratio, int q) {
  const int kw = q % (S.b + 1), khi = q / (S.b + 1), kh = khi - S.b;
  if (kw == 0 && kh < 0) return;  // written by the (-kh, 0) task
  const cf f = band_spectrum(S, sm, kh, kw);
  // Re(ifft2) only sees the Hermitian part of dF, i.e. the partner amplitude averaged over +-k
  // (identical for amplitudes of real images, exact for arbitrary amp_trg arrays).
  const float at = 0.5f * (patch[(size_t)khi * S.nb + (kw + S.b)] + patch[(size_t)(S.nb - 1 - khi) * S.nb + (S.b - kw)]);
  const float a2 = f.re * f.re + f.im * f.im;
  cf d;
  if (a2 > 0.f) {
    const float s = one_minus_ratio * (at / sqrtf(a2) - 1.0f) * S.invHW;
    d = cf{f.re * s, f.im * s};
  } else {
    d = cf{one_minus_ratio * at * S.invHW, 0.f};  // np.angle(0) == 0
  }
  float2* D = smemD(S, sm);
  if (kw == 0) {
    if (kh == 0) d.im = 0.f;  // DC of a real plane
    D[(size_t)(S.nb - 1 - khi) * (S.b + 1)] = make_float2(d.re, -d.im);
  }
  D[(size_t)khi * (S.b + 1) + kw] = make_float2(d.re, d.im);
}

// Band task for the BANK builder: |F| into the global patch at (kh, kw) and (-kh, -kw).
template <class SH>
RD_HD void band_amp(const SH& S, const unsigned char* sm, float* patch, int q) {
  const int kw = q % (S.b + 1), khi = q / (S.b + 1), kh = khi - S.b;
  if (kw == 0 && kh < 0) return;
  const cf f = band_spectrum(S, sm, kh, kw);
  const float a = sqrtf(f.re * f.re + f.im * f.im);
  patch[(size_t)khi * S.nb + (kw + S.b)] = a;
  patch[(size_t)(S.nb - 1 - khi) * S.nb + (S.b - kw)] = a;
}

// Row inverse: task r = (khi, c1): B[k2] = sum_{kw = k2 mod 16} D[kh][kw] * P{a,b}[c1][kw], radix-16
// inverse over k2 -> z'[khi][c1 + N1w * j].
template <int MMAX, class SH>
RD_HD void row_inv(const SH& S, unsigned char* sm, int r) {
  const int khi = r / S.N1w, c1 = r % S.N1w;
  const float2* D = smemD(S, sm);
  const float2* Dp = D + (size_t)khi * (S.b + 1);
  const float2* Dm = D + (size_t)(S.nb - 1 - khi) * (S.b + 1);
  const float4* PT = reinterpret_cast<const float4*>(smemTab(S, sm) + S.tPT) + (size_t)c1 * S.nb + S.b;
  cf va[16], vb[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    cf aa{0, 0}, ab{0, 0};
#pragma unroll
    for (int mi = 0; mi <= 2 * MMAX; ++mi) {
      const int kw = k2 + 16 * (mi - MMAX);
      if (kw >= -S.b && kw <= S.b) {
        cf d;
        if (kw >= 0) { const float2 t = Dp[kw]; d = cf{t.x, t.y}; }
        else { const float2 t = Dm[-kw]; d = cf{t.x, -t.y}; }
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
  float4* Z = smemZ(S, sm) + (size_t)khi * S.pitchZ + c1;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const cf a = va[pos16(j)], bq = vb[pos16(j)];
    Z[S.N1w * j] = make_float4(a.re, a.im, bq.re, bq.im);
  }
}

// I2: task t = (h1 = t / 16, c = t % 16) of strip `strip`: B[k2] = sum_{kh = k2 mod 16} conj(W_H^(h1 kh)) z'[kh][c4],
// radix-16 inverse over k2 -> delta at rows h1 + N1h * j; y = clamp(x + delta) * scale + shift.
template <int MMAX, class SH, class LoadF4, class StoreF4>
RD_HD void col_inv_store(const SH& S, const unsigned char* sm, LoadF4 ld, StoreF4 st, Epilogue ep, int strip, int t) {
  const int h1 = t >> 4, c4 = strip * 16 + (t & 15);
  const float4* Z = reinterpret_cast<const float4*>(sm + S.offZ) + c4;
  const float* tw1 = smemTab(S, sm) + S.tTW1 + 2 * (h1 * S.nb + S.b);
  cf va[16], vb[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    cf aa{0, 0}, ab{0, 0};
#pragma unroll
    for (int mi = 0; mi <= 2 * MMAX; ++mi) {
      const int kh = k2 + 16 * (mi - MMAX);
      if (kh >= -S.b && kh <= S.b) {
        const float4 p = Z[(size_t)(kh + S.b) * S.pitchZ];
        const cf w = ldcf(tw1 + 2 * kh);
        cmacc(aa, cf{p.x, p.y}, w);
        cmacc(ab, cf{p.z, p.w}, w);
      }
    }
    va[k2] = aa;
    vb[k2] = ab;
  }
  fft16<1>(va);
  fft16<1>(vb);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const size_t off = (size_t)(h1 + S.N1h * j) * S.M + c4;
    const float4 x = ld(off);
    const cf a = va[pos16(j)], bq = vb[pos16(j)];
    float4 y;
    y.x = fmaf(fminf(fmaxf(x.x + a.re, ep.lo), ep.hi), ep.scale, ep.shift);
    y.y = fmaf(fminf(fmaxf(x.y + bq.re, ep.lo), ep.hi), ep.scale, ep.shift);
    y.z = fmaf(fminf(fmaxf(x.z + a.im, ep.lo), ep.hi), ep.scale, ep.shift);
    y.w = fmaf(fminf(fmaxf(x.w + bq.im, ep.lo), ep.hi), ep.scale, ep.shift);
    st(off, y);
  }
}

}  // namespace ramdsir
