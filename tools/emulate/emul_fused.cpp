// CPU emulation of the fused RAM kernel's CTA pipeline (TEST TOOL, never shipped or imported by
// the package): runs the exact __host__ __device__ phase bodies of ram_fused_core.cuh task by
// task, with a heap buffer standing in for shared memory, so the index algebra can be checked
// against the numpy oracle in a container without a GPU.
//   g++ -O2 -shared -fPIC -I ramdsir_b200/csrc tools/emulate/emul_fused.cpp -o tools/emulate/libemul.so
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ram_fused_host.h"
#include "ram_multipass_host.h"

using namespace ramdsir;

template <int MMAX>
static void run_plane(const FusedShape& S, unsigned char* sm, const float* x, const float* patch,
                      float omr, float* y, Epilogue ep, float* bank_out) {
  auto ld = [&](size_t off) { const float* p = x + 4 * off; return make_float4(p[0], p[1], p[2], p[3]); };
  auto st = [&](size_t off, float4 v) { float* p = y + 4 * off; p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; };
  const int NT = S.nthreads;
  std::vector<cp> regs_store((size_t)NT * 32);   // 8-byte aligned storage, viewed as cf / cp / cf2 by the phases
  cf* regs = reinterpret_cast<cf*>(regs_store.data());
  for (int strip = 0; strip < S.nstrips; ++strip) {
    for (int t = 0; t < NT; ++t) {
      cf(&va)[16] = *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32]);
      cf(&vb)[16] = *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32 + 16]);
      f1_load_fft(S, ld, strip, t, va, vb);
    }
    for (int half = 0; half < 2; ++half) {
      for (int t = 0; t < NT; ++t) {
        cf(&va)[16] = *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32]);
        cf(&vb)[16] = *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32 + 16]);
        f1_store_half(S, sm, t, half, va, vb);
      }
      for (int u = 0; u < 256; ++u) f1_direct<MMAX>(S, sm, strip, half, u);
    }
  }
  for (int r = 0; r < S.nb * S.N1w; ++r) row_fwd(S, sm, r);
  if (bank_out) {
    for (int q = 0; q < S.nb * (S.b + 1); ++q) band_amp(S, sm, bank_out, q);
    return;
  }
  for (int q = 0; q < S.nb * (S.b + 1); ++q) band_mix(S, sm, patch, omr, q);
  for (int r = 0; r < S.nb * S.N1w; ++r) row_inv<MMAX>(S, sm, r);
  for (int strip = 0; strip < S.nstrips; ++strip)
    for (int t = 0; t < NT; ++t) col_inv_store<MMAX>(S, sm, ld, st, ep, strip, t);
}

template <int MMAX>
static int run_all(const FusedShape& S, const float* src, const float* bank, const long long* pidx,
                   const int* rk, int den, float* out, int B, int C, Epilogue ep, float* bank_out) {
  std::vector<unsigned char> sm(S.smemBytes);
  fused_tables_fill(S, reinterpret_cast<float*>(sm.data() + S.offTab));
  const size_t plane = (size_t)S.H * S.W, psz = (size_t)S.nb * S.nb;
  for (int i = 0; i < B; ++i)
    for (int c = 0; c < C; ++c) {
      const size_t p = (size_t)i * C + c;
      if (bank_out) { run_plane<MMAX>(S, sm.data(), src + p * plane, nullptr, 0.f, nullptr, ep, bank_out + p * psz); continue; }
      const float omr = (float)(den - rk[i]) / (float)den;
      run_plane<MMAX>(S, sm.data(), src + p * plane, bank + ((size_t)pidx[i] * C + c) * psz, omr, out + p * plane, ep, nullptr);
    }
  return 0;
}

extern "C" int emul_fused(const float* src, const float* bank, const long long* pidx, const int* rk, int den,
                          float* out, int B, int C, int H, int W, int b, float lo, float hi, float scale,
                          float shift, float* bank_out) {
  FusedShape S;
  if (!fused_shape_init(S, H, W, b)) return 1;
  Epilogue ep{lo, hi, scale, shift};
  switch (S.MM) {
    case 0: case 1: return run_all<1>(S, src, bank, pidx, rk, den, out, B, C, ep, bank_out);
    case 2: return run_all<2>(S, src, bank, pidx, rk, den, out, B, C, ep, bank_out);
    case 3: return run_all<3>(S, src, bank, pidx, rk, den, out, B, C, ep, bank_out);
    case 4: return run_all<4>(S, src, bank, pidx, rk, den, out, B, C, ep, bank_out);
  }
  return 2;
}

// ---- static-shape (v2) pipeline ------------------------------------------------------------
template <class CFG, class EPI>
static void run_plane_static(unsigned char* sm, const float* x, const float* patch, float omr, float dc_add, float* y,
                             EPI ep, float* bank_out) {
  CFG S;
  auto ld = [&](size_t off) { const float* p = x + 4 * off; return make_float4(p[0], p[1], p[2], p[3]); };
  auto st = [&](size_t off, float4 v, float4) { float* p = y + 4 * off; p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; };
  constexpr int NT = CFG::nthreads;
  std::vector<cp> regs_store((size_t)NT * 32);   // 8-byte aligned storage, viewed as cf / cp / cf2 by the phases
  cf* regs = reinterpret_cast<cf*>(regs_store.data());
  auto A = [&](int t) -> cf(&)[16] { return *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32]); };
  auto Bv = [&](int t) -> cf(&)[16] { return *reinterpret_cast<cf(*)[16]>(&regs[(size_t)t * 32 + 16]); };
  for (int strip = 0; strip < CFG::nstrips; ++strip) {
    // the kernel's packed-pair forward stage (host definitions of FADD2 / FMUL2 / FFMA2: lane by lane)
    auto V = [&](int t) -> cf2(&)[16] { return *reinterpret_cast<cf2(*)[16]>(&regs[(size_t)t * 32]); };
    static_assert(sizeof(cf2) == 2 * sizeof(cf), "a thread's 16 packed pairs fill its 32 emulated complex registers");
    for (int t = 0; t < NT; ++t) s_f1_load_fft_tw_x2<CFG>(ld, sm, strip, t, V(t));
    for (int half = 0; half < 2; ++half) {
      for (int t = 0; t < NT; ++t) s_f1_store_half_x2<CFG>(sm, t, half, V(t));
      for (int u = 0; u < 256; ++u) s_f1_pruned<CFG>(sm, strip, half, u);
    }
  }
  for (int r = 0; r < CFG::nb * CFG::N1w; ++r) s_row_fwd<CFG>(sm, r);
  if (bank_out) {
    for (int q = 0; q < CFG::nb * (CFG::b + 1); ++q) s_band_amp<CFG>(sm, bank_out, q);
    return;
  }
  for (int q = 0; q < CFG::nb * (CFG::b + 1); ++q) s_band_mix<CFG>(sm, patch, omr, dc_add, q);
  for (int r = 0; r < CFG::nb * CFG::N1w; ++r) s_row_inv<CFG>(sm, r);
  (void)S;
  for (int strip = CFG::nstrips - 1; strip >= 0; --strip) {
    for (int par = 0; par < 2; ++par) {
      for (int u = 0; u < 256; ++u) s_inv_pruned<CFG>(sm, strip, par, u);
      for (int t = 0; t < NT; ++t) {
        using CX = typename CFG::cx_fft8;
        CX(&pa)[8] = *reinterpret_cast<CX(*)[8]>(&regs[(size_t)t * 32 + 16 * par]);
        CX(&pb)[8] = *reinterpret_cast<CX(*)[8]>(&regs[(size_t)t * 32 + 16 * par + 8]);
        s_inv_fft8<CFG>(sm, t, par, pa, pb);
      }
    }
    for (int t = 0; t < NT; ++t) {
      cf* r = &regs[(size_t)t * 32];
      using CX = typename CFG::cx_fft8;
      s_inv_combine_store<CFG>(ld, st, ep, strip, t, *reinterpret_cast<CX(*)[8]>(r), *reinterpret_cast<CX(*)[8]>(r + 8),
                               *reinterpret_cast<CX(*)[8]>(r + 16), *reinterpret_cast<CX(*)[8]>(r + 24));
    }
  }
}

template <class CFG>
static int run_all_static(const float* src, const float* bank, const long long* pidx, const int* rk, int den,
                          float* out, int B, int C, Epilogue ep, float* bank_out) {
  std::vector<unsigned char> sm(CFG::smemBytes);
  fused_tables_fill_static<CFG>(reinterpret_cast<float*>(sm.data() + CFG::offTab));
  const size_t plane = (size_t)CFG::H * CFG::W, psz = (size_t)CFG::nb * CFG::nb;
  for (int i = 0; i < B; ++i)
    for (int c = 0; c < C; ++c) {
      const size_t p = (size_t)i * C + c;
      const EpiClamp ec{ep.lo, ep.hi, ep.scale, ep.shift};
      if (bank_out) { run_plane_static<CFG>(sm.data(), src + p * plane, nullptr, 0.f, 0.f, nullptr, ec, bank_out + p * psz); continue; }
      const float omr = (float)(den - rk[i]) / (float)den;
      const float* patch = bank + ((size_t)pidx[i] * C + c) * psz;
      // finite clip range: the two-instruction saturating epilogue, exactly as launch_static() sets it up
      if (std::isfinite(ep.lo) && std::isfinite(ep.hi) && ep.hi > ep.lo) {
        const double span = (double)ep.hi - (double)ep.lo;
        const EpiSat es{(float)(1.0 / span), (float)(span * (double)ep.scale), (float)((double)ep.lo * ep.scale + ep.shift)};
        run_plane_static<CFG>(sm.data(), src + p * plane, patch, omr * es.a, (float)(-(double)ep.lo / span), out + p * plane, es, nullptr);
      } else {
        run_plane_static<CFG>(sm.data(), src + p * plane, patch, omr, 0.f, out + p * plane, ec, nullptr);
      }
    }
  return 0;
}

extern "C" int emul_fused_static(const float* src, const float* bank, const long long* pidx, const int* rk, int den,
                                 float* out, int B, int C, int H, int W, int b, float lo, float hi, float scale,
                                 float shift, float* bank_out) {
  Epilogue ep{lo, hi, scale, shift};
  if (H == 256 && W == 256 && b == 25) return run_all_static<SCfg<256, 256, 25>>(src, bank, pidx, rk, den, out, B, C, ep, bank_out);
  if (H == 384 && W == 384 && b == 38) return run_all_static<SCfg<384, 384, 38>>(src, bank, pidx, rk, den, out, B, C, ep, bank_out);
  return 1;
}

extern "C" int emul_fused_static_smem_bytes(int H, int W, int b) {
  if (H == 256 && W == 256 && b == 25) return SCfg<256, 256, 25>::smemBytes;
  if (H == 384 && W == 384 && b == 38) return SCfg<384, 384, 38>::smemBytes;
  return -1;
}

// ---- multi-pass pipeline (K2) ---------------------------------------------------------------
template <class CFG>
static int run_all_mp(const float* src, const float* bank, const long long* pidx, const int* rk, int den, float* out,
                      int B, int C, Epilogue ep, float* bank_out) {
  std::vector<unsigned char> sm(CFG::smemBytes), smb(CFG::bSmemBytes);
  multipass_tables_ac<CFG>(reinterpret_cast<float*>(sm.data() + CFG::offTab));
  multipass_tables_b<CFG>(reinterpret_cast<float*>(smb.data() + CFG::bOffTab));
  const size_t plane_px = (size_t)CFG::H * CFG::W, psz = (size_t)CFG::nb * CFG::nb;
  std::vector<float4> Zg((size_t)CFG::nb * CFG::M);
  std::vector<cp> regs_store((size_t)256 * 32);
  cf* regs = reinterpret_cast<cf*>(regs_store.data());
  for (int i = 0; i < B; ++i)
    for (int c = 0; c < C; ++c) {
      const size_t pl = (size_t)i * C + c;
      const float* x = src + pl * plane_px;
      auto ld = [&](size_t off) { const float* q = x + 4 * off; return make_float4(q[0], q[1], q[2], q[3]); };
      // pass A
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        for (int p = 0; p < CFG::HP; ++p) {
          for (int t = 0; t < 256; ++t)
            m_f1_load_fft_tw_x2<CFG>(ld, sm.data(), strip, p, t, *reinterpret_cast<cf2(*)[16]>(&regs[(size_t)t * 32]));
          for (int half = 0; half < 2; ++half) {
            for (int t = 0; t < 256; ++t)
              s_f1_store_half_x2<CFG>(sm.data(), t, half, *reinterpret_cast<cf2(*)[16]>(&regs[(size_t)t * 32]));
            for (int u = 0; u < 256; ++u) { if (p == 0) m_f1_pruned<CFG, true>(sm.data(), p, half, u); else m_f1_pruned<CFG, false>(sm.data(), p, half, u); }
          }
        }
        const float4* Zs = reinterpret_cast<const float4*>(sm.data() + CFG::offZ);
        for (int k = 0; k < CFG::nb; ++k)
          for (int cc = 0; cc < 16; ++cc) Zg[(size_t)k * CFG::M + strip * 16 + cc] = Zs[k * CFG::pitchZ + cc];
      }
      // pass B
      const float omr = bank_out ? 0.f : (float)(den - rk[i]) / (float)den;
      const float* patch = bank_out ? nullptr : bank + ((size_t)pidx[i] * C + c) * psz;
      for (int kh = 0; kh <= CFG::b; ++kh) {
        const int nrows = kh == 0 ? 1 : 2;
        float4* R = reinterpret_cast<float4*>(smb.data() + CFG::bOffR);
        float2* D = reinterpret_cast<float2*>(smb.data() + CFG::bOffD);
        const float* tab = reinterpret_cast<const float*>(smb.data() + CFG::bOffTab);
        const int rows[2] = {CFG::b + kh, CFG::b - kh};
        for (int r = 0; r < nrows; ++r)
          for (int m = 0; m < CFG::M; ++m) R[r * CFG::rowPitch + (m % CFG::N1w) + CFG::SP * (m / CFG::N1w)] = Zg[(size_t)rows[r] * CFG::M + m];
        for (int t = 0; t < nrows * CFG::N1w; ++t) m_row_stage1<CFG, -1>(R, tab, t);
        for (int t = 0; t < nrows * 16; ++t) m_row_stage2<CFG, -1>(R, t);
        if (bank_out) {
          for (int t = 0; t < nrows * (CFG::b + 1); ++t)
            m_band<CFG, true>(R, D, tab, nullptr, bank_out + pl * psz, 0.f, 0.f, kh, t / (CFG::b + 1), t % (CFG::b + 1));
          continue;
        }
        std::vector<float> prow(2 * CFG::nb);
        for (int j = 0; j < CFG::nb; ++j) {
          prow[j] = patch[(size_t)(CFG::b + kh) * CFG::nb + j];
          prow[CFG::nb + j] = patch[(size_t)(CFG::b - kh) * CFG::nb + j];
        }
        for (int t = 0; t < nrows * (CFG::b + 1); ++t)
          m_band<CFG, false>(R, D, tab, prow.data(), nullptr, omr, 0.f, kh, t / (CFG::b + 1), t % (CFG::b + 1));
        for (int t = 0; t < 2 * CFG::rowPitch; ++t) R[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < nrows * CFG::nb; ++t) m_put_v<CFG>(R, D, tab, kh, t / CFG::nb, t % CFG::nb);
        for (int t = 0; t < nrows * 16; ++t) m_row_stage2<CFG, 1>(R, t);
        for (int t = 0; t < nrows * CFG::N1w; ++t) m_row_stage1<CFG, 1>(R, tab, t);
        for (int r = 0; r < nrows; ++r)
          for (int m = 0; m < CFG::M; ++m) Zg[(size_t)rows[r] * CFG::M + m] = R[r * CFG::rowPitch + (m % CFG::N1w) + CFG::SP * (m / CFG::N1w)];
      }
      if (bank_out) continue;
      // pass C
      float* y = out + pl * plane_px;
      auto st = [&](size_t off, float4 v) { float* q = y + 4 * off; q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; };
      for (int strip = 0; strip < CFG::nstrips; ++strip) {
        float4* Zs = reinterpret_cast<float4*>(sm.data() + CFG::offZ);
        for (int k = 0; k < CFG::nb; ++k)
          for (int cc = 0; cc < 16; ++cc) Zs[k * CFG::pitchZ + cc] = Zg[(size_t)k * CFG::M + strip * 16 + cc];
        for (int p = 0; p < CFG::HP; ++p) {
          for (int par = 0; par < 2; ++par) {
            for (int u = 0; u < 256; ++u) m_inv_pruned<CFG>(sm.data(), p, par, u);
            for (int t = 0; t < 256; ++t)
              s_inv_fft8<CFG>(sm.data(), t, par, *reinterpret_cast<cp(*)[8]>(&regs[(size_t)t * 32 + 16 * par]),
                              *reinterpret_cast<cp(*)[8]>(&regs[(size_t)t * 32 + 16 * par + 8]));
          }
          for (int t = 0; t < 256; ++t) {
            cf* r = &regs[(size_t)t * 32];
            MCombine<CFG, 0>::run(ld, st, EpiClamp{ep.lo, ep.hi, ep.scale, ep.shift}, t >> 4, p, strip * 16 + (t & 15), *reinterpret_cast<cp(*)[8]>(r),
                                  *reinterpret_cast<cp(*)[8]>(r + 8), *reinterpret_cast<cp(*)[8]>(r + 16), *reinterpret_cast<cp(*)[8]>(r + 24));
          }
        }
      }
    }
  return 0;
}

extern "C" int emul_multipass(const float* src, const float* bank, const long long* pidx, const int* rk, int den,
                              float* out, int B, int C, int H, int W, int b, float lo, float hi, float scale, float shift,
                              float* bank_out) {
  Epilogue ep{lo, hi, scale, shift};
  if (H == 512 && W == 512 && b == 51) return run_all_mp<KCfg<512, 512, 51>>(src, bank, pidx, rk, den, out, B, C, ep, bank_out);
  if (H == 1024 && W == 1024 && b == 102) return run_all_mp<KCfg<1024, 1024, 102>>(src, bank, pidx, rk, den, out, B, C, ep, bank_out);
  return 1;
}

extern "C" int emul_fused_smem_bytes(int H, int W, int b) {
  FusedShape S;
  return fused_shape_init(S, H, W, b) ? S.smemBytes : -1;
}
