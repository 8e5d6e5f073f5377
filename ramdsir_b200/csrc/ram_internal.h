// Internal declarations shared by the translation units of libramdsir.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ramdsir.h"
#include "ram_fused_core.cuh"

namespace ramdsir {

// thread-local error text (ramdsir_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define RD_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct MixArgs {
  const float* src;
  const float* bank;
  const int64_t* pidx;
  const int32_t* rk;
  float* out;
  int B, C, H, W, b, n_bank, ratio_den;
  Epilogue ep;
  unsigned flags = 0;         // RAMDSIR_FLAG_*
  float* clean = nullptr;     // fused dual output (static fused kernels only): x * scale + shift
};

// fused single-CTA-per-plane path (ram_fused.cu)
bool fused_supported(int H, int W, int b);
bool fused_static_shape(int H, int W, int b);
bool fused_dual_supported(int H, int W, int b, unsigned flags);
int fused_prepare(int H, int W, int b);
// the overlap bookkeeping of the static fused launches (ram_fused.cu) must forget a stream as soon as the library
// enqueues anything else on it
void fused_trail_reset(cudaStream_t st);
int u8_to_f32(const unsigned char* x, float* y, size_t n, cudaStream_t st);
int affine_f32(const float* x, float* y, size_t n, float scale, float shift, cudaStream_t st);
int fused_mix(const MixArgs& a, cudaStream_t st);
int fused_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, unsigned flags, cudaStream_t st);
// uint8 sources read directly by the static fused kernels (a.src is ignored)
bool fused_u8_supported(int H, int W, int b);
int fused_mix_u8(const MixArgs& a, const unsigned char* src, cudaStream_t st);
int fused_bank_u8(const unsigned char* imgs, float* bank, int N, int C, int H, int W, int b, cudaStream_t st);

// multi-pass path for large planes (ram_multipass.cu)
bool multipass_supported(int H, int W, int b);
size_t multipass_workspace_bytes(int n_images, int C, int H, int W, int b);
int multipass_mix(const MixArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);
int multipass_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, void* ws, size_t ws_bytes, cudaStream_t st);

// any-shape direct band-pruned DFT path (ram_generic.cu)
bool generic_supported(int H, int W, int b);
size_t generic_workspace_bytes(int n_images, int C, int H, int W, int b);
int generic_mix(const MixArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);
int generic_bank(const float* imgs, float* bank, int N, int C, int H, int W, int b, void* ws, size_t ws_bytes, cudaStream_t st);


// drop-in helpers for the per-sample reference API (ram_generic.cu)
size_t full_amp_workspace_bytes(int N, int C, int H, int W);
int full_amp(const float* imgs, float* amp, int N, int C, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st);
int band_blend(const float* a_src, const float* a_trg, float* out, int C, int H, int W, int b, int ratio_k, int ratio_den, cudaStream_t st);

}  // namespace ramdsir
