// GPU versions of the reference loader's geometric transforms on 8-bit images:
//   Resize            img.resize(size, BILINEAR), mask.resize(size, NEAREST)            dataset/transform.py:163-177
//   RandomScaleCrop   resize by a random factor, then RandomCrop                          dataset/transform.py:180-204
//   RandomCrop        zero / 255 padding to the crop size, random crop                    dataset/transform.py:16-44
//   partner image     Image.open(...).resize((256, 256), BILINEAR)                        dataset/fundus.py:209
// The arithmetic is Pillow's (the reference calls PIL.Image.resize; Pillow is an un-vendored dependency, 12.x here):
//   BILINEAR  two-pass separable convolution, horizontal first, with an 8-bit intermediate image.  Per output
//             coordinate: support = max(scale, 1), window [center - support, center + support] rounded and clipped,
//             triangle weights evaluated in double, normalised, then fixed to 22 fractional bits
//             ((int)(0.5 + w * 2^22)); accumulation starts at 2^21 and the result is clipped to [0, 255] after >> 22
//             (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc).
//   NEAREST   source index = (int)(x0 + k additions of the step), x0 = step / 2, the running double sum included
//             (libImaging/Geometry.c: ImagingScaleAffine).
// The coefficient / index tables are built on the host exactly as Pillow builds them (double arithmetic, same order of
// operations) and uploaded with the call; the kernels do the integer work.  Byte results are bit-exact.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "ram_internal.h"

namespace ramdsir {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

struct Coeffs {
  int ksize = 0;
  std::vector<int32_t> bounds;   // [out][2]: xmin, count
  std::vector<int32_t> kk;       // [out][ksize], 22-bit fixed point
};

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter, box = whole image.
Coeffs bilinear_coeffs(int in_size, int out_size) {
  Coeffs c;
  const double scale = (double)((float)in_size - 0.0f) / out_size;      // (in1 - in0) are floats in Pillow
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  c.ksize = (int)std::ceil(support) * 2 + 1;
  c.bounds.resize((size_t)out_size * 2);
  c.kk.assign((size_t)out_size * c.ksize, 0);
  std::vector<double> k(c.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      const double w = v < 1.0 ? 1.0 - v : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double p = k[x];
      c.kk[(size_t)xx * c.ksize + x] = p < 0 ? (int)(-0.5 + p * (1 << PRECISION_BITS)) : (int)(0.5 + p * (1 << PRECISION_BITS));
    }
    c.bounds[2 * xx] = xmin;
    c.bounds[2 * xx + 1] = xmax;
  }
  return c;
}

// Pillow's nearest-neighbour index table (ImagingScaleAffine): running sum in double, truncation, -1 = outside.
std::vector<int32_t> nearest_table(int in_size, int out_size) {
  std::vector<int32_t> t(out_size);
  const double a = (double)in_size / out_size;
  double xo = 0.0 + a * 0.5;
  for (int x = 0; x < out_size; ++x) {
    const int xin = xo < 0.0 ? -1 : (int)xo;
    t[x] = (xin >= 0 && xin < in_size) ? xin : -1;
    xo += a;
  }
  return t;
}

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRECISION_BITS;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [Hin][Win][C] rows y0 .. y0 + rows - 1  ->  tmp [rows][Wout][C]
__global__ void resize_h_kernel(const unsigned char* __restrict__ src, int Win, int C, int y0, int rows, int Wout, int ksize,
                                const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, unsigned char* __restrict__ tmp) {
  const size_t total = (size_t)rows * Wout * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C), xx = (int)((i / C) % Wout), yy = (int)(i / ((size_t)C * Wout));
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int32_t* k = kk + (size_t)xx * ksize;
    const unsigned char* row = src + ((size_t)(yy + y0) * Win + xmin) * C + ch;
    int ss = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < n; ++x) ss += (int)row[(size_t)x * C] * k[x];
    tmp[i] = clip8(ss);
  }
}

// vertical pass: tmp [rows][Wout][C] (row 0 = source row y0)  ->  dst [Hout][Wout][C]
__global__ void resize_v_kernel(const unsigned char* __restrict__ tmp, int y0, int Wout, int C, int Hout, int ksize,
                                const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, unsigned char* __restrict__ dst) {
  const size_t line = (size_t)Wout * C, total = line * Hout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int yy = (int)(i / line);
    const size_t xc = i - (size_t)yy * line;
    const int ymin = bounds[2 * yy] - y0, n = bounds[2 * yy + 1];
    const int32_t* k = kk + (size_t)yy * ksize;
    int ss = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < n; ++y) ss += (int)tmp[(size_t)(y + ymin) * line + xc] * k[y];
    dst[i] = clip8(ss);
  }
}

__global__ void resize_nearest_kernel(const unsigned char* __restrict__ src, int Win, int C, int Hout, int Wout,
                                      const int32_t* __restrict__ xtab, const int32_t* __restrict__ ytab, unsigned char* __restrict__ dst) {
  const size_t total = (size_t)Hout * Wout * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C), xx = (int)((i / C) % Wout), yy = (int)(i / ((size_t)C * Wout));
    const int xi = xtab[xx], yi = ytab[yy];
    dst[i] = (xi >= 0 && yi >= 0) ? src[((size_t)yi * Win + xi) * C + ch] : (unsigned char)0;
  }
}

// RandomCrop's pad (right / bottom, constant fill) + crop, and HWC -> CHW in the same pass
__global__ void crop_pad_chw_kernel(const unsigned char* __restrict__ src, int H, int W, int C, int x0, int y0, int oh, int ow,
                                    unsigned char fill, unsigned char* __restrict__ dst) {
  const size_t total = (size_t)C * oh * ow;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow), y = (int)((i / ow) % oh), ch = (int)(i / ((size_t)ow * oh));
    const int sx = x + x0, sy = y + y0;
    dst[i] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? src[((size_t)sy * W + sx) * C + ch] : fill;
  }
}

unsigned grid_for(size_t total) {
  size_t b = (total + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  return (unsigned)(b ? b : 1);
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace
}  // namespace ramdsir

using namespace ramdsir;

extern "C" {

int ramdsir_resize_tables_host(int in_size, int out_size, int filter, int32_t* bounds, int32_t* kk, int* ksize) {
  if (in_size < 1 || out_size < 1 || !ksize) {
    set_error("resize_tables_host: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  if (filter == RAMDSIR_FILTER_NEAREST) {
    *ksize = 1;
    if (bounds) {
      const std::vector<int32_t> t = nearest_table(in_size, out_size);
      memcpy(bounds, t.data(), t.size() * 4);
    }
    return RAMDSIR_OK;
  }
  if (filter != RAMDSIR_FILTER_BILINEAR) {
    set_error("resize_tables_host: unknown filter %d", filter);
    return RAMDSIR_ERR_INVALID;
  }
  const Coeffs c = bilinear_coeffs(in_size, out_size);
  *ksize = c.ksize;
  if (bounds) memcpy(bounds, c.bounds.data(), c.bounds.size() * 4);
  if (kk) memcpy(kk, c.kk.data(), c.kk.size() * 4);
  return RAMDSIR_OK;
}

size_t ramdsir_resize_workspace_bytes(int Hin, int Win, int C, int Hout, int Wout) {
  if (Hin < 1 || Win < 1 || C < 1 || Hout < 1 || Wout < 1) return 0;
  // temp image (all source rows in the worst case) + two coefficient tables + two bounds tables
  const double sx = (double)Win / Wout, sy = (double)Hin / Hout;
  const size_t kx = (size_t)std::ceil(sx < 1 ? 1 : sx) * 2 + 1, ky = (size_t)std::ceil(sy < 1 ? 1 : sy) * 2 + 1;
  return al256((size_t)Hin * Wout * C) + al256((size_t)Wout * kx * 4) + al256((size_t)Hout * ky * 4) +
         al256((size_t)Wout * 8) + al256((size_t)Hout * 8);
}

int ramdsir_resize_u8(const uint8_t* src_dev, int Hin, int Win, int C, uint8_t* dst_dev, int Hout, int Wout, int filter,
                      void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!src_dev || !dst_dev || Hin < 1 || Win < 1 || C < 1 || Hout < 1 || Wout < 1) {
    set_error("resize_u8: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  if (filter != RAMDSIR_FILTER_NEAREST && filter != RAMDSIR_FILTER_BILINEAR) {
    set_error("resize_u8: filter must be RAMDSIR_FILTER_NEAREST (0) or RAMDSIR_FILTER_BILINEAR (2), Pillow's numbering");
    return RAMDSIR_ERR_INVALID;
  }
  const size_t need = ramdsir_resize_workspace_bytes(Hin, Win, C, Hout, Wout);
  if (!workspace_dev || workspace_bytes < need) {
    set_error("resize_u8: needs %zu workspace bytes, got %zu", need, workspace_bytes);
    return RAMDSIR_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fused_trail_reset(st);
  unsigned char* ws = static_cast<unsigned char*>(workspace_dev);
  if (Hin == Hout && Win == Wout) {                     // Image.resize returns a copy
    RD_CUDA(cudaMemcpyAsync(dst_dev, src_dev, (size_t)Hin * Win * C, cudaMemcpyDeviceToDevice, st));
    return RAMDSIR_OK;
  }
  if (filter == RAMDSIR_FILTER_NEAREST) {
    const std::vector<int32_t> xt = nearest_table(Win, Wout), yt = nearest_table(Hin, Hout);
    int32_t* d_xt = reinterpret_cast<int32_t*>(ws);
    int32_t* d_yt = reinterpret_cast<int32_t*>(ws + al256((size_t)Wout * 4));
    RD_CUDA(cudaMemcpyAsync(d_xt, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice, st));
    RD_CUDA(cudaMemcpyAsync(d_yt, yt.data(), yt.size() * 4, cudaMemcpyHostToDevice, st));
    resize_nearest_kernel<<<grid_for((size_t)Hout * Wout * C), 256, 0, st>>>(src_dev, Win, C, Hout, Wout, d_xt, d_yt, dst_dev);
    count_launch();
    RD_CUDA(cudaGetLastError());
    return RAMDSIR_OK;
  }
  // bilinear: horizontal pass over the source rows the vertical pass reads, then vertical pass (Pillow's order)
  const bool need_h = Win != Wout, need_v = Hin != Hout;
  const Coeffs ch = bilinear_coeffs(Win, Wout), cv = bilinear_coeffs(Hin, Hout);
  size_t off = 0;
  unsigned char* tmp = ws + off; off += al256((size_t)Hin * Wout * C);
  int32_t* d_kh = reinterpret_cast<int32_t*>(ws + off); off += al256(ch.kk.size() * 4);
  int32_t* d_kv = reinterpret_cast<int32_t*>(ws + off); off += al256(cv.kk.size() * 4);
  int32_t* d_bh = reinterpret_cast<int32_t*>(ws + off); off += al256(ch.bounds.size() * 4);
  int32_t* d_bv = reinterpret_cast<int32_t*>(ws + off); off += al256(cv.bounds.size() * 4);
  if (off > workspace_bytes) {
    set_error("resize_u8: internal workspace layout (%zu) exceeds the reported size (%zu)", off, workspace_bytes);
    return RAMDSIR_ERR_WORKSPACE;
  }
  int y0 = 0, rows = Hin;
  if (need_v) {                                          // only the rows the vertical pass reads
    y0 = cv.bounds[0];
    rows = cv.bounds[2 * (Hout - 1)] + cv.bounds[2 * (Hout - 1) + 1] - y0;
  }
  const unsigned char* vin = src_dev;
  if (need_h) {
    RD_CUDA(cudaMemcpyAsync(d_kh, ch.kk.data(), ch.kk.size() * 4, cudaMemcpyHostToDevice, st));
    RD_CUDA(cudaMemcpyAsync(d_bh, ch.bounds.data(), ch.bounds.size() * 4, cudaMemcpyHostToDevice, st));
    unsigned char* hout = need_v ? tmp : dst_dev;
    resize_h_kernel<<<grid_for((size_t)rows * Wout * C), 256, 0, st>>>(src_dev, Win, C, y0, rows, Wout, ch.ksize, d_bh, d_kh, hout);
    count_launch();
    vin = hout;
  } else {
    vin = src_dev + (size_t)y0 * Win * C;                // the vertical pass indexes rows relative to y0
  }
  if (need_v) {
    RD_CUDA(cudaMemcpyAsync(d_kv, cv.kk.data(), cv.kk.size() * 4, cudaMemcpyHostToDevice, st));
    RD_CUDA(cudaMemcpyAsync(d_bv, cv.bounds.data(), cv.bounds.size() * 4, cudaMemcpyHostToDevice, st));
    resize_v_kernel<<<grid_for((size_t)Hout * Wout * C), 256, 0, st>>>(vin, y0, Wout, C, Hout, cv.ksize, d_bv, d_kv, dst_dev);
    count_launch();
  }
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

int ramdsir_crop_pad_chw_u8(const uint8_t* src_hwc_dev, int H, int W, int C, int x0, int y0, int out_h, int out_w,
                            int fill, uint8_t* dst_chw_dev, void* stream) {
  if (!src_hwc_dev || !dst_chw_dev || H < 1 || W < 1 || C < 1 || out_h < 1 || out_w < 1 || fill < 0 || fill > 255) {
    set_error("crop_pad_chw_u8: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fused_trail_reset(st);
  crop_pad_chw_kernel<<<grid_for((size_t)C * out_h * out_w), 256, 0, st>>>(src_hwc_dev, H, W, C, x0, y0, out_h, out_w,
                                                                           (unsigned char)fill, dst_chw_dev);
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

}  // extern "C"
