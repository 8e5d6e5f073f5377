// How much HBM bandwidth does the fused kernel's ACCESS PATTERN admit?  Copy 768 planes of 256 x 256 float (the bench
// batch) with no arithmetic, (a) linearly, (b) in the kernel's pattern: a persistent CTA walks the four 16-float4
// column strips of its plane one after the other, 256 B per image row at a 1 KB pitch, reads everything, then writes
// the strips back in reverse order.  Build: nvcc -arch=sm_100a -O3 -o strip_copy strip_copy.cu ; run: ./strip_copy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) linear_copy(const float4* __restrict__ x, float4* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = x[i];
}
// mode 0: read strip, write strip (one pass);  mode 1: read all four strips (kept in registers: 64 float4 would not fit,
// so strip by strip with a dependent dummy), then write in reverse order -- here simply two passes: forward read-only
// pass over the plane (sum kept alive), then read again + write (the kernel's second read comes from L2).
template <int MODE>
__global__ void __launch_bounds__(256, 2) strip_copy(const float4* __restrict__ x, float4* __restrict__ y, int nplanes, float* sink) {
  const int t = threadIdx.x, h1 = t >> 4, c = t & 15;
  float acc = 0.f;
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
    const float4* xp = x + (size_t)plane * 256 * 64;
    float4* yp = y + (size_t)plane * 256 * 64;
    if (MODE == 1) {
      for (int strip = 0; strip < 4; ++strip) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __ldg(xp + (size_t)(h1 + 16 * j) * 64 + strip * 16 + c);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += v[j].x + v[j].w;
      }
    }
    for (int strip = 3; strip >= 0; --strip) {
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __ldg(xp + (size_t)(h1 + 16 * j) * 64 + strip * 16 + c);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[j].x += acc * 0.f; yp[(size_t)(h1 + 16 * j) * 64 + strip * 16 + c] = v[j]; }
    }
  }
  if (acc == 123.456f) *sink = acc;
}
int main() {
  const int nplanes = 768, reps = 50, nbuf = 3;
  const size_t n4 = (size_t)nplanes * 256 * 64;
  float4 *x[nbuf], *y[nbuf]; float* sink;
  for (int i = 0; i < nbuf; ++i) { cudaMalloc(&x[i], n4 * 16); cudaMalloc(&y[i], n4 * 16); cudaMemset(x[i], 1, n4 * 16); }
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    for (int i = 0; i < 5; ++i) launch(i);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch(i);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-28s %.4f ms  %.0f GB/s (read + write of the batch)\n", name, ms, 2.0 * n4 * 16 / ms / 1e6);
  };
  timeit("linear copy", [&](int i) { linear_copy<<<148 * 8, 256>>>(x[i % nbuf], y[i % nbuf], n4); });
  timeit("strip pattern, one pass", [&](int i) { strip_copy<0><<<296, 256>>>(x[i % nbuf], y[i % nbuf], nplanes, sink); });
  timeit("strip pattern, read twice", [&](int i) { strip_copy<1><<<296, 256>>>(x[i % nbuf], y[i % nbuf], nplanes, sink); });
  timeit("strip pattern, 592 CTAs", [&](int i) { strip_copy<0><<<592, 256>>>(x[i % nbuf], y[i % nbuf], nplanes, sink); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
