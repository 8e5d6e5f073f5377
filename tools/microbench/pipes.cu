// Instruction-throughput probes for sm_100a (design input for the RAM kernels; not product code).
// Each probe: 148*k CTAs x 256 threads, ITER iterations of NCH independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__global__ void k_fadd(float* out, float c) {
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x+i;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) x[i] = x[i] + c;
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_ffma(float* out, float a, float b) {
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x+i;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) x[i] = fmaf(x[i], a, b);
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// 3 distinct *varying* registers per FFMA
__global__ void k_ffma3(float* out, float a, float b) {
  float x[8], y[8]; for (int i=0;i<8;i++) {x[i]=threadIdx.x+i; y[i]=a*i+b;}
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) x[i] = fmaf(x[i], y[i], y[(i+1)&7]);
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[8], A, B;
  asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b), "f"(b));
  for (int i=0;i<8;i++) asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"((float)threadIdx.x+i), "f"((float)i));
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
  }
  float s=0; for (int i=0;i<8;i++){ float lo,hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s+=lo+hi;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_fadd2(float* out, float a) {
  unsigned long long x[8], A;
  asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a), "f"(a));
  for (int i=0;i<8;i++) asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"((float)threadIdx.x+i), "f"((float)i));
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(A));
  }
  float s=0; for (int i=0;i<8;i++){ float lo,hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s+=lo+hi;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// butterfly-like: pairs of add/sub on varying registers (2 distinct sources)
__global__ void k_bfly(float* out, float c) {
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x*c+i;
  for (int it=0; it<ITER/2; it++) {
#pragma unroll
    for (int i=0;i<8;i+=2) { float a=x[i]+x[i+1], b=x[i]-x[i+1]; x[i]=a; x[i+1]=b*c; }
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_bfly2(float* out, float c) {
  unsigned long long x[8], C;
  asm("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(c), "f"(c));
  unsigned long long M; asm("mov.b64 %0, {%1,%2};" : "=l"(M) : "f"(-1.f), "f"(-1.f));
  for (int i=0;i<8;i++) asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"((float)threadIdx.x*c+i), "f"((float)i));
  for (int it=0; it<ITER/2; it++) {
#pragma unroll
    for (int i=0;i<8;i+=2) {
      unsigned long long a,b;
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(x[i]), "l"(x[i+1]));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(b) : "l"(x[i+1]), "l"(M), "l"(x[i]));
      asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(x[i+1]) : "l"(b), "l"(C));
      x[i]=a;
    }
  }
  float s=0; for (int i=0;i<8;i++){ float lo,hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s+=lo+hi;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_fmnmx(float* out, float c) {
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x+i;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) x[i] = fminf(x[i]*c, 1e30f);
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_shfl(float* out) {
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x+i;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1+(i&3));
  }
  float s=0; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int V> __global__ void k_lds(float* out) {
  extern __shared__ float4 sm4[];
  float* sm=(float*)sm4;
  for (int i=threadIdx.x;i<8192;i+=blockDim.x) sm[i]=i;
  __syncthreads();
  float acc=0; int base=threadIdx.x*V;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) {
      int idx=(base+i*256*V+it*V*32)&8191;
      if (V==1) acc+=sm[idx];
      else if (V==2) { float2 v=*(float2*)&sm[idx]; acc+=v.x+v.y; }
      else { float4 v=*(float4*)&sm[idx]; acc+=v.x+v.y+v.z+v.w; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template<int V> __global__ void k_sts(float* out) {
  extern __shared__ float4 sm4[];
  float* sm=(float*)sm4;
  int base=threadIdx.x*V;
  for (int it=0; it<ITER; it++) {
#pragma unroll
    for (int i=0;i<8;i++) {
      int idx=(base+i*256*V+it*V*32)&8191; float f=it+i;
      if (V==1) sm[idx]=f;
      else if (V==2) *(float2*)&sm[idx]=make_float2(f,f);
      else *(float4*)&sm[idx]=make_float4(f,f,f,f);
    }
  }
  __syncthreads();
  out[blockIdx.x*blockDim.x+threadIdx.x]=sm[threadIdx.x];
}
// mixed: FFMA chain + LDS.128 together (issue-slot sharing)
__global__ void k_mix(float* out, float a, float b) {
  extern __shared__ float4 sm4[];
  float* sm=(float*)sm4;
  for (int i=threadIdx.x;i<8192;i+=blockDim.x) sm[i]=i;
  __syncthreads();
  float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x+i;
  float acc=0;
  for (int it=0; it<ITER; it++) {
    float4 v=*(float4*)&sm[(threadIdx.x*4+it*128)&8191];
#pragma unroll
    for (int i=0;i<8;i++) x[i] = fmaf(x[i], a, b);
#pragma unroll
    for (int i=0;i<8;i++) x[i] = x[i] + v.x;
    acc+=v.y;
  }
  float s=acc; for (int i=0;i<8;i++) s+=x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<class F> float timeit(F f) {
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int clk=0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("dev %s sms %d smem/blk optin %zu l2 %d clock %d kHz\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize, clk);
  int nsm=p.multiProcessorCount; int grid=nsm*8, blk=256;
  float* out; CK(cudaMalloc(&out, (size_t)grid*blk*4));
  double lanes=(double)grid*blk*ITER*8;
  auto rep=[&](const char* n, float ms, double ops_per_lane_iter){
    double per_clk_sm = lanes*ops_per_lane_iter/(ms*1e-3)/nsm/1.0e9; // lane-ops per ns per SM
    printf("%-10s %8.3f ms  %8.1f lane-ops/ns/SM (divide by GHz for per-clk)\n", n, ms, per_clk_sm);
  };
  rep("fadd", timeit([&]{k_fadd<<<grid,blk>>>(out,1.5f);}),1);
  rep("ffma", timeit([&]{k_ffma<<<grid,blk>>>(out,1.0001f,0.5f);}),1);
  rep("ffma3", timeit([&]{k_ffma3<<<grid,blk>>>(out,1.0001f,0.5f);}),1);
  rep("ffma2", timeit([&]{k_ffma2<<<grid,blk>>>(out,1.0001f,0.5f);}),2);
  rep("fadd2", timeit([&]{k_fadd2<<<grid,blk>>>(out,1.5f);}),2);
  rep("bfly", timeit([&]{k_bfly<<<grid,blk>>>(out,0.7f);}),0.5*1.5/1.0); // per pair: 3 ops per 2 elems, ITER/2
  rep("bfly2", timeit([&]{k_bfly2<<<grid,blk>>>(out,0.7f);}),0.5*1.5*2);
  rep("fmul+fmnmx", timeit([&]{k_fmnmx<<<grid,blk>>>(out,1.0001f);}),2);
  rep("shfl", timeit([&]{k_shfl<<<grid,blk>>>(out);}),1);
  cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  rep("lds32", timeit([&]{k_lds<1><<<grid,blk,32768>>>(out);}),1);
  rep("lds64", timeit([&]{k_lds<2><<<grid,blk,32768>>>(out);}),2);
  rep("lds128", timeit([&]{k_lds<4><<<grid,blk,32768>>>(out);}),4);
  rep("sts32", timeit([&]{k_sts<1><<<grid,blk,32768>>>(out);}),1);
  rep("sts64", timeit([&]{k_sts<2><<<grid,blk,32768>>>(out);}),2);
  rep("sts128", timeit([&]{k_sts<4><<<grid,blk,32768>>>(out);}),4);
  { double l2=(double)grid*blk*ITER; float ms=timeit([&]{k_mix<<<grid,blk,32768>>>(out,1.0001f,0.5f);});
    printf("mix(8ffma+8fadd+1lds128) %8.3f ms  %8.1f fp-lane-ops/ns/SM\n", ms, l2*16/(ms*1e-3)/nsm/1e9); }
  // clock under load estimate: fadd kernel known ops -> report with nvidia-smi separately
  return 0;
}
