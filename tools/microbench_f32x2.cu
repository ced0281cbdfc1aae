// microbench_f32x2.cu — does FADD2/FFMA2 (packed f32x2) raise FP32 throughput on B200, or only halve
// the issue slots?  Measures scalar FMUL+FADD chains against packed chains with the same number
// of floating-point operations.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 x, u64 y) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ u64 mulnz(u64 x, u64 y, u64 nz) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(nz)); return r; }

template <int ILP> __global__ void scalar_k(float *out, float a, float b, int iters) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = __fadd_rn(__fmul_rn(v[i], a), b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void packed_k(float *out, float a, float b, int iters, u64 nz) {
  u64 v[ILP / 2];
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) v[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  u64 a2 = pk(a, a), b2 = pk(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP / 2; ++i) v[i] = add2(mulnz(v[i], a2, nz), b2);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000; const int ILP = 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); scalar_k<ILP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * ILP * iters * 148.0 * 8 * 256;
    printf("scalar FMUL+FADD : %.3f ms  %.2f TFLOP/s (%.1f Gop-instr/s)\n", ms, flops / ms / 1e9, flops / ms / 1e6 / 32);
    cudaEventRecord(e0); packed_k<ILP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters, 0x8000000080000000ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2+FADD2: %.3f ms  %.2f TFLOP/s (same flop count)\n", ms, flops / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
