// Micro-benchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a, alone and mixed with integer work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned iop(unsigned a, unsigned b) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[16]; unsigned q[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) q[i] = threadIdx.x + i;
  u64 p[8];
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  u64 ss = pk(s, s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, s);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { a[i] = fma1(a[i], s, s); q[i & 7] = iop(q[i & 7], q[(i + 1) & 7]); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, ss); q[i] = iop(q[i], q[(i + 1) & 7]); q[(i + 3) & 7] = iop(q[(i + 3) & 7], q[i]); }
    } else if (MODE == 4) {          // 16 FFMA + 8 independent integer ops: issue bound if FFMA is 1 slot (24 slots)
#pragma unroll
      for (int i = 0; i < 16; ++i) { a[i] = fma1(a[i], s, s); if (i & 1) q[i >> 1] = iop(q[i >> 1], 0x9e3779b9u); }
    } else if (MODE == 5) {          // 8 FFMA2 + 8 independent integer ops: 16 slots if FFMA2 takes one issue slot, 24 if two
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, ss); q[i] = iop(q[i], 0x9e3779b9u); }
    }
  }
  float r = 0; unsigned z = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; up(p[i], x, y); r += x + y; z ^= q[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)z;
}
template <int MODE> void run(const char* name, float* d, int iters, double fp_lanes_per_iter, double inst_per_iter) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(d, iters, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, iters, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = blocks * 8.0;
  double winst = warps * iters * inst_per_iter;
  double lanes = warps * iters * fp_lanes_per_iter * 32;
  printf("%-28s %8.3f ms  warp-inst/clk/SMSP(@1.965GHz) %.3f  fma-lanes/clk/SM %.1f\n", name, ms,
         winst / (ms * 1e-3 * 1.965e9 * 148 * 4), lanes / (ms * 1e-3 * 1.965e9 * 148));
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  int iters = 4096;
  run<0>("FFMA x16", d, iters, 16, 16);
  run<1>("FFMA2 x8", d, iters, 16, 8);
  run<2>("FFMA x16 + LOP3 x16", d, iters, 16, 32);
  run<3>("FFMA2 x8 + LOP3 x16", d, iters, 16, 24);
  run<4>("FFMA x16 + LOP3 x8 (indep)", d, iters, 16, 24);
  run<5>("FFMA2 x8 + LOP3 x8 (indep)", d, iters, 16, 16);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
