// Micro-benchmark: cost of a thread-block-cluster barrier and of DSMEM reads on sm_100a (2x2 clusters of 256-thread CTAs
// with ~100 KB of dynamic shared memory each, i.e. the residency of the fused tile kernel).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE> __global__ void __launch_bounds__(256, 2) k(float* out, long long* cyc, int iters) {
  extern __shared__ float sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += 256) sm[i] = (float)(i + blockIdx.x);
  cl.sync();
  const unsigned nb = (cl.block_rank() + 1) % cl.num_blocks();
  const float* peer = cl.map_shared_rank(sm, nb);
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) __syncthreads();
    if (MODE == 1) cl.sync();
    if (MODE == 2) { cl.sync(); acc += peer[(tid + it) & 4095]; }                 // one ring-sized DSMEM read per thread
    if (MODE == 3) { __syncthreads(); acc += sm[(tid + it) & 4095]; }
    sm[(tid * 7 + it) & 4095] += 1.f;
  }
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 256 + tid] = acc + sm[tid];
}
template <int MODE> void run(const char* name, float* d, long long* c, int iters) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 * 2 * 4, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = 100 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, d, c, iters);
  cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s %s  cycles/iter (block 0) %.1f\n", name, cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e), (double)h[0] / iters);
}
int main() {
  float* d; long long* c;
  cudaMalloc(&d, 148 * 8 * 256 * 4 * 4); cudaMalloc(&c, 148 * 8 * 8 * 4);
  run<0>("__syncthreads", d, c, 2000);
  run<1>("cluster.sync (4 CTAs)", d, c, 2000);
  run<2>("cluster.sync + DSMEM read", d, c, 2000);
  run<3>("__syncthreads + local smem read", d, c, 2000);
  return 0;
}
