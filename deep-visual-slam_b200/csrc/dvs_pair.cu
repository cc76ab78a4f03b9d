// Two-source tile kernel (dvs_pair_core.cuh) and its launcher: one instantiation per (gradients, input formats).
// Kept in its own translation unit so that the eight variants compile in parallel with the generic kernels.
#include <cuda_runtime.h>

#include <atomic>

#include "dvs_pair_core.cuh"
#include "dvs_host.h"

namespace dvs {

// One tile per CTA, grid = number of tiles.  Persistent CTAs (2 per SM) pulling tiles from a global counter were measured
// 4 % SLOWER at config 2 (1.650 vs 1.589 ms, profiles/r02_experiments.md): CTAs launched together stay phase-locked, so
// the two CTAs of an SM sit in the latency-bound gather phase at the same time, whereas CTAs of a plain grid retire and
// start at different times and overlap gather with arithmetic.
// Resident CTAs per SM the kernel is compiled for: 3 with 28-row tiles (224 threads, <= 97 registers, 74.5 KB of shared
// memory each), 2 with 32-row tiles.
#if !defined(DVS_PAIR_MINBLOCKS)
#define DVS_PAIR_MINBLOCKS (DVS_TILE_ROWS <= 28 ? 3 : 2)
#endif
template <bool GRAD, int IO>
__global__ void __launch_bounds__(NT, DVS_PAIR_MINBLOCKS) fused_pair_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  PairLayout P;
  PairState st;
  const Tile t = make_tile(p, blockIdx.x);

  phase_consts_at<2>(p, t, sm + P.consts(), tid, sm + P.a2());
  pair_phase_load<IO>(p, t, sm, tid, st);
  __syncthreads();
  pair_phase_identity(p, t, sm, tid, st);
  __syncthreads();

  for (int s = 0; s < p.S; ++s) {
    pair_reset_scale_state(st);
    pair_phase_warp<IO>(p, t, sm, tid, s);
    __syncthreads();
    pair_phase_stats<GRAD>(p, t, sm, tid, s, st);
    __syncthreads();
    const bool direct = (p.dh[s] == p.H && p.dw[s] == p.W);
    if (GRAD) {
      pair_phase_grad<IO>(p, t, sm, tid, s, st);
      __syncthreads();
      if (direct) pair_store_gdu_direct(p, t, tid, s, st);
      else pair_stage_gdu(sm, tid, st);
    }
    pair_reduce_write(sm, tid, st);
    __syncthreads();
    if (GRAD && !direct) adjoint_rows_at(P, p, t, sm, tid, s);
    reduce_stage1_at<2>(P, sm, tid);
    __syncthreads();
    if (GRAD && !direct) adjoint_cols_at(P, p, t, sm, tid, s);
    reduce_stage2_at<2>(P, p, t, sm, tid, s);
    // as in fused_tile_kernel: the next warp phase writes only X / DU, which nobody reads any more
  }
}

constexpr int kMaxDev = 64;

template <bool GRAD, int IO>
static cudaError_t launch(const FusedParams& p, int nblk, cudaStream_t st) {
  PairLayout P;
  const size_t bytes = (size_t)P.total() * sizeof(float);
  // the dynamic shared-memory opt-in is a per-device attribute of the function: one flag per (instantiation, device)
  static std::atomic<bool> configured[kMaxDev];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDev || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(fused_pair_kernel<GRAD, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDev) configured[dev].store(true, std::memory_order_release);
  }
  fused_pair_kernel<GRAD, IO><<<nblk, NT, bytes, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pair_kernel(const FusedParams& p, int nblk, cudaStream_t st) {
  switch ((p.want_grad ? 4 : 0) | (p.io_flags & 3)) {
    case 0: return launch<false, 0>(p, nblk, st);
    case 1: return launch<false, 1>(p, nblk, st);
    case 2: return launch<false, 2>(p, nblk, st);
    case 3: return launch<false, 3>(p, nblk, st);
    case 4: return launch<true, 0>(p, nblk, st);
    case 5: return launch<true, 1>(p, nblk, st);
    case 6: return launch<true, 2>(p, nblk, st);
    case 7: return launch<true, 3>(p, nblk, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace dvs
