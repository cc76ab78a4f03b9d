// Fused view-synthesis loss for sm_100a: kernels + C ABI (include/dvsloss.h).
//
// Launch sequence of dvs_photometric_forward (all on the caller's stream, no host sync):
//   1. mean_partial_kernel   partial sums of the up-sampled disparity per (scale, image)      [reads disp once];
//                            when pose parameters are the inputs: the 4x4 matrices (transformation_from_parameters)
//   2. fused_tile_kernel / fused_pair_kernel (dvs_pair.cu)   one CTA per 30x30 tile x image, all scales, loss sums (+ unit gradients)
//   3. postpass_kernel       per (scale, image): fixed-order reduction of the per-CTA partials, pose gradient
//                            dL/dT = K^T dL/dP (-> d/d(axisangle, translation)), smoothness mean-coupling coefficient;
//                            unit gradients of the up-sampled scales (fixed-order sum of the tiles' coarse boxes);
//                            the last block forms loss/s and loss
// dvs_photometric_backward is one elementwise kernel: grad = g_s * (unit - coupling), pose combine.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/dvsloss.h"
#include <cuda_bf16.h>

#include "dvs_fused_core.cuh"
#include "dvs_host.h"
#include "dvs_pose.cuh"

namespace dvs {

constexpr int kMaxDevices = 64;

// ------------------------------------------------------------------------------------------------ 1. disparity mean
// mean(up-sample(d)) is a fixed linear functional of d: sum_ij rw[i] cw[j] d[i,j] / (H W); for the
// exact power-of-two pyramids of the reference the weights are the constant (H/h)(W/w).
// Pose parameters handed to the loss instead of matrices: the pre-pass builds
// T_i = transformation_from_parameters(axisangle_i, translation_i, invert_i) (vo/learner_func.py:29-104) into the workspace,
// the post-pass chains d loss / d T_i back to the six pose numbers -- no separate pose launches.
struct PoseIO {
  const float* aa[kMaxN];     // [B,3] each; aa[0] == null: matrices were given
  const float* tr[kMaxN];
  int invert[kMaxN];
  float* Tws;                 // [N][B][16]
  float* uP;                  // [S][N][B][6] unit gradients (axis-angle, translation)
  int* done;                  // post-pass ticket counter (zeroed here)
};

__global__ void __launch_bounds__(256) mean_partial_kernel(FusedParams p, float* mean_part, PoseIO pose) {
  const int chunk = blockIdx.x, b = blockIdx.y, s = blockIdx.z;
  if ((chunk | s) == 0 && pose.aa[0] && threadIdx.x < p.N) {
    const int i = threadIdx.x;
    pose_matrix(pose.aa[i] + 3 * b, pose.tr[i] + 3 * b, pose.invert[i], pose.Tws + ((size_t)i * p.B + b) * 16);
  }
  if ((chunk | b | s) == 0 && threadIdx.x == 0) *pose.done = 0;
  const int h = p.dh[s], w = p.dw[s], n = h * w;
  const float* d = p.disp[s] + (size_t)b * n;
  const bool exact = (p.H % h == 0) && (p.W % w == 0);
  const int per = (n + kMeanBlocks - 1) / kMeanBlocks;
  const int lo = chunk * per, hi = min(lo + per, n);
  float acc = 0.f;
  const bool bf16 = (p.io_flags & 1) != 0;
  const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(p.disp[s]) + (size_t)b * n;
  auto at = [&](int e) -> float { return bf16 ? __bfloat162float(db[e]) : d[e]; };
  if (exact) {
    // constant weight: plain sum over groups of four consecutive elements with four independent accumulators (several
    // loads in flight per thread), scalar head/tail.  The summation order is a function of the element indices only, so
    // fp32 and bf16 maps of the same values give the same bits; fp32 maps use 128-bit loads when the image is aligned.
    const float cw = (float)((p.H / h) * (p.W / w));
    const int a0 = min((lo + 3) & ~3, hi), a1 = max(a0, hi & ~3);
    const bool vec = !bf16 && (((uintptr_t)d) & 15) == 0;
    const float4* d4 = reinterpret_cast<const float4*>(d);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int e = a0 / 4 + threadIdx.x; e < a1 / 4; e += blockDim.x) {
      float4 v;
      if (vec) v = d4[e];
      else v = make_float4(at(4 * e), at(4 * e + 1), at(4 * e + 2), at(4 * e + 3));
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
    acc = (s0 + s1) + (s2 + s3);
    for (int e = lo + threadIdx.x; e < a0; e += blockDim.x) acc += at(e);
    for (int e = a1 + threadIdx.x; e < hi; e += blockDim.x) acc += at(e);
    acc *= cw;
  } else {
    for (int e = lo + threadIdx.x; e < hi; e += blockDim.x) {
      int i = e / w, j = e - i * w;
      acc = fmaf(up_weight(i, h, p.H) * up_weight(j, w, p.W), at(e), acc);
    }
  }
  __shared__ float red[8];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    mean_part[(s * p.B + b) * kMeanBlocks + chunk] = t;
  }
}

// ------------------------------------------------------------------------------------------------ 2. fused tile kernel
template <int NS, bool GRAD>
__global__ void __launch_bounds__(NT, (NS <= 2 ? 2 : 1)) fused_tile_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const Tile t = make_tile(p, blockIdx.x);
  ThreadState<NS> st;

  phase_consts<NS>(p, t, sm, tid);
  phase_load<NS>(p, t, sm, tid, st);
  __syncthreads();
  phase_identity<NS>(p, t, sm, tid, st);
  __syncthreads();

  for (int s = 0; s < p.S; ++s) {
    reset_scale_state<NS>(st);
    phase_warp<NS>(p, t, sm, tid, s, st);
    __syncthreads();
    phase_stats<NS, GRAD>(p, t, sm, tid, s, st);
    __syncthreads();
    const bool direct = (p.dh[s] == p.H && p.dw[s] == p.W);
    if (GRAD) {
      phase_grad<NS>(p, t, sm, tid, s, st);
      __syncthreads();
      if (direct) store_gdu_direct<NS>(p, t, tid, s, st);
      else stage_gdu<NS>(p, t, sm, tid, st);
    }
    reduce_write<NS>(p, sm, tid, st);
    __syncthreads();
    if (GRAD && !direct) adjoint_rows<NS>(p, t, sm, tid, s);
    reduce_stage1<NS>(p, sm, tid);
    __syncthreads();
    if (GRAD && !direct) adjoint_cols<NS>(p, t, sm, tid, s);
    reduce_stage2<NS>(p, t, sm, tid, s);
    // no barrier needed here: the next phase_warp writes only X/DU, which nobody reads any more
    // (adjoint_cols / reduce_stage2 read the F region, next written after the following barrier) ...
    // ... except adjoint_rows' input DU: all threads passed the barrier after adjoint_rows already.
  }
}

// ------------------------------------------------------------------------------------------------ 3. post-pass
// One launch after the tile kernel:
//   blocks [0, B S)        per (image, scale): fixed-order reduction over the image's tiles (consecutive threads own
//                          consecutive values, groups of threads stride the tiles, groups added in order -> deterministic);
//                          pose moments -> dL/dP -> dL/dT =
//                          K^T dL/dP (-> the six pose parameters when those were the inputs); smoothness mean-coupling
//                          coefficient.  The block that finishes LAST (ticket counter) also forms loss/s and loss from
//                          the per-image sums, again in a fixed order.
//   blocks [B S, ...)      unit gradients of the up-sampled scales: one thread per disparity element gathers the tile
//                          boxes (gather_gdisp).
struct FinishParams {
  int B, H, W, N, S, tiles_per_img, want_grad, gather_blocks;
  float smooth_w;
  const float* part;        // [nblk][S][nv]
  const float* mean_part;   // [S][B][kMeanBlocks]
  const float* K;           // [B,4,4]
  const float* invK;        // [B,4,4]
  float* perimg;            // [S][B][3]
  float* uT;                // [S][N][B][16] or null
  float* coup;              // [S][B] or null
  float* loss_per_scale;    // [S]
  float* loss_total;        // [1]
};
__global__ void __launch_bounds__(256) postpass_kernel(const __grid_constant__ FusedParams p, FinishParams f, PoseIO pose) {
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= f.B * f.S) {
    // ---- gather: grid-stride over the elements of every up-sampled scale
    const int g = blockIdx.x - f.B * f.S;
    for (int s = 0; s < p.S; ++s) {
      const int dh = p.dh[s], dw = p.dw[s];
      if (dh == p.H && dw == p.W) continue;
      const int n = p.B * dh * dw;
      for (int e = g * blockDim.x + tid; e < n; e += f.gather_blocks * blockDim.x) {
        const int b = e / (dh * dw), r = e - b * dh * dw;
        const int I = r / dw, J = r - I * dw;
        p.gdisp[s][e] = gather_gdisp(p, s, b, I, J);
      }
    }
    return;
  }
  const int b = blockIdx.x % f.B, s = blockIdx.x / f.B;
  const int nv = 3 + 12 * f.N;
  __shared__ float res[3 + 12 * kMaxN];
  __shared__ float grp[8][64];
  __shared__ int last;
  {
    // fixed-order sum over the image's tiles: consecutive threads own consecutive values (one coalesced line per tile), G groups
    // of threads stride the tiles with four independent loads in flight, then the groups are added in order
    const int VW = nv <= 32 ? 32 : 64, G = 256 / VW;
    const int v = tid % VW, gq = tid / VW;
    if (v < nv) {
      const size_t stride = (size_t)f.S * nv;
      const float* base = f.part + ((size_t)b * f.tiles_per_img * f.S + s) * nv + v;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int tl = gq;
      for (; tl + 3 * G < f.tiles_per_img; tl += 4 * G) {
        a0 += base[(size_t)tl * stride];
        a1 += base[(size_t)(tl + G) * stride];
        a2 += base[(size_t)(tl + 2 * G) * stride];
        a3 += base[(size_t)(tl + 3 * G) * stride];
      }
      for (; tl < f.tiles_per_img; tl += G) a0 += base[(size_t)tl * stride];
      grp[gq][v] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < nv) {
      float a = 0.f;
      for (int q = 0; q < G; ++q) a += grp[q][tid];
      res[tid] = a;
    }
  }
  __syncthreads();
  if (tid < 3) f.perimg[(s * f.B + b) * 3 + tid] = res[tid];
  if (f.want_grad) {
    if (tid < f.N) {
      // pose moments -> dL/dP -> dL/dT = K^T dL/dP
      float dT[16];
      moments_to_dT(res + 3 + 12 * tid, f.K + b * 16, f.invK + b * 16, dT);
      if (f.uT)
        for (int e = 0; e < 16; ++e) f.uT[(((size_t)s * f.N + tid) * f.B + b) * 16 + e] = dT[e];
      if (pose.aa[0]) {
        float* o = pose.uP + (((size_t)s * f.N + tid) * f.B + b) * 6;
        pose_matrix_grad(dT, pose.aa[tid] + 3 * b, pose.tr[tid] + 3 * b, pose.invert[tid], o, o + 3);
      }
    }
    if (tid == 64) {
      const float* mp = f.mean_part + (s * f.B + b) * kMeanBlocks;
      float mu = 0.f;
      for (int k = 0; k < kMeanBlocks; ++k) mu += mp[k];
      mu = mu / ((float)f.H * (float)f.W);
      float inv = 1.0f / (fmaxf(mu, 0.001f) + 1e-7f);
      float live = mu >= 0.001f ? 1.f : 0.f;
      float kap = f.smooth_w / (float)(1 << s);
      float Nx = (float)f.B * (float)f.H * (float)(f.W - 1), Ny = (float)f.B * (float)(f.H - 1) * (float)f.W;
      // sum_q gn_q * n_q == kappa * (Sx/Nx + Sy/Ny) (Euler: the term is 1-homogeneous in n)
      f.coup[s * f.B + b] = kap * (res[1] / Nx + res[2] / Ny) * inv * live / ((float)f.H * (float)f.W);
    }
  }
  // ---- the last (image, scale) block to get here forms the losses
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(pose.done, 1) == f.B * f.S - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ float ls[kMaxS];
  if (tid < f.S) {
    float ph = 0.f, sx = 0.f, sy = 0.f;
    for (int bb = 0; bb < f.B; ++bb) {
      ph += __ldcg(f.perimg + (tid * f.B + bb) * 3 + 0);
      sx += __ldcg(f.perimg + (tid * f.B + bb) * 3 + 1);
      sy += __ldcg(f.perimg + (tid * f.B + bb) * 3 + 2);
    }
    float kap = f.smooth_w / (float)(1 << tid);
    float Nx = (float)f.B * (float)f.H * (float)(f.W - 1), Ny = (float)f.B * (float)(f.H - 1) * (float)f.W;
    float l = ph / ((float)f.B * (float)f.H * (float)f.W) + kap * (sx / Nx + sy / Ny);
    ls[tid] = l;
    f.loss_per_scale[tid] = l;
  }
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int k = 0; k < f.S; ++k) t += ls[k];
    f.loss_total[0] = t / (float)f.S;
  }
}

// ------------------------------------------------------------------------------------------------ backward (scaling)
struct BackwardParams {
  int B, H, W, N, S;
  int dh[kMaxS], dw[kMaxS];
  const float* g;                 // [S]
  const float* u[kMaxS];
  float* out[kMaxS];
  const float* uT;                // [S][N][B][16]
  const float* coup;              // [S][B]
  float* gT[kMaxN];
  int out_bf16;                   // grad_disp tensors are bf16 (disparities were bf16)
  // pose-parameter mode: uT is [S][N][B][6] and the outputs are gaa[i], gtr[i] [B,3] instead of gT[i] [B,4,4]
  int pose;
  float* gaa[kMaxN];
  float* gtr[kMaxN];
};
__global__ void __launch_bounds__(256) backward_scale_kernel(BackwardParams q) {
  const int s = blockIdx.y;
  if (s == q.S && q.pose) {   // grad_(axisangle, translation)[i][b] = sum_s g_s uP[s][i][b]
    const int n = q.N * q.B * 6;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const int i = e / (q.B * 6), r = e - i * q.B * 6, b = r / 6, k = r - b * 6;
      float a = 0.f;
      for (int kk = 0; kk < q.S; ++kk) a = fmaf(q.g[kk], q.uT[((size_t)kk * q.N + i) * q.B * 6 + r], a);
      if (k < 3) q.gaa[i][b * 3 + k] = a; else q.gtr[i][b * 3 + k - 3] = a;
    }
    return;
  }
  if (s == q.S) {   // pose: grad_T[i][b] = sum_s g_s uT[s][i][b]
    int n = q.N * q.B * 16;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      int i = e / (q.B * 16), r = e - i * q.B * 16;
      float a = 0.f;
      for (int k = 0; k < q.S; ++k) a = fmaf(q.g[k], q.uT[((size_t)k * q.N + i) * q.B * 16 + r], a);
      q.gT[i][r] = a;
    }
    return;
  }
  const int h = q.dh[s], w = q.dw[s];
  const size_t hw = (size_t)h * w, n = (size_t)q.B * hw;
  const float gs = q.g[s];
  const bool exact = (q.H % h == 0) && (q.W % w == 0);
  const float cwc = exact ? (float)((q.H / h) * (q.W / w)) : 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q.out_bf16) {
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(q.out[s]);
    for (size_t e = t0; e < n; e += stride) {
      const int b = (int)(e / hw);
      float wgt = cwc;
      if (!exact) {
        const int r = (int)(e - (size_t)b * hw);
        wgt = up_weight(r / w, h, q.H) * up_weight(r - (r / w) * w, w, q.W);
      }
      ob[e] = __float2bfloat16_rn(gs * (q.u[s][e] - q.coup[s * q.B + b] * wgt));
    }
    return;
  }
  if (exact && (hw & 3) == 0 && ((((uintptr_t)q.u[s]) | ((uintptr_t)q.out[s])) & 15) == 0) {
    // constant up-sampling weight and whole float4s inside one image: 128-bit loads and stores
    const float4* u4 = reinterpret_cast<const float4*>(q.u[s]);
    float4* o4 = reinterpret_cast<float4*>(q.out[s]);
    for (size_t e = t0; e < n / 4; e += stride) {
      const float c = q.coup[s * q.B + (int)((e * 4) / hw)] * cwc;
      float4 v = u4[e];
      v.x = gs * (v.x - c); v.y = gs * (v.y - c); v.z = gs * (v.z - c); v.w = gs * (v.w - c);
      o4[e] = v;
    }
    return;
  }
  for (size_t e = t0; e < n; e += stride) {
    int b = (int)(e / hw);
    float wgt = cwc;
    if (!exact) {
      int r = (int)(e - (size_t)b * hw);
      int i = r / w, j = r - i * w;
      wgt = up_weight(i, h, q.H) * up_weight(j, w, q.W);
    }
    q.out[s][e] = gs * (q.u[s][e] - q.coup[s * q.B + b] * wgt);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
  size_t mean_part, part, perimg, uT, coup, lossbuf, counter, Tws, cpart, total;
  int tiles_x, tiles_y, nblk;
  int cstride, coff[kMaxS], cbw[kMaxS];
};
static WsLayout ws_layout(const DvsShape& sh) {
  WsLayout w;
  w.tiles_x = (sh.W + PITCH_X - 1) / PITCH_X;
  w.tiles_y = (sh.H + PITCH_Y - 1) / PITCH_Y;
  w.nblk = sh.B * w.tiles_x * w.tiles_y;
  size_t o = 0;
  w.mean_part = o; o = align_up(o + sizeof(float) * sh.S * sh.B * kMeanBlocks, 256);
  w.part = o;      o = align_up(o + sizeof(float) * (size_t)w.nblk * sh.S * nvals(sh.N), 256);
  w.perimg = o;    o = align_up(o + sizeof(float) * sh.S * sh.B * 3, 256);
  w.uT = o;        o = align_up(o + sizeof(float) * sh.S * sh.N * sh.B * 16, 256);   // used by backward_recompute
  w.coup = o;      o = align_up(o + sizeof(float) * sh.S * sh.B, 256);
  w.lossbuf = o;   o = align_up(o + sizeof(float) * 8, 256);
  w.counter = o;   o = align_up(o + sizeof(int), 256);
  w.Tws = o;       o = align_up(o + sizeof(float) * sh.N * sh.B * 16, 256);
  w.cstride = 0;
  for (int s = 0; s < sh.S; ++s) {
    const bool direct = sh.dh[s] == sh.H && sh.dw[s] == sh.W;
    w.coff[s] = w.cstride;
    w.cbw[s] = direct ? 0 : coarse_box_extent(sh.dw[s], sh.W, PITCH_X);
    w.cstride += direct ? 0 : coarse_box_extent(sh.dh[s], sh.H, PITCH_Y) * w.cbw[s];
  }
  w.cpart = o;     o = align_up(o + sizeof(float) * (size_t)w.nblk * w.cstride, 256);
  w.total = o;
  return w;
}

static int check_shape(const DvsShape* sh) {
  if (!sh) return DVS_EINVAL;
  if (sh->B < 1 || sh->H < 2 || sh->W < 2) return DVS_EINVAL;
  if (sh->H > 32767 || sh->W > 65535) return DVS_EINVAL;                          // (row << 16) | column position table
  if (sh->N < 1 || sh->N > DVS_MAX_SOURCES || sh->S < 1 || sh->S > DVS_MAX_SCALES) return DVS_EINVAL;
  for (int s = 0; s < sh->S; ++s)
    if (sh->dh[s] < 1 || sh->dw[s] < 1 || sh->dh[s] > sh->H || sh->dw[s] > sh->W) return DVS_EINVAL;
  if ((long long)sh->B * sh->H * sh->W * 3 >= (1LL << 31)) return DVS_EINVAL;   // 32-bit pixel offsets per image set
  return DVS_OK;
}

template <int NS, bool GRAD>
static cudaError_t launch_tile(const FusedParams& p, int nblk, cudaStream_t st) {
  SmemLayout L{NS};
  size_t bytes = (size_t)L.total() * sizeof(float);
  // the dynamic shared-memory opt-in is a per-device attribute of the function: one flag per (instantiation, device)
  static std::atomic<bool> configured[kMaxDevices];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(fused_tile_kernel<NS, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) configured[dev].store(true, std::memory_order_release);
  }
  fused_tile_kernel<NS, GRAD><<<nblk, NT, bytes, st>>>(p);
  return cudaGetLastError();
}

// DVS_GENERIC_KERNEL=1 routes two-source problems through the generic kernel (A/B measurements, regression tests).
static bool use_generic_kernel() {
  static const bool v = [] { const char* e = getenv("DVS_GENERIC_KERNEL"); return e && e[0] == '1'; }();
  return v;
}

static cudaError_t dispatch_tile(const FusedParams& p, int nblk, cudaStream_t st) {
  if (p.N == 2 && (!use_generic_kernel() || p.io_flags)) return launch_pair_kernel(p, nblk, st);
  if (p.io_flags) return cudaErrorNotSupported;     // bf16 / uint8 inputs: two-source kernel only (run_forward rejects it earlier)
  switch (p.N * 2 + (p.want_grad ? 1 : 0)) {
    case 2: return launch_tile<1, false>(p, nblk, st);
    case 3: return launch_tile<1, true>(p, nblk, st);
    case 4: return launch_tile<2, false>(p, nblk, st);
    case 5: return launch_tile<2, true>(p, nblk, st);
    case 6: return launch_tile<3, false>(p, nblk, st);
    case 7: return launch_tile<3, true>(p, nblk, st);
    case 8: return launch_tile<4, false>(p, nblk, st);
    case 9: return launch_tile<4, true>(p, nblk, st);
  }
  return cudaErrorInvalidValue;
}

// Optional timing of the dominant kernel with events on the caller's stream (bench.py roofline leg).
// The events belong to the device that was current when they were created: one pair per device, and the read-back
// reports the pair of the device the last profiled launch ran on.
// Optional per-device step counter of the in-kernel noise generator (dvs_set_noise_counter): added to `offset` by every
// forward call that is not given its own offset_dev, so captured CUDA graphs draw fresh noise on every replay.
static std::atomic<const unsigned long long*> g_noise_ctr[kMaxDevices];

static std::atomic<bool> g_profile{false};
struct ProfileEvents {
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool valid = false;
};
static ProfileEvents g_prof[kMaxDevices];
static std::atomic<int> g_prof_dev{-1};

static int run_forward(const DvsShape* sh, const DvsParams* pr, const float* const* disp, const float* target,
                       const float* const* src, const float* K, const float* inv_K, const float* const* T,
                       const float* const* noise, uint64_t seed, uint64_t offset, float* loss_per_scale,
                       float* loss_total, uint8_t* const* sel, float* const* ugrad_disp, float* uT, float* coup,
                       void* workspace, cudaStream_t st, int io_flags = 0, const PoseIO* pose_in = nullptr,
                       const unsigned long long* offset_dev = nullptr) {
  int rc = check_shape(sh);
  if (rc) return rc;
  if (io_flags && sh->N != 2) return DVS_EINVAL;      // bf16 / uint8 inputs are read by the two-source kernel only
  if (!pr || !disp || !target || !src || !K || !inv_K || (!T && !pose_in) || !loss_per_scale || !loss_total) return DVS_EINVAL;
  if (!workspace || ((uintptr_t)workspace & 255)) return DVS_EWORKSPACE;
  const bool want_grad = ugrad_disp != nullptr;
  if (want_grad && (!coup || (!uT && !(pose_in && pose_in->uP)))) return DVS_EINVAL;
  WsLayout w = ws_layout(*sh);
  char* base = static_cast<char*>(workspace);

  FusedParams p{};
  p.B = sh->B; p.H = sh->H; p.W = sh->W; p.N = sh->N; p.S = sh->S;
  for (int s = 0; s < sh->S; ++s) {
    if (!disp[s]) return DVS_EINVAL;
    p.dh[s] = sh->dh[s]; p.dw[s] = sh->dw[s];
    p.disp[s] = disp[s];
    p.noise[s] = (noise && pr->auto_mask) ? noise[s] : nullptr;
    p.sel[s] = sel ? sel[s] : nullptr;
    p.gdisp[s] = want_grad ? ugrad_disp[s] : nullptr;
    if (want_grad && !ugrad_disp[s]) return DVS_EINVAL;
  }
  PoseIO pose{};
  if (pose_in) pose = *pose_in;
  pose.Tws = reinterpret_cast<float*>(base + w.Tws);
  pose.done = reinterpret_cast<int*>(base + w.counter);
  for (int i = 0; i < sh->N; ++i) {
    if (!src[i]) return DVS_EINVAL;
    p.src[i] = src[i];
    if (pose_in) {
      if (!pose.aa[i] || !pose.tr[i]) return DVS_EINVAL;
      p.T[i] = pose.Tws + (size_t)i * sh->B * 16;
    } else {
      if (!T[i]) return DVS_EINVAL;
      p.T[i] = T[i];
    }
  }
  p.target = target; p.K = K; p.invK = inv_K;
  p.seed = seed; p.offset = offset; p.offset_dev = offset_dev;
  if (!p.offset_dev) {
    int dev = 0;
    DVS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDevices) p.offset_dev = g_noise_ctr[dev].load(std::memory_order_acquire);
  }
  p.min_disp = 1.0f / pr->max_depth;
  p.disp_range = 1.0f / pr->min_depth - 1.0f / pr->max_depth;
  p.ssim_w = pr->ssim_ratio; p.l1_w = 1.0f - pr->ssim_ratio;
  p.smooth_w = pr->smoothness_ratio; p.eps = pr->eps;
  p.auto_mask = pr->auto_mask ? 1 : 0;
  p.want_grad = want_grad ? 1 : 0;
  p.mean_part = reinterpret_cast<float*>(base + w.mean_part);
  p.part = reinterpret_cast<float*>(base + w.part);
  p.tiles_x = w.tiles_x; p.tiles_y = w.tiles_y;
  p.cpart = reinterpret_cast<float*>(base + w.cpart);
  {
    const float npix = 3.0f * (float)sh->B * (float)(sh->H * sh->W);
    p.kF = p.ssim_w / npix;
    p.l1k = p.l1_w / npix;
    for (int s = 0; s < sh->S; ++s) {
      const float kap = p.smooth_w / (float)(1 << s);
      p.kxs[s] = kap / ((float)sh->B * (float)sh->H * (float)(sh->W - 1));
      p.kys[s] = kap / ((float)sh->B * (float)(sh->H - 1) * (float)sh->W);
    }
  }
  p.nblk = w.nblk;
  p.io_flags = io_flags;
  p.cstride = w.cstride;
  for (int s = 0; s < sh->S; ++s) { p.coff[s] = w.coff[s]; p.cbw[s] = w.cbw[s]; }

  mean_partial_kernel<<<dim3(kMeanBlocks, sh->B, sh->S), 256, 0, st>>>(p, reinterpret_cast<float*>(base + w.mean_part), pose);
  DVS_CUDA_TRY(cudaGetLastError());
  ProfileEvents* pe = nullptr;
  if (g_profile.load()) {
    int dev = 0;
    DVS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDevices) {
      pe = &g_prof[dev];
      if (!pe->ev0) {
        DVS_CUDA_TRY(cudaEventCreate(&pe->ev0));
        DVS_CUDA_TRY(cudaEventCreate(&pe->ev1));
      }
      DVS_CUDA_TRY(cudaEventRecord(pe->ev0, st));
    }
  }
  DVS_CUDA_TRY(dispatch_tile(p, w.nblk, st));
  if (pe) {
    DVS_CUDA_TRY(cudaEventRecord(pe->ev1, st));
    pe->valid = true;
    int dev = 0;
    DVS_CUDA_TRY(cudaGetDevice(&dev));
    g_prof_dev.store(dev);
  }

  FinishParams f{};
  f.B = sh->B; f.H = sh->H; f.W = sh->W; f.N = sh->N; f.S = sh->S;
  f.tiles_per_img = w.tiles_x * w.tiles_y; f.want_grad = p.want_grad; f.smooth_w = p.smooth_w;
  f.part = p.part; f.mean_part = p.mean_part; f.K = K; f.invK = inv_K;
  f.perimg = reinterpret_cast<float*>(base + w.perimg);
  f.uT = uT; f.coup = coup;
  f.loss_per_scale = loss_per_scale; f.loss_total = loss_total;
  f.gather_blocks = 0;
  if (want_grad && w.cstride > 0) {
    size_t ntot = 0;
    for (int s = 0; s < sh->S; ++s)
      if (w.cbw[s]) ntot += (size_t)sh->B * sh->dh[s] * sh->dw[s];
    size_t gb = (ntot + 255) / 256;
    f.gather_blocks = (int)(gb > 148 * 16 ? 148 * 16 : gb);
  }
  postpass_kernel<<<sh->B * sh->S + f.gather_blocks, 256, 0, st>>>(p, f, pose);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

static int run_backward(const DvsShape* sh, const float* g, const float* const* u, const float* uT, const float* coup,
                        float* const* grad_disp, float* const* grad_T, cudaStream_t st, int out_bf16 = 0,
                        float* const* grad_aa = nullptr, float* const* grad_tr = nullptr) {
  BackwardParams q{};
  q.out_bf16 = out_bf16;
  q.pose = grad_aa != nullptr;
  q.B = sh->B; q.H = sh->H; q.W = sh->W; q.N = sh->N; q.S = sh->S;
  q.g = g; q.uT = uT; q.coup = coup;
  for (int s = 0; s < sh->S; ++s) {
    if (!u[s] || !grad_disp[s]) return DVS_EINVAL;
    q.dh[s] = sh->dh[s]; q.dw[s] = sh->dw[s]; q.u[s] = u[s]; q.out[s] = grad_disp[s];
  }
  for (int i = 0; i < sh->N; ++i) {
    if (q.pose) {
      if (!grad_aa[i] || !grad_tr || !grad_tr[i]) return DVS_EINVAL;
      q.gaa[i] = grad_aa[i]; q.gtr[i] = grad_tr[i];
    } else {
      if (!grad_T || !grad_T[i]) return DVS_EINVAL;
      q.gT[i] = grad_T[i];
    }
  }
  size_t n0 = (size_t)sh->B * sh->dh[0] * sh->dw[0];
  int gx = (int)((n0 + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 148 * 8) gx = 148 * 8;
  backward_scale_kernel<<<dim3(gx, sh->S + 1), 256, 0, st>>>(q);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

}  // namespace dvs

// ================================================================================================ C ABI
using namespace dvs;

extern "C" int dvs_set_noise_counter(int device, const uint64_t* counter) {
  if (device < 0 || device >= kMaxDevices) return DVS_EINVAL;
  g_noise_ctr[device].store(reinterpret_cast<const unsigned long long*>(counter), std::memory_order_release);
  return DVS_OK;
}

extern "C" int dvs_set_profiling(int enabled) {
  g_profile.store(enabled != 0);
  for (int d = 0; d < kMaxDevices; ++d) g_prof[d].valid = false;
  g_prof_dev.store(-1);
  return DVS_OK;
}

extern "C" int dvs_last_tile_kernel_ms(float* ms) {
  if (!ms) return DVS_EINVAL;
  const int dev = g_prof_dev.load();
  if (dev < 0 || !g_prof[dev].valid) return DVS_EINVAL;
  DVS_CUDA_TRY(cudaEventSynchronize(g_prof[dev].ev1));
  DVS_CUDA_TRY(cudaEventElapsedTime(ms, g_prof[dev].ev0, g_prof[dev].ev1));
  return DVS_OK;
}

extern "C" int dvs_loss_workspace_bytes(const DvsShape* shape, size_t* bytes) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!bytes) return DVS_EINVAL;
  *bytes = ws_layout(*shape).total;
  return DVS_OK;
}

// ugrad_T buffer = [S,N,B,16] unit pose gradients followed by [S,B] smoothness mean-coupling coefficients.
static size_t ut_floats(const DvsShape* sh) { return (size_t)sh->S * sh->N * sh->B * 16; }

extern "C" int dvs_photometric_forward(const DvsShape* shape, const DvsParams* params, const float* const* disp,
                                       const float* target, const float* const* src, const float* K,
                                       const float* inv_K, const float* const* T, const float* const* noise,
                                       uint64_t seed, uint64_t offset, float* loss_per_scale, float* loss_total,
                                       uint8_t* const* sel, float* const* ugrad_disp, float* ugrad_T,
                                       void* workspace, void* stream) {
  if ((ugrad_disp == nullptr) != (ugrad_T == nullptr)) return DVS_EINVAL;
  int rc = check_shape(shape);
  if (rc) return rc;
  float* coup = ugrad_T ? ugrad_T + ut_floats(shape) : nullptr;
  return run_forward(shape, params, disp, target, src, K, inv_K, T, noise, seed, offset, loss_per_scale, loss_total,
                     sel, ugrad_disp, ugrad_T, coup, workspace, static_cast<cudaStream_t>(stream));
}

static int dtype_flags(int disp_dtype, int image_dtype, int* io) {
  if ((disp_dtype != DVS_DTYPE_F32 && disp_dtype != DVS_DTYPE_BF16) || (image_dtype != DVS_DTYPE_F32 && image_dtype != DVS_DTYPE_U8))
    return DVS_EINVAL;
  *io = (disp_dtype == DVS_DTYPE_BF16 ? 1 : 0) | (image_dtype == DVS_DTYPE_U8 ? 2 : 0);
  return DVS_OK;
}

extern "C" int dvs_photometric_forward_ex(const DvsShape* shape, const DvsParams* params, const void* const* disp,
                                          int disp_dtype, const void* target, const void* const* src, int image_dtype,
                                          const float* K, const float* inv_K, const float* const* T,
                                          const float* const* noise, uint64_t seed, uint64_t offset,
                                          float* loss_per_scale, float* loss_total, uint8_t* const* sel,
                                          float* const* ugrad_disp, float* ugrad_T, void* workspace, void* stream) {
  if ((ugrad_disp == nullptr) != (ugrad_T == nullptr)) return DVS_EINVAL;
  int rc = check_shape(shape);
  if (rc) return rc;
  int io = 0;
  rc = dtype_flags(disp_dtype, image_dtype, &io);
  if (rc) return rc;
  float* coup = ugrad_T ? ugrad_T + ut_floats(shape) : nullptr;
  return run_forward(shape, params, reinterpret_cast<const float* const*>(disp), static_cast<const float*>(target),
                     reinterpret_cast<const float* const*>(src), K, inv_K, T, noise, seed, offset, loss_per_scale,
                     loss_total, sel, ugrad_disp, ugrad_T, coup, workspace, static_cast<cudaStream_t>(stream), io);
}

extern "C" int dvs_photometric_backward_ex(const DvsShape* shape, const float* grad_per_scale,
                                           const float* const* ugrad_disp, const float* ugrad_T,
                                           void* const* grad_disp, int grad_dtype, float* const* grad_T, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!grad_per_scale || !ugrad_disp || !ugrad_T || !grad_disp || !grad_T) return DVS_EINVAL;
  if (grad_dtype != DVS_DTYPE_F32 && grad_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  return run_backward(shape, grad_per_scale, ugrad_disp, ugrad_T, ugrad_T + ut_floats(shape),
                      reinterpret_cast<float* const*>(grad_disp), grad_T, static_cast<cudaStream_t>(stream),
                      grad_dtype == DVS_DTYPE_BF16 ? 1 : 0);
}

// ugrad_pose buffer = [S,N,B,6] unit gradients w.r.t. (axis-angle, translation) followed by [S,B] coupling coefficients
static size_t up_floats(const DvsShape* sh) { return (size_t)sh->S * sh->N * sh->B * 6; }

extern "C" int dvs_photometric_forward_pose(const DvsShape* shape, const DvsParams* params, const void* const* disp,
                                            int disp_dtype, const void* target, const void* const* src, int image_dtype,
                                            const float* K, const float* inv_K, const float* const* axisangle,
                                            const float* const* translation, const int32_t* invert,
                                            const float* const* noise, uint64_t seed, uint64_t offset,
                                            const uint64_t* offset_dev, float* loss_per_scale, float* loss_total,
                                            uint8_t* const* sel, float* const* ugrad_disp, float* ugrad_pose,
                                            void* workspace, void* stream) {
  if ((ugrad_disp == nullptr) != (ugrad_pose == nullptr)) return DVS_EINVAL;
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!axisangle || !translation || !invert) return DVS_EINVAL;
  int io = 0;
  rc = dtype_flags(disp_dtype, image_dtype, &io);
  if (rc) return rc;
  PoseIO pose{};
  for (int i = 0; i < shape->N; ++i) { pose.aa[i] = axisangle[i]; pose.tr[i] = translation[i]; pose.invert[i] = invert[i] ? 1 : 0; }
  pose.uP = ugrad_pose;
  float* coup = ugrad_pose ? ugrad_pose + up_floats(shape) : nullptr;
  // the 4x4 unit gradients are not kept in this mode (uT == nullptr); want_grad is signalled by ugrad_disp
  return run_forward(shape, params, reinterpret_cast<const float* const*>(disp), static_cast<const float*>(target),
                     reinterpret_cast<const float* const*>(src), K, inv_K, nullptr, noise, seed, offset, loss_per_scale,
                     loss_total, sel, ugrad_disp, nullptr, coup, workspace, static_cast<cudaStream_t>(stream), io, &pose,
                     reinterpret_cast<const unsigned long long*>(offset_dev));
}

extern "C" int dvs_photometric_backward_pose(const DvsShape* shape, const float* grad_per_scale,
                                             const float* const* ugrad_disp, const float* ugrad_pose,
                                             void* const* grad_disp, int grad_dtype, float* const* grad_axisangle,
                                             float* const* grad_translation, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!grad_per_scale || !ugrad_disp || !ugrad_pose || !grad_disp || !grad_axisangle || !grad_translation) return DVS_EINVAL;
  if (grad_dtype != DVS_DTYPE_F32 && grad_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  return run_backward(shape, grad_per_scale, ugrad_disp, ugrad_pose, ugrad_pose + up_floats(shape),
                      reinterpret_cast<float* const*>(grad_disp), nullptr, static_cast<cudaStream_t>(stream),
                      grad_dtype == DVS_DTYPE_BF16 ? 1 : 0, grad_axisangle, grad_translation);
}

extern "C" int dvs_photometric_backward(const DvsShape* shape, const float* grad_per_scale,
                                        const float* const* ugrad_disp, const float* ugrad_T,
                                        float* const* grad_disp, float* const* grad_T, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!grad_per_scale || !ugrad_disp || !ugrad_T || !grad_disp || !grad_T) return DVS_EINVAL;
  return run_backward(shape, grad_per_scale, ugrad_disp, ugrad_T, ugrad_T + ut_floats(shape), grad_disp, grad_T,
                      static_cast<cudaStream_t>(stream));
}

extern "C" int dvs_photometric_backward_recompute(const DvsShape* shape, const DvsParams* params,
                                                  const float* const* disp, const float* target,
                                                  const float* const* src, const float* K, const float* inv_K,
                                                  const float* const* T, const float* const* noise, uint64_t seed,
                                                  uint64_t offset, const float* grad_per_scale,
                                                  float* const* grad_disp, float* const* grad_T, void* workspace,
                                                  void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!grad_per_scale || !grad_disp || !grad_T) return DVS_EINVAL;
  if (!workspace || ((uintptr_t)workspace & 255)) return DVS_EWORKSPACE;
  WsLayout w = ws_layout(*shape);
  char* base = static_cast<char*>(workspace);
  float* uT = reinterpret_cast<float*>(base + w.uT);
  float* coup = reinterpret_cast<float*>(base + w.coup);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* lp = reinterpret_cast<float*>(base + w.lossbuf);   // losses are recomputed but not returned
  rc = run_forward(shape, params, disp, target, src, K, inv_K, T, noise, seed, offset, lp, lp + DVS_MAX_SCALES,
                   nullptr, grad_disp, uT, coup, workspace, st);
  if (rc) return rc;
  return run_backward(shape, grad_per_scale, grad_disp, uT, coup, grad_disp, grad_T, st);
}
