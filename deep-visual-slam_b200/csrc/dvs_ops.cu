// Granular operators of libdvsloss.so: the reference's public primitives (vo/learner_func.py ==
// model/layers.py), one sm_100a kernel per direction.  The fused loss (dvs_fused.cu) is the hot path;
// these exist so that code calling the primitives one by one (vo/predict.py, eval_traj.py, the ros2
// node, or a learner that was not switched to the fused op) runs on the same library.
//
// All kernels are HBM-bound elementwise / stencil / gather passes: coalesced along W, no atomics,
// reductions are two-stage in a fixed order (run-to-run reproducible).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dvsloss.h"
#include "dvs_fused_core.cuh"
#include "dvs_host.h"
#include "dvs_pose.cuh"

namespace dvs {
namespace {

constexpr int kThreads = 256;

inline int blocks_for(int64_t n, int per = kThreads, int cap = 148 * 16) {
  int64_t b = (n + per - 1) / per;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// deterministic block sum of NV per-thread values; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /* [NV * 8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float a = v[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[k * 8 + warp] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float a = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[k * 8 + w];
      v[k] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------ disp_to_depth
// vo/learner_func.py:16-26.  mul and add are kept separate (no FMA) as in the eager reference.
__global__ void __launch_bounds__(kThreads) d2d_fwd_kernel(const float* __restrict__ disp, float* __restrict__ scaled,
                                                            float* __restrict__ depth, int64_t n, float lo, float range) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = __fadd_rn(lo, __fmul_rn(range, disp[i]));
    if (scaled) scaled[i] = s;
    if (depth) depth[i] = __fdiv_rn(1.0f, s);
  }
}
__global__ void __launch_bounds__(kThreads) d2d_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ gs,
                                                            const float* __restrict__ gd, float* __restrict__ out,
                                                            int64_t n, float range) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = gs ? gs[i] : 0.f;
    if (gd) {
      float d = depth[i];
      g -= gd[i] * d * d;
    }
    out[i] = g * range;
  }
}

// ------------------------------------------------------------------------------------------------ bilinear up-sampling
// F.interpolate(mode="bilinear", align_corners=False); expression shaped like ATen's upsample_bilinear2d.
__global__ void __launch_bounds__(kThreads) up_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int BC,
                                                           int h, int w, int H, int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int64_t n = (int64_t)BC * H * W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(e % W);
    int64_t r = e / W;
    int y = (int)(r % H);
    int64_t pl = r / H;
    const float* p = in + pl * h * w;
    if (h == H && w == W) { out[e] = p[y * w + x]; continue; }
    int y0, y1, x0, x1;
    float ly, lx;
    up_taps(y, sy, h, y0, y1, ly);
    up_taps(x, sx, w, x0, x1, lx);
    float h1 = ly, h0 = 1.f - ly, w1 = lx, w0 = 1.f - lx;
    out[e] = h0 * (w0 * p[y0 * w + x0] + w1 * p[y0 * w + x1]) + h1 * (w0 * p[y1 * w + x0] + w1 * p[y1 * w + x1]);
  }
}
// gather-form adjoint: one thread per coarse element, no atomics
__global__ void __launch_bounds__(kThreads) up_bwd_kernel(const float* __restrict__ go, float* __restrict__ gi, int BC,
                                                           int h, int w, int H, int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float iy = (float)H / (float)h, ix = (float)W / (float)w;
  const int64_t n = (int64_t)BC * h * w;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(e % w);
    int64_t r = e / w;
    int i = (int)(r % h);
    int64_t pl = r / h;
    const float* g = go + pl * H * W;
    if (h == H && w == W) { gi[e] = g[i * W + j]; continue; }
    int ya = imax((int)(((float)i - 1.f) * iy) - 2, 0), yb = imin((int)(((float)i + 1.5f) * iy) + 2, H - 1);
    int xa = imax((int)(((float)j - 1.f) * ix) - 2, 0), xb = imin((int)(((float)j + 1.5f) * ix) + 2, W - 1);
    float acc = 0.f;
    for (int y = ya; y <= yb; ++y) {
      float wy = tap_weight(y, sy, h, i);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int x = xa; x <= xb; ++x) {
        float wx = tap_weight(x, sx, w, j);
        row = fmaf(wx, g[y * W + x], row);
      }
      acc = fmaf(wy, row, acc);
    }
    gi[e] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ BackprojectDepth
// vo/learner_func.py:130-135: cam = depth * (inv_K[:3,:3] (u,v,1)^T), homogeneous 1 appended.
__global__ void __launch_bounds__(kThreads) backproject_fwd_kernel(const float* __restrict__ depth,
                                                                    const float* __restrict__ invK,
                                                                    float* __restrict__ cam, int H, int W) {
  const int b = blockIdx.y, HW = H * W;
  const float* k = invK + b * 16;
  const float k00 = k[0], k01 = k[1], k02 = k[2], k10 = k[4], k11 = k[5], k12 = k[6], k20 = k[8], k21 = k[9], k22 = k[10];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float u = (float)(p % W), v = (float)(p / W);
    float d = depth[(size_t)b * HW + p];
    float* o = cam + (size_t)b * 4 * HW + p;
    o[0] = d * (k00 * u + k01 * v + k02);
    o[HW] = d * (k10 * u + k11 * v + k12);
    o[2 * HW] = d * (k20 * u + k21 * v + k22);
    o[3 * HW] = 1.0f;
  }
}
__global__ void __launch_bounds__(kThreads) backproject_bwd_kernel(const float* __restrict__ gcam,
                                                                    const float* __restrict__ invK,
                                                                    float* __restrict__ gdepth, int H, int W) {
  const int b = blockIdx.y, HW = H * W;
  const float* k = invK + b * 16;
  const float k00 = k[0], k01 = k[1], k02 = k[2], k10 = k[4], k11 = k[5], k12 = k[6], k20 = k[8], k21 = k[9], k22 = k[10];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float u = (float)(p % W), v = (float)(p / W);
    const float* g = gcam + (size_t)b * 4 * HW + p;
    gdepth[(size_t)b * HW + p] =
        g[0] * (k00 * u + k01 * v + k02) + g[HW] * (k10 * u + k11 * v + k12) + g[2 * HW] * (k20 * u + k21 * v + k22);
  }
}

// ------------------------------------------------------------------------------------------------ Project3D
// vo/learner_func.py:148-159.
__device__ __forceinline__ void load_P(const float* K, const float* T, int b, float* P /* smem[12] */) {
  if (threadIdx.x < 12) {
    int j = threadIdx.x / 4, c = threadIdx.x % 4;
    const float* Kb = K + b * 16;
    const float* Tb = T + b * 16;
    float a = 0.f;
    for (int m = 0; m < 4; ++m) a = fmaf(Kb[j * 4 + m], Tb[m * 4 + c], a);
    P[threadIdx.x] = a;
  }
  __syncthreads();
}
__global__ void __launch_bounds__(kThreads) project_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                                const float* __restrict__ T, float* __restrict__ pix,
                                                                int H, int W, float eps) {
  __shared__ float P[12];
  const int b = blockIdx.y, HW = H * W;
  load_P(K, T, b, P);
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const float* x = pts + (size_t)b * 4 * HW + p;
    float X0 = x[0], X1 = x[HW], X2 = x[2 * HW], X3 = x[3 * HW];
    float c0 = P[0] * X0 + P[1] * X1 + P[2] * X2 + P[3] * X3;
    float c1 = P[4] * X0 + P[5] * X1 + P[6] * X2 + P[7] * X3;
    float c2 = P[8] * X0 + P[9] * X1 + P[10] * X2 + P[11] * X3;
    float z = c2 + eps;
    float px = __fdiv_rn(c0, z), py = __fdiv_rn(c1, z);
    float2 o;
    o.x = (__fdiv_rn(px, wm) - 0.5f) * 2.f;
    o.y = (__fdiv_rn(py, hm) - 0.5f) * 2.f;
    reinterpret_cast<float2*>(pix)[(size_t)b * HW + p] = o;
  }
}
// grad_points (optional) + per-block partial sums of dL/dP (optional)
__global__ void __launch_bounds__(kThreads) project_bwd_kernel(const float* __restrict__ gpix, const float* __restrict__ pts,
                                                                const float* __restrict__ K, const float* __restrict__ T,
                                                                float* __restrict__ gpts, float* __restrict__ part,
                                                                int H, int W, float eps) {
  __shared__ float P[12];
  __shared__ float red[12 * 8];
  const int b = blockIdx.y, HW = H * W;
  load_P(K, T, b, P);
  const float kx = 2.f / (float)(W - 1), ky = 2.f / (float)(H - 1);
  float dP[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) dP[k] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const float* x = pts + (size_t)b * 4 * HW + p;
    float X[4] = {x[0], x[HW], x[2 * HW], x[3 * HW]};
    float c0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * X[3];
    float c1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * X[3];
    float c2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * X[3];
    float rz = __fdiv_rn(1.0f, c2 + eps);
    float2 g = reinterpret_cast<const float2*>(gpix)[(size_t)b * HW + p];
    float gpx = g.x * kx, gpy = g.y * ky;
    float gc[3];
    gc[0] = gpx * rz;
    gc[1] = gpy * rz;
    gc[2] = -(gpx * c0 + gpy * c1) * rz * rz;
    if (gpts) {
      float* o = gpts + (size_t)b * 4 * HW + p;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k * HW] = gc[0] * P[k] + gc[1] * P[4 + k] + gc[2] * P[8 + k];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) dP[j * 4 + k] = fmaf(gc[j], X[k], dP[j * 4 + k]);
  }
  if (part) {
    block_sum<12>(dP, red);
    if (threadIdx.x == 0)
      for (int k = 0; k < 12; ++k) part[((size_t)b * gridDim.x + blockIdx.x) * 12 + k] = dP[k];
  }
}
// grad_T[b][m][k] = sum_j K[b][j][m] * (sum over blocks of dP[j][k]); one warp per batch item
__global__ void project_bwd_finish_kernel(const float* __restrict__ part, const float* __restrict__ K,
                                          float* __restrict__ gT, int nblk) {
  __shared__ float dP[12];
  const int b = blockIdx.x;
  if (threadIdx.x < 12) {
    float a = 0.f;
    for (int k = 0; k < nblk; ++k) a += part[((size_t)b * nblk + k) * 12 + threadIdx.x];
    dP[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    int m = threadIdx.x / 4, k = threadIdx.x % 4;
    const float* Kb = K + b * 16;
    float a = 0.f;
    for (int j = 0; j < 3; ++j) a = fmaf(Kb[j * 4 + m], dP[j * 4 + k], a);
    gT[b * 16 + threadIdx.x] = a;
  }
}

// ------------------------------------------------------------------------------------------------ grid_sample (border, align_corners)
struct Taps {
  int x0, y0, x1, y1;
  bool okx, oky;
  float wnw, wne, wsw, wse;     // bilinear weights
  float tx, ty;
  float mx, my;                 // d ix / d gx (0 when clipped), d iy / d gy
};
__device__ __forceinline__ Taps grid_taps(float gx, float gy, int H, int W) {
  Taps t;
  float wm = (float)(W - 1), hm = (float)(H - 1);
  float ix = ((gx + 1.f) / 2.f) * wm, iy = ((gy + 1.f) / 2.f) * hm;
  t.mx = (ix > 0.f && ix < wm) ? wm * 0.5f : 0.f;      // GridSampler.h clip_coordinates_set_grad
  t.my = (iy > 0.f && iy < hm) ? hm * 0.5f : 0.f;
  ix = fminf(fmaxf(ix, 0.f), wm);
  iy = fminf(fmaxf(iy, 0.f), hm);
  float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx; t.y0 = (int)fy;
  t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
  t.okx = t.x1 <= W - 1; t.oky = t.y1 <= H - 1;
  t.tx = ix - fx; t.ty = iy - fy;
  float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy;
  t.wnw = ex * ey; t.wne = t.tx * ey; t.wsw = ex * t.ty; t.wse = t.tx * t.ty;
  return t;
}
__global__ void __launch_bounds__(kThreads) gs_fwd_kernel(const float* __restrict__ src, const float* __restrict__ grid,
                                                           float* __restrict__ out, int C, int H, int W, int Ho, int Wo) {
  const int b = blockIdx.y, HoWo = Ho * Wo, HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HoWo; p += gridDim.x * blockDim.x) {
    float2 g = reinterpret_cast<const float2*>(grid)[(size_t)b * HoWo + p];
    Taps t = grid_taps(g.x, g.y, H, W);
    const int onw = t.y0 * W + t.x0;
    for (int c = 0; c < C; ++c) {
      const float* im = src + ((size_t)b * C + c) * HW;
      float acc = im[onw] * t.wnw;
      if (t.okx) acc += im[onw + 1] * t.wne;
      if (t.oky) acc += im[onw + W] * t.wsw;
      if (t.okx && t.oky) acc += im[onw + W + 1] * t.wse;
      out[((size_t)b * C + c) * HoWo + p] = acc;
    }
  }
}
__global__ void __launch_bounds__(kThreads) gs_bwd_kernel(const float* __restrict__ go, const float* __restrict__ src,
                                                           const float* __restrict__ grid, float* __restrict__ ggrid,
                                                           int C, int H, int W, int Ho, int Wo) {
  const int b = blockIdx.y, HoWo = Ho * Wo, HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HoWo; p += gridDim.x * blockDim.x) {
    float2 g = reinterpret_cast<const float2*>(grid)[(size_t)b * HoWo + p];
    Taps t = grid_taps(g.x, g.y, H, W);
    const int onw = t.y0 * W + t.x0;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* im = src + ((size_t)b * C + c) * HW;
      float nw = im[onw];
      float ne = t.okx ? im[onw + 1] : 0.f;
      float sw = t.oky ? im[onw + W] : 0.f;
      float se = (t.okx && t.oky) ? im[onw + W + 1] : 0.f;
      float o = go[((size_t)b * C + c) * HoWo + p];
      gix += ((ne - nw) * (1.f - t.ty) + (se - sw) * t.ty) * o;
      giy += ((sw - nw) * (1.f - t.tx) + (se - ne) * t.tx) * o;
    }
    float2 r;
    r.x = gix * t.mx;
    r.y = giy * t.my;
    reinterpret_cast<float2*>(ggrid)[(size_t)b * HoWo + p] = r;
  }
}

// ------------------------------------------------------------------------------------------------ SSIM / reprojection loss
// vo/learner_func.py:190-207 and vo/learner_new.py:60-74.
// 3x3 sums around (y,x) of one plane pair with 1-px reflection, straight from global memory (L1-resident rows).
// The granular op reproduces the eager op order of the reference on CUDA exactly: x*x, y*y, x*y are rounded
// products (separate tensors in vo/learner_func.py:199-201), ATen's avg_pool2d adds the window row-major into one
// fp32 accumulator and divides by 9, and no multiply-add is contracted (hence the __f*_rn intrinsics).
__device__ __forceinline__ void sums3x3(const float* __restrict__ X, const float* __restrict__ Y, int y, int x, int H,
                                        int W, float& sx, float& sy, float& sxx, float& syy, float& sxy) {
  sx = sy = sxx = syy = sxy = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    int yy = reflect_clamp(y + dy, H);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      int xx = reflect_clamp(x + dx, W);
      float a = X[yy * W + xx], b = Y[yy * W + xx];
      sx = __fadd_rn(sx, a); sy = __fadd_rn(sy, b);
      sxx = __fadd_rn(sxx, __fmul_rn(a, a)); syy = __fadd_rn(syy, __fmul_rn(b, b)); sxy = __fadd_rn(sxy, __fmul_rn(a, b));
    }
  }
}
// SSIM loss value in the eager op order (vo/learner_func.py:194-207), IEEE division, every op rounded on its own
__device__ __forceinline__ float ssim_value(float sx, float sy, float sxx, float syy, float sxy) {
  float mx = __fdiv_rn(sx, 9.f), my = __fdiv_rn(sy, 9.f);
  float mx2 = __fmul_rn(mx, mx), my2 = __fmul_rn(my, my), mxy = __fmul_rn(mx, my);
  float sgx = __fsub_rn(__fdiv_rn(sxx, 9.f), mx2), sgy = __fsub_rn(__fdiv_rn(syy, 9.f), my2);
  float sgxy = __fsub_rn(__fdiv_rn(sxy, 9.f), mxy);
  float n = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, mx), my), kC1), __fadd_rn(__fmul_rn(2.f, sgxy), kC2));
  float d = __fmul_rn(__fadd_rn(__fadd_rn(mx2, my2), kC1), __fadd_rn(__fadd_rn(sgx, sgy), kC2));
  return fminf(fmaxf(__fdiv_rn(__fsub_rn(1.f, __fdiv_rn(n, d)), 2.f), 0.f), 1.f);
}
// REPROJ=false: out[b,c,y,x] = SSIM loss map.  REPROJ=true: out[b,0,y,x] = w*mean_c SSIM + (1-w)*mean_c |y-x|.
template <bool REPROJ>
__global__ void __launch_bounds__(kThreads) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             float* __restrict__ out, int C, int H, int W, float w) {
  const int b = blockIdx.y, HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int py = p / W, px = p % W;
    float accS = 0.f, accL = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* X = x + ((size_t)b * C + c) * HW;
      const float* Y = y + ((size_t)b * C + c) * HW;
      float sx, sy, sxx, syy, sxy;
      sums3x3(X, Y, py, px, H, W, sx, sy, sxx, syy, sxy);
      float S = ssim_value(sx, sy, sxx, syy, sxy);
      if (REPROJ) {
        accS = __fadd_rn(accS, S);
        accL = __fadd_rn(accL, fabsf(__fsub_rn(Y[p], X[p])));
      } else {
        out[((size_t)b * C + c) * HW + p] = S;
      }
    }
    if (REPROJ)
      out[(size_t)b * HW + p] = __fadd_rn(__fmul_rn(w, __fdiv_rn(accS, (float)C)), __fmul_rn(1.f - w, __fdiv_rn(accL, (float)C)));
  }
}
// Backward w.r.t. the FIRST image argument (SSIM is symmetric: call with the arguments swapped for the other).
// Tile 32x8 outputs; coefficient fields on the 34x10 ring of window centres, images on 36x12.
constexpr int BT_W = 32, BT_H = 8;
template <bool REPROJ>
__global__ void __launch_bounds__(BT_W* BT_H) ssim_bwd_kernel(const float* __restrict__ go, const float* __restrict__ x,
                                                               const float* __restrict__ y, float* __restrict__ gx,
                                                               int C, int H, int W, float w) {
  __shared__ float sx_[BT_H + 4][BT_W + 4], sy_[BT_H + 4][BT_W + 4];
  __shared__ float fa[BT_H + 2][BT_W + 2], fb[BT_H + 2][BT_W + 2], fc[BT_H + 2][BT_W + 2];
  const int bc = blockIdx.z, b = bc / C;
  const int HW = H * W;
  const float* X = x + (size_t)bc * HW;
  const float* Y = y + (size_t)bc * HW;
  const float* G = REPROJ ? go + (size_t)b * HW : go + (size_t)bc * HW;
  const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H;
  const int tid = threadIdx.y * BT_W + threadIdx.x;
  for (int k = tid; k < (BT_H + 4) * (BT_W + 4); k += BT_W * BT_H) {
    int ly = k / (BT_W + 4), lx = k % (BT_W + 4);
    int gy = reflect_clamp(y0 + ly - 2, H), gxx = reflect_clamp(x0 + lx - 2, W);
    sx_[ly][lx] = X[gy * W + gxx];
    sy_[ly][lx] = Y[gy * W + gxx];
  }
  __syncthreads();
  const float scale = REPROJ ? w / (float)C : 1.f;
  for (int k = tid; k < (BT_H + 2) * (BT_W + 2); k += BT_W * BT_H) {
    int ly = k / (BT_W + 2), lx = k % (BT_W + 2);
    int gy = y0 + ly - 1, gxx = x0 + lx - 1;
    float a = 0.f, bq = 0.f, c = 0.f;
    if (gy >= 0 && gy < H && gxx >= 0 && gxx < W) {
      float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          // the staged tile holds reflect_clamp'ed *tile* coordinates; the window of an in-image centre needs
          // reflection about the image border, which for a 1-px overhang equals the staged value at distance 2
          int wy = gy + dy - 1, wx = gxx + dx - 1;
          int ry = reflect_clamp(wy, H) - y0 + 2, rx = reflect_clamp(wx, W) - x0 + 2;
          float u = sx_[ry][rx], v = sy_[ry][rx];
          s1 += u; s2 += v;
          s11 = fmaf(u, u, s11); s22 = fmaf(v, v, s22); s12 = fmaf(u, v, s12);
        }
      float mx = s1 / 9.f, my = s2 / 9.f;
      float sgx = s11 / 9.f - mx * mx, sgy = s22 / 9.f - my * my, sgxy = s12 / 9.f - mx * my;
      float n1 = 2.f * mx * my + kC1, n2 = 2.f * sgxy + kC2;
      float d1 = mx * mx + my * my + kC1, d2 = sgx + sgy + kC2;
      float n = n1 * n2, d = d1 * d2;
      float raw = (1.f - n / d) / 2.f;
      if (raw >= 0.f && raw <= 1.f) {
        float g = G[gy * W + gxx] * scale * (1.f / 9.f);
        float rd = 1.f / d;
        a = -(my * (n2 - n1) - n * rd * mx * (d2 - d1)) * rd * g;
        bq = 0.5f * n * rd / d2 * g;
        c = -n1 * rd * g;
      }
    }
    fa[ly][lx] = a; fb[ly][lx] = bq; fc[ly][lx] = c;
  }
  __syncthreads();
  const int qy = y0 + threadIdx.y, qx = x0 + threadIdx.x;
  if (qy >= H || qx >= W) return;
  float A = 0.f, Bq = 0.f, Cq = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    float wr = 1.f;
    if (dy == 0 && qy == 1) wr = 2.f;
    if (dy == 2 && qy == H - 2) wr = 2.f;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      float wc = wr;
      if (dx == 0 && qx == 1) wc *= 2.f;
      if (dx == 2 && qx == W - 2) wc *= 2.f;
      A = fmaf(wc, fa[threadIdx.y + dy][threadIdx.x + dx], A);
      Bq = fmaf(wc, fb[threadIdx.y + dy][threadIdx.x + dx], Bq);
      Cq = fmaf(wc, fc[threadIdx.y + dy][threadIdx.x + dx], Cq);
    }
  }
  float xv = sx_[threadIdx.y + 2][threadIdx.x + 2], yv = sy_[threadIdx.y + 2][threadIdx.x + 2];
  float g = A + 2.f * xv * Bq + yv * Cq;
  if (REPROJ) g -= (1.f - w) / (float)C * sgn(yv - xv) * G[qy * W + qx];
  gx[(size_t)bc * HW + qy * W + qx] = g;
}

// ------------------------------------------------------------------------------------------------ get_smooth_loss
// vo/learner_func.py:161-174.
__device__ __forceinline__ float edge_w(const float* __restrict__ img, int C, int HW, int p, int q) {
  float a = 0.f;
  for (int c = 0; c < C; ++c) a += fabsf(img[c * HW + p] - img[c * HW + q]);
  return expf(-a / (float)C);
}
__global__ void __launch_bounds__(kThreads) smooth_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                               float* __restrict__ part, int B, int C, int H, int W) {
  __shared__ float red[2 * 8];
  const int HW = H * W;
  const int64_t n = (int64_t)B * HW;
  float acc[2] = {0.f, 0.f};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(e / HW), p = (int)(e % HW), y = p / W, x = p % W;
    const float* d = disp + (size_t)b * HW;
    const float* im = img + (size_t)b * C * HW;
    if (x < W - 1) acc[0] += fabsf(d[p] - d[p + 1]) * edge_w(im, C, HW, p, p + 1);
    if (y < H - 1) acc[1] += fabsf(d[p] - d[p + W]) * edge_w(im, C, HW, p, p + W);
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) { part[blockIdx.x * 2] = acc[0]; part[blockIdx.x * 2 + 1] = acc[1]; }
}
__global__ void smooth_finish_kernel(const float* __restrict__ part, int nblk, float inv_nx, float inv_ny,
                                     float* __restrict__ out) {
  __shared__ float red[2 * 8];
  float acc[2] = {0.f, 0.f};
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) { acc[0] += part[2 * k]; acc[1] += part[2 * k + 1]; }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) out[0] = acc[0] * inv_nx + acc[1] * inv_ny;
}
__global__ void __launch_bounds__(kThreads) smooth_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ disp,
                                                               const float* __restrict__ img, float* __restrict__ gdisp,
                                                               int B, int C, int H, int W, float inv_nx, float inv_ny) {
  const int HW = H * W;
  const int64_t n = (int64_t)B * HW;
  const float g0 = gout[0];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(e / HW), p = (int)(e % HW), y = p / W, x = p % W;
    const float* d = disp + (size_t)b * HW;
    const float* im = img + (size_t)b * C * HW;
    float gxs = 0.f, gys = 0.f, d0 = d[p];
    if (x < W - 1) gxs += sgn(d0 - d[p + 1]) * edge_w(im, C, HW, p, p + 1);
    if (x > 0) gxs -= sgn(d[p - 1] - d0) * edge_w(im, C, HW, p - 1, p);
    if (y < H - 1) gys += sgn(d0 - d[p + W]) * edge_w(im, C, HW, p, p + W);
    if (y > 0) gys -= sgn(d[p - W] - d0) * edge_w(im, C, HW, p - W, p);
    gdisp[e] = g0 * (gxs * inv_nx + gys * inv_ny);
  }
}

// ------------------------------------------------------------------------------------------------ pose matrix
// vo/learner_func.py:29-104 (device functions in dvs_pose.cuh).
__global__ void pose_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, float* __restrict__ M, int B,
                                int invert) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  pose_matrix(aa + 3 * b, tr + 3 * b, invert, M + 16 * b);
}
__global__ void pose_bwd_kernel(const float* __restrict__ gM, const float* __restrict__ aa, const float* __restrict__ tr,
                                float* __restrict__ gaa, float* __restrict__ gtr, int B, int invert) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  pose_matrix_grad(gM + 16 * b, aa + 3 * b, tr + 3 * b, invert, gaa + 3 * b, gtr + 3 * b);
}

constexpr int kProjBlocks = 64;      // partial-sum blocks per image in project3d backward
constexpr int kSmoothBlocks = 592;   // 4 x 148 partial-sum blocks in the smoothness forward

inline bool bad_img(int B, int C, int H, int W) { return B < 1 || C < 1 || H < 2 || W < 2 || (int64_t)C * H * W >= (1LL << 31); }

}  // namespace
}  // namespace dvs

using namespace dvs;
#define ST(s) static_cast<cudaStream_t>(s)
#define LAUNCH_CHECK() DVS_CUDA_TRY(cudaGetLastError())

extern "C" int dvs_disp_to_depth_fwd(const float* disp, float* scaled_disp, float* depth, int64_t n, float min_depth,
                                     float max_depth, void* stream) {
  if (!disp || n < 0 || (!scaled_disp && !depth) || !(min_depth > 0.f) || !(max_depth > 0.f)) return DVS_EINVAL;
  if (n == 0) return DVS_OK;
  double lo = 1.0 / (double)max_depth, hi = 1.0 / (double)min_depth;
  d2d_fwd_kernel<<<blocks_for(n), kThreads, 0, ST(stream)>>>(disp, scaled_disp, depth, n, (float)lo, (float)(hi - lo));
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_disp_to_depth_bwd(const float* depth, const float* grad_scaled, const float* grad_depth,
                                     float* grad_disp, int64_t n, float min_depth, float max_depth, void* stream) {
  if (!grad_disp || n < 0 || (grad_depth && !depth) || !(min_depth > 0.f) || !(max_depth > 0.f)) return DVS_EINVAL;
  if (n == 0) return DVS_OK;
  double lo = 1.0 / (double)max_depth, hi = 1.0 / (double)min_depth;
  d2d_bwd_kernel<<<blocks_for(n), kThreads, 0, ST(stream)>>>(depth, grad_scaled, grad_depth, grad_disp, n, (float)(hi - lo));
  LAUNCH_CHECK();
  return DVS_OK;
}

extern "C" int dvs_upsample_bilinear_fwd(const float* in, float* out, int B, int C, int h, int w, int H, int W,
                                         void* stream) {
  if (!in || !out || B < 1 || C < 1 || h < 1 || w < 1 || H < 1 || W < 1) return DVS_EINVAL;
  up_fwd_kernel<<<blocks_for((int64_t)B * C * H * W), kThreads, 0, ST(stream)>>>(in, out, B * C, h, w, H, W);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_upsample_bilinear_bwd(const float* grad_out, float* grad_in, int B, int C, int h, int w, int H, int W,
                                         void* stream) {
  if (!grad_out || !grad_in || B < 1 || C < 1 || h < 1 || w < 1 || H < 1 || W < 1) return DVS_EINVAL;
  up_bwd_kernel<<<blocks_for((int64_t)B * C * h * w), kThreads, 0, ST(stream)>>>(grad_out, grad_in, B * C, h, w, H, W);
  LAUNCH_CHECK();
  return DVS_OK;
}

extern "C" int dvs_backproject_fwd(const float* depth, const float* inv_K, float* cam_points, int B, int H, int W,
                                   void* stream) {
  if (!depth || !inv_K || !cam_points || bad_img(B, 4, H, W)) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)H * W, kThreads, 148 * 4), B);
  backproject_fwd_kernel<<<g, kThreads, 0, ST(stream)>>>(depth, inv_K, cam_points, H, W);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_backproject_bwd(const float* grad_cam, const float* inv_K, float* grad_depth, int B, int H, int W,
                                   void* stream) {
  if (!grad_cam || !inv_K || !grad_depth || bad_img(B, 4, H, W)) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)H * W, kThreads, 148 * 4), B);
  backproject_bwd_kernel<<<g, kThreads, 0, ST(stream)>>>(grad_cam, inv_K, grad_depth, H, W);
  LAUNCH_CHECK();
  return DVS_OK;
}

extern "C" int dvs_project3d_fwd(const float* points, const float* K, const float* T, float* pix, int B, int H, int W,
                                 float eps, void* stream) {
  if (!points || !K || !T || !pix || bad_img(B, 4, H, W)) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)H * W, kThreads, 148 * 4), B);
  project_fwd_kernel<<<g, kThreads, 0, ST(stream)>>>(points, K, T, pix, H, W, eps);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_project3d_bwd_workspace_bytes(int B, int H, int W, size_t* bytes) {
  if (!bytes || bad_img(B, 4, H, W)) return DVS_EINVAL;
  *bytes = sizeof(float) * (size_t)B * kProjBlocks * 12;
  return DVS_OK;
}
extern "C" int dvs_project3d_bwd(const float* grad_pix, const float* points, const float* K, const float* T,
                                 float* grad_points, float* grad_T, int B, int H, int W, float eps, void* workspace,
                                 void* stream) {
  if (!grad_pix || !points || !K || !T || (!grad_points && !grad_T) || bad_img(B, 4, H, W)) return DVS_EINVAL;
  if (grad_T && !workspace) return DVS_EWORKSPACE;
  int nblk = blocks_for((int64_t)H * W, kThreads, kProjBlocks);
  float* part = grad_T ? static_cast<float*>(workspace) : nullptr;
  project_bwd_kernel<<<dim3(nblk, B), kThreads, 0, ST(stream)>>>(grad_pix, points, K, T, grad_points, part, H, W, eps);
  LAUNCH_CHECK();
  if (grad_T) {
    project_bwd_finish_kernel<<<B, 32, 0, ST(stream)>>>(part, K, grad_T, nblk);
    LAUNCH_CHECK();
  }
  return DVS_OK;
}

extern "C" int dvs_grid_sample_border_fwd(const float* src, const float* grid, float* out, int B, int C, int H, int W,
                                          int Ho, int Wo, void* stream) {
  if (!src || !grid || !out || bad_img(B, C, H, W) || Ho < 1 || Wo < 1) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)Ho * Wo, kThreads, 148 * 4), B);
  gs_fwd_kernel<<<g, kThreads, 0, ST(stream)>>>(src, grid, out, C, H, W, Ho, Wo);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_grid_sample_border_bwd(const float* grad_out, const float* src, const float* grid, float* grad_grid,
                                          int B, int C, int H, int W, int Ho, int Wo, void* stream) {
  if (!grad_out || !src || !grid || !grad_grid || bad_img(B, C, H, W) || Ho < 1 || Wo < 1) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)Ho * Wo, kThreads, 148 * 4), B);
  gs_bwd_kernel<<<g, kThreads, 0, ST(stream)>>>(grad_out, src, grid, grad_grid, C, H, W, Ho, Wo);
  LAUNCH_CHECK();
  return DVS_OK;
}

extern "C" int dvs_ssim_fwd(const float* x, const float* y, float* out, int B, int C, int H, int W, void* stream) {
  if (!x || !y || !out || bad_img(B, C, H, W)) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)H * W, kThreads, 148 * 4), B);
  ssim_fwd_kernel<false><<<g, kThreads, 0, ST(stream)>>>(x, y, out, C, H, W, 0.f);
  LAUNCH_CHECK();
  return DVS_OK;
}
static int ssim_bwd_launch(bool reproj, const float* go, const float* a, const float* b, float* ga, int B, int C, int H,
                           int W, float w, cudaStream_t st) {
  dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H, B * C), blk(BT_W, BT_H);
  if (grid.y > 65535 || grid.z > 65535) return DVS_EINVAL;
  if (reproj) ssim_bwd_kernel<true><<<grid, blk, 0, st>>>(go, a, b, ga, C, H, W, w);
  else ssim_bwd_kernel<false><<<grid, blk, 0, st>>>(go, a, b, ga, C, H, W, w);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_ssim_bwd(const float* grad_out, const float* x, const float* y, float* grad_x, float* grad_y, int B,
                            int C, int H, int W, void* stream) {
  if (!grad_out || !x || !y || (!grad_x && !grad_y) || bad_img(B, C, H, W)) return DVS_EINVAL;
  int rc = DVS_OK;
  if (grad_x) rc = ssim_bwd_launch(false, grad_out, x, y, grad_x, B, C, H, W, 0.f, ST(stream));
  if (rc == DVS_OK && grad_y) rc = ssim_bwd_launch(false, grad_out, y, x, grad_y, B, C, H, W, 0.f, ST(stream));
  return rc;
}

extern "C" int dvs_reprojection_loss_fwd(const float* pred, const float* target, float* out, int B, int C, int H, int W,
                                         float ssim_ratio, void* stream) {
  if (!pred || !target || !out || bad_img(B, C, H, W)) return DVS_EINVAL;
  dim3 g(blocks_for((int64_t)H * W, kThreads, 148 * 4), B);
  ssim_fwd_kernel<true><<<g, kThreads, 0, ST(stream)>>>(pred, target, out, C, H, W, ssim_ratio);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_reprojection_loss_bwd(const float* grad_out, const float* pred, const float* target, float* grad_pred,
                                         int B, int C, int H, int W, float ssim_ratio, void* stream) {
  if (!grad_out || !pred || !target || !grad_pred || bad_img(B, C, H, W)) return DVS_EINVAL;
  return ssim_bwd_launch(true, grad_out, pred, target, grad_pred, B, C, H, W, ssim_ratio, ST(stream));
}

extern "C" int dvs_smooth_loss_workspace_bytes(int B, int H, int W, size_t* bytes) {
  if (!bytes || bad_img(B, 1, H, W)) return DVS_EINVAL;
  *bytes = sizeof(float) * 2 * kSmoothBlocks;
  return DVS_OK;
}
extern "C" int dvs_smooth_loss_fwd(const float* disp, const float* img, float* out, int B, int C, int H, int W,
                                   void* workspace, void* stream) {
  if (!disp || !img || !out || bad_img(B, C, H, W)) return DVS_EINVAL;
  if (!workspace) return DVS_EWORKSPACE;
  int nblk = blocks_for((int64_t)B * H * W, kThreads, kSmoothBlocks);
  float* part = static_cast<float*>(workspace);
  smooth_fwd_kernel<<<nblk, kThreads, 0, ST(stream)>>>(disp, img, part, B, C, H, W);
  LAUNCH_CHECK();
  float inx = 1.0f / ((float)B * (float)H * (float)(W - 1)), iny = 1.0f / ((float)B * (float)(H - 1) * (float)W);
  smooth_finish_kernel<<<1, kThreads, 0, ST(stream)>>>(part, nblk, inx, iny, out);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_smooth_loss_bwd(const float* grad_out, const float* disp, const float* img, float* grad_disp, int B,
                                   int C, int H, int W, void* stream) {
  if (!grad_out || !disp || !img || !grad_disp || bad_img(B, C, H, W)) return DVS_EINVAL;
  float inx = 1.0f / ((float)B * (float)H * (float)(W - 1)), iny = 1.0f / ((float)B * (float)(H - 1) * (float)W);
  smooth_bwd_kernel<<<blocks_for((int64_t)B * H * W), kThreads, 0, ST(stream)>>>(grad_out, disp, img, grad_disp, B, C, H, W,
                                                                                 inx, iny);
  LAUNCH_CHECK();
  return DVS_OK;
}

extern "C" int dvs_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert,
                                   void* stream) {
  if (!axisangle || !translation || !M || B < 1) return DVS_EINVAL;
  pose_fwd_kernel<<<(B + 63) / 64, 64, 0, ST(stream)>>>(axisangle, translation, M, B, invert);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_pose_matrix_bwd(const float* grad_M, const float* axisangle, const float* translation,
                                   float* grad_axisangle, float* grad_translation, int B, int invert, void* stream) {
  if (!grad_M || !axisangle || !translation || !grad_axisangle || !grad_translation || B < 1) return DVS_EINVAL;
  pose_bwd_kernel<<<(B + 63) / 64, 64, 0, ST(stream)>>>(grad_M, axisangle, translation, grad_axisangle, grad_translation,
                                                        B, invert);
  LAUNCH_CHECK();
  return DVS_OK;
}

// ------------------------------------------------------------------------------------------------ image format
// ToTensor of the reference's loader (vo/dataset/common.py:77): uint8 [0,255] -> float32 / 255 (IEEE division, so the
// result is bit-identical to the host-side conversion).  Lets image batches cross PCIe as bytes.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += stride) {
    if (e + 4 <= n && (((uintptr_t)(src + e)) & 3) == 0 && (((uintptr_t)(dst + e)) & 15) == 0) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(src + e);
      *reinterpret_cast<float4*>(dst + e) = make_float4(__fdiv_rn((float)v.x, 255.f), __fdiv_rn((float)v.y, 255.f),
                                                        __fdiv_rn((float)v.z, 255.f), __fdiv_rn((float)v.w, 255.f));
    } else {
      for (int64_t k = e; k < n && k < e + 4; ++k) dst[k] = __fdiv_rn((float)src[k], 255.f);
    }
  }
}
extern "C" int dvs_u8_to_f32(const uint8_t* src, float* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return DVS_EINVAL;
  if (n == 0) return DVS_OK;
  int64_t want = (n / 4 + 255) / 256;
  int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  u8_to_f32_kernel<<<blocks, 256, 0, ST(stream)>>>(src, dst, n);
  LAUNCH_CHECK();
  return DVS_OK;
}


// ------------------------------------------------------------------------------------------------ triplet batcher
// GPU side of MonoDataset.__getitem__ + collation (vo/dataset/common.py:48-92) for frames that are resident on the device
// as decoded, resized uint8 RGB (what _read_image returns): for every sample b and role r (source_left, target_image,
// source_right) copy frame idx[b][r], turn HWC into CHW and either keep the bytes or apply ToTensor (x / 255, exact).
// One thread per output pixel (all three channels), consecutive threads -> consecutive pixels of a row: the three planes
// are written as coalesced rows, the interleaved source bytes are read as 3 consecutive bytes per thread.
template <bool OUT_F32, bool HWC>
__global__ void __launch_bounds__(256) gather_triplets_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ idx,
                                                              void* o0, void* o1, void* o2, int B, int HW) {
  const int b = blockIdx.y, r = blockIdx.z;
  const uint8_t* src = frames + (size_t)idx[b * 3 + r] * 3 * HW;
  void* out = r == 0 ? o0 : (r == 1 ? o1 : o2);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < HW; e += gridDim.x * blockDim.x) {
    uint8_t v[3];
    if (HWC) { v[0] = src[3 * e]; v[1] = src[3 * e + 1]; v[2] = src[3 * e + 2]; }
    else { v[0] = src[e]; v[1] = src[HW + e]; v[2] = src[2 * HW + e]; }
    const size_t o = (size_t)b * 3 * HW + e;
    if (OUT_F32) {
      float* f = static_cast<float*>(out);
      f[o] = __fdiv_rn((float)v[0], 255.f); f[o + HW] = __fdiv_rn((float)v[1], 255.f); f[o + 2 * HW] = __fdiv_rn((float)v[2], 255.f);
    } else {
      uint8_t* u = static_cast<uint8_t*>(out);
      u[o] = v[0]; u[o + HW] = v[1]; u[o + 2 * HW] = v[2];
    }
  }
}
extern "C" int dvs_gather_triplets_u8(const uint8_t* frames, int frames_hwc, const int32_t* idx, void* out_left,
                                      void* out_target, void* out_right, int out_dtype, int B, int H, int W, void* stream) {
  if (!frames || !idx || !out_left || !out_target || !out_right || B < 1 || H < 1 || W < 1) return DVS_EINVAL;
  if (out_dtype != DVS_DTYPE_F32 && out_dtype != DVS_DTYPE_U8) return DVS_EINVAL;
  const int HW = H * W;
  int gx = (HW + 255) / 256;
  if (gx > 148 * 4) gx = 148 * 4;
  const dim3 grid(gx, B, 3);
  const bool f32 = out_dtype == DVS_DTYPE_F32;
  if (f32 && frames_hwc) gather_triplets_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(frames, idx, out_left, out_target, out_right, B, HW);
  else if (f32) gather_triplets_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(frames, idx, out_left, out_target, out_right, B, HW);
  else if (frames_hwc) gather_triplets_kernel<false, true><<<grid, 256, 0, ST(stream)>>>(frames, idx, out_left, out_target, out_right, B, HW);
  else gather_triplets_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(frames, idx, out_left, out_target, out_right, B, HW);
  LAUNCH_CHECK();
  return DVS_OK;
}

// ------------------------------------------------------------------------------------------------ supervised depth: SILog
// Scale-invariant log loss of depth/depth_learner.py:75-95:  d = log(max(pred, 1e-6)) - log(target) over valid pixels,
// loss = sqrt(mean(d^2) - variance_focus * mean(d)^2).  Forward: per-block partial sums of (d, d^2, count) in double (the
// two means nearly cancel under the square root), then one block sums them in a fixed order.  Backward is elementwise:
// d loss / d pred_p = (d_p - vf * mean(d)) / (n * loss * max(pred_p, 1e-6)) where pred_p > 1e-6 and valid, else 0.
__global__ void __launch_bounds__(256) silog_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                            const uint8_t* __restrict__ valid, int64_t n, double* __restrict__ part) {
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    if (!valid[e]) continue;
    const float d = logf(fmaxf(pred[e], 1e-6f)) - logf(target[e]);
    s1 += (double)d; s2 += (double)d * (double)d; cnt += 1.0;
  }
  __shared__ double red[3][8];
  for (int o = 16; o; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; red[2][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
  }
}
// stats[0] = loss, [1] = mean(d), [2] = n, [3] = mean(d^2)   (fp32; kept for the backward)
__global__ void silog_finish_kernel(const double* __restrict__ part, int nblk, float vf, float* __restrict__ stats) {
  __shared__ double tot[3];
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += part[(size_t)b * 3 + threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = tot[2], m1 = tot[0] / n, m2 = tot[1] / n;
    stats[0] = (float)sqrt(m2 - (double)vf * m1 * m1);
    stats[1] = (float)m1; stats[2] = (float)n; stats[3] = (float)m2;
  }
}
__global__ void __launch_bounds__(256) silog_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ stats,
                                                        const float* __restrict__ pred, const float* __restrict__ target,
                                                        const uint8_t* __restrict__ valid, float vf, int64_t n,
                                                        float* __restrict__ gpred) {
  const float k = gout[0] / (stats[2] * stats[0]), m1 = stats[1];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float g = 0.f;
    const float p = pred[e];
    if (valid[e] && p >= 1e-6f) g = k * (logf(p) - logf(target[e]) - vf * m1) / p;     // torch.clamp passes the gradient at the bound
    gpred[e] = g;
  }
}
static int silog_blocks(int64_t n) {
  int64_t b = (n + 256 * 8 - 1) / (256 * 8);
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
extern "C" int dvs_silog_workspace_bytes(int64_t n, size_t* bytes) {
  if (!bytes || n < 1) return DVS_EINVAL;
  *bytes = sizeof(double) * 3 * (size_t)silog_blocks(n) + 256;
  return DVS_OK;
}
extern "C" int dvs_silog_fwd(const float* pred, const float* target, const uint8_t* valid, int64_t n, float variance_focus,
                             float* stats, void* workspace, void* stream) {
  if (!pred || !target || !valid || !stats || n < 1) return DVS_EINVAL;
  if (!workspace || ((uintptr_t)workspace & 255)) return DVS_EWORKSPACE;
  const int nb = silog_blocks(n);
  double* part = static_cast<double*>(workspace);
  silog_partial_kernel<<<nb, 256, 0, ST(stream)>>>(pred, target, valid, n, part);
  LAUNCH_CHECK();
  silog_finish_kernel<<<1, 32, 0, ST(stream)>>>(part, nb, variance_focus, stats);
  LAUNCH_CHECK();
  return DVS_OK;
}
extern "C" int dvs_silog_bwd(const float* grad_out, const float* stats, const float* pred, const float* target,
                             const uint8_t* valid, int64_t n, float variance_focus, float* grad_pred, void* stream) {
  if (!grad_out || !stats || !pred || !target || !valid || !grad_pred || n < 1) return DVS_EINVAL;
  silog_bwd_kernel<<<silog_blocks(n), 256, 0, ST(stream)>>>(grad_out, stats, pred, target, valid, variance_focus, n, grad_pred);
  LAUNCH_CHECK();
  return DVS_OK;
}

// ------------------------------------------------------------------------------------------------ depth map -> world points
// EvalTrajectory.depth_to_pointcloud (vo/eval_traj.py:85-128) and the point cloud of vo/predict.py:81-95 on the device:
// X_world = T [ depth * K^-1 (u, v, 1) ; 1 ] for every pixel, points [H*W,3]; valid[e] = depth > 0 (the host keeps those).
__global__ void __launch_bounds__(256) pointcloud_kernel(const float* __restrict__ depth, const float* __restrict__ invK,
                                                         const float* __restrict__ T, float* __restrict__ pts,
                                                         uint8_t* __restrict__ valid, int H, int W) {
  const int n = H * W;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const float u = (float)(e % W), v = (float)(e / W), z = depth[e];
    const float rx = fmaf(invK[0], u, fmaf(invK[1], v, invK[2])), ry = fmaf(invK[4], u, fmaf(invK[5], v, invK[6])),
                rz = fmaf(invK[8], u, fmaf(invK[9], v, invK[10]));
    const float cx = rx * z, cy = ry * z, cz = rz * z;
    pts[3 * (size_t)e + 0] = fmaf(T[0], cx, fmaf(T[1], cy, fmaf(T[2], cz, T[3])));
    pts[3 * (size_t)e + 1] = fmaf(T[4], cx, fmaf(T[5], cy, fmaf(T[6], cz, T[7])));
    pts[3 * (size_t)e + 2] = fmaf(T[8], cx, fmaf(T[9], cy, fmaf(T[10], cz, T[11])));
    if (valid) valid[e] = z > 0.f;
  }
}
extern "C" int dvs_depth_to_pointcloud(const float* depth, const float* inv_K, const float* T, float* points, uint8_t* valid,
                                       int H, int W, void* stream) {
  if (!depth || !inv_K || !T || !points || H < 1 || W < 1) return DVS_EINVAL;
  int gx = (H * W + 255) / 256;
  if (gx > 148 * 8) gx = 148 * 8;
  pointcloud_kernel<<<gx, 256, 0, ST(stream)>>>(depth, inv_K, T, points, valid, H, W);
  LAUNCH_CHECK();
  return DVS_OK;
}

// ------------------------------------------------------------------------------------------------ network inputs
// What the training step does to a batch before the first convolutions, in one pass: the channels-last copies of the three
// frames (vo/train.py feeds NCHW fp32 tensors from the loader, vo/dataset/common.py:77), the concatenated pose pairs
// [source, target] / [target, source] (vo/learner_new.py:110-123), the encoders' (x - 0.45) / 0.225 (model/resnet_encoder.py
// forward) and the cast autocast applies at conv1.  Reads each frame once, writes the target [B,H,W,3] and one [B,H,W,6] pair per
// source.  The arithmetic is the stock one on CUDA (fp32 subtract, multiply by the fp32 reciprocal of 0.225 -- ATen's tensor / scalar --,
// one round-to-nearest bf16 conversion): same bits.
namespace dvs {
namespace {
template <bool OBF, int N>
__device__ __forceinline__ void store_px(void* base, size_t elem, const float* v) {      // N consecutive elements, 16-byte aligned
  if (OBF) {
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem);
#pragma unroll
    for (int q = 0; q < N / 8; ++q) {
      unsigned int w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[q * 8 + 2 * k], v[q * 8 + 2 * k + 1]);
        w[k] = *reinterpret_cast<const unsigned int*>(&h);
      }
      o[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(base) + elem);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}
// one thread = 8 consecutive pixels of one image row
template <bool OBF>
__global__ void __launch_bounds__(256) pack_net_inputs_kernel(const float* __restrict__ target, const float* const* __restrict__ dummy,
                                                              const float* s0, const float* s1, const float* s2, const float* s3,
                                                              int nsrc, unsigned int src_first_mask, int normalize, void* out_target,
                                                              void* p0, void* p1, void* p2, void* p3, int HW, size_t ngroups) {
  (void)dummy;
  const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ngroups) return;
  const size_t per_img = (size_t)HW / 8;
  const size_t b = gidx / per_img;
  const size_t pix0 = (gidx - b * per_img) * 8;                    // first pixel of the group inside the image
  const float* srcs[4] = {s0, s1, s2, s3};
  void* pairs[4] = {p0, p1, p2, p3};
  auto load8 = [&](const float* img, int c, float* v) {
    const float4* q = reinterpret_cast<const float4*>(img + (b * 3 + c) * (size_t)HW + pix0);
    const float4 a = q[0], d = q[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    if (normalize)
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - 0.45f) * (1.0f / 0.225f);   // ATen's CUDA tensor / scalar: times the fp32 reciprocal
  };
  float t[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) load8(target, c, t[c]);
  {
    float o[24];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[i * 3 + c] = t[c][i];
    store_px<OBF, 24>(out_target, (b * HW + pix0) * 3, o);
  }
  for (int k = 0; k < nsrc; ++k) {
    float s[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) load8(srcs[k], c, s[c]);
    const bool first = (src_first_mask >> k) & 1u;
    float o[48];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        o[i * 6 + c] = first ? s[c][i] : t[c][i];
        o[i * 6 + 3 + c] = first ? t[c][i] : s[c][i];
      }
    store_px<OBF, 48>(pairs[k], (b * HW + pix0) * 6, o);
  }
}
}  // namespace
}  // namespace dvs

extern "C" int dvs_pack_net_inputs(const float* target, const float* const* sources, int num_sources, unsigned int src_first_mask,
                                   int normalize, void* out_target, void* const* out_pairs, int out_dtype, int B, int H, int W,
                                   void* stream) {
  using namespace dvs;
  if (!target || !sources || !out_target || !out_pairs || num_sources < 1 || num_sources > 4 || B < 1 || H < 1 || W < 1)
    return DVS_EINVAL;
  if (out_dtype != DVS_DTYPE_F32 && out_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  if (((size_t)H * W) % 8) return DVS_EINVAL;                          // groups of 8 pixels must not straddle images
  if ((uintptr_t)target & 15) return DVS_EINVAL;
  const float* s[4] = {nullptr, nullptr, nullptr, nullptr};
  void* p[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < num_sources; ++k) {
    s[k] = sources[k];
    p[k] = out_pairs[k];
    if (!s[k] || !p[k] || ((uintptr_t)s[k] & 15)) return DVS_EINVAL;
  }
  const size_t ngroups = (size_t)B * H * W / 8;
  const unsigned int nblk = (unsigned int)((ngroups + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == DVS_DTYPE_BF16)
    pack_net_inputs_kernel<true><<<nblk, 256, 0, st>>>(target, nullptr, s[0], s[1], s[2], s[3], num_sources, src_first_mask, normalize,
                                                       out_target, p[0], p[1], p[2], p[3], H * W, ngroups);
  else
    pack_net_inputs_kernel<false><<<nblk, 256, 0, st>>>(target, nullptr, s[0], s[1], s[2], s[3], num_sources, src_first_mask, normalize,
                                                        out_target, p[0], p[1], p[2], p[3], H * W, ngroups);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}
