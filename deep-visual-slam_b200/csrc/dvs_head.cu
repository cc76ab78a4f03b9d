// Disparity head of DepthNet fused for sm_100a: ReflectionPad2d(1) + Conv2d(C, 1, 3) + Sigmoid in one pass, and its backward.
//
// Reference: the four ("dispconv", s) blocks, model/depthnet.py:57-58,87-88 = Conv3x3 (model/layers.py:120-136) followed by
// nn.Sigmoid.  Stock PyTorch runs them as pad (copy of the C-channel activation) -> cuDNN convolution with ONE output
// channel (a matrix-vector product on the tensor cores' worst shape) -> sigmoid, and three more passes backward.  The op is
// a per-pixel dot product of 9 C channel vectors: HBM bound, no contraction worth a tensor core.  Here each thread owns
// one pixel, reads the 9 channel vectors with 128-bit loads (channels-last, bf16 or fp32; neighbours share taps through
// L1), accumulates in fp32 and writes the sigmoid disparity in the dtype the loss kernel reads (bf16 or fp32).  Backward,
// per INPUT pixel: the nine sums s[t] of d loss / d pre-activation over the outputs that read this pixel through tap t
// (reflection folds the pad ring onto rows / columns 1 and H-2 / W-2), grad_x[c] = sum_t w[c][t] s[t], and
// grad_w[c][t] = sum_pixels x[c] s[t] by a two-stage fixed-order reduction (per-block partials, then one small kernel).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dvsloss.h"
#include "dvs_host.h"

namespace dvs {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxC = 128;   // backward stages 256 x (C + 1) floats of x in shared memory

template <bool BF16>
__device__ __forceinline__ void load8(const void* base, size_t elem, float* v) {
  if (BF16) {
    const uint4 q = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem);
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __uint_as_float(w[k] << 16);
      v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <bool BF16>
__device__ __forceinline__ void store8(void* base, size_t elem, const float* v) {
  if (BF16) {
    uint4 q;
    unsigned int w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      w[k] = *reinterpret_cast<const unsigned int*>(&h);
    }
    q.x = w[0]; q.y = w[1]; q.z = w[2]; q.w = w[3];
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem) = q;
  } else {
    *reinterpret_cast<float4*>(static_cast<float*>(base) + elem) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(static_cast<float*>(base) + elem + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}
template <bool BF16>
__device__ __forceinline__ float load1(const void* base, size_t elem) {
  if (BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[elem]);
  return static_cast<const float*>(base)[elem];
}
template <bool BF16>
__device__ __forceinline__ void store1(void* base, size_t elem, float v) {
  if (BF16) static_cast<__nv_bfloat16*>(base)[elem] = __float2bfloat16_rn(v);
  else static_cast<float*>(base)[elem] = v;
}
__device__ __forceinline__ int reflect1(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * (n - 1) - p : p); }

// ------------------------------------------------------------------------------------------------ forward
// x [B,H,W,C] channels-last, w [C][9] fp32 (tap-minor: w[c*9 + ky*3 + kx], the Conv2d weight [1,C,3,3] as stored), disp [B,1,H,W]
template <bool XBF, bool DBF>
__global__ void __launch_bounds__(kHeadThreads) disp_head_fwd_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, void* __restrict__ disp,
                                                                     int B, int C, int H, int W) {
  extern __shared__ float ws[];                       // [9][C]: tap-major so that a thread walks channels contiguously
  for (int e = threadIdx.x; e < 9 * C; e += blockDim.x) ws[(e % 9) * C + e / 9] = w[e];
  __syncthreads();
  const size_t P = (size_t)B * H * W;
  const float b0 = bias ? bias[0] : 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const size_t img = (p / ((size_t)H * W)) * H * W;
    float acc = b0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ry = reflect1(yy + ky - 1, H);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int rx = reflect1(xx + kx - 1, W);
        const size_t o = (img + (size_t)ry * W + rx) * C;
        const float* wt = ws + (ky * 3 + kx) * C;
        for (int c = 0; c < C; c += 8) {
          float v[8];
          load8<XBF>(x, o + c, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = fmaf(v[k], wt[c + k], acc);
        }
      }
    }
    store1<DBF>(disp, p, 1.0f / (1.0f + __expf(-acc)));
  }
}

// ------------------------------------------------------------------------------------------------ backward
// d loss / d pre-activation of output pixel (yo, xo): g * d * (1 - d), zero outside the image
template <bool GBF, bool DBF>
__device__ __forceinline__ float gpre(const void* g, const void* d, size_t img, int yo, int xo, int H, int W) {
  if (yo < 0 || yo >= H || xo < 0 || xo >= W) return 0.f;
  const size_t o = img + (size_t)yo * W + xo;
  const float dv = load1<DBF>(d, o);
  return load1<GBF>(g, o) * dv * (1.f - dv);
}

// One block = kChunks consecutive chunks of 256 pixels.  Per chunk: thread -> pixel: s[9] (+ own g_pre for the bias), grad_x;
// x and s staged in shared memory; then thread -> (channel, tap) pairs accumulate x[c] * s[t] over the chunk's pixels.
template <bool XBF, bool GBF, bool DBF, bool GXBF>
__global__ void __launch_bounds__(kHeadThreads) disp_head_bwd_kernel(const void* __restrict__ gdisp, const void* __restrict__ disp,
                                                                     const void* __restrict__ x, const float* __restrict__ w,
                                                                     void* __restrict__ gx, float* __restrict__ partial,
                                                                     int B, int C, int H, int W, int chunks_per_block) {
  extern __shared__ float sm[];
  float* ws = sm;                                     // [9][C]
  float* S = ws + 9 * C;                              // [256][10]: s[0..8], own g_pre
  float* X = S + kHeadThreads * 10;                   // [256][C + 1] (padded against bank conflicts)
  const int XP = C + 1;
  for (int e = threadIdx.x; e < 9 * C; e += blockDim.x) ws[(e % 9) * C + e / 9] = w[e];
  const size_t P = (size_t)B * H * W;
  const int npairs = 9 * C + 1;                       // + the bias
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};           // pairs tid, tid + 256, ... (9 * 128 + 1 <= 5 * 256)
  __syncthreads();
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const size_t p = ((size_t)blockIdx.x * chunks_per_block + ch) * kHeadThreads + threadIdx.x;
    float s[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) s[t] = 0.f;
    if (p < P) {
      const int xx = (int)(p % W), yy = (int)((p / W) % H);
      const size_t img = (p / ((size_t)H * W)) * H * W;
      // outputs that read this input pixel through tap (ky, kx): the regular one, plus the ones whose pad-ring tap folds here
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        int yo[2] = {yy - ky + 1, -1};
        if (ky == 0 && yy == 1) yo[1] = 0;
        if (ky == 2 && yy == H - 2) yo[1] = H - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          int xo[2] = {xx - kx + 1, -1};
          if (kx == 0 && xx == 1) xo[1] = 0;
          if (kx == 2 && xx == W - 2) xo[1] = W - 1;
          float a = gpre<GBF, DBF>(gdisp, disp, img, yo[0], xo[0], H, W);
          if (xo[1] >= 0) a += gpre<GBF, DBF>(gdisp, disp, img, yo[0], xo[1], H, W);
          if (yo[1] >= 0) {
            a += gpre<GBF, DBF>(gdisp, disp, img, yo[1], xo[0], H, W);
            if (xo[1] >= 0) a += gpre<GBF, DBF>(gdisp, disp, img, yo[1], xo[1], H, W);
          }
          s[ky * 3 + kx] = a;
        }
      }
      s[9] = gpre<GBF, DBF>(gdisp, disp, img, yy, xx, H, W);
      for (int c = 0; c < C; c += 8) {
        float v[8], o[8];
        load8<XBF>(x, p * C + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          X[threadIdx.x * XP + c + k] = v[k];
          float a = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) a = fmaf(ws[t * C + c + k], s[t], a);
          o[k] = a;
        }
        store8<GXBF>(gx, p * C + c, o);
      }
    } else {
      for (int c = 0; c < C; ++c) X[threadIdx.x * XP + c] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 10; ++t) S[threadIdx.x * 10 + t] = s[t];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int pair = threadIdx.x + k * kHeadThreads;
      if (pair < npairs) {
        const int c = pair / 9, t = pair - c * 9;       // pair == 9 C: the bias (t = 9 against a constant 1)
        float a = acc[k];
        if (pair == 9 * C) {
          for (int q = 0; q < kHeadThreads; ++q) a += S[q * 10 + 9];
        } else {
          for (int q = 0; q < kHeadThreads; ++q) a = fmaf(X[q * XP + c], S[q * 10 + t], a);
        }
        acc[k] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int pair = threadIdx.x + k * kHeadThreads;
    if (pair < npairs) partial[(size_t)blockIdx.x * npairs + pair] = acc[k];
  }
}

// grad_w[c][t] and grad_bias: fixed-order sum of the block partials (one warp per pair, lanes stride the blocks)
__global__ void __launch_bounds__(256) disp_head_reduce_kernel(const float* __restrict__ partial, int nblk, int npairs,
                                                               float* __restrict__ grad_w, float* __restrict__ grad_b) {
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= npairs) return;
  float a = 0.f;
  for (int b = lane; b < nblk; b += 32) a += partial[(size_t)b * npairs + pair];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    if (pair == npairs - 1) { if (grad_b) grad_b[0] = a; }
    else grad_w[pair] = a;
  }
}

static int head_check(const void* x, int x_dtype, int B, int C, int H, int W) {
  if (!x || B < 1 || H < 3 || W < 3 || C < 8 || (C & 7) || C > kHeadMaxC) return DVS_EINVAL;
  if (x_dtype != DVS_DTYPE_F32 && x_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  if ((uintptr_t)x & 15) return DVS_EINVAL;
  return DVS_OK;
}
static int head_chunks(size_t P) {        // pixels per block = 256 * chunks: about 2 400 blocks at 32 x 480 x 640
  size_t c = (P / kHeadThreads + 2399) / 2400;
  return (int)(c < 1 ? 1 : (c > 64 ? 64 : c));
}

}  // namespace dvs

using namespace dvs;

extern "C" int dvs_disp_head_fwd(const void* x, int x_dtype, const float* weight, const float* bias, void* disp, int disp_dtype,
                                 int B, int C, int H, int W, void* stream) {
  int rc = head_check(x, x_dtype, B, C, H, W);
  if (rc) return rc;
  if (!weight || !disp || (disp_dtype != DVS_DTYPE_F32 && disp_dtype != DVS_DTYPE_BF16)) return DVS_EINVAL;
  const size_t P = (size_t)B * H * W;
  int grid = (int)((P + kHeadThreads - 1) / kHeadThreads);
  if (grid > 148 * 32) grid = 148 * 32;
  const size_t smem = sizeof(float) * 9 * C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool xb = x_dtype == DVS_DTYPE_BF16, db = disp_dtype == DVS_DTYPE_BF16;
  if (xb && db) disp_head_fwd_kernel<true, true><<<grid, kHeadThreads, smem, st>>>(x, weight, bias, disp, B, C, H, W);
  else if (xb) disp_head_fwd_kernel<true, false><<<grid, kHeadThreads, smem, st>>>(x, weight, bias, disp, B, C, H, W);
  else if (db) disp_head_fwd_kernel<false, true><<<grid, kHeadThreads, smem, st>>>(x, weight, bias, disp, B, C, H, W);
  else disp_head_fwd_kernel<false, false><<<grid, kHeadThreads, smem, st>>>(x, weight, bias, disp, B, C, H, W);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

extern "C" int dvs_disp_head_bwd_workspace_bytes(int B, int C, int H, int W, size_t* bytes) {
  if (!bytes || B < 1 || H < 3 || W < 3 || C < 8 || (C & 7) || C > kHeadMaxC) return DVS_EINVAL;
  const size_t P = (size_t)B * H * W;
  const int chunks = head_chunks(P);
  const size_t nblk = (P + (size_t)kHeadThreads * chunks - 1) / ((size_t)kHeadThreads * chunks);
  *bytes = sizeof(float) * nblk * (9 * (size_t)C + 1) + 256;
  return DVS_OK;
}

template <bool XBF, bool GBF>
static int head_bwd_launch(const void* gdisp, const void* disp, const void* x, const float* weight, void* gx, float* partial,
                           int B, int C, int H, int W, int chunks, int nblk, size_t smem, cudaStream_t st) {
  // x, grad_x share a dtype (the activation's); grad_disp, disp share a dtype (the head's output)
  auto k = disp_head_bwd_kernel<XBF, GBF, GBF, XBF>;
  DVS_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<nblk, kHeadThreads, smem, st>>>(gdisp, disp, x, weight, gx, partial, B, C, H, W, chunks);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

extern "C" int dvs_disp_head_bwd(const void* grad_disp, const void* disp, int disp_dtype, const void* x, int x_dtype,
                                 const float* weight, void* grad_x, float* grad_weight, float* grad_bias, int B, int C, int H,
                                 int W, void* workspace, void* stream) {
  int rc = head_check(x, x_dtype, B, C, H, W);
  if (rc) return rc;
  if (!grad_disp || !disp || !weight || !grad_x || !grad_weight) return DVS_EINVAL;
  if (disp_dtype != DVS_DTYPE_F32 && disp_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  if (!workspace || ((uintptr_t)workspace & 255) || ((uintptr_t)grad_x & 15)) return DVS_EWORKSPACE;
  const size_t P = (size_t)B * H * W;
  const int chunks = head_chunks(P);
  const int nblk = (int)((P + (size_t)kHeadThreads * chunks - 1) / ((size_t)kHeadThreads * chunks));
  const size_t smem = sizeof(float) * (9 * (size_t)C + kHeadThreads * 10 + (size_t)kHeadThreads * (C + 1));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  const bool xb = x_dtype == DVS_DTYPE_BF16, gb = disp_dtype == DVS_DTYPE_BF16;
  if (xb && gb) rc = head_bwd_launch<true, true>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else if (xb) rc = head_bwd_launch<true, false>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else if (gb) rc = head_bwd_launch<false, true>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else rc = head_bwd_launch<false, false>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  if (rc) return rc;
  const int npairs = 9 * C + 1;
  disp_head_reduce_kernel<<<(npairs + 7) / 8, 256, 0, st>>>(partial, nblk, npairs, grad_weight, grad_bias);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

// ================================================================================================ decoder glue
// The element-wise tail between the two convolutions of a decoder stage (model/depthnet.py:77-84, model/layers.py:106-117,
// 196-199): out = cat([nearest_up2(ELU(x)), skip], channel axis), channels-last, one pass.  Stock PyTorch runs ELU, the
// up-sampling and the concatenation as three kernels with two intermediate tensors; backward likewise (slice copies, the 2x2
// sum of the up-sampling, the ELU derivative).  Pure HBM streaming: every thread moves one 16-byte vector.
namespace dvs {

template <bool BF>
struct Vec16 {
  uint4 raw;
  static constexpr int N = BF ? 8 : 4;
  __device__ float get(int i) const {
    if (BF) {
      const unsigned int w = (&raw.x)[i >> 1];
      return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
    }
    return __uint_as_float((&raw.x)[i]);
  }
  __device__ void set_all(const float* v) {
    if (BF) {
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        (&raw.x)[i] = *reinterpret_cast<const unsigned int*>(&h);
      }
    } else {
      for (int i = 0; i < 4; ++i) (&raw.x)[i] = __float_as_uint(v[i]);
    }
  }
};

__device__ __forceinline__ float elu_value(float a) { return a <= 0.f ? expf(a) - 1.f : a; }          // ATen elu_kernel, alpha = scale = 1
__device__ __forceinline__ float elu_slope(float a) { return a <= 0.f ? expf(a) : 1.f; }              // ATen elu_backward_kernel (is_result = false)

// one thread per 16-byte vector of `out` [B, 2h, 2w, C1 + C2]
template <bool BF>
__global__ void __launch_bounds__(256) elu_up2_cat_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip,
                                                              uint4* __restrict__ out, int h, int w, int v1, int v2,
                                                              size_t nvec) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  const int vt = v1 + v2;
  const size_t pix = e / vt;
  const int v = (int)(e - pix * vt);
  if (v >= v1) {
    out[e] = skip[pix * v2 + (v - v1)];
    return;
  }
  const int W2 = 2 * w, H2 = 2 * h;
  const int X = (int)(pix % W2);
  const size_t r = pix / W2;
  const int Y = (int)(r % H2);
  const size_t b = r / H2;
  Vec16<BF> a;
  a.raw = x[((b * h + (Y >> 1)) * w + (X >> 1)) * v1 + v];
  float f[Vec16<BF>::N];
  for (int i = 0; i < Vec16<BF>::N; ++i) f[i] = elu_value(a.get(i));
  Vec16<BF> o;
  o.set_all(f);
  out[e] = o.raw;
}

// one thread per 16-byte vector of grad_x [B, h, w, C1]: 2x2 sum of grad_out (fixed order) times ELU'(x)
template <bool BF>
__global__ void __launch_bounds__(256) elu_up2_cat_bwd_x_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gout,
                                                                uint4* __restrict__ gx, int h, int w, int v1, int v2,
                                                                size_t nvec) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  const size_t pix = e / v1;
  const int v = (int)(e - pix * v1);
  const int xx = (int)(pix % w);
  const size_t r = pix / w;
  const int yy = (int)(r % h);
  const size_t b = r / h;
  const int vt = v1 + v2, W2 = 2 * w;
  const size_t o00 = ((b * 2 * h + 2 * yy) * W2 + 2 * xx) * vt + v;
  Vec16<BF> g00, g01, g10, g11, a;
  g00.raw = gout[o00];
  g01.raw = gout[o00 + vt];
  g10.raw = gout[o00 + (size_t)W2 * vt];
  g11.raw = gout[o00 + (size_t)W2 * vt + vt];
  a.raw = x[e];
  float f[Vec16<BF>::N];
  for (int i = 0; i < Vec16<BF>::N; ++i) f[i] = ((g00.get(i) + g01.get(i)) + (g10.get(i) + g11.get(i))) * elu_slope(a.get(i));
  Vec16<BF> o;
  o.set_all(f);
  gx[e] = o.raw;
}

// one thread per 16-byte vector of grad_skip [B, 2h, 2w, C2]: the channel slice of grad_out
__global__ void __launch_bounds__(256) cat_bwd_skip_kernel(const uint4* __restrict__ gout, uint4* __restrict__ gskip, int v1,
                                                           int v2, size_t nvec) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  const size_t pix = e / v2;
  const int v = (int)(e - pix * v2);
  gskip[e] = gout[pix * (v1 + v2) + v1 + v];
}

static int glue_check(const void* x, int dtype, int B, int C1, int C2, int h, int w, const void* skip) {
  if (!x || B < 1 || h < 1 || w < 1 || C1 < 1 || C2 < 0) return DVS_EINVAL;
  if (dtype != DVS_DTYPE_F32 && dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  const int per = dtype == DVS_DTYPE_BF16 ? 8 : 4;
  if (C1 % per || C2 % per) return DVS_EINVAL;
  if ((C2 > 0) != (skip != nullptr)) return DVS_EINVAL;
  if (((uintptr_t)x & 15) || ((uintptr_t)skip & 15)) return DVS_EINVAL;
  return DVS_OK;
}

}  // namespace dvs

extern "C" int dvs_elu_up2_cat_fwd(const void* x, const void* skip, void* out, int dtype, int B, int C1, int C2, int h, int w,
                                   void* stream) {
  using namespace dvs;
  int rc = glue_check(x, dtype, B, C1, C2, h, w, skip);
  if (rc) return rc;
  if (!out || ((uintptr_t)out & 15)) return DVS_EINVAL;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int per = bf ? 8 : 4, v1 = C1 / per, v2 = C2 / per;
  const size_t nvec = (size_t)B * 4 * h * w * (v1 + v2);
  const unsigned int nblk = (unsigned int)((nvec + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bf) elu_up2_cat_fwd_kernel<true><<<nblk, 256, 0, st>>>((const uint4*)x, (const uint4*)skip, (uint4*)out, h, w, v1, v2, nvec);
  else elu_up2_cat_fwd_kernel<false><<<nblk, 256, 0, st>>>((const uint4*)x, (const uint4*)skip, (uint4*)out, h, w, v1, v2, nvec);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

extern "C" int dvs_elu_up2_cat_bwd(const void* x, const void* grad_out, void* grad_x, void* grad_skip, int dtype, int B, int C1,
                                   int C2, int h, int w, void* stream) {
  using namespace dvs;
  int rc = glue_check(x, dtype, B, C1, C2, h, w, C2 > 0 ? grad_skip : nullptr);
  if (rc) return rc;
  if (!grad_out || !grad_x || ((uintptr_t)grad_out & 15) || ((uintptr_t)grad_x & 15)) return DVS_EINVAL;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int per = bf ? 8 : 4, v1 = C1 / per, v2 = C2 / per;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nx = (size_t)B * h * w * v1;
  const unsigned int bx = (unsigned int)((nx + 255) / 256);
  if (bf) elu_up2_cat_bwd_x_kernel<true><<<bx, 256, 0, st>>>((const uint4*)x, (const uint4*)grad_out, (uint4*)grad_x, h, w, v1, v2, nx);
  else elu_up2_cat_bwd_x_kernel<false><<<bx, 256, 0, st>>>((const uint4*)x, (const uint4*)grad_out, (uint4*)grad_x, h, w, v1, v2, nx);
  DVS_CUDA_TRY(cudaGetLastError());
  if (v2 > 0) {
    const size_t ns = (size_t)B * 4 * h * w * v2;
    cat_bwd_skip_kernel<<<(unsigned int)((ns + 255) / 256), 256, 0, st>>>((const uint4*)grad_out, (uint4*)grad_skip, v1, v2, ns);
    DVS_CUDA_TRY(cudaGetLastError());
  }
  return DVS_OK;
}
