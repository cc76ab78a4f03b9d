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


// ------------------------------------------------------------------------------------------------ tile kernels
// C = 8 LP channels with LP a power of two (DepthNet: 16, 32, 64, 128).  A block owns a 32 x 8 tile of pixels; LP lanes of a
// warp share one pixel, each owning 8 channels, so a warp-wide 128-bit access covers 32 / LP whole pixels contiguously (the
// thread-per-pixel kernels above stride their lanes by C elements and fetch every activation nine times).  The lane's 72
// weights (or 72 weight-gradient accumulators) live in registers.
//   forward : every activation of the tile + halo (34 x 10) is read ONCE and turned into its nine partial dot products
//             (one per tap) in shared memory; the outputs then add nine scalars each.
//   backward: d loss / d pre-activation of the tile + halo goes to shared memory once; the nine tap sums of a pixel are then
//             shared-memory reads.  grad_x and grad_w are separate kernels (72 registers of weights / of accumulators each).
constexpr int kTileW = 32, kTileH = 8, kHaloW = kTileW + 2, kHaloH = kTileH + 2, kHaloN = kHaloW * kHaloH;

struct HeadTile {
  int b, y0, x0;
};
__device__ __forceinline__ HeadTile head_tile(int t, int tiles_x, int tiles_y) {
  HeadTile r;
  r.b = t / (tiles_x * tiles_y);
  const int q = t - r.b * tiles_x * tiles_y;
  r.y0 = (q / tiles_x) * kTileH;
  r.x0 = (q % tiles_x) * kTileW;
  return r;
}
template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// a lane's 8 channels as loaded (the conversion is deferred so that several loads can be in flight per thread: with one
// dependent load per warp at a time these kernels ran at 1 TB/s, Little's law on 16 warps per SM)
template <bool BF>
struct Raw8 {
  uint4 a, b;                                                // b: second half of 8 floats; unused for bf16
};
template <bool BF>
__device__ __forceinline__ Raw8<BF> raw_load(const void* base, size_t elem) {
  Raw8<BF> r;
  if (BF) {
    r.a = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem);
    r.b = r.a;
  } else {
    r.a = *reinterpret_cast<const uint4*>(static_cast<const float*>(base) + elem);
    r.b = *reinterpret_cast<const uint4*>(static_cast<const float*>(base) + elem + 4);
  }
  return r;
}
template <bool BF>
__device__ __forceinline__ void raw_unpack(const Raw8<BF>& r, float* v) {
  if (BF) {
    const unsigned int w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __uint_as_float(w[k] << 16);
      v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  } else {
    v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
    v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
  }
}
// the lane's 72 weights w[c][t], c = 8 sub .. 8 sub + 7: consecutive in the Conv2d weight [1,C,3,3]
__device__ __forceinline__ void load_lane_weights(const float* __restrict__ w, int sub, float (*wr)[8]) {
  float f[72];
  const float* q = w + sub * 72;
  if (((uintptr_t)w & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const float4 t = reinterpret_cast<const float4*>(q)[i];
      f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 72; ++i) f[i] = q[i];
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[t][k] = f[k * 9 + t];
}

template <bool XBF, bool DBF, int LP>
__global__ void __launch_bounds__(kHeadThreads) disp_head_fwd_tile_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                                          const float* __restrict__ bias, void* __restrict__ disp,
                                                                          int H, int W, int tiles_x, int tiles_y, int xpad) {
  constexpr int C = 8 * LP, PPI = kHeadThreads / LP;         // pixels per pass of the block
  constexpr int NPASS = (kHaloN + PPI - 1) / PPI, U = NPASS < 6 ? NPASS : 6;
  const int Hp = H + 2 * xpad, Wp = W + 2 * xpad;            // x may carry a reflected ring (then only its interior is read)
  __shared__ float pt[9][kHaloN + 4];
  const int sub = threadIdx.x & (LP - 1), pl = threadIdx.x / LP;
  float wr[9][8];
  load_lane_weights(w, sub, wr);
  const HeadTile T = head_tile(blockIdx.x, tiles_x, tiles_y);
  const size_t img = (size_t)T.b * H * W;
  // partial dot products of the halo region; positions outside the image hold the REFLECTED pixel's values
#pragma unroll 1
  for (int p0 = 0; p0 < NPASS; p0 += U) {
    Raw8<XBF> r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {                             // up to six 16-byte loads in flight per thread
      const int q = (p0 + u) * PPI + pl;
      const int qq = q < kHaloN ? q : kHaloN - 1;
      const int ly = qq / kHaloW, lx = qq - ly * kHaloW;
      int gy = reflect1(T.y0 + ly - 1, H), gx = reflect1(T.x0 + lx - 1, W);
      gy = gy < 0 ? 0 : (gy > H - 1 ? H - 1 : gy);            // tiles overhanging the image: any valid address
      gx = gx < 0 ? 0 : (gx > W - 1 ? W - 1 : gx);
      r[u] = raw_load<XBF>(x, (((size_t)T.b * Hp + gy + xpad) * Wp + gx + xpad) * C + sub * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u >= NPASS) break;                             // block-uniform
      const int q = (p0 + u) * PPI + pl;
      float v[8];
      raw_unpack<XBF>(r[u], v);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) a = fmaf(v[k], wr[t][k], a);
        a = group_sum<LP>(a);
        if (q < kHaloN && sub == 0) pt[t][q] = a;
      }
    }
  }
  __syncthreads();
  const int ox = threadIdx.x & 31, oy = threadIdx.x >> 5;
  const int yy = T.y0 + oy, xx = T.x0 + ox;
  if (yy < H && xx < W) {
    float acc = bias ? bias[0] : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc += pt[ky * 3 + kx][(oy + ky) * kHaloW + ox + kx];
    store1<DBF>(disp, img + (size_t)yy * W + xx, 1.0f / (1.0f + __expf(-acc)));
  }
}

// d loss / d pre-activation of the tile + halo -> shared memory (zero outside the image)
template <bool GBF, bool DBF>
__device__ __forceinline__ void stage_gpre(const void* gdisp, const void* disp, const HeadTile& T, size_t img, int H, int W, float* gp) {
  for (int q = threadIdx.x; q < kHaloN; q += kHeadThreads) {
    const int ly = q / kHaloW, lx = q - ly * kHaloW;
    gp[q] = gpre<GBF, DBF>(gdisp, disp, img, T.y0 + ly - 1, T.x0 + lx - 1, H, W);
  }
}
// the same in two steps: the operands of the NEXT tile are fetched into registers while the current tile is processed
struct GpreRegs {
  float g[2], d[2];
};
template <bool GBF, bool DBF>
__device__ __forceinline__ void gpre_fetch(const void* gdisp, const void* disp, const HeadTile& T, size_t img, int H, int W, GpreRegs& r) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = threadIdx.x + i * kHeadThreads;
    r.g[i] = 0.f; r.d[i] = 0.f;
    if (q < kHaloN) {
      const int ly = q / kHaloW, lx = q - ly * kHaloW;
      const int yo = T.y0 + ly - 1, xo = T.x0 + lx - 1;
      if (yo >= 0 && yo < H && xo >= 0 && xo < W) {
        const size_t o = img + (size_t)yo * W + xo;
        r.g[i] = load1<GBF>(gdisp, o);
        r.d[i] = load1<DBF>(disp, o);
      }
    }
  }
}
__device__ __forceinline__ void gpre_store(const GpreRegs& r, float* gp) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = threadIdx.x + i * kHeadThreads;
    if (q < kHaloN) gp[q] = r.g[i] * r.d[i] * (1.f - r.d[i]);
  }
}
// the nine sums s[t] over the outputs that read input pixel (yy, xx) = tile position (oy, ox) through tap t: the regular one
// plus the ones whose pad-ring tap folds here (all within one pixel: the halo)
__device__ __forceinline__ void tap_sums(const float* gp, int oy, int ox, int yy, int xx, int H, int W, float* s) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ry = oy + 1 - ky + 1;                          // halo row of output yy - ky + 1
    const int fy = (ky == 0 && yy == 1) ? oy : ((ky == 2 && yy == H - 2) ? oy + 2 : -1);   // halo row of the folded output
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int rx = ox + 1 - kx + 1;
      const int fx = (kx == 0 && xx == 1) ? ox : ((kx == 2 && xx == W - 2) ? ox + 2 : -1);
      float a = gp[ry * kHaloW + rx];
      if (fx >= 0) a += gp[ry * kHaloW + fx];
      if (fy >= 0) {
        a += gp[fy * kHaloW + rx];
        if (fx >= 0) a += gp[fy * kHaloW + fx];
      }
      s[ky * 3 + kx] = a;
    }
  }
}

// grad_x[c] = sum_t w[c][t] s[t]; blocks persist over the tiles and fetch the next tile's operands while they work
template <bool GBF, bool DBF, bool GXBF, int LP>
__global__ void __launch_bounds__(kHeadThreads) disp_head_bwd_x_tile_kernel(const void* __restrict__ gdisp, const void* __restrict__ disp,
                                                                            const float* __restrict__ w, void* __restrict__ gx, int H,
                                                                            int W, int tiles_x, int tiles_y, int ntiles, int xpad) {
  constexpr int C = 8 * LP, PPI = kHeadThreads / LP;
  const int Hp = H + 2 * xpad, Wp = W + 2 * xpad;            // grad_x in x's layout (the ring, if any, is zeroed by the caller)
  __shared__ float gp[kHaloN];
  const int sub = threadIdx.x & (LP - 1), pl = threadIdx.x / LP;
  float wr[9][8];
  load_lane_weights(w, sub, wr);
  GpreRegs pre;
  if ((int)blockIdx.x < ntiles) {
    const HeadTile T0 = head_tile(blockIdx.x, tiles_x, tiles_y);
    gpre_fetch<GBF, DBF>(gdisp, disp, T0, (size_t)T0.b * H * W, H, W, pre);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const HeadTile T = head_tile(tile, tiles_x, tiles_y);
    const size_t img = (size_t)T.b * H * W;
    __syncthreads();                                          // the previous tile's readers are done with gp
    gpre_store(pre, gp);
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) {
      const HeadTile Tn = head_tile(tile + gridDim.x, tiles_x, tiles_y);
      gpre_fetch<GBF, DBF>(gdisp, disp, Tn, (size_t)Tn.b * H * W, H, W, pre);
    }
#pragma unroll 1
    for (int q0 = 0; q0 < kTileW * kTileH; q0 += PPI) {
      const int q = q0 + pl, oy = q / kTileW, ox = q - oy * kTileW;
      const int yy = T.y0 + oy, xx = T.x0 + ox;
      if (yy >= H || xx >= W) continue;
      float s[9];
      tap_sums(gp, oy, ox, yy, xx, H, W, s);
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(wr[t][k], s[t], a);
        o[k] = a;
      }
      store8<GXBF>(gx, (((size_t)T.b * Hp + yy + xpad) * Wp + xx + xpad) * C + sub * 8, o);
    }
  }
}

// grad_w[c][t] = sum_pixels x[c] s[t], grad_bias = sum_pixels s[centre]: per-lane accumulators over the block's tiles -> fixed-
// order reduction over the lanes that own the same channels, the warps of the block, and (disp_head_reduce_kernel) the blocks
template <bool XBF, bool GBF, bool DBF, int LP>
__global__ void __launch_bounds__(kHeadThreads) disp_head_bwd_w_tile_kernel(const void* __restrict__ gdisp, const void* __restrict__ disp,
                                                                            const void* __restrict__ x, float* __restrict__ partial,
                                                                            int H, int W, int tiles_x, int tiles_y, int ntiles, int xpad) {
  constexpr int C = 8 * LP, PPI = kHeadThreads / LP, NP = 9 * C + 1, NW = kHeadThreads / 32;
  constexpr int NPASS = LP, U = NPASS < 4 ? NPASS : 4;       // activation loads in flight per thread
  const int Hp = H + 2 * xpad, Wp = W + 2 * xpad;
  extern __shared__ float red[];                       // [NW][NP]; the first kHaloN floats double as the g_pre stage
  float* gp = red;
  const int lane = threadIdx.x & 31, sub = threadIdx.x & (LP - 1), pl = threadIdx.x / LP, wid = threadIdx.x >> 5;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  float accb = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const HeadTile T = head_tile(tile, tiles_x, tiles_y);
    const size_t img = (size_t)T.b * H * W;
    Raw8<XBF> r[U];
    auto fetch = [&](int p0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = (p0 + u) * PPI + pl, oy = q / kTileW, ox = q - oy * kTileW;
        int yy = T.y0 + oy, xx = T.x0 + ox;
        yy = yy > H - 1 ? H - 1 : yy; xx = xx > W - 1 ? W - 1 : xx;
        r[u] = raw_load<XBF>(x, (((size_t)T.b * Hp + yy + xpad) * Wp + xx + xpad) * C + sub * 8);
      }
    };
    fetch(0);                                                 // in flight while g_pre is staged
    __syncthreads();
    stage_gpre<GBF, DBF>(gdisp, disp, T, img, H, W, gp);
    __syncthreads();
#pragma unroll 1
    for (int p0 = 0; p0 < NPASS; p0 += U) {
      if (p0) fetch(p0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = (p0 + u) * PPI + pl, oy = q / kTileW, ox = q - oy * kTileW;
        const int yy = T.y0 + oy, xx = T.x0 + ox;
        if (yy >= H || xx >= W) continue;
        float v[8];
        raw_unpack<XBF>(r[u], v);
        float s[9];
        tap_sums(gp, oy, ox, yy, xx, H, W, s);
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(v[k], s[t], acc[t][k]);
        if (sub == 0) accb += s[4];                    // the centre tap never folds: s[4] is the pixel's own g_pre
      }
    }
  }
  __syncthreads();
  // lanes with the same `sub` (different pixels): xor offsets LP .. 16
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = acc[t][k];
#pragma unroll
      for (int o = LP; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[t][k] = a;
    }
#pragma unroll
  for (int o = LP; o < 32; o <<= 1) accb += __shfl_xor_sync(0xffffffffu, accb, o);
  if (lane < LP) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[wid * NP + (sub * 8 + k) * 9 + t] = acc[t][k];
    if (lane == 0) red[wid * NP + 9 * C] = accb;
  }
  __syncthreads();
  for (int pair = threadIdx.x; pair < NP; pair += kHeadThreads) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < NW; ++q) a += red[q * NP + pair];
    partial[(size_t)blockIdx.x * NP + pair] = a;
  }
}

static bool head_grouped(int C) { return C == 8 || C == 16 || C == 32 || C == 64 || C == 128; }
static int head_tiles(int B, int H, int W, int* tx, int* ty) {
  *tx = (W + kTileW - 1) / kTileW;
  *ty = (H + kTileH - 1) / kTileH;
  return B * *tx * *ty;
}
static int head_w_blocks(int ntiles) { return ntiles < 148 * 2 ? ntiles : 148 * 2; }   // persistent over the tiles: 2 blocks per SM

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

extern "C" int dvs_disp_head_fwd(const void* x, int x_dtype, int x_pad, const float* weight, const float* bias, void* disp,
                                 int disp_dtype, int B, int C, int H, int W, void* stream) {
  int rc = head_check(x, x_dtype, B, C, H, W);
  if (rc) return rc;
  if (x_pad < 0 || x_pad > 1 || (x_pad && !head_grouped(C))) return DVS_EINVAL;
  if (!weight || !disp || (disp_dtype != DVS_DTYPE_F32 && disp_dtype != DVS_DTYPE_BF16)) return DVS_EINVAL;
  const size_t P = (size_t)B * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool xb = x_dtype == DVS_DTYPE_BF16, db = disp_dtype == DVS_DTYPE_BF16;
  if (head_grouped(C)) {
    int tx, ty;
    const int grid = head_tiles(B, H, W, &tx, &ty);
#define DVS_HEAD_FWD(LP)                                                                                                      \
  do {                                                                                                                        \
    if (xb && db) disp_head_fwd_tile_kernel<true, true, LP><<<grid, kHeadThreads, 0, st>>>(x, weight, bias, disp, H, W, tx, ty, x_pad);  \
    else if (xb) disp_head_fwd_tile_kernel<true, false, LP><<<grid, kHeadThreads, 0, st>>>(x, weight, bias, disp, H, W, tx, ty, x_pad);  \
    else if (db) disp_head_fwd_tile_kernel<false, true, LP><<<grid, kHeadThreads, 0, st>>>(x, weight, bias, disp, H, W, tx, ty, x_pad);  \
    else disp_head_fwd_tile_kernel<false, false, LP><<<grid, kHeadThreads, 0, st>>>(x, weight, bias, disp, H, W, tx, ty, x_pad);         \
  } while (0)
    switch (C / 8) {
      case 1: DVS_HEAD_FWD(1); break;
      case 2: DVS_HEAD_FWD(2); break;
      case 4: DVS_HEAD_FWD(4); break;
      case 8: DVS_HEAD_FWD(8); break;
      default: DVS_HEAD_FWD(16); break;
    }
#undef DVS_HEAD_FWD
    DVS_CUDA_TRY(cudaGetLastError());
    return DVS_OK;
  }
  int grid = (int)((P + kHeadThreads - 1) / kHeadThreads);
  if (grid > 148 * 32) grid = 148 * 32;
  const size_t smem = sizeof(float) * 9 * C;
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
  size_t nblk = (P + (size_t)kHeadThreads * chunks - 1) / ((size_t)kHeadThreads * chunks);
  if (head_grouped(C)) {
    int tx, ty;
    nblk = (size_t)head_w_blocks(head_tiles(B, H, W, &tx, &ty));
  }
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

extern "C" int dvs_disp_head_bwd(const void* grad_disp, const void* disp, int disp_dtype, const void* x, int x_dtype, int x_pad,
                                 const float* weight, void* grad_x, float* grad_weight, float* grad_bias, int B, int C, int H,
                                 int W, void* workspace, void* stream) {
  int rc = head_check(x, x_dtype, B, C, H, W);
  if (rc) return rc;
  if (x_pad < 0 || x_pad > 1 || (x_pad && !head_grouped(C))) return DVS_EINVAL;
  if (!grad_disp || !disp || !weight || !grad_x || !grad_weight) return DVS_EINVAL;
  if (disp_dtype != DVS_DTYPE_F32 && disp_dtype != DVS_DTYPE_BF16) return DVS_EINVAL;
  if (!workspace || ((uintptr_t)workspace & 255) || ((uintptr_t)grad_x & 15)) return DVS_EWORKSPACE;
  const size_t P = (size_t)B * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  const bool xb = x_dtype == DVS_DTYPE_BF16, gb = disp_dtype == DVS_DTYPE_BF16;
  const int npairs = 9 * C + 1;
  if (head_grouped(C)) {
    int tx, ty;
    const int ntiles = head_tiles(B, H, W, &tx, &ty), grid = head_w_blocks(ntiles);
    const size_t smem_w = sizeof(float) * (kHeadThreads / 32) * npairs;
#define DVS_HEAD_BWD(XB, GB, LP)                                                                                              \
  do {                                                                                                                        \
    disp_head_bwd_x_tile_kernel<GB, GB, XB, LP><<<grid, kHeadThreads, 0, st>>>(grad_disp, disp, weight, grad_x, H, W, tx, ty,     \
                                                                              ntiles, x_pad);                                       \
    disp_head_bwd_w_tile_kernel<XB, GB, GB, LP><<<grid, kHeadThreads, smem_w, st>>>(grad_disp, disp, x, partial, H, W, tx, ty,    \
                                                                                   ntiles, x_pad);                                  \
  } while (0)
#define DVS_HEAD_BWD_LP(LP)                                                                                                   \
  do {                                                                                                                        \
    if (xb && gb) DVS_HEAD_BWD(true, true, LP);                                                                               \
    else if (xb) DVS_HEAD_BWD(true, false, LP);                                                                               \
    else if (gb) DVS_HEAD_BWD(false, true, LP);                                                                               \
    else DVS_HEAD_BWD(false, false, LP);                                                                                      \
  } while (0)
    switch (C / 8) {
      case 1: DVS_HEAD_BWD_LP(1); break;
      case 2: DVS_HEAD_BWD_LP(2); break;
      case 4: DVS_HEAD_BWD_LP(4); break;
      case 8: DVS_HEAD_BWD_LP(8); break;
      default: DVS_HEAD_BWD_LP(16); break;
    }
#undef DVS_HEAD_BWD_LP
#undef DVS_HEAD_BWD
    DVS_CUDA_TRY(cudaGetLastError());
    disp_head_reduce_kernel<<<(npairs + 7) / 8, 256, 0, st>>>(partial, grid, npairs, grad_weight, grad_bias);
    DVS_CUDA_TRY(cudaGetLastError());
    return DVS_OK;
  }
  const int chunks = head_chunks(P);
  const int nblk = (int)((P + (size_t)kHeadThreads * chunks - 1) / ((size_t)kHeadThreads * chunks));
  const size_t smem = sizeof(float) * (9 * (size_t)C + kHeadThreads * 10 + (size_t)kHeadThreads * (C + 1));
  if (xb && gb) rc = head_bwd_launch<true, true>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else if (xb) rc = head_bwd_launch<true, false>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else if (gb) rc = head_bwd_launch<false, true>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  else rc = head_bwd_launch<false, false>(grad_disp, disp, x, weight, grad_x, partial, B, C, H, W, chunks, nblk, smem, st);
  if (rc) return rc;
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

// Activations with a reflected ring.  The decoder's convolutions are reflection-padded (model/layers.py:126-136); the kernels
// that PRODUCE their inputs can write the one-pixel ring themselves (pad = 1: tensor [B, H+2, W+2, C], ring row -1 = row 1,
// row H = row H-2, columns alike), so that the stock convolution runs un-padded on that buffer -- no padded copy and no border
// fix-up convolutions -- and the kernels that consume the convolution's input gradient fold the ring's gradients back.
// IDX: 32-bit index arithmetic when the tensors have fewer than 2^31 vectors (the 64-bit divisions of the pixel decomposition
// cost more than the memory access they address), 64-bit otherwise
template <class IDX>
__device__ __forceinline__ IDX pvec(IDX b, int y, int x, int H, int W, int pad, int vpp, int v) {
  return ((b * (IDX)(H + 2 * pad) + (IDX)(y + pad)) * (IDX)(W + 2 * pad) + (IDX)(x + pad)) * (IDX)vpp + (IDX)v;
}
// the rows (columns) that hold pixel row y: itself, and the ring rows that mirror it
__device__ __forceinline__ int ring_coords(int y, int H, int pad, int* ys) {
  int n = 0;
  ys[n++] = y;
  if (pad) {
    if (y == 1) ys[n++] = -1;
    if (y == H - 2) ys[n++] = H;
  }
  return n;
}
template <class IDX>
__device__ __forceinline__ void store_with_ring(uint4* out, uint4 val, IDX b, int y, int x, int H, int W, int pad, int vpp, int v) {
  int ys[3], xs[3];
  const int ny = ring_coords(y, H, pad, ys), nx = ring_coords(x, W, pad, xs);
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) out[pvec<IDX>(b, ys[i], xs[j], H, W, pad, vpp, v)] = val;
}
// gradient reaching pixel (y, x): its own entry plus those of the ring positions that mirror it (fixed order), in fp32
template <bool BF, class IDX>
__device__ __forceinline__ void load_folded(const uint4* g, IDX b, int y, int x, int H, int W, int pad, int vpp, int v, float* f) {
  constexpr int N = BF ? 8 : 4;
  int ys[3], xs[3];
  const int ny = ring_coords(y, H, pad, ys), nx = ring_coords(x, W, pad, xs);
#pragma unroll
  for (int k = 0; k < N; ++k) f[k] = 0.f;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) {
      Vec16<BF> q;
      q.raw = g[pvec<IDX>(b, ys[i], xs[j], H, W, pad, vpp, v)];
#pragma unroll
      for (int k = 0; k < N; ++k) f[k] += q.get(k);
    }
}

// Per-channel sums over the pixels (the bias gradients) of the glue kernels: every thread keeps its N channels' sums over a
// grid-stride walk whose stride is a multiple of the vectors per pixel (so the thread's channels never change), the block
// adds the threads that own the same channels in thread order, and channel_reduce_kernel adds the blocks in block order.
constexpr int kGlueThreads = 256;
template <int N>
__device__ __forceinline__ void block_channel_sums(const float* acc, int vpp, float* __restrict__ partial, int C) {
  __shared__ float sh[kGlueThreads * 8];
#pragma unroll
  for (int i = 0; i < N; ++i) sh[threadIdx.x * N + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kGlueThreads) {
    const int v = c / N, i = c - v * N;
    float a = 0.f;
    for (int t = v; t < kGlueThreads; t += vpp) a += sh[t * N + i];
    partial[(size_t)blockIdx.x * C + c] = a;
  }
}
__global__ void __launch_bounds__(256) channel_reduce_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float a = 0.f;
  for (int b = lane; b < nblk; b += 32) a += partial[(size_t)b * C + c];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[c] = a;
}

// one thread per 16-byte vector of `out` [B, 2h, 2w, C1 + C2] (+ ring if pad); bias (fp32 [C1], may be null) is added before
// the ELU
template <bool BF, class IDX>
__global__ void __launch_bounds__(256) elu_up2_cat_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip,
                                                              const float* __restrict__ bias, uint4* __restrict__ out, int h, int w,
                                                              int v1, int v2, int pad, IDX nvec) {
  const IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  const int vt = v1 + v2;
  const IDX pix = e / vt;
  const int v = (int)(e - pix * vt);
  const int W2 = 2 * w, H2 = 2 * h;
  const int X = (int)(pix % W2);
  const IDX r = pix / W2;
  const int Y = (int)(r % H2);
  const IDX b = r / H2;
  uint4 val;
  if (v >= v1) {
    val = skip[pix * v2 + (v - v1)];
  } else {
    constexpr int N = Vec16<BF>::N;
    Vec16<BF> a;
    a.raw = x[((b * h + (Y >> 1)) * w + (X >> 1)) * v1 + v];
    float f[N];
    for (int i = 0; i < N; ++i) f[i] = elu_value(a.get(i) + (bias ? bias[v * N + i] : 0.f));
    Vec16<BF> o;
    o.set_all(f);
    val = o.raw;
  }
  store_with_ring<IDX>(out, val, b, Y, X, H2, W2, pad, vt, v);
}

// grad_x [B, h, w, C1]: 2x2 sum of grad_out (fixed order; ring gradients folded in if pad) times ELU'(x + bias); grid-stride,
// per-channel sums of grad_x for the bias gradient (partial != null)
template <bool BF, class IDX>
__global__ void __launch_bounds__(kGlueThreads) elu_up2_cat_bwd_x_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gout,
                                                                         const float* __restrict__ bias, uint4* __restrict__ gx,
                                                                         float* __restrict__ partial, int h, int w, int v1, int v2,
                                                                         int pad, IDX nvec) {
  constexpr int N = Vec16<BF>::N;
  const int vt = v1 + v2, W2 = 2 * w, H2 = 2 * h;
  float acc[N], bv[N];
  const int v = threadIdx.x % v1;                             // constant along the walk: the stride is a multiple of v1
#pragma unroll
  for (int i = 0; i < N; ++i) { acc[i] = 0.f; bv[i] = bias ? bias[v * N + i] : 0.f; }
  for (IDX e = (IDX)blockIdx.x * kGlueThreads + threadIdx.x; e < nvec; e += (IDX)gridDim.x * kGlueThreads) {
    const IDX pix = e / v1;
    const int xx = (int)(pix % w);
    const IDX r = pix / w;
    const int yy = (int)(r % h);
    const IDX b = r / h;
    float g00[N], g01[N], g10[N], g11[N];
    load_folded<BF, IDX>(gout, b, 2 * yy, 2 * xx, H2, W2, pad, vt, v, g00);
    load_folded<BF, IDX>(gout, b, 2 * yy, 2 * xx + 1, H2, W2, pad, vt, v, g01);
    load_folded<BF, IDX>(gout, b, 2 * yy + 1, 2 * xx, H2, W2, pad, vt, v, g10);
    load_folded<BF, IDX>(gout, b, 2 * yy + 1, 2 * xx + 1, H2, W2, pad, vt, v, g11);
    Vec16<BF> a;
    a.raw = x[e];
    float f[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      f[i] = ((g00[i] + g01[i]) + (g10[i] + g11[i])) * elu_slope(a.get(i) + bv[i]);
      acc[i] += f[i];
    }
    Vec16<BF> o;
    o.set_all(f);
    gx[e] = o.raw;
  }
  if (partial) block_channel_sums<N>(acc, v1, partial, v1 * N);
}

// one thread per 16-byte vector of grad_skip [B, 2h, 2w, C2]: the channel slice of grad_out (ring gradients folded in if pad)
template <bool BF, class IDX>
__global__ void __launch_bounds__(256) cat_bwd_skip_kernel(const uint4* __restrict__ gout, uint4* __restrict__ gskip, int H2, int W2,
                                                           int v1, int v2, int pad, IDX nvec) {
  const IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  const IDX pix = e / v2;
  const int v = (int)(e - pix * v2);
  if (!pad) {
    gskip[e] = gout[pix * (v1 + v2) + v1 + v];
    return;
  }
  const int X = (int)(pix % W2);
  const IDX r = pix / W2;
  const int Y = (int)(r % H2);
  const IDX b = r / H2;
  float f[Vec16<BF>::N];
  load_folded<BF, IDX>(gout, b, Y, X, H2, W2, pad, v1 + v2, v1 + v, f);
  Vec16<BF> o;
  o.set_all(f);
  gskip[e] = o.raw;
}

// y = ELU(x + bias[c]) (ConvBlock: the convolution's bias and its ELU, model/layers.py:106-117); y gets the reflected ring if
// pad (then y is [B, H+2, W+2, C]); without pad y may alias x
template <bool BF, class IDX>
__global__ void __launch_bounds__(256) bias_elu_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ bias,
                                                           uint4* __restrict__ y, int H, int W, int vpp, int pad, IDX nvec) {
  const IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  constexpr int N = Vec16<BF>::N;
  const IDX pix = e / vpp;
  const int v = (int)(e - pix * vpp);
  Vec16<BF> a;
  a.raw = x[e];
  float f[N];
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = elu_value(a.get(i) + (bias ? bias[v * N + i] : 0.f));
  Vec16<BF> o;
  o.set_all(f);
  if (!pad) {
    y[e] = o.raw;
    return;
  }
  const int xx = (int)(pix % W);
  const IDX r = pix / W;
  store_with_ring<IDX>(y, o.raw, r / H, (int)(r % H), xx, H, W, pad, vpp, v);
}
// grad_x = grad_y * ELU'(.) written from the OUTPUT y as ATen's in-place ELU does (y <= 0 ? y + 1 : 1: nn.ELU(inplace=True)
// keeps only y); ring gradients folded in if pad; per-channel sums of grad_x for the bias gradient
template <bool BF, class IDX>
__global__ void __launch_bounds__(kGlueThreads) bias_elu_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ gy,
                                                                    uint4* __restrict__ gx, float* __restrict__ partial, int H, int W,
                                                                    int vpp, int pad, IDX nvec) {
  constexpr int N = Vec16<BF>::N;
  float acc[N];
  const int v = threadIdx.x % vpp;
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  for (IDX e = (IDX)blockIdx.x * kGlueThreads + threadIdx.x; e < nvec; e += (IDX)gridDim.x * kGlueThreads) {
    const IDX pix = e / vpp;
    const int xx = (int)(pix % W);
    const IDX r = pix / W;
    const int yy = (int)(r % H);
    const IDX b = r / H;
    Vec16<BF> a;
    a.raw = y[pvec<IDX>(b, yy, xx, H, W, pad, vpp, v)];
    float g[N], f[N];
    load_folded<BF, IDX>(gy, b, yy, xx, H, W, pad, vpp, v, g);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float yv = a.get(i);
      f[i] = g[i] * (yv <= 0.f ? yv + 1.f : 1.f);
      acc[i] += f[i];
    }
    Vec16<BF> o;
    o.set_all(f);
    gx[e] = o.raw;
  }
  if (partial) block_channel_sums<N>(acc, vpp, partial, vpp * N);
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
// grid of the reducing kernels: the walk's stride (grid x 256) must be a multiple of the vectors per pixel (a power of two
// <= 256 here, else the caller is refused), and at most 8 blocks per SM
static bool glue_small(size_t padded_vectors) { return padded_vectors + (size_t)148 * 8 * 256 < ((size_t)1 << 31); }   // incl. the grid-stride overshoot
static int glue_reduce_blocks(size_t nvec, int vpp) {
  if (vpp < 1 || vpp > kGlueThreads || (kGlueThreads % vpp)) return 0;
  size_t n = (nvec + kGlueThreads - 1) / kGlueThreads;
  return (int)(n < 1 ? 1 : (n > 148 * 8 ? 148 * 8 : n));
}

}  // namespace dvs

extern "C" int dvs_glue_workspace_bytes(int C, size_t* bytes) {
  if (!bytes || C < 1) return DVS_EINVAL;
  *bytes = sizeof(float) * (size_t)148 * 8 * C + 256;
  return DVS_OK;
}

extern "C" int dvs_elu_up2_cat_fwd(const void* x, const void* skip, const float* bias, void* out, int dtype, int B, int C1, int C2,
                                   int h, int w, int pad, void* stream) {
  using namespace dvs;
  int rc = glue_check(x, dtype, B, C1, C2, h, w, skip);
  if (rc) return rc;
  if (!out || ((uintptr_t)out & 15) || pad < 0 || pad > 1 || (pad && (2 * h < 3 || 2 * w < 3))) return DVS_EINVAL;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int per = bf ? 8 : 4, v1 = C1 / per, v2 = C2 / per;
  const size_t nvec = (size_t)B * 4 * h * w * (v1 + v2);
  const unsigned int nblk = (unsigned int)((nvec + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define DVS_GLUE_LAUNCH(K, GRID, BLOCK, ...)                                                                     \
  do {                                                                                                          \
    if (small && bf) K<true, unsigned int><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                                 \
    else if (small) K<false, unsigned int><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                                 \
    else if (bf) K<true, size_t><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                                           \
    else K<false, size_t><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__);                                                  \
  } while (0)
  const bool small = glue_small((size_t)B * (2 * h + 2) * (2 * w + 2) * (v1 + v2));
  DVS_GLUE_LAUNCH(elu_up2_cat_fwd_kernel, nblk, 256, (const uint4*)x, (const uint4*)skip, bias, (uint4*)out, h, w, v1, v2, pad, nvec);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

extern "C" int dvs_elu_up2_cat_bwd(const void* x, const void* grad_out, const float* bias, void* grad_x, void* grad_skip,
                                   float* grad_bias, int dtype, int B, int C1, int C2, int h, int w, int pad, void* workspace,
                                   void* stream) {
  using namespace dvs;
  int rc = glue_check(x, dtype, B, C1, C2, h, w, C2 > 0 ? grad_skip : nullptr);
  if (rc) return rc;
  if (!grad_out || !grad_x || ((uintptr_t)grad_out & 15) || ((uintptr_t)grad_x & 15) || pad < 0 || pad > 1) return DVS_EINVAL;
  if (grad_bias && (!workspace || ((uintptr_t)workspace & 255))) return DVS_EWORKSPACE;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int per = bf ? 8 : 4, v1 = C1 / per, v2 = C2 / per;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nx = (size_t)B * h * w * v1;
  const int bx = glue_reduce_blocks(nx, v1);
  if (!bx) return DVS_EINVAL;
  float* partial = grad_bias ? static_cast<float*>(workspace) : nullptr;
  const bool small = glue_small((size_t)B * (2 * h + 2) * (2 * w + 2) * (v1 + v2));
  DVS_GLUE_LAUNCH(elu_up2_cat_bwd_x_kernel, bx, kGlueThreads, (const uint4*)x, (const uint4*)grad_out, bias, (uint4*)grad_x, partial, h, w, v1, v2, pad, nx);
  DVS_CUDA_TRY(cudaGetLastError());
  if (grad_bias) {
    channel_reduce_kernel<<<(C1 + 7) / 8, 256, 0, st>>>(partial, bx, C1, grad_bias);
    DVS_CUDA_TRY(cudaGetLastError());
  }
  if (v2 > 0) {
    const size_t ns = (size_t)B * 4 * h * w * v2;
    const unsigned int nb = (unsigned int)((ns + 255) / 256);
    DVS_GLUE_LAUNCH(cat_bwd_skip_kernel, nb, 256, (const uint4*)grad_out, (uint4*)grad_skip, 2 * h, 2 * w, v1, v2, pad, ns);
    DVS_CUDA_TRY(cudaGetLastError());
  }
  return DVS_OK;
}

extern "C" int dvs_bias_elu_fwd(const void* x, const float* bias, void* y, int dtype, int B, int C, int H, int W, int pad,
                                void* stream) {
  using namespace dvs;
  int rc = glue_check(x, dtype, B, C, 0, H, W, nullptr);
  if (rc) return rc;
  if (!y || ((uintptr_t)y & 15) || pad < 0 || pad > 1 || (pad && (H < 3 || W < 3 || y == x))) return DVS_EINVAL;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int vpp = C / (bf ? 8 : 4);
  const size_t nvec = (size_t)B * H * W * vpp;
  const unsigned int nblk = (unsigned int)((nvec + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool small = glue_small((size_t)B * (H + 2) * (W + 2) * vpp);
  DVS_GLUE_LAUNCH(bias_elu_fwd_kernel, nblk, 256, (const uint4*)x, bias, (uint4*)y, H, W, vpp, pad, nvec);
  DVS_CUDA_TRY(cudaGetLastError());
  return DVS_OK;
}

extern "C" int dvs_bias_elu_bwd(const void* y, const void* grad_y, void* grad_x, float* grad_bias, int dtype, int B, int C, int H,
                                int W, int pad, void* workspace, void* stream) {
  using namespace dvs;
  int rc = glue_check(y, dtype, B, C, 0, H, W, nullptr);
  if (rc) return rc;
  if (!grad_y || !grad_x || ((uintptr_t)grad_y & 15) || ((uintptr_t)grad_x & 15) || pad < 0 || pad > 1) return DVS_EINVAL;
  if (grad_bias && (!workspace || ((uintptr_t)workspace & 255))) return DVS_EWORKSPACE;
  const bool bf = dtype == DVS_DTYPE_BF16;
  const int vpp = C / (bf ? 8 : 4);
  const size_t nvec = (size_t)B * H * W * vpp;
  const int nblk = glue_reduce_blocks(nvec, vpp);
  if (!nblk) return DVS_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = grad_bias ? static_cast<float*>(workspace) : nullptr;
  const bool small = glue_small((size_t)B * (H + 2) * (W + 2) * vpp);
  DVS_GLUE_LAUNCH(bias_elu_bwd_kernel, nblk, kGlueThreads, (const uint4*)y, (const uint4*)grad_y, (uint4*)grad_x, partial, H, W, vpp, pad, nvec);
  DVS_CUDA_TRY(cudaGetLastError());
  if (grad_bias) {
    channel_reduce_kernel<<<(C + 7) / 8, 256, 0, st>>>(partial, nblk, C, grad_bias);
    DVS_CUDA_TRY(cudaGetLastError());
  }
  return DVS_OK;
}
