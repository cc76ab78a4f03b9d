// Fused view-synthesis loss: per-tile phase functions (v2).
//
// One CTA owns a 30x30 block of target pixels of one batch item ("R0") and evaluates SSIM statistics on
// the surrounding 32x32 block ("R1", the window centres whose 3x3 adjoint reaches R0) from warped
// colours on the 34x34 block ("R2").  It walks all S scales with the target / source tiles resident in
// shared memory.  A thread owns four vertically adjacent pixels of one R1 column (lane == column), so
// every shared-memory access of a warp is one conflict-free row segment.
// The phases are written as plain functions of (tid, shared memory, per-thread state) so that the same
// source compiles
//   * under nvcc into the sm_100a kernel in dvs_fused.cu (phases separated by __syncthreads()), and
//   * under g++ into a sequential block emulator used by the CPU-side unit tests (tests/emu),
// which lets the tile logic be checked against the oracle without a GPU.
//
// The kernel is instruction-issue bound (about 40 B of compulsory traffic against several hundred fp32
// operations per warped pixel), so the arithmetic is arranged for few instructions rather than few bytes:
//   * projection  c = D * (A_i (u,v,1)) + p_i  with A_i = (K T_i)[:, :3] inv_K hoisted per tile
//   * bilinear taps as lerps (value and both slopes share the differences)
//   * SSIM on 9-sums with the 1/9 and 1/81 factors folded into the constants, separable 3x3 sums shared
//     between the four pixels of a thread, target-side sums shared between the sources
//   * pose gradient accumulated as 12 moments  sum g_c * D * (u, v, 1), sum g_c  (mapped to dL/dT in the
//     finish kernel)
//
// Reference arithmetic being replaced (see oracle/closed_form.py for the explicit formulas):
//   vo/learner_new.py:136-170   up-sample, disp_to_depth, BackprojectDepth, Project3D, border grid_sample
//   vo/learner_new.py:60-74     SSIM + L1 (vo/learner_func.py:177-207)
//   vo/learner_new.py:199-257   identity automask + noise, min over sources, smoothness, scale sum
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DVS_HD __host__ __device__ __forceinline__
#define DVS_UNROLL _Pragma("unroll")
#define DVS_NOUNROLL _Pragma("unroll 1")
#else
#define DVS_HD inline
#define DVS_UNROLL
#define DVS_NOUNROLL
#endif

namespace dvs {

constexpr int kMaxS = 4;
constexpr int kMaxN = 4;
constexpr int TW = 32;                 // R1 width  (one lane per column)
#if !defined(DVS_TILE_ROWS)
#define DVS_TILE_ROWS 32
#endif
constexpr int TH = DVS_TILE_ROWS;      // R1 height (a multiple of 4: TH / 4 warps x 4 rows)
constexpr int PITCH_X = TW - 2;        // R0 width  = tile pitch in x
constexpr int PITCH_Y = TH - 2;        // R0 height = tile pitch in y
constexpr int RW2 = TW + 2;            // R2 width
constexpr int RH2 = TH + 2;            // R2 height
constexpr int PW = RW2;                // plane row stride (floats)
constexpr int PLANE = RH2 * PW;        // 1156 floats; plane index of R2 pixel (ly,lx) is (ly+1)*PW + lx+1
constexpr int kWarps = TH / 4;
constexpr int NT = 32 * kWarps;        // threads per CTA
constexpr int kMeanBlocks = 16;        // partial sums per (scale, batch item) in the disparity-mean pre-pass
constexpr float kC1 = 0.0001f, kC2 = 0.0009f;
constexpr float kK1 = 81.0f * 0.0001f, kK2 = 81.0f * 0.0009f;   // SSIM constants on 9-sums
constexpr int kSelNone = 255;
constexpr int kTbufCols = 36;          // >= max coarse columns touched by 30 fine columns (+ slack)
constexpr int kTbufFloats = TH * kTbufCols;   // one row per fine row of R0 (rows = TH - 2)

// ------------------------------------------------------------------------------------------------
struct FusedParams {
  int B, H, W, N, S;
  int dh[kMaxS], dw[kMaxS];
  const float* disp[kMaxS];
  const float* target;
  const float* src[kMaxN];
  const float* K;
  const float* invK;
  const float* T[kMaxN];
  const float* noise[kMaxS];       // null -> hash generator
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;   // optional device counter added to `offset` (advances under CUDA-graph replay)
  float min_disp, disp_range;      // scaled = min_disp + disp_range*disp
  float ssim_w, l1_w, smooth_w, eps;
  int auto_mask;
  int want_grad;
  // outputs
  float* gdisp[kMaxS];             // unit gradients d loss/s / d disp[s] (written directly when full resolution,
                                   // else gathered from cpart by gather_gdisp in a fixed order)
  unsigned char* sel[kMaxS];       // optional argmin maps
  // workspace
  const float* mean_part;          // [S][B][kMeanBlocks] partial sums of the up-sampled disparity
  float* part;                     // [nblk][S][3 + 12N] per-block partial sums
  // up-sampled scales: every tile leaves the adjoint of the bilinear up-sample restricted to its own pixels as a small
  // coarse box (rows i0..i1 x columns j0..j1 of disp[s], row pitch cbw[s]) at cpart[blk * cstride + coff[s]]
  float* cpart;
  int cstride, coff[kMaxS], cbw[kMaxS];
  int tiles_x, tiles_y;
  int nblk;
  // input formats (two-source kernel, dvs_pair_core.cuh): bit 0 = disparities are bf16, bit 1 = images are uint8; the
  // pointers above are then to be read as unsigned short / unsigned char with the same element offsets
  int io_flags;
  // loop-invariant scalars of the two-source kernel, divided once on the host (IEEE, same values as on the device):
  // kF = ssim_w / (3 B H W), l1k = l1_w / (3 B H W), kxs[s] / kys[s] = smooth_w / 2^s / (B H (W-1)) resp. (B (H-1) W)
  float kF, l1k, kxs[kMaxS], kys[kMaxS];
};

// Upper bound of the coarse rows (columns) touched by the `fine` fine rows (columns) of one tile: the clamped source
// coordinate spans (fine - 1) * in/out, plus the two taps.
DVS_HD int coarse_box_extent(int in, int out, int fine) { int e = ((fine - 1) * in) / out + 3; return e < in ? e : in; }

DVS_HD unsigned long long noise_offset(const FusedParams& p) { return p.offset + (p.offset_dev ? *p.offset_dev : 0ull); }

// per-block partial sums: [0] photometric, [1] smooth-x, [2] smooth-y, then per source 12 pose moments
//   M[r*4 + 0..2] = sum g_c[r] * D * (u, v, 1),  M[r*4 + 3] = sum g_c[r]      (r = row of the 3x4 projection)
DVS_HD int nvals(int N) { return 3 + 12 * N; }

// ------------------------------------------------------------------------------------------------ math
DVS_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
DVS_HD float exp_fast(float x) {
#if defined(__CUDA_ARCH__)
  return __expf(x);
#else
  return expf(x);
#endif
}
DVS_HD float sat01(float x) {
#if defined(__CUDA_ARCH__)
  return __saturatef(x);
#else
  return x != x ? 0.f : fminf(fmaxf(x, 0.f), 1.f);
#endif
}
// Packed pairs of fp32 (Blackwell FADD2 / FMUL2 / FFMA2: two lanes per issued instruction; same lane throughput as
// the scalar forms, half the issue slots -- profiles/tools/ffma2_bench.cu).  The kernel is issue bound, so the row sums
// and the per-pixel SSIM algebra are written on pairs.  Host build (block emulator): plain scalar arithmetic.
struct f2 {
  float x, y;
};
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long f2_pk(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 f2_up(unsigned long long r) {
  f2 a;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_up(r);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_up(r);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_up(r);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)), "l"(f2_pk(c)));
  return f2_up(r);
}
#else
inline f2 add2(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
inline f2 sub2(f2 a, f2 b) { return f2{a.x - b.x, a.y - b.y}; }
inline f2 mul2(f2 a, f2 b) { return f2{a.x * b.x, a.y * b.y}; }
inline f2 fma2(f2 a, f2 b, f2 c) { return f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif

DVS_HD float sgn(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }
DVS_HD int imin(int a, int b) { return a < b ? a : b; }
DVS_HD int imax(int a, int b) { return a > b ? a : b; }

DVS_HD int reflect_clamp(int g, int n) {
  g = g < 0 ? -g : g;
  g = g >= n ? 2 * (n - 1) - g : g;
  return imin(imax(g, 0), n - 1);
}
DVS_HD int pidx(int ly, int lx) { return (ly + 1) * PW + (lx + 1); }

// ATen bilinear (align_corners=False) source taps for one output coordinate (UpSample.h:259-311).
DVS_HD void up_taps(int dst, float scale, int in, int& i0, int& i1, float& lam) {
  float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = imin((int)s, in - 1);
  i1 = imin(i0 + 1, in - 1);
  lam = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
}
// Weight of coarse index `coarse` in the up-sampled value at fine index `fine`: the taps above are the hat
// function around the clamped source coordinate.
DVS_HD float tap_weight(int fine, float scale, int in, int coarse) {
  float s = fminf(fmaxf(scale * ((float)fine + 0.5f) - 0.5f, 0.f), (float)(in - 1));
  return fmaxf(1.f - fabsf(s - (float)coarse), 0.f);
}
// Total weight that coarse index i receives from all `out` fine positions (== out/in for exact ratios).
DVS_HD float up_weight(int i, int in, int out) {
  if (out % in == 0) return (float)(out / in);
  float scale = (float)in / (float)out, w = 0.f;
  int lo = imax((int)(((float)i - 1.f) / scale) - 2, 0);
  int hi = imin((int)(((float)i + 2.f) / scale) + 2, out - 1);
  for (int o = lo; o <= hi; ++o) {
    int a, b;
    float l;
    up_taps(o, scale, in, a, b, l);
    if (a == i) w += 1.f - l;
    if (b == i) w += l;
  }
  return w;
}

// Counter-based standard normals for the automask tie-break when no noise tensor is given: one Box-Muller
// pair per (pixel, scale, source pair).
DVS_HD unsigned int hash32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
DVS_HD void hash_normal2(unsigned long long seed, unsigned long long offset, unsigned int idx, unsigned int stream,
                         float& n0, float& n1) {
  unsigned int k = (unsigned int)(seed ^ (seed >> 32)) + 0x9e3779b9U * (unsigned int)(offset + stream);
  unsigned int h1 = hash32(idx ^ k);
  unsigned int h2 = hash32(h1 + 0x68bc21ebU + stream);
  float u1 = ((float)(h1 >> 8) + 1.0f) * (1.0f / 16777217.0f);   // (0,1)
  float u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);
#if defined(__CUDA_ARCH__)
  float r = sqrtf(-2.0f * __logf(u1)), sn, cs;
  __sincosf(6.28318530718f * u2, &sn, &cs);
#else
  float r = sqrtf(-2.0f * logf(u1)), sn = sinf(6.28318530718f * u2), cs = cosf(6.28318530718f * u2);
#endif
  n0 = r * cs;
  n1 = r * sn;
}

// ------------------------------------------------------------------------------------------------ shared memory
// planes: Y[3] | X[3N] | F[9] | DU | WX | WY | POS (ints) | SEL(bytes, PLANE of them) | consts
struct SmemLayout {
  int N;
  DVS_HD int y(int c) const { return c * PLANE; }
  DVS_HD int x(int i, int c) const { return (3 + i * 3 + c) * PLANE; }
  DVS_HD int f(int k) const { return (3 + 3 * N + k) * PLANE; }
  DVS_HD int du() const { return (12 + 3 * N) * PLANE; }
  DVS_HD int wx() const { return (13 + 3 * N) * PLANE; }
  DVS_HD int wy() const { return (14 + 3 * N) * PLANE; }
  DVS_HD int pos() const { return (15 + 3 * N) * PLANE; }            // (ry << 16) | rx of every R2 pixel
  DVS_HD int sel() const { return (16 + 3 * N) * PLANE; }            // PLANE bytes = PLANE/4 floats
  DVS_HD int consts() const { return sel() + PLANE / 4; }
  DVS_HD int total() const { return consts() + 8 + 12 * kMaxN + 8 * kMaxS; }
  // scratch for block reductions / up-sample adjoint: aliases X (and the head of F) once those are dead
  DVS_HD int scratch() const { return x(0, 0); }
  DVS_HD int tbuf() const { return f(9) - kTbufFloats; }              // the tail of F
  DVS_HD int rbuf() const { return f(9) - kTbufFloats - 512; }        // 512 floats before it
};
// consts block: [0..3] inv_mu[s]; [8+12i ..] A_i (3x3 row-major) then p_i (3); then per scale 8 ints for the up-sample
// adjoint: coarse box i0, i1, j0, j1 of the tile and the integer ratios H/dh, W/dw (0 when not integer)
constexpr int kC_invmu = 0, kC_A = 8, kC_adj = 8 + 12 * kMaxN;

template <int NS>
struct ThreadState {
  float ident[NS][4];      // identity reprojection loss of the own pixels (scale independent)
  int flags;               // bit j: pixel j inside the image; bit 4+j: pixel j belongs to R0 (own)
  float acc[3];            // photometric sum, smooth-x sum, smooth-y sum of the current scale
  float dM[NS][12];        // pose-gradient moments of the current scale
  float gdu[4];            // d loss / d disp_up of the own pixels, current scale
  int tags;                // selected source of the 4 pixels, one byte each (kSelNone: identity / outside)
};

struct Tile {
  int b, gx0, gy0;         // batch item; global coords of R1 (0,0)
  int blk;                 // linear block id
};

DVS_HD Tile make_tile(const FusedParams& p, int blk) {
  Tile t;
  t.blk = blk;
  int per = p.tiles_x * p.tiles_y;
  t.b = blk / per;
  int r = blk - t.b * per;
  int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
  t.gx0 = tx * PITCH_X - 1;
  t.gy0 = ty * PITCH_Y - 1;
  return t;
}

// thread -> pixels: column cx = lane, rows r0 .. r0+3 of R1
DVS_HD void quad_coords(int tid, int& r0, int& cx) { r0 = (tid >> 5) << 2; cx = tid & 31; }

struct CoarseBox {
  int i0, i1, j0, j1;   // inclusive coarse ranges touched by the tile
  int fy0, fy1, fx0, fx1;  // fine ranges of R0 clipped to the image (inclusive)
};
DVS_HD CoarseBox coarse_box(const FusedParams& p, const Tile& t, int s) {
  CoarseBox c;
  c.fy0 = t.gy0 + 1; c.fy1 = imin(t.gy0 + TH - 2, p.H - 1);
  c.fx0 = t.gx0 + 1; c.fx1 = imin(t.gx0 + TW - 2, p.W - 1);
  int a, b;
  float l;
  up_taps(c.fy0, (float)p.dh[s] / (float)p.H, p.dh[s], a, b, l); c.i0 = a;
  up_taps(c.fy1, (float)p.dh[s] / (float)p.H, p.dh[s], a, b, l); c.i1 = b;
  up_taps(c.fx0, (float)p.dw[s] / (float)p.W, p.dw[s], a, b, l); c.j0 = a;
  up_taps(c.fx1, (float)p.dw[s] / (float)p.W, p.dw[s], a, b, l); c.j1 = b;
  return c;
}

// ------------------------------------------------------------------------------------------------ phase 0
// constants of the tile: A_i = (K T_i)[:3,:3] inv_K[:3,:3], p_i = (K T_i)[:3,3], 1/(clamp(mean disp)+1e-7) per scale.
// `a2` (two-source kernel): additionally the same numbers interleaved per source, a2[2 e + i] = (A_i | p_i)[e].
template <int NS>
DVS_HD void phase_consts_at(const FusedParams& p, const Tile& t, float* c, int tid, float* a2);
template <int NS>
DVS_HD void phase_consts(const FusedParams& p, const Tile& t, float* sm, int tid, float* a2 = nullptr) {
  SmemLayout L{NS};
  phase_consts_at<NS>(p, t, sm + L.consts(), tid, a2);
}
// `c`: the constants block of the tile (any layout)
template <int NS>
DVS_HD void phase_consts_at(const FusedParams& p, const Tile& t, float* c, int tid, float* a2) {
  if (tid < 12 * NS) {
    int i = tid / 12, e = tid - i * 12;
    const float* Kb = p.K + t.b * 16;
    const float* Tb = p.T[i] + t.b * 16;
    const float* iK = p.invK + t.b * 16;
    // a few hundred double-precision operations per tile: keeps A within one fp32 rounding of the exact product,
    // so the hoisting adds no coordinate error beyond the reference's own fp32 evaluation
    if (e < 9) {
      int r = e / 3, k = e - r * 3;
      double a = 0.0;
      for (int m = 0; m < 3; ++m) {
        double P = 0.0;                                    // (K T)[r][m]
        for (int n = 0; n < 4; ++n) P += (double)Kb[r * 4 + n] * (double)Tb[n * 4 + m];
        a += P * (double)iK[m * 4 + k];
      }
      c[kC_A + 12 * i + e] = (float)a;
      if (a2) a2[2 * e + i] = (float)a;
    } else {
      int r = e - 9;
      double P = 0.0;
      for (int n = 0; n < 4; ++n) P += (double)Kb[r * 4 + n] * (double)Tb[n * 4 + 3];
      c[kC_A + 12 * i + e] = (float)P;
      if (a2) a2[2 * e + i] = (float)P;
    }
  } else if (tid >= 64 && tid < 64 + p.S) {
    int s = tid - 64;
    const float* mp = p.mean_part + (s * p.B + t.b) * kMeanBlocks;
    float m = 0.f;
    for (int k = 0; k < kMeanBlocks; ++k) m += mp[k];
    m = m / ((float)p.H * (float)p.W);
    c[kC_invmu + s] = 1.0f / (fmaxf(m, 0.001f) + 1e-7f);
  } else if (tid >= 96 && tid < 96 + p.S) {
    int s = tid - 96;
    int* a = reinterpret_cast<int*>(c + kC_adj + 8 * s);
    CoarseBox cb = coarse_box(p, t, s);
    a[0] = cb.i0; a[1] = cb.i1; a[2] = cb.j0; a[3] = cb.j1;
    a[4] = (p.H % p.dh[s] == 0) ? p.H / p.dh[s] : 0;
    a[5] = (p.W % p.dw[s] == 0) ? p.W / p.dw[s] : 0;
  }
}

// load target (-> Y) and sources (-> X, for the identity terms) on R2 with reflection; zero F; per-thread pixel flags.
template <int NS>
DVS_HD void phase_load(const FusedParams& p, const Tile& t, float* sm, int tid, ThreadState<NS>& st) {
  SmemLayout L{NS};
  const int HW = p.H * p.W;
  int* posp = reinterpret_cast<int*>(sm + L.pos());
  DVS_NOUNROLL
  for (int k = tid; k < PLANE; k += NT) {
    int ly = k / PW - 1, lx = k % PW - 1;
    int gy = reflect_clamp(t.gy0 + ly, p.H), gx = reflect_clamp(t.gx0 + lx, p.W);
    posp[k] = (gy << 16) | gx;           // reflected image coordinates, re-used by the warp phase of every scale
    int o = gy * p.W + gx;
    const float* tg = p.target + (size_t)t.b * 3 * HW + o;
    sm[L.y(0) + k] = tg[0];
    sm[L.y(1) + k] = tg[HW];
    sm[L.y(2) + k] = tg[2 * HW];
    if (p.auto_mask) {
      DVS_UNROLL
      for (int i = 0; i < NS; ++i) {
        const float* sg = p.src[i] + (size_t)t.b * 3 * HW + o;
        sm[L.x(i, 0) + k] = sg[0];
        sm[L.x(i, 1) + k] = sg[HW];
        sm[L.x(i, 2) + k] = sg[2 * HW];
      }
    }
  }
#if defined(__CUDA_ARCH__)
  float4* f4 = reinterpret_cast<float4*>(sm + L.f(0));          // 9 planes of 1156 floats: 16-byte aligned, 2601 float4
  for (int k = tid; k < 9 * PLANE / 4; k += NT) f4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#else
  for (int k = tid; k < 9 * PLANE; k += NT) sm[L.f(0) + k] = 0.f;
#endif
  for (int k = tid; k < PLANE / 4; k += NT) reinterpret_cast<unsigned int*>(sm + L.sel())[k] = 0xffffffffu;   // kSelNone
  int r0, cx;
  quad_coords(tid, r0, cx);
  int fl = 0;
  const int gx = t.gx0 + cx;
  for (int j = 0; j < 4; ++j) {
    int gy = t.gy0 + r0 + j;
    bool in = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    bool own = in && (r0 + j) >= 1 && (r0 + j) <= TH - 2 && cx >= 1 && cx <= TW - 2;
    fl |= (in ? 1 : 0) << j;
    fl |= (own ? 1 : 0) << (4 + j);
  }
  st.flags = fl;
}

// ------------------------------------------------------------------------------------------------ 3x3 sums, SSIM
// Neighbourhood of the four pixels: rows m = 0..5 <-> R1 rows r0-1 .. r0+4, columns k = 0..2 <-> cx-1 .. cx+1.
struct YN {
  f2 vp[3][3];             // target values: vp[m2][k] = rows (2 m2, 2 m2 + 1), column k
  float sy[4];             // 9-sum of y
  float ysq[4];            // sy^2
  float ty[4];             // 9*sum(y^2) + 81*C2
};
DVS_HD float yn_at(const YN& q, int m, int k) { return (m & 1) ? q.vp[m >> 1][k].y : q.vp[m >> 1][k].x; }
struct XS {
  float sx[4], sxx[4], sxy[4], xc[4];
};
// vertical 3-sums of 6 row values (three row pairs) for the 4 pixels (two shared partial sums)
DVS_HD void vsum4(const f2* h, float* s) {
  float u12 = h[0].y + h[1].x, u34 = h[1].y + h[2].x;
  s[0] = h[0].x + u12;
  s[1] = u12 + h[1].y;
  s[2] = h[1].x + u34;
  s[3] = u34 + h[2].y;
}
DVS_HD void load_yn(const float* Y, int base, YN& q) {
  f2 hy[3], hyy[3];
  for (int m2 = 0; m2 < 3; ++m2) {
    const float* r0 = Y + base + (2 * m2 - 1) * PW;
    const float* r1 = r0 + PW;
    f2 a{r0[-1], r1[-1]}, b{r0[0], r1[0]}, c{r0[1], r1[1]};
    q.vp[m2][0] = a; q.vp[m2][1] = b; q.vp[m2][2] = c;
    hy[m2] = add2(add2(a, b), c);
    hyy[m2] = fma2(c, c, fma2(b, b, mul2(a, a)));
  }
  float syy[4];
  vsum4(hy, q.sy);
  vsum4(hyy, syy);
  for (int j = 0; j < 4; ++j) {
    q.ysq[j] = q.sy[j] * q.sy[j];
    q.ty[j] = fmaf(9.f, syy[j], kK2);
  }
}
DVS_HD void stats_x(const float* X, int base, const YN& y, XS& q) {
  f2 hx[3], hxx[3], hxy[3];
  for (int m2 = 0; m2 < 3; ++m2) {
    const float* r0 = X + base + (2 * m2 - 1) * PW;
    const float* r1 = r0 + PW;
    f2 a{r0[-1], r1[-1]}, b{r0[0], r1[0]}, c{r0[1], r1[1]};
    if (m2 == 0) q.xc[0] = b.y;
    if (m2 == 1) { q.xc[1] = b.x; q.xc[2] = b.y; }
    if (m2 == 2) q.xc[3] = b.x;
    hx[m2] = add2(add2(a, b), c);
    hxx[m2] = fma2(c, c, fma2(b, b, mul2(a, a)));
    hxy[m2] = fma2(c, y.vp[m2][2], fma2(b, y.vp[m2][1], mul2(a, y.vp[m2][0])));
  }
  vsum4(hx, q.sx);
  vsum4(hxx, q.sxx);
  vsum4(hxy, q.sxy);
}

// SSIM loss value clamp((1 - n/d)/2, 0, 1) from 9-sums.  With S* the sums, the reference's
//   n = (2 mx my + C1)(2 sxy + C2), d = (mx^2 + my^2 + C1)(sx + sy + C2)
// equal N1 N2 / 81^2 and D1 D2 / 81^2 of the quantities below (the factors cancel in n/d).
struct SsimTerms {
  float N1, N2, D1, D2, R, rd;
};
DVS_HD void ssim_terms(float sx, float sxx, float sxy, float sy, float ysq, float ty, SsimTerms& t) {
  float pr = sx * sy;
  float e = fmaf(sx, sx, ysq);
  t.N1 = fmaf(2.f, pr, kK1);
  t.N2 = fmaf(-2.f, pr, fmaf(18.f, sxy, kK2));
  t.D1 = e + kK1;
  t.D2 = fmaf(9.f, sxx, ty) - e;
  t.rd = rcp_fast(t.D1 * t.D2);
  t.R = (t.N1 * t.N2) * t.rd;
}
DVS_HD float ssim_value(float sx, float sxx, float sxy, float sy, float ysq, float ty) {
  SsimTerms t;
  ssim_terms(sx, sxx, sxy, sy, ysq, ty, t);
  return sat01(fmaf(-0.5f, t.R, 0.5f));
}
// d S / d (sum x), d S / d (sum x^2), d S / d (sum x y) times `scale`; zero where the clamp is active
// (torch.clamp passes the gradient at exactly 0 and 1).  Returns the loss value.
DVS_HD float ssim_coefs(float sx, float sxx, float sxy, float sy, float ysq, float ty, float scale, float& al, float& be,
                        float& ga) {
  SsimTerms t;
  ssim_terms(sx, sxx, sxy, sy, ysq, ty, t);
  float rk = (t.R >= -1.f && t.R <= 1.f) ? t.rd * scale : 0.f;
  float w = fmaf(-(t.R * sx), t.D2 - t.D1, sy * (t.N2 - t.N1));
  al = -w * rk;
  be = 4.5f * (t.R * t.D1) * rk;
  ga = -9.f * t.N1 * rk;
  return sat01(fmaf(-0.5f, t.R, 0.5f));
}

// reprojection losses ssim_w*mean_c SSIM + l1_w*mean_c |y-x| of the own pixels for all sources; X planes of
// source i start at xoff + 3*i*PLANE.  Optionally the smoothness edge weights' |dy| sums.
template <int NS, bool EDGES>
DVS_HD void quad_reproj(const float* sm, int xoff, int yoff, int base, float ssim_w3, float l1_w3, float (*r)[4], float* ax,
                        float* ay) {
  float rs[NS][4], rl[NS][4];
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int j = 0; j < 4; ++j) rs[i][j] = rl[i][j] = 0.f;
  for (int c = 0; c < 3; ++c) {
    YN yn;
    load_yn(sm + yoff + c * PLANE, base, yn);
    if (EDGES)
      for (int j = 0; j < 4; ++j) {
        ax[j] += fabsf(yn_at(yn, j + 1, 1) - yn_at(yn, j + 1, 2));
        ay[j] += fabsf(yn_at(yn, j + 1, 1) - yn_at(yn, j + 2, 1));
      }
    DVS_UNROLL
    for (int i = 0; i < NS; ++i) {
      XS xs;
      stats_x(sm + xoff + (3 * i + c) * PLANE, base, yn, xs);
      for (int j = 0; j < 4; ++j) {
        rs[i][j] += ssim_value(xs.sx[j], xs.sxx[j], xs.sxy[j], yn.sy[j], yn.ysq[j], yn.ty[j]);
        rl[i][j] += fabsf(yn_at(yn, j + 1, 1) - xs.xc[j]);
      }
    }
  }
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int j = 0; j < 4; ++j) r[i][j] = fmaf(ssim_w3, rs[i][j], l1_w3 * rl[i][j]);
}

// Same, and in the same pass the SSIM coefficient fields (NS <= 2): the fields of every source but the last go
// straight to the F planes, those of the last source are held in registers until the selection is known
// (`hold[c][f][j]`); the caller then overwrites F where the last source won.  With two equally good sources the
// per-pixel selection is salt-and-pepper, so nearly every thread needs both anyway and a second statistics pass
// for "the selected source only" would cost a full pass.
template <int NS>
DVS_HD void quad_reproj_coefs(float* sm, int xoff, int yoff, int foff, int base, float ssim_w3, float l1_w3, float kF,
                              float (*r)[4], float (*hold)[3][4]) {
  float rs[NS][4], rl[NS][4];
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int j = 0; j < 4; ++j) rs[i][j] = rl[i][j] = 0.f;
  DVS_UNROLL
  for (int c = 0; c < 3; ++c) {
    YN yn;
    load_yn(sm + yoff + c * PLANE, base, yn);
    DVS_UNROLL
    for (int i = 0; i < NS; ++i) {
      XS xs;
      stats_x(sm + xoff + (3 * i + c) * PLANE, base, yn, xs);
      for (int j = 0; j < 4; ++j) {
        float al, be, ga;
        rs[i][j] += ssim_coefs(xs.sx[j], xs.sxx[j], xs.sxy[j], yn.sy[j], yn.ysq[j], yn.ty[j], kF, al, be, ga);
        rl[i][j] += fabsf(yn_at(yn, j + 1, 1) - xs.xc[j]);
        if (i == NS - 1) {
          hold[c][0][j] = al; hold[c][1][j] = be; hold[c][2][j] = ga;
        } else {
          int o = foff + (c * 3) * PLANE + base + j * PW;
          sm[o] = al; sm[o + PLANE] = be; sm[o + 2 * PLANE] = ga;
        }
      }
    }
  }
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int j = 0; j < 4; ++j) r[i][j] = fmaf(ssim_w3, rs[i][j], l1_w3 * rl[i][j]);
}

// ------------------------------------------------------------------------------------------------ phase 1
// identity terms + smoothness edge weights of the own pixels (scale independent).
template <int NS>
DVS_HD void phase_identity(const FusedParams& p, const Tile& t, float* sm, int tid, ThreadState<NS>& st) {
  SmemLayout L{NS};
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base = pidx(r0, cx);
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.auto_mask) {
    quad_reproj<NS, true>(sm, L.x(0, 0), L.y(0), base, sw3, lw3, st.ident, ax, ay);
  } else {
    for (int c = 0; c < 3; ++c)
      for (int j = 0; j < 4; ++j) {
        int o = base + j * PW;
        float y0 = sm[L.y(c) + o];
        ax[j] += fabsf(y0 - sm[L.y(c) + o + 1]);
        ay[j] += fabsf(y0 - sm[L.y(c) + o + PW]);
      }
    DVS_UNROLL
    for (int i = 0; i < NS; ++i)
      for (int j = 0; j < 4; ++j) st.ident[i][j] = 0.f;
  }
  const int gx = t.gx0 + cx;
  for (int j = 0; j < 4; ++j) {
    int gy = t.gy0 + r0 + j, o = base + j * PW;
    bool in = (st.flags >> j) & 1;
    sm[L.wx() + o] = (in && gx < p.W - 1) ? exp_fast(-ax[j] * (1.f / 3.f)) : 0.f;
    sm[L.wy() + o] = (in && gy < p.H - 1) ? exp_fast(-ay[j] * (1.f / 3.f)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ geometry
// Sampling position of one target pixel in source i: c = D * (A (u,v,1)) + p, pixel = c.xy / (c.z + eps),
// border-clipped like F.grid_sample(padding_mode="border", align_corners=True).  The cell origin is clamped to
// W-2 / H-2 so that the four taps are always in range (the weights make that exact: tx = 1 on the last column).
struct Proj {
  float q[3];            // A (u,v,1) = d c / d D
  float px, py, rz;      // un-clipped pixel coordinates, 1/(z+eps)
  float tx, ty;
  int o;                 // offset of the north-west tap inside a plane
};
DVS_HD void project(const float* A, float u, float v, float D, float eps, int H, int W, Proj& r) {
  r.q[0] = fmaf(A[0], u, fmaf(A[1], v, A[2]));
  r.q[1] = fmaf(A[3], u, fmaf(A[4], v, A[5]));
  r.q[2] = fmaf(A[6], u, fmaf(A[7], v, A[8]));
  float c0 = fmaf(D, r.q[0], A[9]), c1 = fmaf(D, r.q[1], A[10]), c2 = fmaf(D, r.q[2], A[11]);
  r.rz = rcp_fast(c2 + eps);
  r.px = c0 * r.rz;
  r.py = c1 * r.rz;
  float ix = fminf(fmaxf(r.px, 0.f), (float)(W - 1)), iy = fminf(fmaxf(r.py, 0.f), (float)(H - 1));
  int x0 = imin((int)ix, W - 2), y0 = imin((int)iy, H - 2);
  r.tx = ix - (float)x0;
  r.ty = iy - (float)y0;
  r.o = y0 * W + x0;
}
DVS_HD float bilerp(const float* __restrict__ img, int W, const Proj& r) {
  const float* a = img + r.o;
  float nw = a[0], ne = a[1], sw = a[W], se = a[W + 1];
  float top = fmaf(r.tx, ne - nw, nw), bot = fmaf(r.tx, se - sw, sw);
  return fmaf(r.ty, bot - top, top);
}
// d value / d ix and d value / d iy of the bilinear interpolation
DVS_HD void bilerp_slopes(const float* __restrict__ img, int W, const Proj& r, float& dx, float& dy) {
  const float* a = img + r.o;
  float nw = a[0], ne = a[1], sw = a[W], se = a[W + 1];
  float dt = ne - nw, db = se - sw;
  dx = fmaf(r.ty, db - dt, dt);
  float top = fmaf(r.tx, dt, nw), bot = fmaf(r.tx, db, sw);
  dy = bot - top;
}

// Up-sampled disparity at one fine pixel, split into "issue the loads" and "combine" so that the warp phase can fetch
// the disparity of its NEXT pixel while it works on the current one.
struct DispTaps {
  float a, b, c, e, lx, ly;
};
DVS_HD void disp_taps_load(const float* __restrict__ d, int dh, int dw, float sy, float sx, bool direct, int ry, int rx,
                           DispTaps& q) {
  if (direct) {
    q.a = q.b = q.c = q.e = d[ry * dw + rx];
    q.lx = q.ly = 0.f;
    return;
  }
  int y0, y1, x0, x1;
  up_taps(ry, sy, dh, y0, y1, q.ly);
  up_taps(rx, sx, dw, x0, x1, q.lx);
  q.a = d[y0 * dw + x0]; q.b = d[y0 * dw + x1]; q.c = d[y1 * dw + x0]; q.e = d[y1 * dw + x1];
}
DVS_HD float disp_taps_value(const DispTaps& q, bool direct) {
  if (direct) return q.a;
  float top = q.a * (1.f - q.lx) + q.b * q.lx;
  float bot = q.c * (1.f - q.lx) + q.e * q.lx;
  return top * (1.f - q.ly) + bot * q.ly;
}

// ------------------------------------------------------------------------------------------------ phase W
// warp every source onto R2 for scale s; store up-sampled disparity.
template <int NS>
DVS_HD void phase_warp(const FusedParams& p, const Tile& t, float* sm, int tid, int s, const ThreadState<NS>& st) {
  SmemLayout L{NS};
  const float* c = sm + L.consts();
  const int HW = p.H * p.W;
  const int dh = p.dh[s], dw = p.dw[s];
  const float* d = p.disp[s] + (size_t)t.b * dh * dw;
  const bool direct = dh == p.H && dw == p.W;
  const float scy = (float)dh / (float)p.H, scx = (float)dw / (float)p.W;
  const int* posp = reinterpret_cast<const int*>(sm + L.pos());
  int pk = posp[tid];                                   // tid < PLANE always (NT <= PLANE)
  DispTaps dt;
  disp_taps_load(d, dh, dw, scy, scx, direct, pk >> 16, pk & 0xffff, dt);
  DVS_NOUNROLL
  for (int k = tid; k < PLANE; k += NT) {
    const int rx = pk & 0xffff, ry = pk >> 16;
    const float du = disp_taps_value(dt, direct);
    if (k + NT < PLANE) {                                // disparity of the next pixel: in flight during this one
      pk = posp[k + NT];
      disp_taps_load(d, dh, dw, scy, scx, direct, pk >> 16, pk & 0xffff, dt);
    }
    sm[L.du() + k] = du;
    float D = rcp_fast(fmaf(du, p.disp_range, p.min_disp));
    float u = (float)rx, v = (float)ry;
    // Three explicit stages so that the 12 N tap loads of a pixel are all in flight before the first one is used:
    // left to itself the compiler finishes source 0 (loads -> lerps -> shared-memory stores) before it starts the
    // loads of source 1, because it cannot prove that the stores do not alias the images.
    Proj pr[NS];
    float tap[NS][3][4];
    DVS_UNROLL
    for (int i = 0; i < NS; ++i) project(c + kC_A + 12 * i, u, v, D, p.eps, p.H, p.W, pr[i]);
    DVS_UNROLL
    for (int i = 0; i < NS; ++i) {
      const float* im = p.src[i] + (size_t)t.b * 3 * HW + pr[i].o;
      DVS_UNROLL
      for (int ch = 0; ch < 3; ++ch) {
        const float* a = im + ch * HW;
        tap[i][ch][0] = a[0]; tap[i][ch][1] = a[1]; tap[i][ch][2] = a[p.W]; tap[i][ch][3] = a[p.W + 1];
      }
    }
    DVS_UNROLL
    for (int i = 0; i < NS; ++i)
      DVS_UNROLL
      for (int ch = 0; ch < 3; ++ch) {
        const float* q = tap[i][ch];
        float top = fmaf(pr[i].tx, q[1] - q[0], q[0]), bot = fmaf(pr[i].tx, q[3] - q[2], q[2]);
        sm[L.x(i, ch) + k] = fmaf(pr[i].ty, bot - top, top);
      }
  }
}

// ------------------------------------------------------------------------------------------------ phase S
// own pixels: reprojection losses of all sources, automask + min, loss sums, smoothness (+ its gradient),
// selection tags, and (GRAD) the SSIM coefficient fields of the selected source into the F planes.
template <int NS, bool GRAD>
DVS_HD void phase_stats(const FusedParams& p, const Tile& t, float* sm, int tid, int s, ThreadState<NS>& st) {
  SmemLayout L{NS};
  const float* cst = sm + L.consts();
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base = pidx(r0, cx);
  const int gx = t.gx0 + cx, gyb = t.gy0 + r0;
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  const int HW = p.H * p.W;
  const int fl = st.flags;

  // N <= 2 with gradients: one statistics pass, target-side sums shared between the sources (quad_reproj_coefs).
  // N  > 2 with gradients: one source at a time (rolled loop below), its coefficient fields go to F where it takes the lead.
  constexpr bool kSinglePass = GRAD && NS <= 2;
  constexpr bool kRolled = GRAD && NS > 2;
  const float kF = p.ssim_w / (3.0f * (float)p.B * (float)HW);
  float r[NS][4];
  float hold[3][3][4];
  if constexpr (kSinglePass) quad_reproj_coefs<NS>(sm, L.x(0, 0), L.y(0), L.f(0), base, sw3, lw3, kF, r, hold);
  else if constexpr (!kRolled) quad_reproj<NS, false>(sm, L.x(0, 0), L.y(0), base, sw3, lw3, r, nullptr, nullptr);

  float best[4];
  int tag[4], chan[4];
  for (int j = 0; j < 4; ++j) {
    best[j] = 3.0e38f;
    tag[j] = kSelNone;
    chan[j] = 0;
  }
  if (p.auto_mask) {
    DVS_UNROLL
    for (int i = 0; i < NS; i += 2)
      for (int j = 0; j < 4; ++j) {
        float n0 = 0.f, n1 = 0.f;
        if ((fl >> j) & 1) {
          int gy = gyb + j;
          if (p.noise[s]) {
            n0 = p.noise[s][((size_t)(t.b * NS + i) * p.H + gy) * p.W + gx];
            if (i + 1 < NS) n1 = p.noise[s][((size_t)(t.b * NS + i + 1) * p.H + gy) * p.W + gx];
          } else {
            hash_normal2(p.seed, noise_offset(p), (unsigned)((t.b * p.H + gy) * p.W + gx), (unsigned)(s * kMaxN + i), n0, n1);
          }
        }
        float v0 = fmaf(n0, 0.00001f, st.ident[i][j]);
        if (v0 < best[j]) { best[j] = v0; chan[j] = i; }
        if (i + 1 < NS) {
          float v1 = fmaf(n1, 0.00001f, st.ident[i + 1 < NS ? i + 1 : i][j]);
          if (v1 < best[j]) { best[j] = v1; chan[j] = i + 1; }
        }
      }
  }
  const int off = p.auto_mask ? NS : 0;
  if constexpr (kRolled) {
    DVS_NOUNROLL
    for (int i = 0; i < NS; ++i) {
      float rs[4] = {0.f, 0.f, 0.f, 0.f}, rl[4] = {0.f, 0.f, 0.f, 0.f};
      DVS_UNROLL
      for (int c = 0; c < 3; ++c) {
        YN yn;
        XS xs;
        load_yn(sm + L.y(c), base, yn);
        stats_x(sm + L.x(0, c) + i * 3 * PLANE, base, yn, xs);
        for (int j = 0; j < 4; ++j) {
          rs[j] += ssim_coefs(xs.sx[j], xs.sxx[j], xs.sxy[j], yn.sy[j], yn.ysq[j], yn.ty[j], kF, hold[c][0][j], hold[c][1][j],
                              hold[c][2][j]);
          rl[j] += fabsf(yn_at(yn, j + 1, 1) - xs.xc[j]);
        }
      }
      for (int j = 0; j < 4; ++j) {
        const float rij = fmaf(sw3, rs[j], lw3 * rl[j]);
        if (rij < best[j]) {
          best[j] = rij; chan[j] = off + i; tag[j] = i;
          const int o = L.f(0) + base + j * PW;
          for (int c = 0; c < 3; ++c)
            for (int f = 0; f < 3; ++f) sm[o + (c * 3 + f) * PLANE] = hold[c][f][j];
        }
      }
    }
  } else {
    DVS_UNROLL
    for (int i = 0; i < NS; ++i)
      for (int j = 0; j < 4; ++j)
        if (r[i][j] < best[j]) { best[j] = r[i][j]; chan[j] = off + i; tag[j] = i; }
  }

  // loss sums, selection output, tags
  int tags = 0;
  unsigned char* selp = reinterpret_cast<unsigned char*>(sm + L.sel());
  for (int j = 0; j < 4; ++j) {
    if (!((fl >> j) & 1)) tag[j] = kSelNone;
    tags |= tag[j] << (8 * j);
    selp[base + j * PW] = (unsigned char)tag[j];
    if ((fl >> (4 + j)) & 1) {
      st.acc[0] += best[j];
      if (p.sel[s]) p.sel[s][((size_t)t.b * p.H + gyb + j) * p.W + gx] = (unsigned char)chan[j];
    }
  }
  st.tags = tags;

  // smoothness on the normalised up-sampled disparity (own pixels)
  const float inv_mu = cst[kC_invmu + s];
  const float kap = p.smooth_w / (float)(1 << s);
  const float kx = kap / ((float)p.B * (float)p.H * (float)(p.W - 1));
  const float ky = kap / ((float)p.B * (float)(p.H - 1) * (float)p.W);
  const float* DU = sm + L.du();
  const float* WX = sm + L.wx();
  const float* WY = sm + L.wy();
  for (int j = 0; j < 4; ++j) {
    st.gdu[j] = 0.f;
    if (!((fl >> (4 + j)) & 1)) continue;
    int o = base + j * PW;
    // difference first, then normalise: a*m - b*m would be contracted into an FMA whose result is the
    // rounding error of a*m when a == b (flat, border-clamped regions of the up-sampled map) and the
    // sign() below must see an exact zero there.
    float d0 = DU[o];
    float dxr = (d0 - DU[o + 1]) * inv_mu, dxl = (DU[o - 1] - d0) * inv_mu;
    float dyd = (d0 - DU[o + PW]) * inv_mu, dyu = (DU[o - PW] - d0) * inv_mu;
    float wxr = WX[o], wxl = WX[o - 1], wyd = WY[o], wyu = WY[o - PW];
    st.acc[1] += fabsf(dxr) * wxr;
    st.acc[2] += fabsf(dyd) * wyd;
    if (GRAD) {
      float gn = kx * (sgn(dxr) * wxr - sgn(dxl) * wxl) + ky * (sgn(dyd) * wyd - sgn(dyu) * wyu);
      st.gdu[j] = gn * inv_mu;
    }
  }
  if (!GRAD) return;

  if constexpr (kSinglePass) {
    // F already holds the fields of source 0 (N == 2); the last source's go in where it was selected
    for (int j = 0; j < 4; ++j) {
      if (tag[j] != NS - 1) continue;
      int o = L.f(0) + base + j * PW;
      for (int c = 0; c < 3; ++c)
        for (int f = 0; f < 3; ++f) sm[o + (c * 3 + f) * PLANE] = hold[c][f][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------ phase G
// own pixels, per source: pooled adjoint of the coefficient fields -> d loss / d warped colour -> sampling
// coordinates -> depth / pose moments.  Accumulates st.gdu and st.dM.
DVS_HD float pick4(const float* a, int j) { return j == 0 ? a[0] : (j == 1 ? a[1] : (j == 2 ? a[2] : a[3])); }

// The loops over sources and over the four pixels are deliberately NOT unrolled: the kernel is sensitive to its
// instruction footprint (the per-scale loop must stay resident in the instruction cache: at 140 KB of SASS a third
// of all stall samples were "no instruction").  Per-pixel values are picked with selects instead of indexing.
template <int NS>
DVS_HD void phase_grad(const FusedParams& p, const Tile& t, float* sm, int tid, int s, ThreadState<NS>& st) {
  SmemLayout L{NS};
  const float* cst = sm + L.consts();
  if (!(st.flags >> 4)) return;                       // no own pixel
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base = pidx(r0, cx);
  const int gx = t.gx0 + cx, gyb = t.gy0 + r0;
  const int HW = p.H * p.W;
  const unsigned char* selp = reinterpret_cast<const unsigned char*>(sm + L.sel());
  const float l1k = p.l1_w / (3.0f * (float)p.B * (float)HW);
  // reflection adjoint: the pad ring mirrors row/column 1 (and H-2 / W-2), so a pixel there appears twice in the
  // window of the centre on the image border next to it
  const float wl = (gx == 1) ? 2.f : 1.f, wr = (gx == p.W - 2) ? 2.f : 1.f;
  const bool edge_rows = (gyb <= 1 && gyb + 3 >= 1) || (gyb <= p.H - 2 && gyb + 3 >= p.H - 2);
  const float u = (float)gx;

  int tg[6][3];
  for (int m = 0; m < 6; ++m)
    for (int k = 0; k < 3; ++k) tg[m][k] = selp[base + (m - 1) * PW + k - 1];

  DVS_NOUNROLL
  for (int i = 0; i < NS; ++i) {
    f2 mk[3][3];                                        // masks on row pairs (2 m2, 2 m2 + 1)
    bool any = false;
    for (int m = 0; m < 6; ++m)
      for (int k = 0; k < 3; ++k) {
        bool hit = tg[m][k] == i;
        any = any || hit;
        float w = hit ? (k == 0 ? wl : (k == 2 ? wr : 1.f)) : 0.f;
        if (m & 1) mk[m >> 1][k].y = w; else mk[m >> 1][k].x = w;
      }
    if (!any) continue;
    float G[3][4];
    DVS_UNROLL
    for (int c = 0; c < 3; ++c) {
      float pooled[3][4];
      DVS_UNROLL
      for (int f = 0; f < 3; ++f) {
        const float* F = sm + L.f(c * 3 + f) + base;
        f2 h[3];
        for (int m2 = 0; m2 < 3; ++m2) {
          const float* r0 = F + (2 * m2 - 1) * PW;
          const float* r1 = r0 + PW;
          h[m2] = fma2(f2{r0[1], r1[1]}, mk[m2][2], fma2(f2{r0[0], r1[0]}, mk[m2][1], mul2(f2{r0[-1], r1[-1]}, mk[m2][0])));
        }
        vsum4(h, pooled[f]);
        if (edge_rows) {
          const float hs[6] = {h[0].x, h[0].y, h[1].x, h[1].y, h[2].x, h[2].y};
          for (int j = 0; j < 4; ++j) {
            if (gyb + j == 1) pooled[f][j] += hs[j];
            if (gyb + j == p.H - 2) pooled[f][j] += hs[j + 2];
          }
        }
      }
      const float* X = sm + L.x(0, c) + i * 3 * PLANE + base;
      const float* Y = sm + L.y(c) + base;
      for (int j = 0; j < 4; ++j) {
        float x = X[j * PW], y = Y[j * PW];
        float g = fmaf(2.f * x, pooled[1][j], fmaf(y, pooled[2][j], pooled[0][j]));
        if (tg[j + 1][1] == i) g -= l1k * sgn(y - x);
#if defined(DVS_FAULT_GRAD_SCALE)
        g *= DVS_FAULT_GRAD_SCALE;   // fault-injection builds of the tests only: the parity gates must catch a 1 % error
#endif
        G[c][j] = g;
      }
    }
    // chain through the bilinear gather and the projection (taps re-read; they are L1/L2 resident)
    const float* A = cst + kC_A + 12 * i;
    const float* im = p.src[i] + (size_t)t.b * 3 * HW;
    float M[12];
    for (int k = 0; k < 12; ++k) M[k] = 0.f;
    float ga[4] = {0.f, 0.f, 0.f, 0.f};
    DVS_NOUNROLL
    for (int j = 0; j < 4; ++j) {
      float g0 = pick4(G[0], j), g1 = pick4(G[1], j), g2 = pick4(G[2], j);
      if (!((st.flags >> (4 + j)) & 1)) continue;
      if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
      const float v = (float)(gyb + j);
      const float D = rcp_fast(fmaf(sm[L.du() + base + j * PW], p.disp_range, p.min_disp));
      Proj pr;
      project(A, u, v, D, p.eps, p.H, p.W, pr);
      float dx, dy, gix, giy;
      bilerp_slopes(im, p.W, pr, dx, dy);
      gix = g0 * dx; giy = g0 * dy;
      bilerp_slopes(im + HW, p.W, pr, dx, dy);
      gix = fmaf(g1, dx, gix); giy = fmaf(g1, dy, giy);
      bilerp_slopes(im + 2 * HW, p.W, pr, dx, dy);
      gix = fmaf(g2, dx, gix); giy = fmaf(g2, dy, giy);
      // ATen clip_coordinates_set_grad: zero gradient when the coordinate was clipped (border included)
      if (!(pr.px > 0.f && pr.px < (float)(p.W - 1))) gix = 0.f;
      if (!(pr.py > 0.f && pr.py < (float)(p.H - 1))) giy = 0.f;
      float gc0 = gix * pr.rz, gc1 = giy * pr.rz;
      float gc2 = -(gc0 * pr.px + gc1 * pr.py);
      float gD = fmaf(gc0, pr.q[0], fmaf(gc1, pr.q[1], gc2 * pr.q[2]));
      float gd = gD * (-p.disp_range) * (D * D);
      ga[0] += j == 0 ? gd : 0.f; ga[1] += j == 1 ? gd : 0.f; ga[2] += j == 2 ? gd : 0.f; ga[3] += j == 3 ? gd : 0.f;
      float w0 = gc0 * D, w1 = gc1 * D, w2 = gc2 * D;
      M[0] = fmaf(w0, u, M[0]); M[1] = fmaf(w0, v, M[1]); M[2] += w0; M[3] += gc0;
      M[4] = fmaf(w1, u, M[4]); M[5] = fmaf(w1, v, M[5]); M[6] += w1; M[7] += gc1;
      M[8] = fmaf(w2, u, M[8]); M[9] = fmaf(w2, v, M[9]); M[10] += w2; M[11] += gc2;
    }
    for (int j = 0; j < 4; ++j) st.gdu[j] += ga[j];
    DVS_UNROLL
    for (int ii = 0; ii < NS; ++ii)
      if (ii == i)
        for (int k = 0; k < 12; ++k) st.dM[ii][k] += M[k];
  }
}

// ------------------------------------------------------------------------------------------------ disparity gradient
// Scale whose disparity map is full resolution: direct store of the own pixels.
template <int NS>
DVS_HD void store_gdu_direct(const FusedParams& p, const Tile& t, int tid, int s, const ThreadState<NS>& st) {
  int r0, cx;
  quad_coords(tid, r0, cx);
  for (int j = 0; j < 4; ++j)
    if ((st.flags >> (4 + j)) & 1) p.gdisp[s][((size_t)t.b * p.H + t.gy0 + r0 + j) * p.W + t.gx0 + cx] = st.gdu[j];
}
// Otherwise: put the own-pixel gradients in the (now dead) DU plane, zero elsewhere ...
template <int NS>
DVS_HD void stage_gdu(const FusedParams& p, const Tile& t, float* sm, int tid, const ThreadState<NS>& st) {
  SmemLayout L{NS};
  int r0, cx;
  quad_coords(tid, r0, cx);
  for (int j = 0; j < 4; ++j) sm[L.du() + pidx(r0 + j, cx)] = ((st.flags >> (4 + j)) & 1) ? st.gdu[j] : 0.f;
}
// ... then the adjoint of the bilinear up-sample restricted to this tile, separably:
// (a) rows of R0 x coarse columns into tbuf, (b) coarse rows x coarse columns -> atomic add to global.
// Fine indices whose up-sampling taps can include coarse index J: for an integer ratio f = out/in the hat function
// around J covers [J f - f/2, J f + 3f/2 - 1] (one spare element on each side for odd f; the border clamps only
// shrink it); otherwise a generous float bound.
DVS_HD void fine_range(int J, int f, float inv, int& lo, int& hi) {
  if (f > 0) {
    lo = J * f - f / 2 - 1;
    hi = J * f + (3 * f) / 2;
  } else {
    lo = (int)(((float)J - 1.f) * inv) - 2;
    hi = (int)(((float)J + 1.5f) * inv) + 2;
  }
}
template <class LT>
DVS_HD void adjoint_rows_at(const LT& L, const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  const int* a = reinterpret_cast<const int*>(sm + L.consts() + kC_adj + 8 * s);
  const int j0 = a[2], ncj = a[3] - a[2] + 1, fxi = a[5];
  const int fy0 = t.gy0 + 1, fx0 = t.gx0 + 1, fx1 = imin(t.gx0 + TW - 2, p.W - 1);
  const int nfy = imin(t.gy0 + TH - 2, p.H - 1) - fy0 + 1;
  const float scale = (float)p.dw[s] / (float)p.W, inv = (float)p.W / (float)p.dw[s];
  const float rn = 1.0f / (float)ncj;
  for (int k = tid; k < nfy * ncj; k += NT) {
    int y = (int)(((float)k + 0.5f) * rn);              // k / ncj (exact for these small integers)
    int Jl = k - y * ncj, J = j0 + Jl;
    // fine columns that can touch coarse column J (superset, exact weight inside)
    int xa, xb;
    fine_range(J, fxi, inv, xa, xb);
    xa = imax(xa, fx0); xb = imin(xb, fx1);
    float acc = 0.f;
    const float* row = sm + L.du() + pidx(y + 1, -t.gx0);
    for (int x = xa; x <= xb; ++x) acc = fmaf(tap_weight(x, scale, p.dw[s], J), row[x], acc);
    sm[L.tbuf() + y * kTbufCols + Jl] = acc;
  }
}
template <int NS>
DVS_HD void adjoint_rows(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  adjoint_rows_at(SmemLayout{NS}, p, t, sm, tid, s);
}
template <class LT>
DVS_HD void adjoint_cols_at(const LT& L, const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  const int* a = reinterpret_cast<const int*>(sm + L.consts() + kC_adj + 8 * s);
  const int nci = a[1] - a[0] + 1, ncj = a[3] - a[2] + 1, fyi = a[4], I0 = a[0];
  const int fy0 = t.gy0 + 1, fy1 = imin(t.gy0 + TH - 2, p.H - 1);
  const float scale = (float)p.dh[s] / (float)p.H, inv = (float)p.H / (float)p.dh[s];
  const float rn = 1.0f / (float)ncj;
  for (int k = tid; k < nci * ncj; k += NT) {
    int Il = (int)(((float)k + 0.5f) * rn);
    int Jl = k - Il * ncj, I = I0 + Il;
    int ya, yb;
    fine_range(I, fyi, inv, ya, yb);
    ya = imax(ya, fy0); yb = imin(yb, fy1);
    float acc = 0.f;
    for (int y = ya; y <= yb; ++y)
      acc = fmaf(tap_weight(y, scale, p.dh[s], I), sm[L.tbuf() + (y - fy0) * kTbufCols + Jl], acc);
    p.cpart[(size_t)t.blk * p.cstride + p.coff[s] + Il * p.cbw[s] + Jl] = acc;     // own box slot: no atomics
  }
}
template <int NS>
DVS_HD void adjoint_cols(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  adjoint_cols_at(SmemLayout{NS}, p, t, sm, tid, s);
}

// d loss/s / d disp[s][b, I, J] of an up-sampled scale: the sum, in a fixed order (tile row, then tile column), of
// the boxes of the <= 4 (integer ratios; a few more otherwise) tiles whose pixels have a tap on (I, J).  Bit-reproducible.
DVS_HD float gather_gdisp(const FusedParams& p, int s, int b, int I, int J) {
  const int dh = p.dh[s], dw = p.dw[s];
  const float sy = (float)dh / (float)p.H, sx = (float)dw / (float)p.W;
  int ya, yb, xa, xb;
  fine_range(I, (p.H % dh == 0) ? p.H / dh : 0, (float)p.H / (float)dh, ya, yb);
  fine_range(J, (p.W % dw == 0) ? p.W / dw : 0, (float)p.W / (float)dw, xa, xb);
  const int ty0 = imax(ya, 0) / PITCH_Y, ty1 = imin(imin(yb, p.H - 1) / PITCH_Y, p.tiles_y - 1);
  const int tx0 = imax(xa, 0) / PITCH_X, tx1 = imin(imin(xb, p.W - 1) / PITCH_X, p.tiles_x - 1);
  float acc = 0.f;
  for (int ty = ty0; ty <= ty1; ++ty) {
    int i0, i1, a, bb;
    float l;
    up_taps(ty * PITCH_Y, sy, dh, i0, bb, l);
    up_taps(imin(ty * PITCH_Y + PITCH_Y - 1, p.H - 1), sy, dh, a, i1, l);
    if (I < i0 || I > i1) continue;
    for (int tx = tx0; tx <= tx1; ++tx) {
      int j0, j1;
      up_taps(tx * PITCH_X, sx, dw, j0, bb, l);
      up_taps(imin(tx * PITCH_X + PITCH_X - 1, p.W - 1), sx, dw, a, j1, l);
      if (J < j0 || J > j1) continue;
      const size_t blk = ((size_t)b * p.tiles_y + ty) * p.tiles_x + tx;
      acc += p.cpart[blk * p.cstride + p.coff[s] + (I - i0) * p.cbw[s] + (J - j0)];
    }
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ block reduction
// (a) every thread writes its nv partial values; (b) kWarps threads per value sum 32 entries each;
// (c) one thread per value sums the kWarps and writes the block partial.  Deterministic.
template <int NS>
DVS_HD void reduce_write(const FusedParams& p, float* sm, int tid, const ThreadState<NS>& st) {
  SmemLayout L{NS};
  constexpr int nv = 3 + 12 * NS;
  float* sc = sm + L.scratch() + tid * nv;
  sc[0] = st.acc[0]; sc[1] = st.acc[1]; sc[2] = st.acc[2];
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int k = 0; k < 12; ++k) sc[3 + 12 * i + k] = st.dM[i][k];
}
template <int NS, class LT>
DVS_HD void reduce_stage1_at(const LT& L, float* sm, int tid) {
  constexpr int nv = 3 + 12 * NS;
  for (int w = tid; w < nv * kWarps; w += NT) {
    int g = w / nv, v = w - g * nv;            // consecutive lanes -> consecutive words
    float a = 0.f;
    for (int k = 0; k < 32; ++k) a += sm[L.scratch() + (g * 32 + k) * nv + v];
    sm[L.rbuf() + w] = a;
  }
}
template <int NS>
DVS_HD void reduce_stage1(const FusedParams& p, float* sm, int tid) {
  reduce_stage1_at<NS>(SmemLayout{NS}, sm, tid);
}
template <int NS, class LT>
DVS_HD void reduce_stage2_at(const LT& L, const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  constexpr int nv = 3 + 12 * NS;
  if (tid < nv) {
    float a = 0.f;
    for (int g = 0; g < kWarps; ++g) a += sm[L.rbuf() + g * nv + tid];
    p.part[((size_t)t.blk * p.S + s) * nv + tid] = a;
  }
}
template <int NS>
DVS_HD void reduce_stage2(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  reduce_stage2_at<NS>(SmemLayout{NS}, p, t, sm, tid, s);
}

template <int NS>
DVS_HD void reset_scale_state(ThreadState<NS>& st) {
  st.acc[0] = st.acc[1] = st.acc[2] = 0.f;
  DVS_UNROLL
  for (int i = 0; i < NS; ++i)
    for (int k = 0; k < 12; ++k) st.dM[i][k] = 0.f;
}

// pose moments of one (image, scale, source) -> d loss / d T (4x4 row-major, last row zero):
//   d/dP[r][k<3] = sum_j inv_K[k][j] M[r][j],  d/dP[r][3] = M[r][3],  d/dT = K[:3,:]^T d/dP
DVS_HD void moments_to_dT(const float* M, const float* Kb, const float* iK, float* dT) {
  float dP[12];
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k)
      dP[r * 4 + k] = fmaf(iK[k * 4 + 0], M[r * 4 + 0], fmaf(iK[k * 4 + 1], M[r * 4 + 1], iK[k * 4 + 2] * M[r * 4 + 2]));
    dP[r * 4 + 3] = M[r * 4 + 3];
  }
  for (int m = 0; m < 4; ++m)
    for (int k = 0; k < 4; ++k) {
      float a = 0.f;
      for (int j = 0; j < 3; ++j) a = fmaf(Kb[j * 4 + m], dP[j * 4 + k], a);
      dT[m * 4 + k] = a;
    }
}

}  // namespace dvs
