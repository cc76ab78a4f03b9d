// Fused view-synthesis loss: per-tile phase functions.
//
// One CTA owns a 30x30 block of target pixels of one batch item ("R0") and evaluates SSIM statistics on
// the surrounding 32x32 block ("R1", the window centres whose 3x3 adjoint reaches R0) from warped
// colours on the 34x34 block ("R2").  It walks all S scales with the target / source tiles resident in
// shared memory.  The phases are written as plain functions of (tid, shared memory, per-thread state) so
// that the same source compiles
//   * under nvcc into the sm_100a kernel in dvs_fused.cu (phases separated by __syncthreads()), and
//   * under g++ into a sequential block emulator used by the CPU-side unit tests (tests/emu),
// which lets the tile logic be checked against the oracle without a GPU.
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
#else
#define DVS_HD inline
#endif

namespace dvs {

constexpr int kMaxS = 4;
constexpr int kMaxN = 4;
constexpr int TW = 32;                 // R1 width  (8 quads of 4 pixels)
constexpr int TH = 32;                 // R1 height
constexpr int PITCH_X = TW - 2;        // R0 width  = tile pitch in x
constexpr int PITCH_Y = TH - 2;        // R0 height = tile pitch in y
constexpr int RW2 = TW + 2;            // R2 width
constexpr int RH2 = TH + 2;            // R2 height
constexpr int PW = 36;                 // plane row stride (floats); col lx lives at lx+4, lx in [-1,TW]
constexpr int PLANE = RH2 * PW + 4;    // 1228 floats; the +4 absorbs lx==TW of the last row
constexpr int NT = 256;                // threads per CTA == quads in R1
constexpr int kMeanBlocks = 16;        // partial sums per (scale, batch item) in the disparity-mean pre-pass
constexpr float kC1 = 0.0001f, kC2 = 0.0009f;
constexpr int kSelNone = 255;

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
  float min_disp, disp_range;      // scaled = min_disp + disp_range*disp
  float ssim_w, l1_w, smooth_w, eps;
  int auto_mask;
  int want_grad;
  // outputs
  float* gdisp[kMaxS];             // unit gradients d loss/s / d disp[s] (accumulated; pre-zeroed when up-sampled)
  unsigned char* sel[kMaxS];       // optional argmin maps
  // workspace
  const float* mean_part;          // [S][B][kMeanBlocks] partial sums of the up-sampled disparity
  float* part;                     // [nblk][S][3 + 12N] per-block partial sums
  int tiles_x, tiles_y;
};

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
DVS_HD float sgn(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }
DVS_HD int imin(int a, int b) { return a < b ? a : b; }
DVS_HD int imax(int a, int b) { return a > b ? a : b; }

DVS_HD int reflect_clamp(int g, int n) {
  g = g < 0 ? -g : g;
  g = g >= n ? 2 * (n - 1) - g : g;
  return imin(imax(g, 0), n - 1);
}
DVS_HD int pidx(int ly, int lx) { return (ly + 1) * PW + (lx + 4); }

// ATen bilinear (align_corners=False) source taps for one output coordinate (UpSample.h:259-311).
DVS_HD void up_taps(int dst, float scale, int in, int& i0, int& i1, float& lam) {
  float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = imin((int)s, in - 1);
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
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

DVS_HD float bilinear_disp(const float* __restrict__ d, int dh, int dw, int H, int W, int ry, int rx) {
  if (dh == H && dw == W) return d[ry * dw + rx];
  int y0, y1, x0, x1;
  float ly, lx;
  up_taps(ry, (float)dh / (float)H, dh, y0, y1, ly);
  up_taps(rx, (float)dw / (float)W, dw, x0, x1, lx);
  float a = d[y0 * dw + x0], b = d[y0 * dw + x1], c = d[y1 * dw + x0], e = d[y1 * dw + x1];
  float top = a * (1.f - lx) + b * lx;
  float bot = c * (1.f - lx) + e * lx;
  return top * (1.f - ly) + bot * ly;
}

// Counter-based standard normal for the automask tie-break when no noise tensor is given.
DVS_HD unsigned int hash32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
DVS_HD float hash_normal(unsigned long long seed, unsigned long long offset, unsigned int idx, unsigned int stream) {
  unsigned int k = (unsigned int)(seed ^ (seed >> 32)) + 0x9e3779b9U * (unsigned int)(offset + stream);
  unsigned int h1 = hash32(idx ^ k);
  unsigned int h2 = hash32(h1 + 0x68bc21ebU + stream);
  float u1 = ((float)(h1 >> 8) + 1.0f) * (1.0f / 16777217.0f);   // (0,1)
  float u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);
#if defined(__CUDA_ARCH__)
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
#else
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530718f * u2);
#endif
}

// ------------------------------------------------------------------------------------------------ shared memory
// planes: Y[3] | X[3N] | F[9] | DU | WX | WY | SEL(bytes, PLANE of them) | consts
struct SmemLayout {
  int N;
  DVS_HD int y(int c) const { return c * PLANE; }
  DVS_HD int x(int i, int c) const { return (3 + i * 3 + c) * PLANE; }
  DVS_HD int f(int k) const { return (3 + 3 * N + k) * PLANE; }
  DVS_HD int du() const { return (12 + 3 * N) * PLANE; }
  DVS_HD int wx() const { return (13 + 3 * N) * PLANE; }
  DVS_HD int wy() const { return (14 + 3 * N) * PLANE; }
  DVS_HD int sel() const { return (15 + 3 * N) * PLANE; }            // PLANE bytes = PLANE/4 floats
  DVS_HD int consts() const { return sel() + PLANE / 4; }            // kConstFloats floats
  DVS_HD int total() const { return consts() + 64 + 16 * kMaxN; }
  // scratch for block reductions / up-sample adjoint: aliases X (+F) once those are dead
  DVS_HD int scratch() const { return x(0, 0); }
  DVS_HD int tbuf() const { return f(9) - 1152; }                     // last 1152 floats of F
  DVS_HD int rbuf() const { return f(9) - 1152 - 512; }               // 512 floats before it
};
// consts block: [0..8] inv_K 3x3, [9..12] inv_mu[s], [13..16] unused, [20+12i ..] P_i (3x4)
constexpr int kC_iK = 0, kC_invmu = 9, kC_P = 20;

template <int NS>
struct ThreadState {
  float ident[NS][4];   // identity reprojection loss of the own quad (scale independent)
  float acc[3];            // photometric sum, smooth-x sum, smooth-y sum of the current scale
  float dP[NS][12];     // pose-gradient accumulators of the current scale
  float gdu[4];            // d loss / d disp_up of the own quad, current scale
  unsigned char tag[4];    // selected source of the own quad (kSelNone: identity / outside)
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

// ------------------------------------------------------------------------------------------------ phase 0
// constants of the tile: inv_K, P_i = (K T_i)[:3,:], 1/(clamp(mean disp)+1e-7) per scale.
template <int NS>
DVS_HD void phase_consts(const FusedParams& p, const Tile& t, float* sm, int tid) {
  SmemLayout L{NS};
  float* c = sm + L.consts();
  if (tid < 9) {
    int r = tid / 3, k = tid - r * 3;
    c[kC_iK + tid] = p.invK[t.b * 16 + r * 4 + k];
  } else if (tid >= 32 && tid < 32 + 12 * NS) {
    int e = tid - 32, i = e / 12, j = (e % 12) / 4, k = e % 4;
    const float* Kb = p.K + t.b * 16;
    const float* Tb = p.T[i] + t.b * 16;
    float a = 0.f;
    for (int m = 0; m < 4; ++m) a = fmaf(Kb[j * 4 + m], Tb[m * 4 + k], a);
    c[kC_P + e] = a;
  } else if (tid >= 96 && tid < 96 + p.S) {
    int s = tid - 96;
    const float* mp = p.mean_part + (s * p.B + t.b) * kMeanBlocks;
    float m = 0.f;
    for (int k = 0; k < kMeanBlocks; ++k) m += mp[k];
    m = m / ((float)p.H * (float)p.W);
    c[kC_invmu + s] = 1.0f / (fmaxf(m, 0.001f) + 1e-7f);
  }
}

// load target (-> Y) and sources (-> X, for the identity terms) on R2 with reflection; zero F.
template <int NS>
DVS_HD void phase_load(const FusedParams& p, const Tile& t, float* sm, int tid) {
  SmemLayout L{NS};
  const int HW = p.H * p.W;
  for (int k = tid; k < RW2 * RH2; k += NT) {
    int ly = k / RW2 - 1, lx = k % RW2 - 1;
    int gy = reflect_clamp(t.gy0 + ly, p.H), gx = reflect_clamp(t.gx0 + lx, p.W);
    int o = gy * p.W + gx, q = pidx(ly, lx);
    const float* tg = p.target + (size_t)t.b * 3 * HW + o;
    sm[L.y(0) + q] = tg[0];
    sm[L.y(1) + q] = tg[HW];
    sm[L.y(2) + q] = tg[2 * HW];
    if (p.auto_mask) {
      for (int i = 0; i < NS; ++i) {
        const float* sg = p.src[i] + (size_t)t.b * 3 * HW + o;
        sm[L.x(i, 0) + q] = sg[0];
        sm[L.x(i, 1) + q] = sg[HW];
        sm[L.x(i, 2) + q] = sg[2 * HW];
      }
    }
  }
  for (int k = tid; k < 9 * PLANE; k += NT) sm[L.f(0) + k] = 0.f;
}

// ------------------------------------------------------------------------------------------------ quad SSIM
struct QuadStats {
  float sx[4], sxx[4], sxy[4];
};
struct QuadY {
  float yv[3][6];
  float sy[4], syy[4];
};

DVS_HD void load_row6(const float* pl, int o, float* v) {
  v[0] = pl[o - 1];
  v[1] = pl[o];
  v[2] = pl[o + 1];
  v[3] = pl[o + 2];
  v[4] = pl[o + 3];
  v[5] = pl[o + 4];
}

DVS_HD void quad_y(const float* Y, int base, QuadY& q) {
  float cy[6], cyy[6];
  for (int r = 0; r < 3; ++r) {
    load_row6(Y, base + (r - 1) * PW, q.yv[r]);
    for (int k = 0; k < 6; ++k) {
      float v = q.yv[r][k];
      cy[k] = r ? cy[k] + v : v;
      cyy[k] = r ? fmaf(v, v, cyy[k]) : v * v;
    }
  }
  for (int j = 0; j < 4; ++j) {
    q.sy[j] = cy[j] + cy[j + 1] + cy[j + 2];
    q.syy[j] = cyy[j] + cyy[j + 1] + cyy[j + 2];
  }
}

DVS_HD void quad_x(const float* X, int base, const QuadY& y, QuadStats& q, float* xc) {
  float cx[6], cxx[6], cxy[6];
  for (int r = 0; r < 3; ++r) {
    float xv[6];
    load_row6(X, base + (r - 1) * PW, xv);
    if (r == 1) { xc[0] = xv[1]; xc[1] = xv[2]; xc[2] = xv[3]; xc[3] = xv[4]; }
    for (int k = 0; k < 6; ++k) {
      float v = xv[k];
      cx[k] = r ? cx[k] + v : v;
      cxx[k] = r ? fmaf(v, v, cxx[k]) : v * v;
      cxy[k] = r ? fmaf(v, y.yv[r][k], cxy[k]) : v * y.yv[r][k];
    }
  }
  for (int j = 0; j < 4; ++j) {
    q.sx[j] = cx[j] + cx[j + 1] + cx[j + 2];
    q.sxx[j] = cxx[j] + cxx[j + 1] + cxx[j + 2];
    q.sxy[j] = cxy[j] + cxy[j + 1] + cxy[j + 2];
  }
}

// SSIM loss value from 9-sums; optionally the coefficient fields (a=dS/dm(x), b=dS/dm(x^2), c=dS/dm(xy)).
template <bool COEF>
DVS_HD float ssim_from_sums(float sx, float sxx, float sxy, float sy, float syy, float& a, float& b, float& c) {
  const float i9 = 1.0f / 9.0f;
  float mx = sx * i9, my = sy * i9;
  float mxy = mx * my;
  float mx2 = mx * mx, my2 = my * my;
  float n1 = fmaf(2.f, mxy, kC1);
  float n2 = fmaf(2.f, fmaf(sxy, i9, -mxy), kC2);
  float d1 = mx2 + my2 + kC1;
  float d2 = fmaf(sxx + syy, i9, -(mx2 + my2)) + kC2;
  float n = n1 * n2, d = d1 * d2;
  float rd = rcp_fast(d);
  float raw = fmaf(-0.5f * n, rd, 0.5f);
  float S = fminf(fmaxf(raw, 0.f), 1.f);
  if (COEF) {
    float live = (raw >= 0.f && raw <= 1.f) ? 1.f : 0.f;
    float nrd = n * rd;
    // a = -[my (n2-n1) - nrd mx (d2-d1)] / d ; b = 0.5 nrd / d2 ; c = -n1/d
    a = -(my * (n2 - n1) - nrd * mx * (d2 - d1)) * rd * live;
    b = 0.5f * nrd * rcp_fast(d2) * live;
    c = -n1 * rd * live;
  }
  return S;
}

// reprojection loss (ssim_w*mean_c SSIM + l1_w*mean_c |y-x|) of the own quad for image planes X[3]
DVS_HD void quad_reproj(const float* sm, int xoff, int yoff, int base, float ssim_w3, float l1_w3, float* r) {
  for (int j = 0; j < 4; ++j) r[j] = 0.f;
  for (int c = 0; c < 3; ++c) {
    QuadY qy;
    QuadStats qs;
    float xc[4], a, b, cc;
    quad_y(sm + yoff + c * PLANE, base, qy);
    quad_x(sm + xoff + c * PLANE, base, qy, qs, xc);
    for (int j = 0; j < 4; ++j) {
      float S = ssim_from_sums<false>(qs.sx[j], qs.sxx[j], qs.sxy[j], qy.sy[j], qy.syy[j], a, b, cc);
      r[j] = fmaf(ssim_w3, S, fmaf(l1_w3, fabsf(qy.yv[1][j + 1] - xc[j]), r[j]));
    }
  }
}

DVS_HD void quad_coords(int tid, int& qr, int& qc) { qr = tid >> 3; qc = (tid & 7) << 2; }

// ------------------------------------------------------------------------------------------------ phase 1
// identity terms + smoothness edge weights of the own quad (scale independent).
template <int NS>
DVS_HD void phase_identity(const FusedParams& p, const Tile& t, float* sm, int tid, ThreadState<NS>& st) {
  SmemLayout L{NS};
  int qr, qc;
  quad_coords(tid, qr, qc);
  int base = pidx(qr, qc);
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  if (p.auto_mask)
    for (int i = 0; i < NS; ++i) quad_reproj(sm, L.x(i, 0), L.y(0), base, sw3, lw3, st.ident[i]);
  int gy = t.gy0 + qr;
  for (int j = 0; j < 4; ++j) {
    int gx = t.gx0 + qc + j, o = base + j;
    bool in = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    float ax = 0.f, ay = 0.f;
    for (int c = 0; c < 3; ++c) {
      float y0 = sm[L.y(c) + o];
      ax += fabsf(y0 - sm[L.y(c) + o + 1]);
      ay += fabsf(y0 - sm[L.y(c) + o + PW]);
    }
    sm[L.wx() + o] = (in && gx < p.W - 1) ? exp_fast(-ax * (1.f / 3.f)) : 0.f;
    sm[L.wy() + o] = (in && gy < p.H - 1) ? exp_fast(-ay * (1.f / 3.f)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ geometry
struct Geo {
  float D, ray[3], cam[3];
};
DVS_HD void pixel_geo(const float* c, float du, int rx, int ry, float min_disp, float range, Geo& g) {
  float u = (float)rx, v = (float)ry;
  g.D = rcp_fast(fmaf(du, range, min_disp));
  for (int k = 0; k < 3; ++k) {
    g.ray[k] = fmaf(c[kC_iK + 3 * k], u, fmaf(c[kC_iK + 3 * k + 1], v, c[kC_iK + 3 * k + 2]));
    g.cam[k] = g.D * g.ray[k];
  }
}
struct Proj {
  float px, py, rz;      // un-clipped pixel coordinates, 1/(z+eps)
  float tx, ty;
  int o00, o01, o10, o11;
  float m01, m10, m11;   // tap validity (0/1): taps beyond the last row/column contribute 0
  float livex, livey;    // 0 when the coordinate was clipped (ATen clip_coordinates_set_grad)
};
DVS_HD void project(const float* P, const Geo& g, float eps, int H, int W, Proj& q) {
  float c0 = fmaf(P[0], g.cam[0], fmaf(P[1], g.cam[1], fmaf(P[2], g.cam[2], P[3])));
  float c1 = fmaf(P[4], g.cam[0], fmaf(P[5], g.cam[1], fmaf(P[6], g.cam[2], P[7])));
  float c2 = fmaf(P[8], g.cam[0], fmaf(P[9], g.cam[1], fmaf(P[10], g.cam[2], P[11])));
  q.rz = rcp_fast(c2 + eps);
  q.px = c0 * q.rz;
  q.py = c1 * q.rz;
  float wm = (float)(W - 1), hm = (float)(H - 1);
  q.livex = (q.px > 0.f && q.px < wm) ? 1.f : 0.f;
  q.livey = (q.py > 0.f && q.py < hm) ? 1.f : 0.f;
  float ix = fminf(fmaxf(q.px, 0.f), wm), iy = fminf(fmaxf(q.py, 0.f), hm);
  int x0 = (int)ix, y0 = (int)iy;
  q.tx = ix - (float)x0;
  q.ty = iy - (float)y0;
  int okx = x0 + 1 <= W - 1, oky = y0 + 1 <= H - 1;
  int x1 = x0 + okx, y1 = y0 + oky;
  q.o00 = y0 * W + x0; q.o01 = y0 * W + x1; q.o10 = y1 * W + x0; q.o11 = y1 * W + x1;
  q.m01 = (float)okx; q.m10 = (float)oky; q.m11 = (float)(okx & oky);
}
// bilinear value (and, if SLOPE, d value / d ix, d value / d iy) of one channel plane
template <bool SLOPE>
DVS_HD float sample(const float* __restrict__ img, const Proj& q, float& dx, float& dy) {
  float nw = img[q.o00], ne = img[q.o01] * q.m01, sw = img[q.o10] * q.m10, se = img[q.o11] * q.m11;
  float wx1 = q.tx, wx0 = 1.f - q.tx, wy1 = q.ty, wy0 = 1.f - q.ty;
  if (SLOPE) {
    dx = ((ne - nw) * wy0 + (se - sw) * wy1) * q.livex;
    dy = ((sw - nw) * wx0 + (se - ne) * wx1) * q.livey;
  }
  return nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1);
}

// ------------------------------------------------------------------------------------------------ phase W
// warp every source onto R2 for scale s; store up-sampled disparity.
template <int NS>
DVS_HD void phase_warp(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  SmemLayout L{NS};
  const float* c = sm + L.consts();
  const int HW = p.H * p.W;
  const float* d = p.disp[s] + (size_t)t.b * p.dh[s] * p.dw[s];
  for (int k = tid; k < RW2 * RH2; k += NT) {
    int ly = k / RW2 - 1, lx = k % RW2 - 1;
    int ry = reflect_clamp(t.gy0 + ly, p.H), rx = reflect_clamp(t.gx0 + lx, p.W);
    int q = pidx(ly, lx);
    float du = bilinear_disp(d, p.dh[s], p.dw[s], p.H, p.W, ry, rx);
    sm[L.du() + q] = du;
    Geo g;
    pixel_geo(c, du, rx, ry, p.min_disp, p.disp_range, g);
    for (int i = 0; i < NS; ++i) {
      Proj pr;
      project(c + kC_P + 12 * i, g, p.eps, p.H, p.W, pr);
      const float* im = p.src[i] + (size_t)t.b * 3 * HW;
      float dx, dy;
      sm[L.x(i, 0) + q] = sample<false>(im, pr, dx, dy);
      sm[L.x(i, 1) + q] = sample<false>(im + HW, pr, dx, dy);
      sm[L.x(i, 2) + q] = sample<false>(im + 2 * HW, pr, dx, dy);
    }
  }
}

// ------------------------------------------------------------------------------------------------ phase S
// own quad: reprojection losses of all sources, automask + min, loss sums, smoothness (+ its gradient),
// selection tags, and (GRAD) the SSIM coefficient fields of the selected source into the F planes.
template <int NS, bool GRAD>
DVS_HD void phase_stats(const FusedParams& p, const Tile& t, float* sm, int tid, int s, ThreadState<NS>& st) {
  SmemLayout L{NS};
  const float* cst = sm + L.consts();
  int qr, qc;
  quad_coords(tid, qr, qc);
  const int base = pidx(qr, qc);
  const int gy = t.gy0 + qr;
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  const int HW = p.H * p.W;

  float best[4];
  int tag[4], chan[4];
  bool inimg[4], own[4];
  for (int j = 0; j < 4; ++j) {
    int gx = t.gx0 + qc + j;
    inimg[j] = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    own[j] = inimg[j] && qr >= 1 && qr <= TH - 2 && (qc + j) >= 1 && (qc + j) <= TW - 2;
    best[j] = 3.0e38f;
    tag[j] = kSelNone;
    chan[j] = 0;
  }
  if (p.auto_mask) {
    for (int i = 0; i < NS; ++i)
      for (int j = 0; j < 4; ++j) {
        float v = st.ident[i][j];
        if (inimg[j]) {
          int gx = t.gx0 + qc + j;
          float nz;
          if (p.noise[s])
            nz = p.noise[s][((size_t)(t.b * NS + i) * p.H + gy) * p.W + gx];
          else
            nz = hash_normal(p.seed, p.offset, (unsigned)((t.b * p.H + gy) * p.W + gx), (unsigned)(s * kMaxN + i));
          v = fmaf(nz, 0.00001f, v);
        }
        if (v < best[j]) { best[j] = v; chan[j] = i; }
      }
  }
  const int off = p.auto_mask ? NS : 0;
  for (int i = 0; i < NS; ++i) {
    float r[4];
    quad_reproj(sm, L.x(i, 0), L.y(0), base, sw3, lw3, r);
    for (int j = 0; j < 4; ++j)
      if (r[j] < best[j]) { best[j] = r[j]; chan[j] = off + i; tag[j] = i; }
  }
  // loss sums, selection output, tags
  for (int j = 0; j < 4; ++j) {
    if (!inimg[j]) tag[j] = kSelNone;
    st.tag[j] = (unsigned char)tag[j];
    if (own[j]) {
      st.acc[0] += best[j];
      if (p.sel[s]) p.sel[s][((size_t)t.b * p.H + gy) * p.W + t.gx0 + qc + j] = (unsigned char)chan[j];
    }
  }
  unsigned char* selp = reinterpret_cast<unsigned char*>(sm + L.sel());
  for (int j = 0; j < 4; ++j) selp[base + j] = st.tag[j];

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
    if (!own[j]) continue;
    int o = base + j;
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

  // coefficient fields of the selected source(s) of this quad -> F planes (pre-scaled)
  const float kF = p.ssim_w / (27.0f * (float)p.B * (float)HW);
  for (int i = 0; i < NS; ++i) {
    bool any = false;
    for (int j = 0; j < 4; ++j) any = any || (tag[j] == i);
    if (!any) continue;
    for (int c = 0; c < 3; ++c) {
      QuadY qy;
      QuadStats qs;
      float xc[4];
      quad_y(sm + L.y(c), base, qy);
      quad_x(sm + L.x(i, c), base, qy, qs, xc);
      for (int j = 0; j < 4; ++j) {
        if (tag[j] != i) continue;
        float a, b, cc;
        ssim_from_sums<true>(qs.sx[j], qs.sxx[j], qs.sxy[j], qy.sy[j], qy.syy[j], a, b, cc);
        sm[L.f(c * 3 + 0) + base + j] = a * kF;
        sm[L.f(c * 3 + 1) + base + j] = b * kF;
        sm[L.f(c * 3 + 2) + base + j] = cc * kF;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ phase G
// own quad, per source: pooled adjoint of the coefficient fields -> d loss / d warped colour -> sampling
// coordinates -> depth / pose.  Accumulates st.gdu and st.dP.
template <int NS>
DVS_HD void phase_grad(const FusedParams& p, const Tile& t, float* sm, int tid, int s, ThreadState<NS>& st) {
  SmemLayout L{NS};
  const float* cst = sm + L.consts();
  int qr, qc;
  quad_coords(tid, qr, qc);
  if (qr < 1 || qr > TH - 2) return;
  const int base = pidx(qr, qc);
  const int gy = t.gy0 + qr;
  if (gy >= p.H) return;
  const int HW = p.H * p.W;
  const unsigned char* selp = reinterpret_cast<const unsigned char*>(sm + L.sel());
  const float wup = (gy == 1) ? 2.f : 1.f, wdn = (gy == p.H - 2) ? 2.f : 1.f;
  const float l1k = p.l1_w / (3.0f * (float)p.B * (float)HW);

  unsigned char tg[3][6];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 6; ++k) tg[r][k] = selp[base + (r - 1) * PW + k - 1];

  for (int i = 0; i < NS; ++i) {
    float m[3][6];
    bool any = false;
    for (int r = 0; r < 3; ++r) {
      float wr = r == 0 ? wup : (r == 2 ? wdn : 1.f);
      for (int k = 0; k < 6; ++k) {
        bool hit = tg[r][k] == i;
        any = any || hit;
        m[r][k] = hit ? wr : 0.f;
      }
    }
    if (!any) continue;
    float G[3][4];
    for (int c = 0; c < 3; ++c) {
      float pooled[3][4];
      for (int f = 0; f < 3; ++f) {
        const float* F = sm + L.f(c * 3 + f);
        float cs[6];
        for (int r = 0; r < 3; ++r) {
          float v[6];
          load_row6(F, base + (r - 1) * PW, v);
          for (int k = 0; k < 6; ++k) cs[k] = r ? fmaf(v[k], m[r][k], cs[k]) : v[k] * m[r][k];
        }
        for (int j = 0; j < 4; ++j) {
          int gx = t.gx0 + qc + j;
          float v = cs[j] + cs[j + 1] + cs[j + 2];
          if (gx == 1) v += cs[j];
          if (gx == p.W - 2) v += cs[j + 2];
          pooled[f][j] = v;
        }
      }
      for (int j = 0; j < 4; ++j) {
        float x = sm[L.x(i, c) + base + j], y = sm[L.y(c) + base + j];
        float g = fmaf(2.f * x, pooled[1][j], fmaf(y, pooled[2][j], pooled[0][j]));
        if (tg[1][j + 1] == i) g -= l1k * sgn(y - x);
        G[c][j] = g;
      }
    }
    // chain through the bilinear gather and the projection (taps re-read; they are L1/L2 resident)
    const float* P = cst + kC_P + 12 * i;
    const float* im = p.src[i] + (size_t)t.b * 3 * HW;
    for (int j = 0; j < 4; ++j) {
      int lx = qc + j, gx = t.gx0 + lx;
      if (lx < 1 || lx > TW - 2 || gx >= p.W) continue;
      Geo g;
      pixel_geo(cst, sm[L.du() + base + j], gx, gy, p.min_disp, p.disp_range, g);
      Proj pr;
      project(P, g, p.eps, p.H, p.W, pr);
      float gix = 0.f, giy = 0.f;
      for (int c = 0; c < 3; ++c) {
        float dx, dy;
        sample<true>(im + c * HW, pr, dx, dy);
        gix = fmaf(G[c][j], dx, gix);
        giy = fmaf(G[c][j], dy, giy);
      }
      float gc[3];
      gc[0] = gix * pr.rz;
      gc[1] = giy * pr.rz;
      gc[2] = -(gix * pr.px + giy * pr.py) * pr.rz;
      float gD = 0.f;
      for (int k = 0; k < 3; ++k) {
        float gcam = fmaf(gc[0], P[k], fmaf(gc[1], P[4 + k], gc[2] * P[8 + k]));
        gD = fmaf(gcam, g.ray[k], gD);
      }
      st.gdu[j] = fmaf(gD * (-p.disp_range), g.D * g.D, st.gdu[j]);
      for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) st.dP[i][r * 4 + k] = fmaf(gc[r], g.cam[k], st.dP[i][r * 4 + k]);
        st.dP[i][r * 4 + 3] += gc[r];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ disparity gradient
// Scale whose disparity map is full resolution: direct store of the own pixels.
template <int NS>
DVS_HD void store_gdu_direct(const FusedParams& p, const Tile& t, int tid, int s, const ThreadState<NS>& st) {
  int qr, qc;
  quad_coords(tid, qr, qc);
  int gy = t.gy0 + qr;
  if (qr < 1 || qr > TH - 2 || gy >= p.H) return;
  for (int j = 0; j < 4; ++j) {
    int lx = qc + j, gx = t.gx0 + lx;
    if (lx < 1 || lx > TW - 2 || gx >= p.W) continue;
    p.gdisp[s][((size_t)t.b * p.H + gy) * p.W + gx] = st.gdu[j];
  }
}
// Otherwise: put the own-pixel gradients in the (now dead) DU plane, zero elsewhere ...
template <int NS>
DVS_HD void stage_gdu(const FusedParams& p, const Tile& t, float* sm, int tid, const ThreadState<NS>& st) {
  SmemLayout L{NS};
  int qr, qc;
  quad_coords(tid, qr, qc);
  int gy = t.gy0 + qr;
  for (int j = 0; j < 4; ++j) {
    int lx = qc + j, gx = t.gx0 + lx;
    bool own = qr >= 1 && qr <= TH - 2 && lx >= 1 && lx <= TW - 2 && gy < p.H && gx < p.W;
    sm[L.du() + pidx(qr, lx)] = own ? st.gdu[j] : 0.f;
  }
}
// ... then the adjoint of the bilinear up-sample restricted to this tile, separably:
// (a) rows of R0 x coarse columns into tbuf, (b) coarse rows x coarse columns -> atomic add to global.
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
constexpr int kTbufCols = 36;   // >= max coarse columns touched by 30 fine columns (+ slack), rows = 30
DVS_HD float tap_weight(int fine, float scale, int in, int coarse) {
  int a, b;
  float l;
  up_taps(fine, scale, in, a, b, l);
  float w = 0.f;
  if (a == coarse) w += 1.f - l;
  if (b == coarse) w += l;
  return w;
}
template <int NS>
DVS_HD void adjoint_rows(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  SmemLayout L{NS};
  CoarseBox cb = coarse_box(p, t, s);
  int ncj = cb.j1 - cb.j0 + 1, nfy = cb.fy1 - cb.fy0 + 1;
  float scale = (float)p.dw[s] / (float)p.W;
  float inv = (float)p.W / (float)p.dw[s];
  for (int k = tid; k < nfy * ncj; k += NT) {
    int y = k / ncj, J = cb.j0 + k % ncj;
    // fine columns that can touch coarse column J (conservative superset, exact test inside)
    int xa = imax((int)(((float)J - 1.f) * inv) - 2, cb.fx0), xb = imin((int)(((float)J + 1.5f) * inv) + 2, cb.fx1);
    float acc = 0.f;
    for (int x = xa; x <= xb; ++x) {
      float w = tap_weight(x, scale, p.dw[s], J);
      acc = fmaf(w, sm[L.du() + pidx(cb.fy0 + y - t.gy0, x - t.gx0)], acc);
    }
    sm[L.tbuf() + y * kTbufCols + (J - cb.j0)] = acc;
  }
}
DVS_HD void atomic_add_f32(float* a, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(a, v);
#else
  *a += v;
#endif
}
template <int NS>
DVS_HD void adjoint_cols(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  SmemLayout L{NS};
  CoarseBox cb = coarse_box(p, t, s);
  int ncj = cb.j1 - cb.j0 + 1, nci = cb.i1 - cb.i0 + 1;
  float scale = (float)p.dh[s] / (float)p.H;
  float inv = (float)p.H / (float)p.dh[s];
  for (int k = tid; k < nci * ncj; k += NT) {
    int I = cb.i0 + k / ncj, Jl = k % ncj;
    int ya = imax((int)(((float)I - 1.f) * inv) - 2, cb.fy0), yb = imin((int)(((float)I + 1.5f) * inv) + 2, cb.fy1);
    float acc = 0.f;
    for (int y = ya; y <= yb; ++y) {
      float w = tap_weight(y, scale, p.dh[s], I);
      acc = fmaf(w, sm[L.tbuf() + (y - cb.fy0) * kTbufCols + Jl], acc);
    }
    atomic_add_f32(p.gdisp[s] + ((size_t)t.b * p.dh[s] + I) * p.dw[s] + cb.j0 + Jl, acc);
  }
}

// ------------------------------------------------------------------------------------------------ block reduction
// (a) every thread writes its nv partial values; (b) 8 threads per value sum 32 entries each;
// (c) one thread per value sums the 8 and writes the block partial.  Deterministic.
template <int NS>
DVS_HD void reduce_write(const FusedParams& p, float* sm, int tid, const ThreadState<NS>& st) {
  SmemLayout L{NS};
  constexpr int nv = 3 + 12 * NS;
  float* sc = sm + L.scratch() + tid * nv;
  sc[0] = st.acc[0]; sc[1] = st.acc[1]; sc[2] = st.acc[2];
  for (int i = 0; i < NS; ++i)
    for (int k = 0; k < 12; ++k) sc[3 + 12 * i + k] = st.dP[i][k];
}
template <int NS>
DVS_HD void reduce_stage1(const FusedParams& p, float* sm, int tid) {
  SmemLayout L{NS};
  constexpr int nv = 3 + 12 * NS;
  for (int w = tid; w < nv * 8; w += NT) {
    int v = w >> 3, g = w & 7;
    float a = 0.f;
    for (int k = 0; k < 32; ++k) a += sm[L.scratch() + (g * 32 + k) * nv + v];
    sm[L.rbuf() + w] = a;
  }
}
template <int NS>
DVS_HD void reduce_stage2(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  SmemLayout L{NS};
  constexpr int nv = 3 + 12 * NS;
  if (tid < nv) {
    float a = 0.f;
    for (int g = 0; g < 8; ++g) a += sm[L.rbuf() + tid * 8 + g];
    p.part[((size_t)t.blk * p.S + s) * nv + tid] = a;
  }
}

template <int NS>
DVS_HD void reset_scale_state(ThreadState<NS>& st) {
  st.acc[0] = st.acc[1] = st.acc[2] = 0.f;
  for (int i = 0; i < NS; ++i)
    for (int k = 0; k < 12; ++k) st.dP[i][k] = 0.f;
}

}  // namespace dvs
