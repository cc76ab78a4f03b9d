// Fused view-synthesis loss, two-source specialisation ("pair" kernel): per-tile phase functions.
//
// Same tile geometry, shared-memory planes, reductions and up-sample adjoint as dvs_fused_core.cuh (one CTA per 30x30
// block of target pixels, lane == column, a thread owns four vertically adjacent pixels).  What changes is the
// arithmetic layout: the reference's configuration has exactly two source frames (vo/learner_new.py:148,206,214 hard-code
// [-1, +1]), and everything that is computed "per source" -- projection, bilinear taps, 3x3 sums of x, x^2, xy, the SSIM
// algebra, the gradient chain, the pose moments -- is carried as a PAIR (source 0, source 1) in one 64-bit register and
// issued as Blackwell packed fp32 instructions (FADD2 / FMUL2 / FFMA2: two lanes per issue slot).  The kernel is
// issue-slot bound (profiles/r01d_*), so halving the slots of the per-source arithmetic is the lever; the warped colours
// of the two sources are interleaved in shared memory (float2 per pixel: one 64-bit load feeds both lanes).
//
//   * statistics: two pixels at a time (rolled loop over the halves of the quad): 4 rows of horizontal sums for 2
//     pixels; keeps the live state small (no spills at 128 registers) and the loop body short (instruction cache)
//   * coefficient pooling: F holds the fields of the selected source and exact zeros elsewhere, so one sweep yields
//     the pooled fields of source 0 (masked) and of both (unmasked, same operation order); source 1 = difference
//   * gradient chain: both sources of a pixel at once, taps re-gathered (L1 resident)
//
// Reference arithmetic: vo/learner_new.py:60-74,132-258; formulas as in oracle/closed_form.py.  Operation order
// follows dvs_fused_core.cuh so that the two kernels agree to the last bit wherever the algorithm is the same.
#pragma once
#include "dvs_fused_core.cuh"

namespace dvs {

// Shared memory of the two-source kernel (floats), 94 KB with 32-row tiles: two CTAs per SM fit the 196 KB carve-out, which
// leaves 60 KB of L1 for the gathers (the generic layout's 103 KB forces the 228 KB carve-out and 28 KB of L1).  No position
// table (the warp phase recomputes the reflected image coordinates), the identity terms live in registers, and the nine
// coefficient planes cover R1 only (row pitch TW).
//   Y[3] | X2[3] (float2) | F[9] (R1) | DU | WX | WY | SEL (bytes) | consts | A2
constexpr int FP = TH * TW;            // floats of one coefficient plane; R1 pixel (ly, lx) is at ly * TW + lx
struct PairLayout {
  DVS_HD int y(int c) const { return c * PLANE; }
  DVS_HD int x2(int c) const { return (3 + 2 * c) * PLANE; }                       // float2 plane of channel c
  DVS_HD int f(int k) const { return 9 * PLANE + k * FP; }
  DVS_HD int du() const { return f(9); }
  DVS_HD int wx() const { return du() + PLANE; }                                   // edge weights of the smoothness term
  DVS_HD int wy() const { return du() + 2 * PLANE; }
  DVS_HD int sel() const { return du() + 3 * PLANE; }                              // PLANE bytes, indexed like an R2 plane
  DVS_HD int consts() const { return sel() + (PLANE + 3) / 4; }
  DVS_HD int a2() const { return (consts() + 8 + 12 * kMaxN + 8 * kMaxS + 1) & ~1; }   // 12 float2: (A_0[e], A_1[e]); 8-byte aligned
  DVS_HD int total() const { return a2() + 24; }
  // scratch for the block reductions / the up-sample adjoint: aliases X2 (and the head of F), and the tail of F, once
  // those are dead
  DVS_HD int scratch() const { return x2(0); }
  DVS_HD int tbuf() const { return f(9) - kTbufFloats; }
  DVS_HD int rbuf() const { return f(9) - kTbufFloats - 512; }
};
static_assert(NT * (3 + 12 * 2) + kTbufFloats + 512 <= 6 * PLANE + 9 * FP,
              "reduction scratch (from the start of X2) and tbuf / rbuf (at the end of F) must not overlap");

DVS_HD f2 ld2(const float* p) {
#if defined(__CUDA_ARCH__)
  float2 v = *reinterpret_cast<const float2*>(p);
  return f2{v.x, v.y};
#else
  return f2{p[0], p[1]};
#endif
}
DVS_HD void st2(float* p, f2 v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
#else
  p[0] = v.x; p[1] = v.y;
#endif
}
DVS_HD f2 bc2(float a) { return f2{a, a}; }

// ------------------------------------------------------------------------------------------------ input formats
// IO bit 0: disparities are bf16 (what DepthNet emits under bf16 autocast, vo/train.py:177-181); bit 1: images are uint8
// (what the loader decodes, vo/dataset/common.py:39-46,77).  Both are converted on load: bf16 -> fp32 is exact, and
// x / 255 is the correctly rounded quotient ToTensor computes (one multiply + one FMA correction step; checked for all
// 256 inputs in tests/test_abi.py), so the results are bit-identical to running on the fp32-expanded tensors.
constexpr int kIoDispBf16 = 1, kIoImgU8 = 2;

DVS_HD float u8_to_unit(unsigned char v) {
  const float x = (float)v, r = 1.0f / 255.0f;
  const float q = x * r;
  return fmaf(fmaf(-q, 255.0f, x), r, q);
}
DVS_HD float bf16_bits_to_float(unsigned short h) {
  union { unsigned int u; float f; } c;
  c.u = (unsigned int)h << 16;
  return c.f;
}
template <bool U8>
struct ImgPtr {
  const void* p;
  DVS_HD float at(int i) const {
    if (U8) return u8_to_unit(static_cast<const unsigned char*>(p)[i]);
    return static_cast<const float*>(p)[i];
  }
  DVS_HD ImgPtr off(size_t e) const {
    if (U8) return ImgPtr{static_cast<const unsigned char*>(p) + e};
    return ImgPtr{static_cast<const float*>(p) + e};
  }
};
// A pointer the compiler must treat as an opaque base (global address space): taps are then addressed as base + 32-bit
// offset (one widening multiply-add each) instead of being re-derived from the kernel parameters as 64-bit add chains.
template <class T>
DVS_HD const T* opaque_global(const T* q) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+l"(q));
  __builtin_assume(__isGlobal(q));
#endif
  return q;
}
template <bool BF16>
DVS_HD float ld_disp(const float* d, int i) {
  if (BF16) return bf16_bits_to_float(reinterpret_cast<const unsigned short*>(d)[i]);
  return d[i];
}
template <bool BF16>
DVS_HD const float* disp_image(const float* d, size_t e) {       // start of image b of a disparity map
  if (BF16) return reinterpret_cast<const float*>(reinterpret_cast<const unsigned short*>(d) + e);
  return d + e;
}
template <bool BF16>
DVS_HD void disp_taps_load_t(const float* d, int dh, int dw, float sy, float sx, bool direct, int ry, int rx, DispTaps& q) {
  if (direct) {
    q.a = q.b = q.c = q.e = ld_disp<BF16>(d, ry * dw + rx);
    q.lx = q.ly = 0.f;
    return;
  }
  int y0, y1, x0, x1;
  up_taps(ry, sy, dh, y0, y1, q.ly);
  up_taps(rx, sx, dw, x0, x1, q.lx);
  q.a = ld_disp<BF16>(d, y0 * dw + x0); q.b = ld_disp<BF16>(d, y0 * dw + x1);
  q.c = ld_disp<BF16>(d, y1 * dw + x0); q.e = ld_disp<BF16>(d, y1 * dw + x1);
}

// Each source image is addressed from one base pointer
// the compiler must treat as opaque (otherwise it re-derives every tap address as a 64-bit add chain from the kernel
// parameters, ~50 instructions per pixel) with 32-bit element offsets: one add + one widening multiply-add per (plane, row).
template <bool U8>
struct SrcPlanes {
  ImgPtr<U8> b0, b1;
  DVS_HD SrcPlanes(const FusedParams& p, int b, int HW) {
    b0 = ImgPtr<U8>{p.src[0]}.off((size_t)b * 3 * HW);
    b1 = ImgPtr<U8>{p.src[1]}.off((size_t)b * 3 * HW);
    b0.p = opaque_global(b0.p);
    b1.p = opaque_global(b1.p);
  }
};

struct PairState {
  f2 ident[4];             // identity reprojection terms of the own pixels (scale independent), (source 0, source 1)
  int flags;               // bit j: pixel j inside the image; bit 4+j: pixel j belongs to R0 (own)
  float acc[3];            // photometric sum, smooth-x sum, smooth-y sum of the current scale
  f2 M[12];                // pose-gradient moments of the current scale, (source 0, source 1)
  float gdu[4];            // d loss / d disp_up of the own pixels, current scale
  int tags;                // selected source of the 4 pixels, one byte each (kSelNone: identity / outside)
};

// ------------------------------------------------------------------------------------------------ load
// target -> Y, sources interleaved -> X2 (for the identity terms), zero F, selection plane, pixel flags.
template <int IO>
DVS_HD void pair_phase_load(const FusedParams& p, const Tile& t, float* sm, int tid, PairState& st) {
  PairLayout P;
  const PairLayout& L = P;
  const int HW = p.H * p.W;
  constexpr bool U8 = (IO & kIoImgU8) != 0;
  ImgPtr<U8> tgt = ImgPtr<U8>{p.target}.off((size_t)t.b * 3 * HW);
  tgt.p = opaque_global(tgt.p);
  const SrcPlanes<U8> im(p, t.b, HW);
  const ImgPtr<U8> sr0 = im.b0, sr1 = im.b1;
  DVS_NOUNROLL
  for (int k = tid; k < PLANE; k += NT) {
    int ly = k / PW - 1, lx = k % PW - 1;
    int gy = reflect_clamp(t.gy0 + ly, p.H), gx = reflect_clamp(t.gx0 + lx, p.W);
    int o = gy * p.W + gx;
    sm[L.y(0) + k] = tgt.at(o);
    sm[L.y(1) + k] = tgt.at(o + HW);
    sm[L.y(2) + k] = tgt.at(o + 2 * HW);
    if (p.auto_mask) {
      st2(sm + P.x2(0) + 2 * k, f2{sr0.at(o), sr1.at(o)});
      st2(sm + P.x2(1) + 2 * k, f2{sr0.at(o + HW), sr1.at(o + HW)});
      st2(sm + P.x2(2) + 2 * k, f2{sr0.at(o + 2 * HW), sr1.at(o + 2 * HW)});
    }
  }
  // the coefficient planes are rewritten on all of R1 by every statistics phase: nothing to clear
  for (int k = tid; k < (PLANE + 3) / 4; k += NT) reinterpret_cast<unsigned int*>(sm + L.sel())[k] = 0xffffffffu;   // kSelNone
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

// ------------------------------------------------------------------------------------------------ statistics of two pixels
// SSIM from 9-sums for both sources of one pixel; same operation order as ssim_terms / ssim_coefs.
struct SsimPair {
  f2 N1, N2, D1, D2, R, rd;
};
DVS_HD void ssim_terms2(f2 sx, f2 sxx, f2 sxy, float sy, float ysq, float ty, SsimPair& t) {
  const f2 pr = mul2(sx, bc2(sy));
  const f2 e = fma2(sx, sx, bc2(ysq));
  t.N1 = fma2(bc2(2.f), pr, bc2(kK1));
  t.N2 = fma2(bc2(-2.f), pr, fma2(bc2(18.f), sxy, bc2(kK2)));
  t.D1 = add2(e, bc2(kK1));
  t.D2 = sub2(fma2(bc2(9.f), sxx, bc2(ty)), e);
  const f2 dd = mul2(t.D1, t.D2);
  t.rd = f2{rcp_fast(dd.x), rcp_fast(dd.y)};
  t.R = mul2(mul2(t.N1, t.N2), t.rd);
}
DVS_HD f2 ssim_value2(const SsimPair& t) {
  return f2{sat01(fmaf(-0.5f, t.R.x, 0.5f)), sat01(fmaf(-0.5f, t.R.y, 0.5f))};
}
DVS_HD void ssim_coefs2(const SsimPair& t, f2 sx, float sy, float scale, f2& al, f2& be, f2& ga) {
  const f2 rk = f2{fabsf(t.R.x) <= 1.f ? t.rd.x * scale : 0.f, fabsf(t.R.y) <= 1.f ? t.rd.y * scale : 0.f};
  const f2 nrk = f2{-rk.x, -rk.y};
  // w = -(R sx) (D2 - D1) + sy (N2 - N1)   (the sign moved onto the difference: exact)
  const f2 w = fma2(mul2(t.R, sx), sub2(t.D1, t.D2), mul2(bc2(sy), sub2(t.N2, t.N1)));
  al = mul2(w, nrk);
  be = mul2(mul2(bc2(4.5f), mul2(t.R, t.D1)), rk);
  ga = mul2(mul2(bc2(-9.f), t.N1), rk);
}

// Reprojection terms r = ssim_w3 * sum_c SSIM + l1_w3 * sum_c |y - x| of both sources for the two pixels at R1 rows
// (ra, ra + 1), column cx (base = pidx(ra, cx), fbase = ra * TW + cx); X2 holds what is compared with the target (warped
// colours, or the raw sources for the identity terms).  COEFS: the SSIM coefficient fields of source 0 go to F, those of
// source 1 to `hold`.  EDGES: the four |dy| sums per pixel behind the smoothness edge weights, e[j] = {right, left, down,
// up} neighbour (sum over the channels of |y - y_neighbour|, accumulated in channel order like the reference's mean).
template <bool COEFS, bool EDGES>
DVS_HD void pair_half(float* sm, int base, int fbase, float ssim_w3, float l1_w3, float kF, f2* r, float (*hold)[3][2],
                      float (*e)[4]) {
  PairLayout P;
  const PairLayout& L = P;
  f2 rs[2] = {f2{0.f, 0.f}, f2{0.f, 0.f}}, rl[2] = {f2{0.f, 0.f}, f2{0.f, 0.f}};
  DVS_UNROLL
  for (int c = 0; c < 3; ++c) {
    const float* Y = sm + L.y(c) + base;
    const float* X = sm + P.x2(c) + 2 * base;
    f2 hx[4], hxx[4], hxy[4], xc[2];
    float hy[4], hyy[4], yc[4];
    float yv[4][3];
    DVS_UNROLL
    for (int m = 0; m < 4; ++m) {
      const float* yr = Y + (m - 1) * PW;
      yv[m][0] = yr[-1]; yv[m][1] = yr[0]; yv[m][2] = yr[1];
    }
    // target-side row sums on row pairs (0,1), (2,3): packed, same operation order as the scalar form
    DVS_UNROLL
    for (int m2 = 0; m2 < 2; ++m2) {
      const f2 a{yv[2 * m2][0], yv[2 * m2 + 1][0]}, b{yv[2 * m2][1], yv[2 * m2 + 1][1]}, d{yv[2 * m2][2], yv[2 * m2 + 1][2]};
      const f2 s1 = add2(add2(a, b), d), s2 = fma2(d, d, fma2(b, b, mul2(a, a)));
      hy[2 * m2] = s1.x; hy[2 * m2 + 1] = s1.y;
      hyy[2 * m2] = s2.x; hyy[2 * m2 + 1] = s2.y;
    }
    DVS_UNROLL
    for (int m = 0; m < 4; ++m) {
      const float* xr = X + 2 * (m - 1) * PW;
      const float a = yv[m][0], b = yv[m][1], d = yv[m][2];
      const f2 xa = ld2(xr - 2), xb = ld2(xr), xd = ld2(xr + 2);
      hx[m] = add2(add2(xa, xb), xd);
      hxx[m] = fma2(xd, xd, fma2(xb, xb, mul2(xa, xa)));
      hxy[m] = fma2(xd, bc2(d), fma2(xb, bc2(b), mul2(xa, bc2(a))));
      yc[m] = b;
      if (m == 1) xc[0] = xb;
      if (m == 2) xc[1] = xb;
      if (EDGES && (m == 1 || m == 2)) {
        e[m - 1][0] += fabsf(b - d);                      // centre - right
        e[m - 1][1] += fabsf(a - b);                      // left - centre (the left neighbour's "centre - right")
      }
    }
    if (EDGES) {
      e[0][2] += fabsf(yc[1] - yc[2]); e[0][3] += fabsf(yc[0] - yc[1]);     // centre - down, up - centre
      e[1][2] += fabsf(yc[2] - yc[3]); e[1][3] += fabsf(yc[1] - yc[2]);
    }
    // vertical 3-sums, shared middle partial (same order as vsum4)
    const float uy = hy[1] + hy[2], uyy = hyy[1] + hyy[2];
    const f2 ux = add2(hx[1], hx[2]), uxx = add2(hxx[1], hxx[2]), uxy = add2(hxy[1], hxy[2]);
    DVS_UNROLL
    for (int j = 0; j < 2; ++j) {
      const float sy = j ? uy + hy[3] : hy[0] + uy;
      const float syy = j ? uyy + hyy[3] : hyy[0] + uyy;
      const f2 sx = j ? add2(ux, hx[3]) : add2(hx[0], ux);
      const f2 sxx = j ? add2(uxx, hxx[3]) : add2(hxx[0], uxx);
      const f2 sxy = j ? add2(uxy, hxy[3]) : add2(hxy[0], uxy);
      const float ysq = sy * sy, ty = fmaf(9.f, syy, kK2);
      SsimPair t;
      ssim_terms2(sx, sxx, sxy, sy, ysq, ty, t);
      rs[j] = add2(rs[j], ssim_value2(t));
      const f2 d = sub2(bc2(yc[j + 1]), xc[j]);
      rl[j] = f2{rl[j].x + fabsf(d.x), rl[j].y + fabsf(d.y)};
      if (COEFS) {
        f2 al, be, ga;
        ssim_coefs2(t, sx, sy, kF, al, be, ga);
        float* F = sm + L.f(c * 3) + fbase + j * TW;
        F[0] = al.x; F[FP] = be.x; F[2 * FP] = ga.x;
        hold[c][0][j] = al.y; hold[c][1][j] = be.y; hold[c][2][j] = ga.y;
      }
    }
  }
  DVS_UNROLL
  for (int j = 0; j < 2; ++j) r[j] = fma2(bc2(ssim_w3), rs[j], mul2(bc2(l1_w3), rl[j]));
}

// ------------------------------------------------------------------------------------------------ identity
// identity reprojection terms of the own pixels (scale independent; kept in registers).
DVS_HD void pair_phase_identity(const FusedParams& p, const Tile& t, float* sm, int tid, PairState& st) {
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base0 = pidx(r0, cx);
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  DVS_UNROLL
  for (int j = 0; j < 4; ++j) st.ident[j] = f2{0.f, 0.f};
  PairLayout L;
  const int gx = t.gx0 + cx;
  DVS_NOUNROLL
  for (int h = 0; h < 2; ++h) {
    f2 r[2] = {f2{0.f, 0.f}, f2{0.f, 0.f}};
    float e[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int base = base0 + 2 * h * PW;
    if (p.auto_mask) {
      pair_half<false, true>(sm, base, 0, sw3, lw3, 0.f, r, nullptr, e);
    } else {
      for (int c = 0; c < 3; ++c)
        for (int j = 0; j < 2; ++j) {
          const int o = base + j * PW;
          const float y0 = sm[L.y(c) + o];
          e[j][0] += fabsf(y0 - sm[L.y(c) + o + 1]);
          e[j][2] += fabsf(y0 - sm[L.y(c) + o + PW]);
        }
    }
    if (h == 0) { st.ident[0] = r[0]; st.ident[1] = r[1]; }
    else { st.ident[2] = r[0]; st.ident[3] = r[1]; }
    for (int j = 0; j < 2; ++j) {
      const int jj = 2 * h + j, gy = t.gy0 + r0 + jj, o = base + j * PW;
      const bool in = (st.flags >> jj) & 1;
      sm[L.wx() + o] = (in && gx < p.W - 1) ? exp_fast(-e[j][0] * (1.f / 3.f)) : 0.f;
      sm[L.wy() + o] = (in && gy < p.H - 1) ? exp_fast(-e[j][2] * (1.f / 3.f)) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ geometry (pairs)
struct Proj2 {
  f2 q[3];             // A (u,v,1) = d c / d D
  f2 px, py, rz;       // un-clipped pixel coordinates, 1/(z+eps)
  f2 tx, ty;
  int o0, o1;          // offsets of the north-west taps inside a plane
};
// A: the 12 (A_0[e], A_1[e]) pairs (registers in the warp phase, re-read from shared memory in the gradient phase)
DVS_HD void project2(const f2* A, float u, float v, float D, float eps, int H, int W, Proj2& r) {
  const f2 u2 = bc2(u), v2 = bc2(v), D2 = bc2(D);
  r.q[0] = fma2(A[0], u2, fma2(A[1], v2, A[2]));
  r.q[1] = fma2(A[3], u2, fma2(A[4], v2, A[5]));
  r.q[2] = fma2(A[6], u2, fma2(A[7], v2, A[8]));
  const f2 c0 = fma2(D2, r.q[0], A[9]), c1 = fma2(D2, r.q[1], A[10]), c2 = fma2(D2, r.q[2], A[11]);
  r.rz = f2{rcp_fast(c2.x + eps), rcp_fast(c2.y + eps)};
  r.px = mul2(c0, r.rz);
  r.py = mul2(c1, r.rz);
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  const f2 ix = f2{fminf(fmaxf(r.px.x, 0.f), wm), fminf(fmaxf(r.px.y, 0.f), wm)};
  const f2 iy = f2{fminf(fmaxf(r.py.x, 0.f), hm), fminf(fmaxf(r.py.y, 0.f), hm)};
  const int x0 = imin((int)ix.x, W - 2), x1 = imin((int)ix.y, W - 2);
  const int y0 = imin((int)iy.x, H - 2), y1 = imin((int)iy.y, H - 2);
  r.tx = sub2(ix, f2{(float)x0, (float)x1});
  r.ty = sub2(iy, f2{(float)y0, (float)y1});
  r.o0 = y0 * W + x0;
  r.o1 = y1 * W + x1;
}
// the 24 taps of one pixel (both sources, three channels): issue only
template <bool U8>
DVS_HD void gather_taps2(const SrcPlanes<U8>& im, int o0, int o1, int HW, int W, f2 (*tap)[4]) {
  DVS_UNROLL
  for (int ch = 0; ch < 3; ++ch) {
    const int a0 = o0 + ch * HW, a1 = o1 + ch * HW;
    tap[ch][0] = f2{im.b0.at(a0), im.b1.at(a1)};
    tap[ch][1] = f2{im.b0.at(a0 + 1), im.b1.at(a1 + 1)};
    tap[ch][2] = f2{im.b0.at(a0 + W), im.b1.at(a1 + W)};
    tap[ch][3] = f2{im.b0.at(a0 + W + 1), im.b1.at(a1 + W + 1)};
  }
}
DVS_HD void lerp_store2(float* sm, int x2off, int k, f2 tx, f2 ty, const f2 (*tap)[4]) {
  DVS_UNROLL
  for (int ch = 0; ch < 3; ++ch) {
    const f2 top = fma2(tx, sub2(tap[ch][1], tap[ch][0]), tap[ch][0]);
    const f2 bot = fma2(tx, sub2(tap[ch][3], tap[ch][2]), tap[ch][2]);
    st2(sm + x2off + 2 * ch * PLANE + 2 * k, fma2(ty, sub2(bot, top), top));
  }
}

// ------------------------------------------------------------------------------------------------ phase W
// warp both sources onto R2 for scale s (interleaved float2 planes); store the up-sampled disparity.
// One pixel per iteration: all 24 image taps are issued before the first interpolation, and the disparity taps of the NEXT
// pixel are fetched one iteration ahead.
template <int IO>
DVS_HD void pair_phase_warp(const FusedParams& p, const Tile& t, float* sm, int tid, int s) {
  PairLayout P;
  const PairLayout& L = P;
  const int HW = p.H * p.W;
  const int dh = p.dh[s], dw = p.dw[s];
  constexpr bool U8 = (IO & kIoImgU8) != 0, BF = (IO & kIoDispBf16) != 0;
  const float* d = opaque_global(disp_image<BF>(p.disp[s], (size_t)t.b * dh * dw));
  const bool direct = dh == p.H && dw == p.W;
  const float scy = (float)dh / (float)p.H, scx = (float)dw / (float)p.W;
  const SrcPlanes<U8> im(p, t.b, HW);
  const int x2off = P.x2(0);

  // R2 coordinates of pixel k = tid + NT it, advanced incrementally (NT = dq PW + dr); the reflected image coordinates
  // are recomputed from them (no position table in shared memory)
  constexpr int dq = NT / PW, dr = NT % PW;
  int ly = tid / PW, lx = tid - ly * PW;
  int ry = reflect_clamp(t.gy0 + ly - 1, p.H), rx = reflect_clamp(t.gx0 + lx - 1, p.W);
  DispTaps dt;
  disp_taps_load_t<BF>(d, dh, dw, scy, scx, direct, ry, rx, dt);
  f2 A[12];                                              // projection constants: in registers for the whole phase
  DVS_UNROLL
  for (int e = 0; e < 12; ++e) A[e] = ld2(sm + P.a2() + 2 * e);
  DVS_NOUNROLL
  for (int k = tid; k < PLANE; k += NT) {
    const float u = (float)rx, v = (float)ry;
    const float du = disp_taps_value(dt, direct);
    if (k + NT < PLANE) {                                // disparity of the next pixel: in flight during this one
      ly += dq; lx += dr;
      if (lx >= PW) { lx -= PW; ly += 1; }
      ry = reflect_clamp(t.gy0 + ly - 1, p.H); rx = reflect_clamp(t.gx0 + lx - 1, p.W);
      disp_taps_load_t<BF>(d, dh, dw, scy, scx, direct, ry, rx, dt);
    }
    sm[L.du() + k] = du;
    Proj2 pr;
    project2(A, u, v, rcp_fast(fmaf(du, p.disp_range, p.min_disp)), p.eps, p.H, p.W, pr);
    f2 tap[3][4];
    gather_taps2(im, pr.o0, pr.o1, HW, p.W, tap);   // all 24 tap loads of the pixel before the first use
    lerp_store2(sm, x2off, k, pr.tx, pr.ty, tap);
  }
}

// ------------------------------------------------------------------------------------------------ phase S
template <bool GRAD>
DVS_HD void pair_phase_stats(const FusedParams& p, const Tile& t, float* sm, int tid, int s, PairState& st) {
  PairLayout P;
  const PairLayout& L = P;
  const float* cst = sm + L.consts();
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base0 = pidx(r0, cx);
  const int gx = t.gx0 + cx, gyb = t.gy0 + r0;
  const float sw3 = p.ssim_w * (1.f / 3.f), lw3 = p.l1_w * (1.f / 3.f);
  const int fl = st.flags;
  const float kF = p.kF;
  unsigned char* selp = reinterpret_cast<unsigned char*>(sm + L.sel());
  const int off = p.auto_mask ? 2 : 0;
  const float inv_mu = cst[kC_invmu + s];
  const float kx = p.kxs[s], ky = p.kys[s];
  const float* DU = sm + L.du();
  int tags = 0;
  float photo = 0.f, smx = 0.f, smy = 0.f;
  float g01[2] = {0.f, 0.f}, g23[2] = {0.f, 0.f};        // smoothness part of d loss / d disp_up of the four pixels

  DVS_NOUNROLL
  for (int h = 0; h < 2; ++h) {
    const int base = base0 + 2 * h * PW;
    f2 r[2];
    float hold[3][3][2];
    pair_half<GRAD, false>(sm, base, (r0 + 2 * h) * TW + cx, sw3, lw3, kF, r, hold, nullptr);
    DVS_UNROLL
    for (int j = 0; j < 2; ++j) {
      const int jj = 2 * h + j;
      const bool in = (fl >> jj) & 1;
      const bool own = (fl >> (4 + jj)) & 1;
      float best = 3.0e38f;
      int tag = kSelNone, chan = 0;
      if (p.auto_mask) {
        float n0 = 0.f, n1 = 0.f;
        const f2 idv = h ? st.ident[2 + j] : st.ident[j];
        // The in-kernel generator is bounded (|n| <= sqrt(48 ln 2) = 5.77, i.e. 5.77e-5 after scaling): where a
        // reprojection term beats both identity terms by more than that, no draw can change the outcome (minimum,
        // argmin and loss value are the reprojection's), so the draw is skipped.  Given noise tensors are always read.
        const bool need = p.noise[s] != nullptr || !(fminf(r[j].x, r[j].y) < fminf(idv.x, idv.y) - 6.0e-5f);
        if (in && need) {
          const int gy = gyb + jj;
          if (p.noise[s]) {
            n0 = p.noise[s][((size_t)(t.b * 2) * p.H + gy) * p.W + gx];
            n1 = p.noise[s][((size_t)(t.b * 2 + 1) * p.H + gy) * p.W + gx];
          } else {
            hash_normal2(p.seed, noise_offset(p), (unsigned)((t.b * p.H + gy) * p.W + gx), (unsigned)(s * kMaxN), n0, n1);
          }
        }
        const float v0 = fmaf(n0, 0.00001f, idv.x);
        const float v1 = fmaf(n1, 0.00001f, idv.y);
        if (v0 < best) { best = v0; chan = 0; }
        if (v1 < best) { best = v1; chan = 1; }
      }
      if (r[j].x < best) { best = r[j].x; chan = off; tag = 0; }
      if (r[j].y < best) { best = r[j].y; chan = off + 1; tag = 1; }
      if (!in) tag = kSelNone;
      tags |= tag << (8 * jj);
      selp[base + j * PW] = (unsigned char)tag;
      if (own) {
        photo += best;
        if (p.sel[s]) p.sel[s][((size_t)t.b * p.H + gyb + jj) * p.W + gx] = (unsigned char)chan;
      }
      if (GRAD && tag != 0) {
        // F holds source 0's fields: replace them by source 1's where it won, by exact zeros where neither did
        float* F = sm + L.f(0) + (r0 + jj) * TW + cx;
        DVS_UNROLL
        for (int c = 0; c < 3; ++c)
          DVS_UNROLL
          for (int f = 0; f < 3; ++f) F[(c * 3 + f) * FP] = tag == 1 ? hold[c][f][j] : 0.f;
      }
      // smoothness on the normalised up-sampled disparity (own pixels); edge weights exp(-mean_c |dy|) from the sums
      // pair_half collected (zero across the image border, learner_func.py:161-174)
      float gsm = 0.f;
      if (own) {
        const int o = base + j * PW;
        const float wxr = sm[L.wx() + o], wxl = sm[L.wx() + o - 1], wyd = sm[L.wy() + o], wyu = sm[L.wy() + o - PW];
        // difference first, then normalise (see phase_stats)
        const float d0 = DU[o];
        const float dxr = (d0 - DU[o + 1]) * inv_mu, dxl = (DU[o - 1] - d0) * inv_mu;
        const float dyd = (d0 - DU[o + PW]) * inv_mu, dyu = (DU[o - PW] - d0) * inv_mu;
        smx += fabsf(dxr) * wxr;
        smy += fabsf(dyd) * wyd;
        if (GRAD) {
          const float gn = kx * (sgn(dxr) * wxr - sgn(dxl) * wxl) + ky * (sgn(dyd) * wyd - sgn(dyu) * wyu);
          gsm = gn * inv_mu;
        }
      }
      if (h == 0) g01[j] = gsm; else g23[j] = gsm;
    }
  }
  st.tags = tags;
  st.acc[0] += photo;
  st.acc[1] += smx;
  st.acc[2] += smy;
  st.gdu[0] = g01[0]; st.gdu[1] = g01[1]; st.gdu[2] = g23[0]; st.gdu[3] = g23[1];
}

// ------------------------------------------------------------------------------------------------ phase G
// own pixels: pooled adjoint of the coefficient fields for both sources in one sweep -> d loss / d warped colour ->
// sampling coordinates -> depth / pose moments.  Accumulates st.gdu and st.M.
template <int IO>
DVS_HD void pair_phase_grad(const FusedParams& p, const Tile& t, float* sm, int tid, int s, PairState& st) {
  PairLayout P;
  const PairLayout& L = P;
  constexpr bool U8 = (IO & kIoImgU8) != 0;
  if (!(st.flags >> 4)) return;                       // no own pixel
  int r0, cx;
  quad_coords(tid, r0, cx);
  const int base = pidx(r0, cx);
  const int gx = t.gx0 + cx, gyb = t.gy0 + r0;
  const int HW = p.H * p.W;
  const unsigned char* selp = reinterpret_cast<const unsigned char*>(sm + L.sel());
  const float l1k = p.l1k;
  // reflection adjoint: the pad ring mirrors row/column 1 (and H-2 / W-2)
  const float wl = (gx == 1) ? 2.f : 1.f, wr = (gx == p.W - 2) ? 2.f : 1.f;
  const bool edge_rows = (gyb <= 1 && gyb + 3 >= 1) || (gyb <= p.H - 2 && gyb + 3 >= p.H - 2);
  const float u = (float)gx;

  // masks of source 0 on row pairs (2 m2, 2 m2 + 1), and "any reprojection selected in the neighbourhood"
  f2 mk[3][3];
  bool any = false;
  for (int m = 0; m < 6; ++m)
    for (int k = 0; k < 3; ++k) {
      const int tg = selp[base + (m - 1) * PW + k - 1];
      any = any || tg != kSelNone;
      const float w = (tg == 0) ? (k == 0 ? wl : (k == 2 ? wr : 1.f)) : 0.f;
      if (m & 1) mk[m >> 1][k].y = w; else mk[m >> 1][k].x = w;
    }
  if (!any) return;                                    // F is all zero around the own pixels: no photometric gradient
  const f2 wl2 = bc2(wl), wr2 = bc2(wr), one2 = bc2(1.f);

  // d loss / d warped colour of the own pixels, (source 0, source 1): parked in the pixel's own X2 slot for the chain below
  // (selecting among 12 register pairs by the loop index cost more issue slots than three 64-bit shared-memory loads)
  DVS_UNROLL
  for (int c = 0; c < 3; ++c) {
    f2 pooled[3][4];                                    // (source 0, source 1) per field and pixel
    DVS_UNROLL
    for (int f = 0; f < 3; ++f) {
      const float* F = sm + L.f(c * 3 + f) + r0 * TW + cx;     // coefficient planes cover R1 only, row pitch TW
      f2 hT[3], h0[3];
      for (int m2 = 0; m2 < 3; ++m2) {
        const float* ra = F + (2 * m2 - 1) * TW;
        const float* rb = ra + TW;
        const f2 lft{ra[-1], rb[-1]}, mid{ra[0], rb[0]}, rgt{ra[1], rb[1]};
        // same operation order in both sums: where every selected neighbour chose source 0 they are bit-equal
        hT[m2] = fma2(rgt, wr2, fma2(mid, one2, mul2(lft, wl2)));
        h0[m2] = fma2(rgt, mk[m2][2], fma2(mid, mk[m2][1], mul2(lft, mk[m2][0])));
      }
      float pT[4], p0[4];
      vsum4(hT, pT);
      vsum4(h0, p0);
      if (edge_rows) {
        const float hsT[6] = {hT[0].x, hT[0].y, hT[1].x, hT[1].y, hT[2].x, hT[2].y};
        const float hs0[6] = {h0[0].x, h0[0].y, h0[1].x, h0[1].y, h0[2].x, h0[2].y};
        for (int j = 0; j < 4; ++j) {
          if (gyb + j == 1) { pT[j] += hsT[j]; p0[j] += hs0[j]; }
          if (gyb + j == p.H - 2) { pT[j] += hsT[j + 2]; p0[j] += hs0[j + 2]; }
        }
      }
      for (int j = 0; j < 4; ++j) pooled[f][j] = f2{p0[j], pT[j] - p0[j]};
    }
    float* X = sm + P.x2(c) + 2 * base;
    const float* Y = sm + L.y(c) + base;
    for (int j = 0; j < 4; ++j) {
      const f2 x = ld2(X + 2 * j * PW);
      const float y = Y[j * PW];
      f2 g = fma2(add2(x, x), pooled[1][j], fma2(bc2(y), pooled[2][j], pooled[0][j]));
      const int tg = (st.tags >> (8 * j)) & 0xff;
      if (tg == 0) g.x -= l1k * sgn(y - x.x);
      if (tg == 1) g.y -= l1k * sgn(y - x.y);
#if defined(DVS_FAULT_GRAD_SCALE)
      g = mul2(g, bc2(DVS_FAULT_GRAD_SCALE));
#endif
      st2(X + 2 * j * PW, g);      // over the warped colour itself: only this thread reads its own pixels' X2 entries here
    }
  }

  // chain through the bilinear gather and the projection (taps re-read; they are L1/L2 resident)
  const SrcPlanes<U8> im(p, t.b, HW);
  DVS_NOUNROLL
  for (int j = 0; j < 4; ++j) {
    if (!((st.flags >> (4 + j)) & 1)) continue;
    const f2 g0 = ld2(sm + P.x2(0) + 2 * (base + j * PW));
    const f2 g1 = ld2(sm + P.x2(1) + 2 * (base + j * PW));
    const f2 g2 = ld2(sm + P.x2(2) + 2 * (base + j * PW));
    if (g0.x == 0.f && g1.x == 0.f && g2.x == 0.f && g0.y == 0.f && g1.y == 0.f && g2.y == 0.f) continue;
    const float v = (float)(gyb + j);
    const float D = rcp_fast(fmaf(sm[L.du() + base + j * PW], p.disp_range, p.min_disp));
    f2 A[12];                                            // re-read per pixel: held across the loop they cost more in registers
    DVS_UNROLL
    for (int e = 0; e < 12; ++e) A[e] = ld2(sm + P.a2() + 2 * e);
    Proj2 pr;
    project2(A, u, v, D, p.eps, p.H, p.W, pr);
    f2 tap[3][4];
    gather_taps2(im, pr.o0, pr.o1, HW, p.W, tap);
    f2 gix = f2{0.f, 0.f}, giy = f2{0.f, 0.f};
    DVS_UNROLL
    for (int ch = 0; ch < 3; ++ch) {
      const f2 gc = ch == 0 ? g0 : (ch == 1 ? g1 : g2);
      const f2 dtp = sub2(tap[ch][1], tap[ch][0]), dbt = sub2(tap[ch][3], tap[ch][2]);
      const f2 dx = fma2(pr.ty, sub2(dbt, dtp), dtp);
      const f2 top = fma2(pr.tx, dtp, tap[ch][0]), bot = fma2(pr.tx, dbt, tap[ch][2]);
      const f2 dy = sub2(bot, top);
      if (ch == 0) { gix = mul2(gc, dx); giy = mul2(gc, dy); }
      else { gix = fma2(gc, dx, gix); giy = fma2(gc, dy, giy); }
    }
    // ATen clip_coordinates_set_grad: zero gradient when the coordinate was clipped (border included)
    const float wm = (float)(p.W - 1), hm = (float)(p.H - 1);
    if (!(pr.px.x > 0.f && pr.px.x < wm)) gix.x = 0.f;
    if (!(pr.px.y > 0.f && pr.px.y < wm)) gix.y = 0.f;
    if (!(pr.py.x > 0.f && pr.py.x < hm)) giy.x = 0.f;
    if (!(pr.py.y > 0.f && pr.py.y < hm)) giy.y = 0.f;
    const f2 gc0 = mul2(gix, pr.rz), gc1 = mul2(giy, pr.rz);
    const f2 sneg = fma2(gc0, pr.px, mul2(gc1, pr.py));
    const f2 gc2 = f2{-sneg.x, -sneg.y};
    const f2 gD = fma2(gc0, pr.q[0], fma2(gc1, pr.q[1], mul2(gc2, pr.q[2])));
    const f2 gd = mul2(mul2(gD, bc2(-p.disp_range)), bc2(D * D));
    const float gsum = gd.x + gd.y;
    st.gdu[0] += j == 0 ? gsum : 0.f; st.gdu[1] += j == 1 ? gsum : 0.f;
    st.gdu[2] += j == 2 ? gsum : 0.f; st.gdu[3] += j == 3 ? gsum : 0.f;
    const f2 D2 = bc2(D), u2 = bc2(u), v2 = bc2(v);
    const f2 w0 = mul2(gc0, D2), w1 = mul2(gc1, D2), w2 = mul2(gc2, D2);
    st.M[0] = fma2(w0, u2, st.M[0]); st.M[1] = fma2(w0, v2, st.M[1]); st.M[2] = add2(st.M[2], w0); st.M[3] = add2(st.M[3], gc0);
    st.M[4] = fma2(w1, u2, st.M[4]); st.M[5] = fma2(w1, v2, st.M[5]); st.M[6] = add2(st.M[6], w1); st.M[7] = add2(st.M[7], gc1);
    st.M[8] = fma2(w2, u2, st.M[8]); st.M[9] = fma2(w2, v2, st.M[9]); st.M[10] = add2(st.M[10], w2); st.M[11] = add2(st.M[11], gc2);
  }
}

// ------------------------------------------------------------------------------------------------ glue to the shared tail
DVS_HD void pair_store_gdu_direct(const FusedParams& p, const Tile& t, int tid, int s, const PairState& st) {
  int r0, cx;
  quad_coords(tid, r0, cx);
  for (int j = 0; j < 4; ++j)
    if ((st.flags >> (4 + j)) & 1) p.gdisp[s][((size_t)t.b * p.H + t.gy0 + r0 + j) * p.W + t.gx0 + cx] = st.gdu[j];
}
DVS_HD void pair_stage_gdu(float* sm, int tid, const PairState& st) {
  PairLayout L;
  int r0, cx;
  quad_coords(tid, r0, cx);
  for (int j = 0; j < 4; ++j) sm[L.du() + pidx(r0 + j, cx)] = ((st.flags >> (4 + j)) & 1) ? st.gdu[j] : 0.f;
}
DVS_HD void pair_reduce_write(float* sm, int tid, const PairState& st) {
  PairLayout L;
  constexpr int nv = 3 + 12 * 2;
  float* sc = sm + L.scratch() + tid * nv;
  sc[0] = st.acc[0]; sc[1] = st.acc[1]; sc[2] = st.acc[2];
  DVS_UNROLL
  for (int k = 0; k < 12; ++k) {
    sc[3 + k] = st.M[k].x;
    sc[15 + k] = st.M[k].y;
  }
}
DVS_HD void pair_reset_scale_state(PairState& st) {
  st.acc[0] = st.acc[1] = st.acc[2] = 0.f;
  DVS_UNROLL
  for (int k = 0; k < 12; ++k) st.M[k] = f2{0.f, 0.f};
}

}  // namespace dvs
