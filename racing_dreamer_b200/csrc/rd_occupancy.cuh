// rd_occupancy.cuh -- K3: the 'lidar_occupancy' observation, bit-exact with the reference.
//
// Replaces OccupancyMapObs.step [REF dreamer/wrappers.py:390-408]:
//   (pr,pc) = to_pixel(pose); crop M[pr-110:pr+110, pc-110:pc+110] as uint8;
//   scipy.ndimage.rotate(crop, rad2deg(2*pi - yaw))  -- cubic B-spline prefilter (mirror boundary) + affine
//     resampling with constant-0 outside, result rounded to uint8;
//   centre crop 200x200; PIL.Image.resize((64,64)) -- bicubic, two fixed-point passes with a uint8 image between.
//
// One CTA per env (persistent loop), everything in shared memory (no global scratch):
//   A  crop        220x220 bits of the drivable grid -> smem (funnel-shifted words)
//   B  prefilter   columns then rows (scipy's order), float32 recursion, lines cut into 2-4 segments, into a 223x227 float32
//                  image whose 1+2 border rows/columns hold the mirrored values, so the 4x4 taps need no index math
//   A' uniform map  8x8-cell blocks that are, with their 8 neighbours, all drivable / all not: pixels landing there are
//                  that constant exactly and skip the taps (about 60-70 % of the pixels of a track crop)
//   A" tile classes  every 8x4 tile of the mid image is classified by ONE lane (32 tiles per warp instruction): its four
//                  corners bound the source coordinates of its 32 pixels (each coordinate is a rounded sum of a monotone
//                  function of the row and one of the column), so a tile whose bounding blocks are all uniform with
//                  one class -- or that lies outside the crop -- is that constant for every pixel by the per-pixel
//                  rule; its plane bytes are written on the spot.  Only the mixed tiles (about half) go to stage C.
//   C  rotation    one pixel per thread, an 8x4 pixel tile per warp; source coordinates in float64 with scipy's exact operation order (they decide floor() and
//                  the inside test); value in float32; the rounded pixel is stored as two warp-ballot bit planes
//   C' exactness   a pixel whose float32 value lies within OCC_EPS of a rounding threshold (k + 0.5) is re-evaluated in
//                  float64 by its warp from the banded prefilter operator H (coef = H X H^T), which reproduces
//                  scipy's float64 result to ~1e-15.  float32 error is < 2e-6 (bound in DESIGN.md), OCC_EPS = 2e-5, and
//                  about one image in 40 has such a pixel, so the output is the reference's, bit for bit.
//   D  resize      Pillow's 22-bit fixed-point bicubic, horizontal then vertical, uint8 intermediate, in smem
#pragma once
#include "rd_common.cuh"
#include <cstring>

#ifndef OCC_THREADS
#define OCC_THREADS 640
#endif
#define OCC_ROWS 223        // 220 + mirrored border (index -1 and 220, 221)
#define OCC_PITCH 227       // row pitch in floats: odd (row pass conflict-free) and = 3 mod 32 (rotated 8x4 tiles spread over the banks)
#define OCC_XW 7            // 32-bit words per crop row
#define OCC_KSIZE 15        // Pillow: ceil(2*3.125)*2+1 taps per output pixel
#define OCC_PREC_BITS 22
#define OCC_EPS 2.0e-5f
#define OCC_BAND 40         // |H[r][k]| < 1.3e-23 beyond this distance from the diagonal
#define OCC_BANDW (2 * OCC_BAND + 1)
#define OCC_NB 28           // 8x8-cell blocks per crop side (uniform-region map)

// Pillow precompute_coeffs + normalize_coeffs_8bpc for 200 -> 64, bicubic.  25 input pixels map onto 8 output pixels
// exactly (scale 3.125), so output pixel xx + 8 has the weights of xx with its window moved by 25: apart from the two
// clipped windows at either end there are only 8 distinct weight sets -- OCC_NPH = 12 "phases" -- and the tables are
// small enough (4.4 KB) to stay in shared memory for the whole launch instead of being copied in per env.
#define OCC_NPH 12
#define OCC_LUT_ROWS (32 + 32 + 8)   // 5-bit masks of taps 0..4 and 5..9, 3-bit mask of taps 10..12 (windows are <= 13 taps)
struct OccTables {
  // horizontal pass on BINARY rows: lut[row][ph] = sum of kk[ph][5g + j] over the set bits j of the mask m of tap group g
  // (row = 32 g + m), so the <= 13-tap sum of one output pixel is three table reads on the window's bit mask (exact
  // integer arithmetic); the phase is the fastest index: 32 consecutive output pixels with the same mask read 8
  // consecutive words
  int32_t lut[OCC_LUT_ROWS * OCC_NPH];
  int32_t kk[OCC_NPH * OCC_KSIZE];
  uint8_t xmin[RD_OCC_OUT];
  uint8_t xnum[RD_OCC_OUT];
  uint8_t phase[RD_OCC_OUT];
};

struct OccScratch {
  OccTables* tables = nullptr;
  double* hband = nullptr;    // [220][OCC_BANDW] banded 1-D prefilter operator (float64)
  float eps = OCC_EPS;        // RD_OCC_EPS overrides it (tests widen it to force pixels through the exact path)
};

struct OccGeom {              // per-env geometry, computed by one thread
  double c, s, off0, off1;
  int o0_first, o1_first;     // output index of centre-crop pixel (0,0)
  int pr, pc;
  int env, mode;              // env index; 0 = compute, 1 = reset observation (zeros), 2 = leave the buffer untouched
  // float32 bounds of the source coordinates over an 8x4 tile (stage A"): c0 in [t0 + lo0, t0 + hi0] with
  // t0 = fb0 + a0*fc + b0*fs for the tile's first pixel (a0, b0), likewise c1 with t1 = fb1 - a0*fs + b0*fc.  The four
  // corners of the affine map give the extremes; OCC_TILE_MARGIN covers float32 rounding (< 1e-4 for |c| < 512) and
  // the float64 roundings of the per-pixel rule, so a tile is only ever classified conservatively.
  float fb0, fb1, fc, fs, lo0, hi0, lo1, hi1;
};
#define OCC_TILE_MARGIN 1.0e-3f
struct OccPose { double x, y, yaw; int env, mode; };

// smem carve-up (bytes)
#define OCC_SM_COEF 0
#define OCC_SM_COEF_BYTES ((OCC_ROWS * OCC_PITCH * 4 + 15) & ~15)      // 202,496
#define OCC_SM_XBITS (OCC_SM_COEF + OCC_SM_COEF_BYTES)
#define OCC_SM_XBITS_BYTES (RD_OCC_IN * OCC_XW * 4)                   // 6,160
#define OCC_SM_PLANES (OCC_SM_XBITS + OCC_SM_XBITS_BYTES)
#define OCC_SM_PLANES_BYTES (2 * (RD_OCC_MID * RD_OCC_MID / 32) * 4)  // 10,000
#define OCC_SM_EDGE (OCC_SM_PLANES + OCC_SM_PLANES_BYTES)              // near-edge bitmap, same layout as the crop bits
#define OCC_SM_EDGE_BYTES (RD_OCC_IN * OCC_XW * 4)                    // 6,160
#define OCC_SM_TLIST (OCC_SM_EDGE + OCC_SM_EDGE_BYTES)                // packed (ty, tx) of the mixed 8x4 tiles of the mid image (u16)
#define OCC_N_TILES (RD_OCC_MID * RD_OCC_MID / 32)                    // 1250
#define OCC_SM_TLIST_BYTES ((OCC_N_TILES * 2 + 15) & ~15)             // 2,512
#define OCC_SM_TAB (OCC_SM_TLIST + OCC_SM_TLIST_BYTES)                // resize tables, resident for the whole launch
#define OCC_SM_TAB_BYTES ((sizeof(OccTables) + 15) & ~15)             // 4,368
#define OCC_SM_TOTAL (OCC_SM_TAB + OCC_SM_TAB_BYTES)                  // 231,696 of 232,448 (227 KB per CTA, 64 B static)
static_assert(OCC_SM_TOTAL + 256 <= 232448, "k_occupancy shared memory (dynamic + static)");
// after the rotation the coefficient image is dead; its space holds the uint8 intermediate of the resize
#define OCC_SM_TMP 0

__device__ __forceinline__ int occ_mirror(int idx, int len) {
  if (idx < 0) idx = -idx;
  if (idx >= len) idx = 2 * len - 2 - idx;
  return idx;
}

__device__ __forceinline__ void occ_weights(double x, double (&w)[4]) {
  const double y = x - floor(x), z = 1.0 - y;
  w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
  w[0] = z * z * z / 6.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

// float32 cubic B-spline prefilter of the 220 lines of one axis (mirror boundary, pole z = sqrt(3)-2, gain 6), one WARP
// per PAIR of adjacent lines, every sample in registers, the two lines packed in float2 (FFMA2: one instruction, two
// lines):
//   scipy's recursion is  y[i] = 6 x[i] + z y[i-1]  (causal, y[0] = the mirror sum  sum_i z^i 6 x[i])  followed by
//   w[n-1] = (z y[n-2] + y[n-1]) z / (z^2 - 1),  w[i] = z (w[i+1] - y[i])  (anticausal).
// The inputs are scaled by -z (on top of the gain), so that the anticausal step is ONE fma, w[i] = z w[i+1] + y'[i]
// with y' = -z y; every quantity of the causal pass carries that factor (the recursion is linear).
// Lane l owns samples 7l-4 .. 7l+2: the line's last sample (219) is the last slot of lane 31, so the anticausal
// initialisation touches one register; the four slots in front of sample 0 (lane 0) hold zeros and are never stored.
// Each lane runs both recursions on its seven samples with a zero incoming state, and the true incoming state is
// restored from its neighbours' end values: the state decays by z^7 = -1e-4 per lane, so three neighbours (z^21 = 1e-12)
// reproduce the full-line recursion far below float32 resolution.  The mirror initialisation enters as the state c in
// front of lane 0's slot 0 that makes the causal value at sample 0 equal the mirror sum M: four zero samples later the
// state is z^4 c = (M - x0) / z.
// Element i of line L lives at img[L * LINE_STRIDE + i * ELEM_STRIDE].  FROM_BITS: the line is crop ROW L and its input
// the crop bits (L, i) instead of the image -- the first pass reads the binary crop directly, seven bits of one funnel
// shift per lane and line.  (scipy filters axis 0 first; the operator is separable, so the order only moves float32
// rounding, which the exactness stage C' bounds.)
#define OCC_LANE_N 7
#define OCC_LANE_LEAD 4
__device__ __forceinline__ float2 occ_f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 occ_shfl2(float2 v, int src) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ float2 occ_shfl2_up(float2 v) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 occ_shfl2_down(float2 v) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
template <bool FROM_BITS, int LINE_STRIDE, int ELEM_STRIDE>
__device__ __forceinline__ void occ_prefilter_axis(float* img, const uint32_t* xb) {
  constexpr int n = RD_OCC_IN;
  static_assert(n % 2 == 0 && 32 * OCC_LANE_N - OCC_LANE_LEAD == n, "220 samples = 32 lanes of 7 minus 4 leading slots");
  constexpr float z = -0.26794919243112270647f, gain = 6.0f, xg = -z * gain;
  constexpr float z2 = z * z, z3 = z2 * z, z4 = z3 * z, z5 = z4 * z, z6 = z5 * z, z7 = z6 * z;
  constexpr float zm5 = 1.0f / z5, zm9 = 1.0f / (z5 * z4);
  const float zp[OCC_LANE_N] = {z, z2, z3, z4, z5, z6, z7};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = lane * OCC_LANE_N - OCC_LANE_LEAD;
  const bool first = lane == 0, tail = lane == 31;
  const float2 Z = occ_f2(z), Z7 = occ_f2(z7), zero = occ_f2(0.0f);
  for (int L = 2 * warp; L < n; L += 2 * (OCC_THREADS / 32)) {
    float* p = img + L * LINE_STRIDE + i0 * ELEM_STRIDE;
    float2 x[OCC_LANE_N];
    if (FROM_BITS) {   // bits i0 .. i0+6 of crop rows L, L+1 (lane 0: four zero slots, then bits 0..2)
      const uint32_t* ra = xb + L * OCC_XW;
      const uint32_t* rb = ra + OCC_XW;
      const int s = first ? 0 : i0, wj = s >> 5;
      uint32_t wa = __funnelshift_r(ra[wj], wj + 1 < OCC_XW ? ra[wj + 1] : 0u, s & 31);
      uint32_t wb = __funnelshift_r(rb[wj], wj + 1 < OCC_XW ? rb[wj + 1] : 0u, s & 31);
      if (first) { wa <<= OCC_LANE_LEAD; wb <<= OCC_LANE_LEAD; }
#pragma unroll
      for (int j = 0; j < OCC_LANE_N; ++j) x[j] = make_float2(((wa >> j) & 1u) ? xg : 0.0f, ((wb >> j) & 1u) ? xg : 0.0f);
    } else {
#pragma unroll
      for (int j = 0; j < OCC_LANE_N; ++j)
        x[j] = (i0 + j >= 0) ? __fmul2_rn(make_float2(p[j * ELEM_STRIDE], p[j * ELEM_STRIDE + LINE_STRIDE]), occ_f2(xg)) : zero;
    }
    // ---- causal pass (zero incoming state) and the mirror sum ----
    float2 u[OCC_LANE_N];
    u[0] = x[0];
#pragma unroll
    for (int j = 1; j < OCC_LANE_N; ++j) u[j] = __ffma2_rn(Z, u[j - 1], x[j]);
    float2 poly = x[OCC_LANE_N - 1];
#pragma unroll
    for (int j = OCC_LANE_N - 2; j >= 0; --j) poly = __ffma2_rn(Z, poly, x[j]);
    // lanes 0..3 hold samples 0..23 behind four zeros: poly0 + z^7 (poly1 + z^7 (poly2 + z^7 poly3)) = z^4 M (z^24 = 2e-14)
    const float2 p1 = occ_shfl2(poly, 1), p2 = occ_shfl2(poly, 2), p3 = occ_shfl2(poly, 3);
    const float2 m4 = __ffma2_rn(Z7, __ffma2_rn(Z7, __ffma2_rn(Z7, p3, p2), p1), occ_shfl2(poly, 0));
    // c = (M - x0) / z^5 = z^-9 (z^4 M) - z^-5 x0
    const float2 c = __ffma2_rn(m4, occ_f2(zm9), __fmul2_rn(occ_shfl2(x[OCC_LANE_LEAD], 0), occ_f2(-zm5)));
    float2 c1 = occ_shfl2_up(u[OCC_LANE_N - 1]);
    c1 = first ? c : c1;
    float2 c2 = occ_shfl2_up(c1);
    c2 = first ? zero : c2;
    float2 c3 = occ_shfl2_up(c2);
    c3 = first ? zero : c3;
    const float2 sin_ = __ffma2_rn(Z7, __ffma2_rn(Z7, c3, c2), c1);
    float2 y[OCC_LANE_N];
#pragma unroll
    for (int j = 0; j < OCC_LANE_N; ++j) y[j] = __ffma2_rn(occ_f2(zp[j]), sin_, u[j]);
    // ---- anticausal pass: w[i] = z w[i+1] + y[i]; the line ends at slot 6 of lane 31 ----
    const float2 w_end = __fmul2_rn(__ffma2_rn(Z, y[OCC_LANE_N - 2], y[OCC_LANE_N - 1]), occ_f2(-1.0f / (z * z - 1.0f)));
    float2 v[OCC_LANE_N];
    v[OCC_LANE_N - 1] = tail ? w_end : y[OCC_LANE_N - 1];
#pragma unroll
    for (int j = OCC_LANE_N - 2; j >= 0; --j) v[j] = __ffma2_rn(Z, v[j + 1], y[j]);
    float2 d1 = occ_shfl2_down(v[0]);
    d1 = tail ? zero : d1;
    float2 d2 = occ_shfl2_down(d1);
    d2 = tail ? zero : d2;
    float2 d3 = occ_shfl2_down(d2);
    d3 = tail ? zero : d3;
    const float2 tin = __ffma2_rn(Z7, __ffma2_rn(Z7, d3, d2), d1);   // 0 for lane 31: its values are final already
#pragma unroll
    for (int j = 0; j < OCC_LANE_N; ++j) {
      const float2 o = __ffma2_rn(occ_f2(zp[OCC_LANE_N - 1 - j]), tin, v[j]);
      if (i0 + j >= 0) { p[j * ELEM_STRIDE] = o.x; p[j * ELEM_STRIDE + LINE_STRIDE] = o.y; }
    }
  }
  __syncthreads();
}

// Source coordinates of mid pixel (a, b) with scipy's operation order (shift first, then one product per output axis):
//   c0 = (off0 + o0*c) + o1*s ,  c1 = (off1 + o0*(-s)) + o1*c .
__device__ __forceinline__ void occ_coords(const OccGeom& g, int a, int b, double& c0, double& c1) {
  const double o0 = (double)(g.o0_first + a), o1 = (double)(g.o1_first + b);
  c0 = __dadd_rn(__dadd_rn(g.off0, __dmul_rn(o0, g.c)), __dmul_rn(o1, g.s));
  c1 = __dadd_rn(__dadd_rn(g.off1, __dmul_rn(o0, -g.s)), __dmul_rn(o1, g.c));
}

// Exact float64 value of one rotated pixel (a, b), computed by ONE WARP from coef = H X H^T; every lane returns the
// rounded pixel.  Lane l owns crop columns L0 + l, L0 + l + 32, L0 + l + 64 of the <= 84-column window:
//   A: T[p][col] = sum_k H[rp][k] X[k][col]      (4 tap rows x 81 band entries, X from the crop bits)
//   B: coef[p][q] = sum_col H[cq][col] T[p][col]  (warp reduction)
//   C: the 4x4 taps in scipy's accumulation order and its rounding.
// No shared scratch, no CTA barrier: about one image in 40 needs it, so its cost (~10k instructions) is irrelevant, but
// it must not make the other warps wait.
__device__ __noinline__ uint32_t occ_exact_pixel(const OccGeom& g, int a, int b, const uint32_t* xb,
                                                 const double* __restrict__ hband) {
  const int lane = threadIdx.x & 31;
  double c0, c1;
  occ_coords(g, a, b, c0, c1);
  const int s0 = (int)floor(c0) - 1, s1 = (int)floor(c1) - 1;
  int rp[4], cq[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { rp[k] = occ_mirror(s0 + k, RD_OCC_IN); cq[k] = occ_mirror(s1 + k, RD_OCC_IN); }
  const int cmin = min(min(cq[0], cq[1]), min(cq[2], cq[3])), cmax = max(max(cq[0], cq[1]), max(cq[2], cq[3]));
  const int L0 = max(0, cmin - OCC_BAND), L1 = min(RD_OCC_IN - 1, cmax + OCC_BAND);  // L1 - L0 + 1 <= 84 < 96
  double T[4][3];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = rp[p];
    const double* h = hband + (size_t)r * OCC_BANDW + (OCC_BAND - r);
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    const int k0 = max(0, r - OCC_BAND), k1 = min(RD_OCC_IN - 1, r + OCC_BAND);
    for (int k = k0; k <= k1; ++k) {
      const double hk = __ldg(h + k);
      const uint32_t* row = xb + k * OCC_XW;
      const int l0 = L0 + lane, l1 = l0 + 32, l2 = l0 + 64;
      if (l0 <= L1 && ((row[l0 >> 5] >> (l0 & 31)) & 1u)) t0 += hk;
      if (l1 <= L1 && ((row[l1 >> 5] >> (l1 & 31)) & 1u)) t1 += hk;
      if (l2 <= L1 && ((row[l2 >> 5] >> (l2 & 31)) & 1u)) t2 += hk;
    }
    T[p][0] = t0; T[p][1] = t1; T[p][2] = t2;
  }
  double cf[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = cq[q];
    const double* h = hband + (size_t)c * OCC_BANDW + (OCC_BAND - c);
    double hq[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int l = L0 + lane + 32 * j;
      hq[j] = (l <= L1 && l >= c - OCC_BAND && l <= c + OCC_BAND) ? __ldg(h + l) : 0.0;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      double sum = hq[0] * T[p][0] + hq[1] * T[p][1] + hq[2] * T[p][2];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      cf[p][q] = sum;
    }
  }
  double w0[4], w1[4];
  occ_weights(c0, w0);
  occ_weights(c1, w1);
  double t = 0.0;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = cf[p][q];
      v *= w0[p];
      v *= w1[q];
      t += v;
    }
  double tv = t > 0.0 ? t + 0.5 : 0.0;
  tv = tv > 255.0 ? 255.0 : tv;
  return (uint32_t)(uint8_t)tv;
}

// The pose (and reset mode) of the env in work slot `slot`: three dependent global loads, issued one env ahead.
__device__ __forceinline__ OccPose occ_load_pose(const OriginRec* __restrict__ recs, const double* __restrict__ poses,
                                                 const double2* __restrict__ f2, int n_state,
                                                 const int32_t* __restrict__ order, int slot) {
  OccPose p;
  p.env = order ? __ldg(order + slot) : slot;
  p.mode = 0;
  if (poses) { p.x = poses[3 * slot]; p.y = poses[3 * slot + 1]; p.yaw = poses[3 * slot + 2]; }
  else {
    const double2 xy = f2[p.env];                                   // env state groups (rd_dynamics.cuh): 0 = (x, y),
    p.x = xy.x; p.y = xy.y;
    p.yaw = f2[(size_t)2 * n_state + p.env].x;                      // 2 = (yaw, yaw_rate)
    p.mode = recs[p.env].was_reset;
  }
  return p;
}
// crop position and scipy.ndimage.rotate's output geometry for one pose [REF dreamer/wrappers.py:396-403]
__device__ __forceinline__ void occ_make_geom(const DevMap& m, const OccPose& p, OccGeom& g) {
  const int col = (int)floor((p.x - m.ox) * m.inv_res);
  const int rup = (int)floor((p.y - m.oy) * m.inv_res);
  g.pr = m.full_h - 1 - rup;
  g.pc = col;
  const double ang = 2.0 * 3.141592653589793 - p.yaw;
  double s, c;
  sincos(ang, &s, &c);
  const double N = (double)RD_OCC_IN;
  const double b0[4] = {0.0, s * N, c * N, c * N + s * N};
  const double b1[4] = {0.0, c * N, -s * N, -s * N + c * N};
  double mn0 = b0[0], mx0 = b0[0], mn1 = b1[0], mx1 = b1[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    mn0 = fmin(mn0, b0[k]); mx0 = fmax(mx0, b0[k]); mn1 = fmin(mn1, b1[k]); mx1 = fmax(mx1, b1[k]);
  }
  const int oh = (int)((mx0 - mn0) + 0.5), ow = (int)((mx1 - mn1) + 0.5);
  const double oc0 = (oh - 1) / 2.0, oc1 = (ow - 1) / 2.0, ic = (N - 1.0) / 2.0;
  g.c = c; g.s = s;
  g.off0 = ic - (c * oc0 + s * oc1);
  g.off1 = ic - (-s * oc0 + c * oc1);
  g.o0_first = oh / 2 - RD_OCC_MID / 2;
  g.o1_first = ow / 2 - RD_OCC_MID / 2;
  g.env = p.env;
  g.mode = p.mode;
  const double f0 = (double)g.o0_first, f1 = (double)g.o1_first;
  g.fb0 = (float)(g.off0 + f0 * c + f1 * s);
  g.fb1 = (float)(g.off1 - f0 * s + f1 * c);
  g.fc = (float)c;
  g.fs = (float)s;
  const float ac = 3.0f * g.fc, as = 3.0f * g.fs, bc = 7.0f * g.fc, bs = 7.0f * g.fs;   // rows a0..a0+3, columns b0..b0+7
  g.lo0 = fminf(0.0f, ac) + fminf(0.0f, bs) - OCC_TILE_MARGIN;
  g.hi0 = fmaxf(0.0f, ac) + fmaxf(0.0f, bs) + OCC_TILE_MARGIN;
  g.lo1 = fminf(0.0f, -as) + fminf(0.0f, bc) - OCC_TILE_MARGIN;
  g.hi1 = fmaxf(0.0f, -as) + fmaxf(0.0f, bc) + OCC_TILE_MARGIN;
}

__global__ void __launch_bounds__(OCC_THREADS, 1)
k_occupancy(const DevMap* __restrict__ maps, int map_id, const OriginRec* __restrict__ recs,
            const double* __restrict__ poses, const double2* __restrict__ f2, int n_state,
            const int32_t* __restrict__ order, int n_env, const double* __restrict__ hband,
            const OccTables* __restrict__ tables, float eps, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* coef = reinterpret_cast<float*>(smem + OCC_SM_COEF);
  uint32_t* xb = reinterpret_cast<uint32_t*>(smem + OCC_SM_XBITS);
  uint32_t* planes = reinterpret_cast<uint32_t*>(smem + OCC_SM_PLANES);
  uint32_t* eb = reinterpret_cast<uint32_t*>(smem + OCC_SM_EDGE);   // near-edge bitmap (see stage A')
  uint8_t* tmp = smem + OCC_SM_TMP;
  OccTables* tb = reinterpret_cast<OccTables*>(smem + OCC_SM_TAB);
  uint16_t* tlist = reinterpret_cast<uint16_t*>(smem + OCC_SM_TLIST);
  // Geometry is double-buffered and computed ONE ENV AHEAD: the pose loads are issued at the top of an iteration by the
  // first lane of the last warp, which turns them into the next env's geometry at the start of the prefilter stage (its
  // warp has one line pair less than the first ten there), so neither the global-load latency nor the float64 sincos
  // chain of a single thread holds the other 639 threads at a barrier.
  __shared__ OccGeom gbuf[2];
  __shared__ int any_hi;
  __shared__ int n_mixed;

  const DevMap& m = maps[map_id];
  const int tid = threadIdx.x, lane = tid & 31;
  const bool geom_thread = tid == OCC_THREADS - 32;
  for (int i = tid; i < (int)(sizeof(OccTables) / 4); i += OCC_THREADS)     // resident for every env of this CTA
    reinterpret_cast<int32_t*>(tb)[i] = __ldg(reinterpret_cast<const int32_t*>(tables) + i);
  if (geom_thread && (int)blockIdx.x < n_env) occ_make_geom(m, occ_load_pose(recs, poses, f2, n_state, order, blockIdx.x), gbuf[0]);
  __syncthreads();

  int it = 0;
  for (int slot = blockIdx.x; slot < n_env; slot += gridDim.x, ++it) {
    const OccGeom& geom = gbuf[it & 1];
    OccGeom& geom_next = gbuf[(it + 1) & 1];
    const bool have_next = geom_thread && slot + (int)gridDim.x < n_env;
    OccPose next_pose;
    if (have_next) next_pose = occ_load_pose(recs, poses, f2, n_state, order, slot + gridDim.x);   // in flight until stage B
    const int env = geom.env, mode = geom.mode;
    uint8_t* dst = out + (size_t)(poses ? slot : env) * (RD_OCC_OUT * RD_OCC_OUT);
    if (mode >= 1) {
      // 2: frozen / not reset by this call: leave the buffer untouched; 1: the reset observation is all zeros
      // [REF dreamer/wrappers.py:410-414]
      if (mode == 1)
        for (int i = tid; i < RD_OCC_OUT * RD_OCC_OUT / 4; i += OCC_THREADS) reinterpret_cast<uint32_t*>(dst)[i] = 0u;
      __syncthreads();                       // nobody still reads the buffer the next geometry goes into
      if (have_next) occ_make_geom(m, next_pose, geom_next);
      __syncthreads();
      continue;
    }
    __syncthreads();  // previous env's smem fully consumed
    if (tid == 0) { any_hi = 0; n_mixed = 0; }   // first used two barriers further down

    // ---- A: crop bits -> smem.  word (i, j) = crop columns 32j..32j+31 of crop row i ----
    for (int t = tid; t < RD_OCC_IN * OCC_XW; t += OCC_THREADS) {
      const int i = t / OCC_XW, j = t - i * OCC_XW;
      const int r_img = geom.pr - RD_OCC_IN / 2 + i;
      const int cy = (m.full_h - 1 - r_img) - m.row0;
      const int cx = geom.pc - RD_OCC_IN / 2 + 32 * j - m.col0;      // map column of bit 0 of this word
      uint32_t w = 0u;
      if (cy >= 0 && cy < m.h && cx > -32 && cx < m.rw * 32) {
        const uint32_t* row = m.bits + (size_t)cy * m.rw;
        const int wi = cx >> 5, sh = cx & 31;                        // arithmetic shift: floor for negative cx
        const uint32_t lo = (wi >= 0 && wi < m.rw) ? __ldg(row + wi) : 0u;
        const uint32_t hi = (wi + 1 >= 0 && wi + 1 < m.rw) ? __ldg(row + wi + 1) : 0u;
        w = __funnelshift_r(lo, hi, sh);
      }
      if (j == OCC_XW - 1) w &= (1u << (RD_OCC_IN - 32 * (OCC_XW - 1))) - 1u;  // only 28 columns in the last word
      xb[t] = w;
    }
    __syncthreads();
    // ---- A': near-edge bitmap.  Bit (i0, i1) = the 4x4 cells [i0-1, i0+2] x [i1-1, i1+2] around a source point with
    // floor coordinates (i0, i1) -- exactly the cells its cubic-spline taps touch; cells beyond the crop mirror cells of the
    // same window -- are NOT all equal.  Where the bit is clear the rotated pixel needs no arithmetic: the cardinal
    // cubic spline has |eta| mass 0.0933 outside a window of this size and total mass 1.549 per axis, so the interpolated
    // value lies within 1.549^2 - (1.549 - 0.0933)^2 = 0.28 of the window's constant c in {0, 1} and rounds to c
    // ((uint8)(v + 0.5) = 1 for v in [0.72, 1.28], 0 for v <= 0.28).  Only a band of +-2 cells along the track walls has
    // the bit set.  Row stage: AND / OR over columns c-1 .. c+2; column stage: over rows r-1 .. r+2. ----
    // Two passes (the reductions are separable); the row results live in the not-yet-written coefficient image.
    {
      uint32_t* row_all = reinterpret_cast<uint32_t*>(coef);
      uint32_t* row_any = row_all + RD_OCC_IN * OCC_XW;
      const uint32_t last_mask = (1u << (RD_OCC_IN - 32 * (OCC_XW - 1))) - 1u;       // 28 valid columns in the last word
      for (int t = tid; t < RD_OCC_IN * OCC_XW; t += OCC_THREADS) {
        const int i = t / OCC_XW, j = t - i * OCC_XW;
        const uint32_t* row = xb + i * OCC_XW;
        const uint32_t w = row[j];
        const uint32_t lo = j > 0 ? row[j - 1] : 0u, hi = j < OCC_XW - 1 ? row[j + 1] : 0u;
        // neighbours of column c: c-1 (bit c of w << 1 | carry), c+1, c+2
        const uint32_t m1 = (w << 1) | (lo >> 31), p1 = (w >> 1) | (hi << 31), p2 = (w >> 2) | (hi << 30);
        // columns outside [0, 219] are not part of the window: neutral element of each reduction there
        const uint32_t vw = j == OCC_XW - 1 ? last_mask : 0xffffffffu;                // valid columns of this word
        const uint32_t vm1 = j == 0 ? 0xfffffffeu : 0xffffffffu;                      // column -1 does not exist
        const uint32_t vp1 = j == OCC_XW - 1 ? (last_mask >> 1) : 0xffffffffu;        // column 220 does not exist
        const uint32_t vp2 = j == OCC_XW - 1 ? (last_mask >> 2) : 0xffffffffu;        // columns 220, 221 do not exist
        row_all[t] = (w | ~vw) & (m1 | ~vm1) & (p1 | ~vp1) & (p2 | ~vp2);
        row_any[t] = (w & vw) | (m1 & vm1) | (p1 & vp1) | (p2 & vp2);
      }
      __syncthreads();
      for (int t = tid; t < RD_OCC_IN * OCC_XW; t += OCC_THREADS) {
        const int i = t / OCC_XW, j = t - i * OCC_XW;
        uint32_t all1 = 0xffffffffu, any1 = 0u;
#pragma unroll
        for (int dr = -1; dr <= 2; ++dr) {
          const int r = i + dr;
          if (r < 0 || r >= RD_OCC_IN) continue;                                     // clipped window
          all1 &= row_all[t + dr * OCC_XW];
          any1 |= row_any[t + dr * OCC_XW];
        }
        uint32_t e = any1 & ~all1;
        if (j == OCC_XW - 1) e &= last_mask;
        eb[t] = e;
      }
    }
    __syncthreads();
    // ---- A": tile classes (see the header).  One lane per tile; uniform / outside tiles get their plane bytes here,
    // mixed tiles are appended to tlist (order irrelevant). ----
    {
      uint8_t* plane0 = reinterpret_cast<uint8_t*>(planes);
      uint8_t* plane1 = plane0 + RD_OCC_MID * RD_OCC_MID / 8;
      for (int base = 0; base < OCC_N_TILES; base += OCC_THREADS) {
        const int t = base + tid;
        int cls_t = 2;
        const int ty = t / (RD_OCC_MID / 8), tx = t - ty * (RD_OCC_MID / 8);
        if (t < OCC_N_TILES) {
          const int a0 = 4 * ty, b0 = 8 * tx;
          const float fa = (float)a0, fb = (float)b0;
          const float t0 = fmaf(fa, geom.fc, fmaf(fb, geom.fs, geom.fb0));
          const float t1 = fmaf(-fa, geom.fs, fmaf(fb, geom.fc, geom.fb1));
          const float lo0 = t0 + geom.lo0, hi0 = t0 + geom.hi0, lo1 = t1 + geom.lo1, hi1 = t1 + geom.hi1;
          const float top = (float)(RD_OCC_IN - 1);
          if (hi0 < 0.0f || lo0 > top || hi1 < 0.0f || lo1 > top) {
            cls_t = 0;                                             // every pixel outside the crop: 0
          } else if (lo0 >= 0.0f && hi0 <= top && lo1 >= 0.0f && hi1 <= top) {
            // every pixel of the tile has its floor coordinates in [r0, r1] x [q0, q1] (<= 9 x 9 cells): no near-edge bit
            // there means every pixel is the crop bit of its own cell, and those cells are all equal
            const int r0 = (int)floorf(lo0), r1 = (int)floorf(hi0), q0 = (int)floorf(lo1), q1 = (int)floorf(hi1);
            const uint32_t cmask = (q1 - q0 + 1 >= 32) ? 0xffffffffu : ((1u << (q1 - q0 + 1)) - 1u);
            const int wj = q0 >> 5, sh = q0 & 31;
            uint32_t any_e = 0u;
            for (int r = r0; r <= r1; ++r) {
              const uint32_t* row = eb + r * OCC_XW;
              any_e |= __funnelshift_r(row[wj], wj + 1 < OCC_XW ? row[wj + 1] : 0u, sh) & cmask;
            }
            if (!any_e) cls_t = (int)((xb[r0 * OCC_XW + wj] >> sh) & 1u);
          }
          if (cls_t != 2) {
            const uint8_t v = cls_t ? 0xFF : 0x00;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int byte = ((a0 + r) * RD_OCC_MID + b0) >> 3;
              plane0[byte] = v;
              plane1[byte] = 0;
            }
          }
        }
        const uint32_t mixed = __ballot_sync(0xffffffffu, cls_t == 2 && t < OCC_N_TILES);
        int pos = 0;
        if (lane == 0 && mixed) pos = atomicAdd(&n_mixed, __popc(mixed));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (cls_t == 2 && t < OCC_N_TILES) tlist[pos + __popc(mixed & ((1u << lane) - 1u))] = (uint16_t)((ty << 8) | tx);   // ty < 50, tx < 25
      }
    }
    // ---- B: prefilter, axis 0 (columns) then axis 1 (rows), float32; image stored at padded index (r+1, c+1) ----
    __syncthreads();
    if (have_next) occ_make_geom(m, next_pose, geom_next);                   // see gbuf
    occ_prefilter_axis<true, OCC_PITCH, 1>(coef + OCC_PITCH + 1, xb);        // lines = crop rows, from the bits
    occ_prefilter_axis<false, 1, OCC_PITCH>(coef + OCC_PITCH + 1, nullptr);  // lines = columns, in place
    // mirrored border: columns -1, 220, 221 of rows 0..219, then rows -1, 220, 221 of all 223 columns
    for (int t = tid; t < RD_OCC_IN * 3; t += OCC_THREADS) {
      const int r = t / 3, k = t - r * 3;
      float* row = coef + (r + 1) * OCC_PITCH;
      if (k == 0) row[0] = row[2];                                   // col -1 <- col 1
      else if (k == 1) row[RD_OCC_IN + 1] = row[RD_OCC_IN - 1];      // col 220 <- col 218
      else row[RD_OCC_IN + 2] = row[RD_OCC_IN - 2];                  // col 221 <- col 217
    }
    __syncthreads();
    for (int t = tid; t < OCC_ROWS * 3; t += OCC_THREADS) {
      const int c = t / 3, k = t - c * 3;
      if (k == 0) coef[c] = coef[2 * OCC_PITCH + c];                                        // row -1 <- row 1
      else if (k == 1) coef[(RD_OCC_IN + 1) * OCC_PITCH + c] = coef[(RD_OCC_IN - 1) * OCC_PITCH + c];  // 220 <- 218
      else coef[(RD_OCC_IN + 2) * OCC_PITCH + c] = coef[(RD_OCC_IN - 2) * OCC_PITCH + c];              // 221 <- 217
    }
    __syncthreads();

    // ---- C: rotation.  One pixel per thread and round; a warp owns an 8-wide x 4-high tile of the mid image ----
    const int n_pix = RD_OCC_MID * RD_OCC_MID;
    const int warps = OCC_THREADS / 32;
    uint8_t* plane0 = reinterpret_cast<uint8_t*>(planes);
    uint8_t* plane1 = plane0 + n_pix / 8;
    const int n_list = n_mixed;
    uint32_t hi_acc = 0u;
    for (int k = tid >> 5; k < n_list; k += warps) {              // the mixed tiles only
      const int tile = tlist[k];
      const int ty = tile >> 8, tx = tile & 255;
      {
        const int a = ty * 4 + (lane >> 3), b = tx * 8 + (lane & 7);
        double c0, c1;
        occ_coords(geom, a, b, c0, c1);
        uint32_t val = 0u;
        bool ambiguous = false;
        const bool inside = !(c0 < 0.0 || c0 > (double)(RD_OCC_IN - 1) || c1 < 0.0 || c1 > (double)(RD_OCC_IN - 1));
        const double f0 = floor(c0), f1 = floor(c1);
        const int i0 = inside ? (int)f0 : 0, i1 = inside ? (int)f1 : 0;
        const int wpos = i0 * OCC_XW + (i1 >> 5);
        const bool near_edge = inside && ((eb[wpos] >> (i1 & 31)) & 1u);
        if (!near_edge) {
          val = inside ? ((xb[wpos] >> (i1 & 31)) & 1u) : 0u;        // constant 4x4 neighbourhood (or outside the crop: 0)
        } else {
          const float y0 = (float)(c0 - f0), y1 = (float)(c1 - f1);
          const float z0 = 1.0f - y0, z1 = 1.0f - y1;
          const float y02 = y0 * y0, z02 = z0 * z0, y12 = y1 * y1, z12 = z1 * z1;
          float w0[4], w1[4];
          // cubic B-spline weights: (3t^3 - 6t^2 + 4)/6 = t^2 (t/2 - 1) + 2/3 ;  t^3/6
          w0[1] = fmaf(y02, fmaf(0.5f, y0, -1.0f), 2.0f / 3.0f);
          w0[2] = fmaf(z02, fmaf(0.5f, z0, -1.0f), 2.0f / 3.0f);
          w0[0] = z02 * z0 * (1.0f / 6.0f);
          w0[3] = y02 * y0 * (1.0f / 6.0f);
          w1[1] = fmaf(y12, fmaf(0.5f, y1, -1.0f), 2.0f / 3.0f);
          w1[2] = fmaf(z12, fmaf(0.5f, z1, -1.0f), 2.0f / 3.0f);
          w1[0] = z12 * z1 * (1.0f / 6.0f);
          w1[3] = y12 * y1 * (1.0f / 6.0f);
          // taps rows f0-1..f0+2 -> padded rows f0..f0+3; same for columns
          const float* base = coef + i0 * OCC_PITCH + i1;
          float v = 0.0f;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float* row = base + p * OCC_PITCH;
            float r = row[0] * w1[0];
            r = fmaf(row[1], w1[1], r);
            r = fmaf(row[2], w1[2], r);
            r = fmaf(row[3], w1[3], r);
            v = fmaf(r, w0[p], v);
          }
          const float t = v + 0.5f;
          const float k = floorf(t);
          const float d = t - k;
          val = t >= 1.0f ? (uint32_t)fminf(k, 3.0f) : 0u;      // (uint8)(v + 0.5) for v > 0, else 0; provably <= 3
          ambiguous = (d < eps || d > 1.0f - eps) && t > 0.5f;    // too close to a rounding threshold for float32
        }
        // float64 re-evaluation of the (rare) ambiguous pixels, one after the other, by the whole warp
        for (uint32_t todo = __ballot_sync(0xffffffffu, ambiguous); todo; todo &= todo - 1u) {
          const int src = __ffs(todo) - 1;
          const uint32_t exact = occ_exact_pixel(geom, __shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src), xb, hband);
          if (lane == src) val = exact > 3u ? 3u : exact;
        }
        // tile row i (8 pixels) is one byte of the row-major bit planes
        const uint32_t b0 = __ballot_sync(0xffffffffu, val & 1u), b1 = __ballot_sync(0xffffffffu, val & 2u);
        if ((lane & 7) == 0) {
          const int byte = (a * RD_OCC_MID + tx * 8) >> 3;
          plane0[byte] = (uint8_t)(b0 >> (lane & 24));
          plane1[byte] = (uint8_t)(b1 >> (lane & 24));
        }
        hi_acc |= b1;
      }
    }
    if (hi_acc && lane == 0) any_hi = 1;
    __syncthreads();

    // ---- D: Pillow bicubic 200 -> 64.  The coefficient image is dead: its space takes the uint8 intermediate.
    // Horizontal pass straight on the bit planes: the <= 13-tap window of an output pixel is a bit mask, its fixed-point
    // sum three table reads (plane 1, the "value >= 2" plane, is empty for every real map; if any pixel set it, its
    // windows are added with weight 2).  A thread keeps its output column: window position and phase are loop constants.
    {
      static_assert(OCC_THREADS % RD_OCC_OUT == 0, "one output column per thread");
      const bool hi_plane = any_hi != 0;
      const int xx = tid & (RD_OCC_OUT - 1);
      const int x0 = tb->xmin[xx], xn = tb->xnum[xx];
      const uint32_t mask = (1u << xn) - 1u;
      const int32_t* lut = tb->lut + tb->phase[xx];
      for (int yy = tid / RD_OCC_OUT; yy < RD_OCC_MID; yy += OCC_THREADS / RD_OCC_OUT) {
        const int bitpos = yy * RD_OCC_MID + x0;
        const uint32_t m0 = __funnelshift_r(planes[bitpos >> 5], planes[(bitpos >> 5) + 1], bitpos & 31) & mask;
        int32_t ss = (1 << (OCC_PREC_BITS - 1)) + lut[(m0 & 31u) * OCC_NPH] + lut[(32 + ((m0 >> 5) & 31u)) * OCC_NPH] +
                     lut[(64 + (m0 >> 10)) * OCC_NPH];
        if (hi_plane) {
          const uint32_t* p1 = planes + (n_pix >> 5);
          const uint32_t m1 = __funnelshift_r(p1[bitpos >> 5], p1[min((bitpos >> 5) + 1, (n_pix >> 5) - 1)], bitpos & 31) & mask;
          ss += 2 * (lut[(m1 & 31u) * OCC_NPH] + lut[(32 + ((m1 >> 5) & 31u)) * OCC_NPH] + lut[(64 + (m1 >> 10)) * OCC_NPH]);
        }
        ss >>= OCC_PREC_BITS;
        tmp[yy * RD_OCC_OUT + xx] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
      }
    }
    __syncthreads();
    // vertical pass: four adjacent output pixels per thread (one 32-bit read of the uint8 intermediate per tap, one
    // 32-bit store of the result)
    for (int i = tid; i < RD_OCC_OUT * RD_OCC_OUT / 4; i += OCC_THREADS) {
      const int yy = i / (RD_OCC_OUT / 4), x4 = i - yy * (RD_OCC_OUT / 4);
      int32_t s0 = 1 << (OCC_PREC_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
      const int y0 = tb->xmin[yy], yn = tb->xnum[yy];
      const int32_t* k = tb->kk + tb->phase[yy] * OCC_KSIZE;
      const uint32_t* col = reinterpret_cast<const uint32_t*>(tmp) + y0 * (RD_OCC_OUT / 4) + x4;
      for (int t = 0; t < yn; ++t) {
        const uint32_t w = col[t * (RD_OCC_OUT / 4)];
        const int32_t kt = k[t];
        s0 += (int32_t)(w & 255u) * kt;
        s1 += (int32_t)((w >> 8) & 255u) * kt;
        s2 += (int32_t)((w >> 16) & 255u) * kt;
        s3 += (int32_t)(w >> 24) * kt;
      }
      s0 >>= OCC_PREC_BITS; s1 >>= OCC_PREC_BITS; s2 >>= OCC_PREC_BITS; s3 >>= OCC_PREC_BITS;
      const uint32_t b0 = (uint32_t)(s0 < 0 ? 0 : (s0 > 255 ? 255 : s0)), b1 = (uint32_t)(s1 < 0 ? 0 : (s1 > 255 ? 255 : s1));
      const uint32_t b2 = (uint32_t)(s2 < 0 ? 0 : (s2 > 255 ? 255 : s2)), b3 = (uint32_t)(s3 < 0 ? 0 : (s3 > 255 ? 255 : s3));
      reinterpret_cast<uint32_t*>(dst)[i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
  }
}

// ---- host side ----
static inline double occ_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// returns false if the resize geometry does not have the expected structure (cannot happen for 200 -> 64)
static inline bool occ_build_tables(OccTables& t) {
  const int in_size = RD_OCC_MID, out_size = RD_OCC_OUT;
  const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  double k[OCC_KSIZE];
  int32_t kk[RD_OCC_OUT][OCC_KSIZE];
  int n_phase = 0;
  std::memset(&t, 0, sizeof(t));
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale, ss = 1.0 / filterscale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) { double w = occ_bicubic((x + xmin - center + 0.5) * ss); k[x] = w; ww += w; }
    for (x = 0; x < xmax; ++x) if (ww != 0.0) k[x] /= ww;
    for (; x < OCC_KSIZE; ++x) k[x] = 0;
    for (x = 0; x < OCC_KSIZE; ++x)
      kk[xx][x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << OCC_PREC_BITS)) : (int32_t)(0.5 + k[x] * (1 << OCC_PREC_BITS));
    if (xmax > 13 || xmin > 255) return false;
    t.xmin[xx] = (uint8_t)xmin;
    t.xnum[xx] = (uint8_t)xmax;
    // phase = index of the first output pixel with the same window length and weights
    int ph = -1;
    for (int q = 0; q < xx && ph < 0; ++q)
      if (t.xnum[q] == t.xnum[xx] && std::memcmp(kk[q], kk[xx], sizeof(kk[q])) == 0) ph = t.phase[q];
    if (ph < 0) {
      if (n_phase == OCC_NPH) return false;
      ph = n_phase++;
      for (x = 0; x < OCC_KSIZE; ++x) t.kk[ph * OCC_KSIZE + x] = kk[xx][x];
      for (int g = 0; g < 3; ++g)
        for (int m = 0; m < (g < 2 ? 32 : 8); ++m) {
          int32_t acc = 0;
          for (int j = 0; j < 5; ++j)
            if ((m >> j) & 1) acc += kk[xx][5 * g + j];
          t.lut[(g * 32 + m) * OCC_NPH + ph] = acc;
        }
    }
    t.phase[xx] = (uint8_t)ph;
  }
  return true;
}

// float64 1-D prefilter (scipy's recursion: gain 6, pole sqrt(3)-2, mirror boundary), host copy used to tabulate H
static inline void occ_prefilter_host(double* p, int n) {
  const double z = -0.26794919243112270647, gain = 6.0;
  double zi = 1.0, acc = 0.0;
  for (int i = 0; i < 48 && i < n; ++i) { acc += zi * (gain * p[i]); zi *= z; }
  double prev = acc;
  p[0] = prev;
  for (int i = 1; i < n; ++i) { prev = gain * p[i] + z * prev; p[i] = prev; }
  double cur = (z * p[n - 2] + p[n - 1]) * z / (z * z - 1.0);
  p[n - 1] = cur;
  for (int i = n - 2; i >= 0; --i) { cur = z * (cur - p[i]); p[i] = cur; }
}

// hband[r][k - r + OCC_BAND] = H[r][k], H = the prefilter as a linear operator on length-220 lines (column k of H is
// the prefilter's response to the unit impulse e_k)
static inline void occ_build_hband(std::vector<double>& hb) {
  const int n = RD_OCC_IN;
  hb.assign((size_t)n * OCC_BANDW, 0.0);
  std::vector<double> e(n);
  for (int k = 0; k < n; ++k) {
    std::fill(e.begin(), e.end(), 0.0);
    e[k] = 1.0;
    occ_prefilter_host(e.data(), n);
    for (int r = std::max(0, k - OCC_BAND); r <= std::min(n - 1, k + OCC_BAND); ++r) hb[(size_t)r * OCC_BANDW + (k - r + OCC_BAND)] = e[r];
  }
}

static inline void occ_free(OccScratch& sc) {
  cudaFree(sc.tables);
  cudaFree(sc.hband);
  sc = OccScratch{};
}

// returns a cudaError_t value (0 = ok)
static inline int occ_launch(OccScratch& sc, const DevMap* d_maps, int map_id, const DevMap& hm, const OriginRec* recs,
                             const double* poses, const double2* f2, int n_state, const int32_t* order, int n_env,
                             uint8_t* out, int sm_count, cudaStream_t s, int64_t* launches) {
  (void)hm;
  const size_t smem = OCC_SM_TOTAL;
  cudaError_t e = cudaFuncSetAttribute(k_occupancy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (!sc.tables) {
    OccTables t;
    if (!occ_build_tables(t)) return (int)cudaErrorInvalidValue;
    e = cudaMalloc(&sc.tables, sizeof(OccTables));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(sc.tables, &t, sizeof(OccTables), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
  }
  if (!sc.hband) {
    if (const char* ev = std::getenv("RD_OCC_EPS")) { const float v = (float)std::atof(ev); if (v >= OCC_EPS && v <= 0.49f) sc.eps = v; }
    std::vector<double> hb;
    occ_build_hband(hb);
    e = cudaMalloc(&sc.hband, sizeof(double) * hb.size());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(sc.hband, hb.data(), sizeof(double) * hb.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = n_env < sm_count ? n_env : sm_count;  // one CTA per SM (221 KB of shared memory each)
  if (grid < 1) return 0;
  k_occupancy<<<grid, OCC_THREADS, smem, s>>>(d_maps, map_id, recs, poses, f2, n_state, order, n_env, sc.hband, sc.tables, sc.eps, out);
  (*launches)++;
  return (int)cudaGetLastError();
}
