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
//   C  rotation    one pixel per thread, an 8x4 pixel tile per warp; source coordinates in float64 with scipy's exact operation order (they decide floor() and
//                  the inside test); value in float32; the rounded pixel is stored as two warp-ballot bit planes
//   C' exactness   a pixel whose float32 value lies within OCC_EPS of a rounding threshold (k + 0.5) is re-evaluated in
//                  float64 by its warp from the banded prefilter operator H (coef = H X H^T), which reproduces
//                  scipy's float64 result to ~1e-15.  float32 error is < 2e-6 (bound in DESIGN.md), OCC_EPS = 2e-5, and
//                  about one image in 40 has such a pixel, so the output is the reference's, bit for bit.
//   D  resize      Pillow's 22-bit fixed-point bicubic, horizontal then vertical, uint8 intermediate, in smem
#pragma once
#include "rd_common.cuh"

#ifndef OCC_THREADS
#define OCC_THREADS 512
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

struct OccTables {            // Pillow precompute_coeffs + normalize_coeffs_8bpc for 200 -> 64, bicubic
  int32_t kk[RD_OCC_OUT * OCC_KSIZE];
  int32_t xmin[RD_OCC_OUT];
  int32_t xnum[RD_OCC_OUT];
  // horizontal pass on BINARY rows: lut[g][m][xx] = sum of kk[xx][5g + j] over the set bits j of the 5-bit mask m, so the
  // 15-tap sum of one output pixel is three table reads on the window's bit mask (exact integer arithmetic); xx is the
  // fastest index so that a warp (32 consecutive xx) reads 32 consecutive words when its lanes see the same mask
  int32_t lut[3 * 32 * RD_OCC_OUT];
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
};

// smem carve-up (bytes)
#define OCC_SM_COEF 0
#define OCC_SM_COEF_BYTES ((OCC_ROWS * OCC_PITCH * 4 + 15) & ~15)      // 202,496
#define OCC_SM_XBITS (OCC_SM_COEF + OCC_SM_COEF_BYTES)
#define OCC_SM_XBITS_BYTES (RD_OCC_IN * OCC_XW * 4)                   // 6,160
#define OCC_SM_PLANES (OCC_SM_XBITS + OCC_SM_XBITS_BYTES)
#define OCC_SM_PLANES_BYTES (2 * (RD_OCC_MID * RD_OCC_MID / 32) * 4)  // 10,000
#define OCC_SM_RC (OCC_SM_PLANES + OCC_SM_PLANES_BYTES)               // per-row / per-column coordinate terms (float64); 16-byte aligned
#define OCC_SM_RC_BYTES (4 * RD_OCC_MID * 8)                          // 6,400
#define OCC_SM_UNI (OCC_SM_RC + OCC_SM_RC_BYTES)                      // block class maps (2 x 28 x 28 bytes)
#define OCC_SM_UNI_BYTES (2 * OCC_NB * OCC_NB)                        // 1,568
#define OCC_SM_TOTAL (OCC_SM_UNI + OCC_SM_UNI_BYTES)                  // 226,624 of 232,448
// after the rotation the coefficient image is dead; its space holds the uint8 images and tables of the resize
#define OCC_SM_TMP 0
#define OCC_SM_TAB (OCC_SM_TMP + RD_OCC_MID * RD_OCC_OUT)

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

// float32 cubic B-spline prefilter of the 220 lines of one axis (mirror boundary, pole z = sqrt(3)-2, gain 6), in place,
// by the whole CTA.  Element i of line L lives at img[L * line_stride + i * elem_stride].  FROM_BITS: the input is the
// crop bit (i, L) instead of the image (the column pass reads the binary crop directly).
// Each line is cut into OCC_NSEG segments handled by different threads: the recursion's memory decays like |z|^k, so a
// segment that starts 24 samples early from zero reproduces the full-line recursion to |z|^24 = 2e-14 relative -- far
// below float32 resolution -- and the serial chain per thread is 4x (2x) shorter.  Warm-ups only read; barriers
// separate them from the in-place main loops.
#define OCC_NSEG (OCC_THREADS / 256)
#define OCC_SEG (RD_OCC_IN / OCC_NSEG)
#define OCC_WARM 24
template <bool FROM_BITS>
__device__ __forceinline__ void occ_prefilter_axis(float* img, int line_stride, int elem_stride, const uint32_t* xb) {
  const int n = RD_OCC_IN;
  const float z = -0.26794919243112270647f, gain = 6.0f;
  const int tid = threadIdx.x;
  const int seg = tid / RD_OCC_IN, L = tid - seg * RD_OCC_IN;
  const bool active = seg < OCC_NSEG;
  const int i0 = seg * OCC_SEG, i1 = i0 + OCC_SEG;
  float* p = img + L * line_stride;
  auto in = [&](int i) -> float {
    if (FROM_BITS) return (float)((xb[i * OCC_XW + (L >> 5)] >> (L & 31)) & 1u);
    return p[i * elem_stride];
  };
  float prev = 0.0f;
  if (active) {
    if (seg == 0) {  // scipy's mirror initialisation: sum_i z^i c[i] (z^24 ~ 2e-14 truncation)
      float zi = 1.0f;
#pragma unroll 4
      for (int i = 0; i < OCC_WARM; ++i) { prev = fmaf(zi, gain * in(i), prev); zi *= z; }
    } else {
#pragma unroll 4
      for (int i = i0 - OCC_WARM; i < i0; ++i) prev = fmaf(z, prev, gain * in(i));
    }
  }
  __syncthreads();
  if (active) {
    int i = i0;
    if (seg == 0) { p[0] = prev; i = 1; }
#pragma unroll 4
    for (; i < i1; ++i) {
      prev = fmaf(z, prev, gain * in(i));
      p[i * elem_stride] = prev;
    }
  }
  __syncthreads();
  float cur = 0.0f;
  if (active) {
    if (seg == OCC_NSEG - 1) {
      cur = (z * p[(n - 2) * elem_stride] + p[(n - 1) * elem_stride]) * (z / (z * z - 1.0f));
    } else {
#pragma unroll 4
      for (int i = i1 + OCC_WARM - 1; i >= i1; --i) cur = z * (cur - p[i * elem_stride]);
    }
  }
  __syncthreads();
  if (active) {
    int i = i1 - 1;
    if (seg == OCC_NSEG - 1) { p[(n - 1) * elem_stride] = cur; i = n - 2; }
#pragma unroll 4
    for (; i >= i0; --i) {
      cur = z * (cur - p[i * elem_stride]);
      p[i * elem_stride] = cur;
    }
  }
  __syncthreads();
}

// Source coordinates of mid pixel (a, b) with scipy's operation order (shift first, then one product per output axis):
//   c0 = (off0 + o0*c) + o1*s ,  c1 = (off1 + o0*(-s)) + o1*c .
// The per-row terms (off + o0*..) and per-column terms (o1*..) are tabulated once per env: rc[0..199] row term of c0,
// rc[200..399] row term of c1, rc[400..599] column term of c0, rc[600..799] column term of c1 -- same operations, same bits.
__device__ __forceinline__ void occ_coord_tables(const OccGeom& g, double* rc, int i) {
  const double o0 = (double)(g.o0_first + i), o1 = (double)(g.o1_first + i);
  rc[i] = __dadd_rn(g.off0, __dmul_rn(o0, g.c));
  rc[RD_OCC_MID + i] = __dadd_rn(g.off1, __dmul_rn(o0, -g.s));
  rc[2 * RD_OCC_MID + i] = __dmul_rn(o1, g.s);
  rc[3 * RD_OCC_MID + i] = __dmul_rn(o1, g.c);
}
__device__ __forceinline__ void occ_coords(const double* rc, int a, int b, double& c0, double& c1) {
  c0 = __dadd_rn(rc[a], rc[2 * RD_OCC_MID + b]);
  c1 = __dadd_rn(rc[RD_OCC_MID + a], rc[3 * RD_OCC_MID + b]);
}

// Exact float64 value of one rotated pixel (a, b), computed by ONE WARP from coef = H X H^T; every lane returns the
// rounded pixel.  Lane l owns crop columns L0 + l, L0 + l + 32, L0 + l + 64 of the <= 84-column window:
//   A: T[p][col] = sum_k H[rp][k] X[k][col]      (4 tap rows x 81 band entries, X from the crop bits)
//   B: coef[p][q] = sum_col H[cq][col] T[p][col]  (warp reduction)
//   C: the 4x4 taps in scipy's accumulation order and its rounding.
// No shared scratch, no CTA barrier: about one image in 40 needs it, so its cost (~10k instructions) is irrelevant, but
// it must not make the other warps wait.
__device__ __noinline__ uint32_t occ_exact_pixel(const double* rc, int a, int b, const uint32_t* xb,
                                                 const double* __restrict__ hband) {
  const int lane = threadIdx.x & 31;
  double c0, c1;
  occ_coords(rc, a, b, c0, c1);
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

__global__ void __launch_bounds__(OCC_THREADS, 1)
k_occupancy(const DevMap* __restrict__ maps, int map_id, const OriginRec* __restrict__ recs,
            const double* __restrict__ poses, const double2* __restrict__ f2, int n_state,
            const int32_t* __restrict__ order, int n_env, const double* __restrict__ hband,
            const OccTables* __restrict__ tables, float eps, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* coef = reinterpret_cast<float*>(smem + OCC_SM_COEF);
  uint32_t* xb = reinterpret_cast<uint32_t*>(smem + OCC_SM_XBITS);
  uint32_t* planes = reinterpret_cast<uint32_t*>(smem + OCC_SM_PLANES);
  double* rc = reinterpret_cast<double*>(smem + OCC_SM_RC);
  uint8_t* cls = smem + OCC_SM_UNI;                 // per 8x8 block: 0 all non-drivable, 1 all drivable, 2 mixed
  uint8_t* uni = cls + OCC_NB * OCC_NB;             // same, but only if the 8 neighbouring blocks agree (else 2)
  uint8_t* tmp = smem + OCC_SM_TMP;
  OccTables* tb = reinterpret_cast<OccTables*>(smem + OCC_SM_TAB);
  __shared__ OccGeom geom;
  __shared__ int any_hi;

  const DevMap& m = maps[map_id];
  const int tid = threadIdx.x, lane = tid & 31;

  for (int slot = blockIdx.x; slot < n_env; slot += gridDim.x) {
    const int env = order ? __ldg(order + slot) : slot;
    double x, y, yaw;
    int mode = 0;
    if (poses) { x = poses[3 * slot]; y = poses[3 * slot + 1]; yaw = poses[3 * slot + 2]; }
    else {
      const double2 xy = f2[env];                                   // env state groups (rd_dynamics.cuh): 0 = (x, y),
      x = xy.x; y = xy.y;
      yaw = f2[(size_t)2 * n_state + env].x;                        // 2 = (yaw, yaw_rate)
      mode = recs[env].was_reset;
    }
    uint8_t* dst = out + (size_t)(poses ? slot : env) * (RD_OCC_OUT * RD_OCC_OUT);
    if (mode >= 2) continue;                 // frozen / not reset by this call: leave the buffer untouched
    if (mode == 1) {                         // reset observation is all zeros [REF dreamer/wrappers.py:410-414]
      for (int i = tid; i < RD_OCC_OUT * RD_OCC_OUT / 4; i += OCC_THREADS) reinterpret_cast<uint32_t*>(dst)[i] = 0u;
      continue;
    }
    __syncthreads();  // previous env's smem fully consumed
    if (tid == 0) {
      const int col = (int)floor((x - m.ox) * m.inv_res);
      const int rup = (int)floor((y - m.oy) * m.inv_res);
      geom.pr = m.full_h - 1 - rup;
      geom.pc = col;
      const double ang = 2.0 * 3.141592653589793 - yaw;
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
      geom.c = c; geom.s = s;
      geom.off0 = ic - (c * oc0 + s * oc1);
      geom.off1 = ic - (-s * oc0 + c * oc1);
      geom.o0_first = oh / 2 - RD_OCC_MID / 2;
      geom.o1_first = ow / 2 - RD_OCC_MID / 2;
      any_hi = 0;
    }
    __syncthreads();
    if (tid < RD_OCC_MID) occ_coord_tables(geom, rc, tid);

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
    // ---- A': uniform-region map.  A rotated pixel whose source cell lies in a block that is, together with its eight
    // neighbours, entirely drivable (or entirely not) sees a constant crop within >= 8 cells; the spline there equals
    // that constant to < 1e-3 (|pole|^7 tail), so the pixel is exactly that constant and needs no taps. ----
    for (int t = tid; t < OCC_NB * OCC_NB; t += OCC_THREADS) {
      const int bi = t / OCC_NB, bj = t - bi * OCC_NB;
      const uint32_t valid = bj == OCC_NB - 1 ? 0x0Fu : 0xFFu;   // crop columns 220..223 do not exist
      uint32_t all_and = valid, all_or = 0u;
      for (int rr = 0; rr < 8; ++rr) {
        const int r = 8 * bi + rr;
        if (r >= RD_OCC_IN) break;
        const uint32_t byte = (xb[r * OCC_XW + (bj >> 2)] >> ((bj & 3) * 8)) & valid;
        all_and &= byte;
        all_or |= byte;
      }
      cls[t] = all_or == 0u ? 0 : (all_and == valid ? 1 : 2);
    }
    __syncthreads();
    for (int t = tid; t < OCC_NB * OCC_NB; t += OCC_THREADS) {
      const int bi = t / OCC_NB, bj = t - bi * OCC_NB;
      const uint8_t c = cls[t];
      bool same = c != 2;
      for (int di = -1; di <= 1 && same; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
          const int i = bi + di, j = bj + dj;
          if (i < 0 || i >= OCC_NB || j < 0 || j >= OCC_NB) continue;   // beyond the crop the prefilter mirrors the block itself
          if (cls[i * OCC_NB + j] != c) same = false;
        }
      uni[t] = same ? c : 2;
    }
    // ---- B: prefilter, axis 0 (columns) then axis 1 (rows), float32; image stored at padded index (r+1, c+1) ----
    __syncthreads();
    occ_prefilter_axis<true>(coef + OCC_PITCH + 1, 1, OCC_PITCH, xb);       // axis 0: line = column
    occ_prefilter_axis<false>(coef + OCC_PITCH + 1, OCC_PITCH, 1, nullptr);  // axis 1: line = row
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
    const int n_tiles = n_pix / 32;                                  // 50 x 25 tiles
    const int warps = OCC_THREADS / 32;
    const int n_rounds = (n_tiles + warps - 1) / warps;
    uint8_t* plane0 = reinterpret_cast<uint8_t*>(planes);
    uint8_t* plane1 = plane0 + n_pix / 8;
    for (int round = 0; round < n_rounds; ++round) {
      const int tile = round * warps + (tid >> 5);
      const int ty = tile / (RD_OCC_MID / 8), tx = tile - ty * (RD_OCC_MID / 8);
      if (ty < RD_OCC_MID / 4) {
        const int a = ty * 4 + (lane >> 3), b = tx * 8 + (lane & 7);
        double c0, c1;
        occ_coords(rc, a, b, c0, c1);
        uint32_t val = 0u;
        bool ambiguous = false;
        const bool inside = !(c0 < 0.0 || c0 > (double)(RD_OCC_IN - 1) || c1 < 0.0 || c1 > (double)(RD_OCC_IN - 1));
        const double f0 = floor(c0), f1 = floor(c1);
        const int i0 = inside ? (int)f0 : 0, i1 = inside ? (int)f1 : 0;
        const uint32_t u = inside ? (uint32_t)uni[(i0 >> 3) * OCC_NB + (i1 >> 3)] : 0u;
        if (u != 2u) {
          val = u;                                                   // constant neighbourhood (or outside the crop: 0)
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
          const uint32_t exact = occ_exact_pixel(rc, __shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src), xb, hband);
          if (lane == src) val = exact > 3u ? 3u : exact;
        }
        // tile row i (8 pixels) is one byte of the row-major bit planes
        const uint32_t b0 = __ballot_sync(0xffffffffu, val & 1u), b1 = __ballot_sync(0xffffffffu, val & 2u);
        if ((lane & 7) == 0) {
          const int byte = (a * RD_OCC_MID + tx * 8) >> 3;
          plane0[byte] = (uint8_t)(b0 >> (lane & 24));
          plane1[byte] = (uint8_t)(b1 >> (lane & 24));
        }
        if (lane == 0 && b1) any_hi = 1;
      }
    }
    __syncthreads();

    // ---- D: Pillow bicubic 200 -> 64.  The coefficient image is dead: its space takes the tables and the uint8
    // intermediate.  Horizontal pass straight on the bit planes: the <= 15-tap window of an output pixel is a bit mask,
    // its fixed-point sum three table reads (plane 1, the "value >= 2" plane, is empty for every real map; if any
    // pixel set it, its windows are added with weight 2).
    for (int i = tid; i < (int)(sizeof(OccTables) / 4); i += OCC_THREADS)
      reinterpret_cast<int32_t*>(tb)[i] = __ldg(reinterpret_cast<const int32_t*>(tables) + i);
    __syncthreads();
    {
      const bool hi_plane = any_hi != 0;
      for (int i = tid; i < RD_OCC_MID * RD_OCC_OUT; i += OCC_THREADS) {
        const int yy = i / RD_OCC_OUT, xx = i - yy * RD_OCC_OUT;
        const int x0 = tb->xmin[xx], xn = tb->xnum[xx];
        const int bitpos = yy * RD_OCC_MID + x0;
        const uint32_t mask = (1u << xn) - 1u;
        const int32_t* lut = tb->lut + xx;
        const uint32_t m0 = __funnelshift_r(planes[bitpos >> 5], planes[(bitpos >> 5) + 1], bitpos & 31) & mask;
        int32_t ss = (1 << (OCC_PREC_BITS - 1)) + lut[(m0 & 31u) * RD_OCC_OUT] + lut[(32 + ((m0 >> 5) & 31u)) * RD_OCC_OUT] +
                     lut[(64 + (m0 >> 10)) * RD_OCC_OUT];
        if (hi_plane) {
          const uint32_t* p1 = planes + (n_pix >> 5);
          const uint32_t m1 = __funnelshift_r(p1[bitpos >> 5], p1[min((bitpos >> 5) + 1, (n_pix >> 5) - 1)], bitpos & 31) & mask;
          ss += 2 * (lut[(m1 & 31u) * RD_OCC_OUT] + lut[(32 + ((m1 >> 5) & 31u)) * RD_OCC_OUT] + lut[(64 + (m1 >> 10)) * RD_OCC_OUT]);
        }
        ss >>= OCC_PREC_BITS;
        tmp[yy * RD_OCC_OUT + xx] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
      }
    }
    __syncthreads();
    for (int i = tid; i < RD_OCC_OUT * RD_OCC_OUT; i += OCC_THREADS) {
      const int yy = i / RD_OCC_OUT, xx = i - yy * RD_OCC_OUT;
      int32_t ss = 1 << (OCC_PREC_BITS - 1);
      const int y0 = tb->xmin[yy], yn = tb->xnum[yy];
      const int32_t* k = tb->kk + yy * OCC_KSIZE;
      for (int t = 0; t < yn; ++t) ss += (int32_t)tmp[(y0 + t) * RD_OCC_OUT + xx] * k[t];
      ss >>= OCC_PREC_BITS;
      dst[i] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
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

static inline void occ_build_tables(OccTables& t) {
  const int in_size = RD_OCC_MID, out_size = RD_OCC_OUT;
  const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  double k[OCC_KSIZE];
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
      t.kk[xx * OCC_KSIZE + x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << OCC_PREC_BITS)) : (int32_t)(0.5 + k[x] * (1 << OCC_PREC_BITS));
    t.xmin[xx] = xmin;
    t.xnum[xx] = xmax;
    for (int g = 0; g < 3; ++g)
      for (int m = 0; m < 32; ++m) {
        int32_t acc = 0;
        for (int j = 0; j < 5; ++j)
          if ((m >> j) & 1) acc += t.kk[xx * OCC_KSIZE + 5 * g + j];
        t.lut[(g * 32 + m) * RD_OCC_OUT + xx] = acc;
      }
  }
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
    occ_build_tables(t);
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
