// rd_occupancy.cuh -- K3: the 'lidar_occupancy' observation, bit-exact with the reference.
//
// Replaces OccupancyMapObs.step [REF dreamer/wrappers.py:390-408]:
//   (pr,pc) = to_pixel(pose); crop M[pr-110:pr+110, pc-110:pc+110] as uint8;
//   scipy.ndimage.rotate(crop, rad2deg(2*pi - yaw))  -- cubic B-spline prefilter (mirror boundary) + affine
//     resampling with constant-0 outside, result rounded to uint8;
//   centre crop 200x200; PIL.Image.resize((64,64)) -- bicubic, two fixed-point passes with a uint8 image between.
//
// One CTA per env (persistent loop).  Phases, all inside one kernel:
//   B  column prefilter: thread = column, float64 recursion straight from the bit grid -> coefficient scratch
//   C  row prefilter:    thread = row, in place on the scratch (L1/L2-resident, 387 KB per CTA)
//   D  rotation:         32x32 output tiles; the tile's source bounding box is staged in shared memory,
//                        each thread evaluates 4 pixels (4x4 B-spline taps, float64), thresholds to {0,1,..}
//   E  Pillow resize:    integer 22-bit fixed point, horizontal then vertical, uint8 intermediate, in smem
// float64 is required: interpolated values come within ~3e-5 of the 0.5 rounding threshold, float32 would
// flip pixels (SURVEY.md §7 hard part 1).
#pragma once
#include "rd_common.cuh"

#define OCC_THREADS 256
#define OCC_TILE 32
#define OCC_BOX 56          // max side of a tile's source bounding box (32*sqrt(2)+3+margins)
#define OCC_BOX_PITCH 57    // odd pitch (in doubles) to spread banks
#define OCC_KSIZE 15        // Pillow: ceil(2*3.125)*2+1 taps per output pixel
#define OCC_PREC_BITS 22

struct OccTables {            // Pillow precompute_coeffs + normalize_coeffs_8bpc for 200 -> 64, bicubic
  int32_t kk[RD_OCC_OUT * OCC_KSIZE];
  int32_t xmin[RD_OCC_OUT];
  int32_t xnum[RD_OCC_OUT];
};

struct OccScratch {
  double* coef = nullptr;     // [ctas][220*220]
  OccTables* tables = nullptr;
  int ctas = 0;
};

struct OccGeom {              // per-env geometry, computed by one thread
  double c, s, off0, off1;
  int o0_first, o1_first;     // output index of centre-crop pixel (0,0)
  int pr, pc;
};

__device__ __forceinline__ int occ_mirror(int idx, int len) {
  const int s2 = 2 * len - 2;
  if (idx < 0) {
    idx = s2 * (int)(-idx / s2) + idx;
    idx = idx <= 1 - len ? idx + s2 : -idx;
  } else if (idx >= len) {
    idx -= s2 * (int)(idx / s2);
    if (idx >= len) idx = s2 - idx;
  }
  return idx;
}

__device__ __forceinline__ void occ_weights(double x, double (&w)[4]) {
  const double y = x - floor(x), z = 1.0 - y;
  w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
  w[0] = z * z * z / 6.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

// in-place cubic B-spline prefilter of one line held at p[0], p[stride], ... (mirror boundary, pole sqrt(3)-2)
__device__ __forceinline__ void occ_prefilter_line(double* p, int stride) {
  const int n = RD_OCC_IN;
  const double z = -0.26794919243112270647;  // sqrt(3) - 2
  const double gain = 6.0;                   // (1-z)(1-1/z)
  // causal init: sum_{i<=n-2} z^i c_i (+ mirror terms of relative size z^(n-1) ~ 1e-125, below rounding)
  double zi = 1.0, acc = 0.0;
#pragma unroll 4
  for (int i = 0; i < 48; ++i) { acc += zi * (gain * p[(size_t)i * stride]); zi *= z; }
  double prev = acc;
  p[0] = prev;
  for (int i = 1; i < n; ++i) {
    prev = gain * p[(size_t)i * stride] + z * prev;
    p[(size_t)i * stride] = prev;
  }
  double cur = (z * p[(size_t)(n - 2) * stride] + p[(size_t)(n - 1) * stride]) * z / (z * z - 1.0);
  p[(size_t)(n - 1) * stride] = cur;
  for (int i = n - 2; i >= 0; --i) {
    cur = z * (cur - p[(size_t)i * stride]);
    p[(size_t)i * stride] = cur;
  }
}

__global__ void __launch_bounds__(OCC_THREADS)
k_occupancy(const DevMap* __restrict__ maps, int map_id, const OriginRec* __restrict__ recs,
            const double* __restrict__ poses, const double* __restrict__ f64, int n_state,
            const int32_t* __restrict__ order, int n_env, double* __restrict__ scratch_all,
            const OccTables* __restrict__ tables, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* box = reinterpret_cast<double*>(smem);                                  // OCC_BOX * OCC_BOX_PITCH doubles
  uint8_t* mid = smem + sizeof(double) * OCC_BOX * OCC_BOX_PITCH;                  // 200*200
  uint8_t* tmp = mid + RD_OCC_MID * RD_OCC_MID;                                    // 200 rows * 64
  OccTables* tb = reinterpret_cast<OccTables*>(tmp + RD_OCC_MID * RD_OCC_OUT);     // 4-byte aligned: sizes above are multiples of 8
  __shared__ OccGeom geom;
  __shared__ int box_r0, box_c0, box_h, box_w;

  const DevMap& m = maps[map_id];
  double* coef = scratch_all + (size_t)blockIdx.x * RD_OCC_IN * RD_OCC_IN;
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(OccTables) / 4); i += OCC_THREADS)
    reinterpret_cast<int32_t*>(tb)[i] = reinterpret_cast<const int32_t*>(tables)[i];

  for (int slot = blockIdx.x; slot < n_env; slot += gridDim.x) {
    const int env = order ? __ldg(order + slot) : slot;
    double x, y, yaw;
    int mode = 0;
    if (poses) { x = poses[3 * slot]; y = poses[3 * slot + 1]; yaw = poses[3 * slot + 2]; }
    else {
      x = f64[(size_t)RD_S_X * n_state + env]; y = f64[(size_t)RD_S_Y * n_state + env];
      yaw = f64[(size_t)RD_S_YAW * n_state + env];
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
    }
    __syncthreads();

    // ---- B: crop + column prefilter (thread = column) ----
    if (tid < RD_OCC_IN) {
      const int cx = geom.pc - RD_OCC_IN / 2 + tid - m.col0;
      const bool col_ok = cx >= 0 && cx < m.w;
      for (int i = 0; i < RD_OCC_IN; ++i) {
        const int r_img = geom.pr - RD_OCC_IN / 2 + i;
        const int cy = (m.full_h - 1 - r_img) - m.row0;
        uint32_t bit = 0;
        if (col_ok && cy >= 0 && cy < m.h) bit = (__ldg(m.bits + (size_t)cy * m.rw + (cx >> 5)) >> (cx & 31)) & 1u;
        coef[(size_t)i * RD_OCC_IN + tid] = (double)bit;
      }
      occ_prefilter_line(coef + tid, RD_OCC_IN);
    }
    __syncthreads();
    // ---- C: row prefilter (thread = row) ----
    if (tid < RD_OCC_IN) occ_prefilter_line(coef + (size_t)tid * RD_OCC_IN, 1);
    __syncthreads();

    // ---- D: rotation, tile by tile ----
    const double c = geom.c, s = geom.s, off0 = geom.off0, off1 = geom.off1;
    const int ntile = (RD_OCC_MID + OCC_TILE - 1) / OCC_TILE;
    for (int tile = 0; tile < ntile * ntile; ++tile) {
      const int ta = (tile / ntile) * OCC_TILE, tb0 = (tile % ntile) * OCC_TILE;
      const int th = min(OCC_TILE, RD_OCC_MID - ta), tw = min(OCC_TILE, RD_OCC_MID - tb0);
      if (tid == 0) {
        double lo0 = 1e30, hi0 = -1e30, lo1 = 1e30, hi1 = -1e30;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double o0 = (double)(geom.o0_first + ta + ((k & 1) ? th - 1 : 0));
          const double o1 = (double)(geom.o1_first + tb0 + ((k & 2) ? tw - 1 : 0));
          const double c0 = off0 + o0 * c + o1 * s, c1 = off1 + o0 * -s + o1 * c;
          lo0 = fmin(lo0, c0); hi0 = fmax(hi0, c0); lo1 = fmin(lo1, c1); hi1 = fmax(hi1, c1);
        }
        int r0 = max((int)floor(lo0) - 3, 0), r1 = min((int)floor(hi0) + 4, RD_OCC_IN - 1);
        int q0 = max((int)floor(lo1) - 3, 0), q1 = min((int)floor(hi1) + 4, RD_OCC_IN - 1);
        box_r0 = r0; box_c0 = q0;
        box_h = max(min(r1 - r0 + 1, OCC_BOX), 0);
        box_w = max(min(q1 - q0 + 1, OCC_BOX), 0);
      }
      __syncthreads();
      const int bh = box_h, bw = box_w, br0 = box_r0, bc0 = box_c0;
      for (int i = tid; i < bh * bw; i += OCC_THREADS) {
        const int r = i / bw, q = i - r * bw;
        box[r * OCC_BOX_PITCH + q] = coef[(size_t)(br0 + r) * RD_OCC_IN + bc0 + q];
      }
      __syncthreads();
      for (int i = tid; i < th * tw; i += OCC_THREADS) {
        const int a = i / tw, b = i - a * tw;
        const double o0 = (double)(geom.o0_first + ta + a), o1 = (double)(geom.o1_first + tb0 + b);
        // scipy accumulates shift first, then one product per output axis (no contraction)
        double c0 = __dadd_rn(__dadd_rn(off0, __dmul_rn(o0, c)), __dmul_rn(o1, s));
        double c1 = __dadd_rn(__dadd_rn(off1, __dmul_rn(o0, -s)), __dmul_rn(o1, c));
        double t = 0.0;
        if (!(c0 < 0.0 || c0 > (double)(RD_OCC_IN - 1) || c1 < 0.0 || c1 > (double)(RD_OCC_IN - 1))) {
          const int s0 = (int)floor(c0) - 1, s1 = (int)floor(c1) - 1;
          double w0[4], w1[4];
          occ_weights(c0, w0);
          occ_weights(c1, w1);
          int jj[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) jj[q] = occ_mirror(s1 + q, RD_OCC_IN) - bc0;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int ii = occ_mirror(s0 + p, RD_OCC_IN) - br0;
            const double* row = box + ii * OCC_BOX_PITCH;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double cf = row[jj[q]];
              cf *= w0[p];
              cf *= w1[q];
              t += cf;
            }
          }
        }
        double tv = t > 0.0 ? t + 0.5 : 0.0;
        tv = tv > 255.0 ? 255.0 : tv;
        mid[(ta + a) * RD_OCC_MID + tb0 + b] = (uint8_t)tv;
      }
      __syncthreads();
    }

    // ---- E: Pillow bicubic 200 -> 64: horizontal pass to uint8, then vertical ----
    for (int i = tid; i < RD_OCC_MID * RD_OCC_OUT; i += OCC_THREADS) {
      const int yy = i / RD_OCC_OUT, xx = i - yy * RD_OCC_OUT;
      int32_t ss = 1 << (OCC_PREC_BITS - 1);
      const int x0 = tb->xmin[xx], xn = tb->xnum[xx];
      const uint8_t* src = mid + yy * RD_OCC_MID + x0;
      const int32_t* k = tb->kk + xx * OCC_KSIZE;
      for (int t = 0; t < xn; ++t) ss += (int32_t)src[t] * k[t];
      ss >>= OCC_PREC_BITS;
      tmp[yy * RD_OCC_OUT + xx] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
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
  }
}

static inline void occ_free(OccScratch& sc) {
  cudaFree(sc.coef);
  cudaFree(sc.tables);
  sc = OccScratch{};
}

// returns a cudaError_t value (0 = ok)
static inline int occ_launch(OccScratch& sc, const DevMap* d_maps, int map_id, const DevMap& hm, const OriginRec* recs,
                             const double* poses, const double* f64, int n_state, const int32_t* order, int n_env,
                             uint8_t* out, int sm_count, cudaStream_t s, int64_t* launches) {
  (void)hm;
  const size_t smem = sizeof(double) * OCC_BOX * OCC_BOX_PITCH + RD_OCC_MID * RD_OCC_MID + RD_OCC_MID * RD_OCC_OUT + sizeof(OccTables);
  cudaError_t e = cudaFuncSetAttribute(k_occupancy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_occupancy, OCC_THREADS, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  const int max_ctas = sm_count * per_sm;
  if (!sc.tables) {
    OccTables t;
    occ_build_tables(t);
    e = cudaMalloc(&sc.tables, sizeof(OccTables));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(sc.tables, &t, sizeof(OccTables), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
  }
  if (sc.ctas < max_ctas) {
    cudaFree(sc.coef);
    sc.coef = nullptr;
    sc.ctas = 0;
    e = cudaMalloc(&sc.coef, sizeof(double) * RD_OCC_IN * RD_OCC_IN * (size_t)max_ctas);
    if (e != cudaSuccess) return (int)e;
    sc.ctas = max_ctas;
  }
  const int grid = n_env < max_ctas ? n_env : max_ctas;
  if (grid < 1) return 0;
  k_occupancy<<<grid, OCC_THREADS, smem, s>>>(d_maps, map_id, recs, poses, f64, n_state, order, n_env, sc.coef, sc.tables, out);
  (*launches)++;
  return (int)cudaGetLastError();
}
