// rd_lidar.cuh -- K1: LiDAR ray cast through the bit-packed drivable grid.
//
// Replaces the 'lidar' sensor of the reference scenario [REF dreamer/scenarios/max_progress/austria.yml:7]
// (pybullet.rayTestBatch inside racecar_gym).  Beam i points at yaw + fov/2 - i*fov/(n-1): beam 0 = +135 deg
// (left), last = -135 deg [REF dreamer/tools.py:84-86]; 1080 beams [REF dreamer/dream.py:66]; 15 m cap
// [REF dreamer/tools.py:274].
//
// Work decomposition: one warp per (env, group of 32 ADJACENT beams) so the 32 march lengths are
// correlated; persistent CTAs stride over the items.  The track's bit grid (14-120 KB) is brought into
// shared memory once per CTA with one 1-D bulk copy (TMA, cp.async.bulk + mbarrier); rows are an odd number
// of 32-bit words so vertically adjacent cells sit in different banks.  The march itself is integer-only
// (a 32-bit error term decides x- vs y-crossing), so it is bit-reproducible against the CPU oracle.
#pragma once
#include "rd_common.cuh"

__device__ __forceinline__ float rd_finish_range(const LidarParams& lp, float r, const OriginRec& rec, uint32_t beam) {
  if (lp.noise > 0.0f) {
    uint32_t c[4] = {rec.gid, rec.episode, rec.step, beam};
    philox4x32_10(c, lp.key0, lp.key1);
    float u = __fmul_rn((float)(c[0] >> 8), 1.0f / 16777216.0f);
    float t = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
    float f = __fadd_rn(1.0f, __fmul_rn(lp.noise, t));
    r = __fmul_rn(r, f);
  }
  r = r < lp.range_min ? lp.range_min : r;
  r = r > lp.range_max ? lp.range_max : r;
  if (lp.normalize) r = __fsub_rn(__fdiv_rn(r, lp.range_max), 0.5f);  // [REF dreamer/tools.py:274]
  return r;
}

// smem layout: [0,16) mbarrier | beam table 2*n_beams f64 (cos then sin) | bit grid
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_lidar(const DevMap* __restrict__ maps, int map_id, const OriginRec* __restrict__ recs,
        const int32_t* __restrict__ env_order, int n_env, LidarParams lp, const double* __restrict__ beam_tab,
        float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  double* tab = reinterpret_cast<double*>(smem + 16);
  const int tab_bytes = ((2 * lp.n_beams * 8) + 15) & ~15;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + 16 + tab_bytes);

  const DevMap& m = maps[map_id];
  const int rw = m.rw;
  if (threadIdx.x == 0) rd_mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = (uint32_t)m.bits_bytes;
    rd_mbar_expect_tx(bar, total);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(m.bits);
    unsigned char* dst = reinterpret_cast<unsigned char*>(bits);
    for (uint32_t off = 0; off < total; off += 32768u) {
      uint32_t nbytes = total - off < 32768u ? total - off : 32768u;
      rd_bulk_g2s(dst + off, src + off, nbytes, bar);
    }
  }
  for (int i = threadIdx.x; i < 2 * lp.n_beams; i += WARPS * 32) tab[i] = __ldg(beam_tab + i);
  __syncthreads();
  rd_mbar_wait(bar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total_items = (long long)n_env * lp.groups;
  for (long long item = (long long)blockIdx.x * WARPS + warp; item < total_items; item += (long long)gridDim.x * WARPS) {
    const int slot = (int)(item / lp.groups);
    const int g = (int)(item - (long long)slot * lp.groups);
    const int env = env_order ? __ldg(env_order + slot) : slot;
    const OriginRec rec = recs[env];
    if (rec.was_reset == 2) continue;  // frozen env: outputs stay as they are
    const int beam = g * 32 + lane;
    if (beam >= lp.n_beams) continue;
    float r;
    if (!rec.valid) {
      r = 0.0f;
    } else {
      const double ca = tab[beam], sa = tab[lp.n_beams + beam];
      const double dx = __dsub_rn(__dmul_rn(rec.c, ca), __dmul_rn(rec.s, sa));
      const double dy = __dadd_rn(__dmul_rn(rec.s, ca), __dmul_rn(rec.c, sa));
      const int DX = __double2int_rn(dx * (double)(1 << RD_DIR_BITS));
      const int DY = __double2int_rn(dy * (double)(1 << RD_DIR_BITS));
      const int adx = abs(DX), ady = abs(DY);
      const int ix0 = rec.px >> RD_SUB_BITS, iy0 = rec.py >> RD_SUB_BITS;
      const int fx = rec.px & (RD_SUB - 1), fy = rec.py & (RD_SUB - 1);
      const int bx = DX > 0 ? RD_SUB - fx : fx;
      const int by = DY > 0 ? RD_SUB - fy : fy;
      int e = (int)((long long)bx * ady - (long long)by * adx);
      if (ady == 0) e = -1;
      const long long lx = (lp.rsub * adx) >> RD_DIR_BITS, ly = (lp.rsub * ady) >> RD_DIR_BITS;
      const int nx = (adx != 0 && lx >= bx) ? (int)((lx - bx) >> RD_SUB_BITS) + 1 : 0;
      const int ny = (ady != 0 && ly >= by) ? (int)((ly - by) >> RD_SUB_BITS) + 1 : 0;
      const int n0 = nx + ny;
      int n = n0;
      const int stepx = DX > 0 ? 1 : -1;
      const int steprow = DY > 0 ? rw : -rw;
      const int ex = ady << RD_SUB_BITS, ey = adx << RD_SUB_BITS;
      int ix = ix0, row = iy0 * rw;
      bool hit = false, lastx = false;
      while (n > 0) {
        lastx = e < 0;
        e += lastx ? ex : -ey;
        ix += lastx ? stepx : 0;
        row += lastx ? 0 : steprow;
        const uint32_t word = bits[row + (ix >> 5)];
        --n;
        if (!((word >> (ix & 31)) & 1u)) { hit = true; break; }
      }
      if (hit) {
        int num, den;
        const int xs = abs(ix - ix0);            // x-crossings taken; the rest of the n0-n steps were y-crossings
        if (lastx) { num = bx + (xs - 1) * RD_SUB; den = adx; }
        else       { num = by + ((n0 - n) - xs - 1) * RD_SUB; den = ady; }
        r = __fmul_rn(__fdiv_rn((float)num, (float)den), lp.scale);
      } else {
        r = lp.range_max;
      }
    }
    out[(size_t)env * lp.n_beams + beam] = rd_finish_range(lp, r, rec, (uint32_t)beam);
  }
}

// Sensor-origin record from a pose (stage entry rd_lidar_cast / rd_occupancy_obs, and the step kernels).
__device__ __forceinline__ void rd_make_origin(const DevMap& m, double x, double y, double yaw, double lidar_offset,
                                               OriginRec& rec) {
  double s, c;
  sincos(yaw, &s, &c);
  const double sx = x + lidar_offset * c;
  const double sy = y + lidar_offset * s;
  const double u = (sx - m.ox) * m.inv_res;
  const double v = (sy - m.oy) * m.inv_res;
  const double pu = floor(u * (double)RD_SUB), pv = floor(v * (double)RD_SUB);
  bool ok = (pu > -1.0e12 && pu < 1.0e12 && pv > -1.0e12 && pv < 1.0e12);
  long long PX = 0, PY = 0;
  if (ok) {
    PX = (long long)pu - (long long)m.col0 * RD_SUB;
    PY = (long long)pv - (long long)m.row0 * RD_SUB;
    ok = PX >= 0 && PY >= 0 && (PX >> RD_SUB_BITS) < m.w && (PY >> RD_SUB_BITS) < m.h;
  }
  if (ok) ok = rd_drivable_at(m, (int)(PX >> RD_SUB_BITS), (int)(PY >> RD_SUB_BITS)) != 0;
  rec.px = ok ? (int32_t)PX : 0;
  rec.py = ok ? (int32_t)PY : 0;
  rec.valid = ok ? 1 : 0;
  rec.c = c;
  rec.s = s;
}

__global__ void k_origin_from_poses(const DevMap* __restrict__ maps, const int32_t* __restrict__ map_ids,
                                    const double* __restrict__ poses, int n, double lidar_offset,
                                    OriginRec* __restrict__ recs) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int mid = map_ids ? map_ids[e] : 0;
  OriginRec rec;
  rd_make_origin(maps[mid], poses[3 * e], poses[3 * e + 1], poses[3 * e + 2], lidar_offset, rec);
  rec.gid = (uint32_t)e; rec.episode = 0u; rec.step = 0u; rec.was_reset = 0; rec.pad = mid;
  recs[e] = rec;
}
