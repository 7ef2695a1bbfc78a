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

#ifndef RD_LIDAR_CHUNK
#define RD_LIDAR_CHUNK 2   // work items a warp draws from the global counter at a time (see k_lidar)
#endif
// GPI = beam groups (32 adjacent beams each) per work item: the item's decode, its env look-up and its origin
// record are shared by gpi consecutive groups of one env.  Measured on B200 (profiles/r3o_lidar_groups_per_item.txt):
// two groups per item are 4-8 % faster for long launches (fewer draws from the work counter, half the set-up) and 5 %
// slower at 4096 envs, where a warp only gets ~20 groups and the coarser items leave a longer tail; the host picks.

// atomicAdd whose result is NOT needed right away.  For an atomic on a provably warp-uniform address ptxas emits its
// warp-aggregation pattern (leader ATOMG + SHFL of the result right behind it), which waits for the L2 round trip on the
// spot and defeats drawing the next work chunk ahead of time; the (always zero) threadIdx.y offset hides the uniformity.
__device__ __forceinline__ unsigned rd_atom_add(unsigned int* p, unsigned v) {
  unsigned old;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p + threadIdx.y), "r"(v) : "memory");  // blockDim.y == 1
  return old;
}

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
  if (lp.normalize == 1) r = __fsub_rn(__fdiv_rn(r, lp.range_max), 0.5f);  // [REF dreamer/tools.py:274]
  // NormalizeObservations: (x - low) * (1 / (high - low)), float64, then float32 [REF baselines single_agent.py:92-99]
  else if (lp.normalize == 2) r = (float)__dmul_rn(__dsub_rn((double)r, lp.norm_lo), lp.norm_sc);
  return r;
}

// Range (metres) at which the beam (DX, DY) * 2^-18 from (px, py) enters the body box of another car of the world, or
// +inf.  float32 slab test in the other car's sensor frame; every operand is an exactly representable integer (or a
// per-launch constant) and every operation a single IEEE operation in a fixed order, so the CPU oracle's float32
// restatement produces the same bits.  [NEW-SPEC: racecar_gym's rayTestBatch hits the other racecars' collision
// shapes; SURVEY.md §8-f3 "rays hitting cars"]
__device__ __forceinline__ float rd_car_hit(const LidarParams& lp, int px, int py, int DX, int DY, const OriginRec& q) {
  const float inf = __int_as_float(0x7f800000);
  const int rx = px - q.px, ry = py - q.py;
  const int arx = rx < 0 ? -rx : rx, ary = ry < 0 ? -ry : ry;
  if (!(q.valid & 2) || arx > lp.car_reach || ary > lp.car_reach) return inf;
  const float ox = __fmul_rn((float)rx, 1.0f / (float)RD_SUB), oy = __fmul_rn((float)ry, 1.0f / (float)RD_SUB);
  const float c = (float)q.c, s = (float)q.s;
  const float dxf = __fmul_rn((float)DX, 1.0f / (float)(1 << RD_DIR_BITS));
  const float dyf = __fmul_rn((float)DY, 1.0f / (float)(1 << RD_DIR_BITS));
  const float u0 = __fadd_rn(__fmul_rn(ox, c), __fmul_rn(oy, s));
  const float v0 = __fsub_rn(__fmul_rn(oy, c), __fmul_rn(ox, s));
  const float du = __fadd_rn(__fmul_rn(dxf, c), __fmul_rn(dyf, s));
  const float dv = __fsub_rn(__fmul_rn(dyf, c), __fmul_rn(dxf, s));
  float tmin = 0.0f, tmax = inf;
  if (du != 0.0f) {
    const float t1 = __fdiv_rn(__fsub_rn(lp.car_ulo, u0), du), t2 = __fdiv_rn(__fsub_rn(lp.car_uhi, u0), du);
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  } else if (u0 < lp.car_ulo || u0 > lp.car_uhi) {
    return inf;
  }
  if (dv != 0.0f) {
    const float t1 = __fdiv_rn(__fsub_rn(-lp.car_hw, v0), dv), t2 = __fdiv_rn(__fsub_rn(lp.car_hw, v0), dv);
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  } else if (v0 < -lp.car_hw || v0 > lp.car_hw) {
    return inf;
  }
  return tmin <= tmax ? __fmul_rn(tmin, lp.res) : inf;
}

// smem layout: [0,16) mbarrier | beam table 2*n_beams f64 (cos then sin) | bit grid | block clearance field
// The tail of the item list is handed out through a global counter (ctr[0]); the last CTA to finish (ticket ctr[1])
// re-arms both for the next launch on the stream.
template <int WARPS, int GPI, bool CARS>
__global__ void __launch_bounds__(WARPS * 32, CARS ? (WARPS == 16 ? 2 : 1) : (WARPS == 32 ? 1 : (WARPS == 24 ? 2 : 3)))
k_lidar(const DevMap* __restrict__ maps, int map_id, const OriginRec* __restrict__ recs,
        const int32_t* __restrict__ env_order, int n_env, LidarParams lp, const double* __restrict__ beam_tab,
        float* __restrict__ out, unsigned int* __restrict__ ctr) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  double* tab = reinterpret_cast<double*>(smem + 16);
  const int tab_bytes = ((2 * lp.n_beams * 8) + 15) & ~15;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + 16 + tab_bytes);

  const DevMap& m = maps[map_id];
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
  // first draw from the work counter: in flight while the map arrives
  const int lane = threadIdx.x & 31;
  unsigned ahead = 0;
  bool have_ahead = true;
  if (lane == 0) ahead = rd_atom_add(ctr, (unsigned)RD_LIDAR_CHUNK);
  MarchGrid grid;
  grid.bits = bits;
  grid.coarse = reinterpret_cast<const uint8_t*>(bits) + m.coarse_off;
  grid.rw = m.rw;
  grid.cw = m.cw;
  grid.cshift = m.cshift;
  __syncthreads();
  rd_mbar_wait(bar, 0);
  // everything above only read launch-invariant data; the origin records come from the kernel in front of this one
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const unsigned total_items = (unsigned)n_env * (unsigned)lp.units;
  // Scheduling: items are handed out through a global counter, RD_LIDAR_CHUNK at a time, so warps that drew long rays
  // do not hold the kernel up.  Measured on B200 (Austria, 4096 envs): chunk 2 beats 4, 8, a guided (shrinking) chunk
  // and a static round-robin with a dynamic tail (profiles/r01_lidar_variants.txt).
  unsigned item = 0, last = 0;
  for (;;) {
    if (item >= last) {
      const unsigned chunk = RD_LIDAR_CHUNK;
      if (!have_ahead && lane == 0) ahead = rd_atom_add(ctr, chunk);
      item = __shfl_sync(0xffffffffu, ahead, 0);
      have_ahead = false;
      if (item >= total_items) break;
      last = min(item + chunk, total_items);
    }
    {
      unsigned slot;
      int g;
      if (lp.centre_first) {
        // Longest-first order: items are (beam group, env) with the groups taken from the centre of the scan outwards
        // -- the beams that look along the track march furthest, the side beams end at the nearby wall -- so the items
        // drawn last are the cheap ones and the warps finish closer together (the end-of-launch barrier held 11 % of the
        // warps' time in the env-major order, profiles/r01x_ncu_k_lidar_lines.txt).
        unsigned r = __umulhi(item, lp.envs_magic);          // item / n_env, envs_magic = ceil(2^32 / n_env) (0: n_env == 1)
        if (!lp.envs_magic) r = item;
        if (item - r * (unsigned)n_env >= (unsigned)n_env) --r;
        slot = item - r * (unsigned)n_env;
        const int c = lp.units >> 1;
        g = (r & 1u) ? c - (int)((r + 1u) >> 1) : c + (int)(r >> 1);
      } else {
        // item / groups by a multiply: groups_magic = ceil(2^32 / groups) (exact for item < 2^32 / groups; one fix-up
        // step covers the rest of the 31-bit range)
        slot = lp.groups_magic ? __umulhi(item, lp.groups_magic) : item;   // magic 0: one unit per env
        if (item - slot * (unsigned)lp.units >= (unsigned)lp.units) --slot;        // the estimate never falls short
        g = (int)(item - slot * (unsigned)lp.units);
      }
      const int env = env_order ? __ldg(env_order + slot) : (int)slot;
      const OriginRec rec = recs[env];
#pragma unroll
      for (int h = 0; h < GPI; ++h) {
      const int beam = (g * GPI + h) * 32 + lane;
      if (rec.was_reset != 2 && beam < lp.n_beams) {  // was_reset == 2: frozen env, outputs stay as they are
        float r;
        if (!(rec.valid & 1)) {
          r = 0.0f;
        } else {
          const double ca = tab[beam], sa = tab[lp.n_beams + beam];
          const double dx = __dsub_rn(__dmul_rn(rec.c, ca), __dmul_rn(rec.s, sa));
          const double dy = __dadd_rn(__dmul_rn(rec.s, ca), __dmul_rn(rec.c, sa));
          const int DX = __double2int_rn(dx * (double)(1 << RD_DIR_BITS));
          const int DY = __double2int_rn(dy * (double)(1 << RD_DIR_BITS));
          const MarchResult mr = rd_march(grid, rec.px, rec.py, DX, DY, (long long)lp.rsub, nullptr);
          r = mr.hit ? __fmul_rn(__fdiv_rn((float)mr.num, (float)mr.den), lp.scale) : lp.range_max;
          if (CARS) {  // worlds: the other cars of this world (their records sit next to this one)
            // Warp-level rejection first: the group's beams fill the wedge between its first and last direction (W = the
            // points clockwise of d_a and counter-clockwise of d_b, < 8 degrees wide); a car whose bounding disc lies
            // entirely on the far side of either edge line cannot be hit by any beam of the group.  64-bit integers,
            // conservative radius -- most groups skip every car, so worlds cost little more than single cars.
            const unsigned act = __activemask();
            const int la = __ffs(act) - 1, lb = 31 - __clz(act);
            const int dax = __shfl_sync(act, DX, la), day = __shfl_sync(act, DY, la);
            const int dbx = __shfl_sync(act, DX, lb), dby = __shfl_sync(act, DY, lb);
            const long long rr = (long long)lp.car_radius_sub * ((1ll << RD_DIR_BITS) + 4);
            const int base = env - env % lp.agents;
            for (int j = 0; j < lp.agents; ++j) {
              if (base + j == env) continue;
              const OriginRec q = recs[base + j];
              const long long ox = (long long)q.px - rec.px, oy = (long long)q.py - rec.py;   // me -> the other car
              const long long ca = (long long)dax * oy - (long long)day * ox;                // > 0: left of d_a
              const long long cb = (long long)dbx * oy - (long long)dby * ox;                // < 0: right of d_b
              if ((q.valid & 2) && (ca > rr || cb < -rr)) continue;
              r = fminf(r, rd_car_hit(lp, rec.px, rec.py, DX, DY, q));
            }
          }
        }
        const float rf = rd_finish_range(lp, r, rec, (uint32_t)beam);
        if (lp.f16) reinterpret_cast<__half*>(out)[(size_t)env * lp.n_beams + beam] = __float2half_rn(rf);   // Collect at precision 16
        else out[(size_t)env * lp.n_beams + beam] = rf;
      }
      }
    }
    ++item;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(ctr + 1, 1u);
    if (ticket == gridDim.x - 1) { ctr[0] = 0u; ctr[1] = 0u; __threadfence(); }
  }
}

// Sensor-origin record from a pose and the cosine / sine of its heading (the step kernels already hold them).
__device__ __forceinline__ void rd_make_origin_cs(const DevMap& m, double x, double y, double c, double s,
                                                  double lidar_offset, OriginRec& rec) {
  const double sx = x + lidar_offset * c;
  const double sy = y + lidar_offset * s;
  const double u = (sx - m.ox) * m.inv_res;
  const double v = (sy - m.oy) * m.inv_res;
  const double pu = floor(u * (double)RD_SUB), pv = floor(v * (double)RD_SUB);
  bool ok = (pu > -1.0e12 && pu < 1.0e12 && pv > -1.0e12 && pv < 1.0e12);
  long long PX = 0, PY = 0;
  bool pos = false;   // px/py representable (other cars of the world need them even when the sensor left the track)
  if (ok) {
    PX = (long long)pu - (long long)m.col0 * RD_SUB;
    PY = (long long)pv - (long long)m.row0 * RD_SUB;
    pos = PX > -(1ll << 30) && PX < (1ll << 30) && PY > -(1ll << 30) && PY < (1ll << 30);
    ok = PX >= 0 && PY >= 0 && (PX >> RD_SUB_BITS) < m.w && (PY >> RD_SUB_BITS) < m.h;
  }
  if (ok) ok = rd_drivable_at(m, (int)(PX >> RD_SUB_BITS), (int)(PY >> RD_SUB_BITS)) != 0;
  rec.px = pos ? (int32_t)PX : 0;
  rec.py = pos ? (int32_t)PY : 0;
  rec.valid = (ok ? 1 : 0) | (pos ? 2 : 0);   // bit0: origin inside a drivable cell; bit1: px/py hold the position
  rec.c = c;
  rec.s = s;
}
// ... from a pose alone (stage entries rd_lidar_cast / rd_occupancy_obs)
__device__ __forceinline__ void rd_make_origin(const DevMap& m, double x, double y, double yaw, double lidar_offset,
                                               OriginRec& rec) {
  double s, c;
  sincos(yaw, &s, &c);
  rd_make_origin_cs(m, x, y, c, s, lidar_offset, rec);
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
