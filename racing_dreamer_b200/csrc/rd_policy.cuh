// rd_policy.cuh -- on-device policies: the follow-the-gap controller, one warp per env.
//
// Replaces AgentNode.laserscan_callback + publish_drive_from_heading + PID.calculate
// [REF ros_agent/agents/follow_the_gap/src/agent.py:128-193, 200-238, 45-55] for a batch of envs whose scans are
// already in HBM (the rows k_lidar just wrote), so a closed-loop rollout never leaves the device.
//
// Per env (one warp, lanes over beams; the forward arc of the scan and its "adjusted" copy live in shared memory):
//   clip to the lookahead distance -> |first difference| -> candidates = differences that are the maximum of their
//   10-degree window (reflected border), exceed 9 x the window median (clamped border) and 0.2 m -> every candidate
//   lowers the adjusted ranges inside the arc a vehicle-width chord subtends at the nearer side -> the 83.3rd
//   percentile of the adjusted ranges (np.percentile's linear rule) -> heading = mean angle of the beams at or above
//   it, heading distance = their mean range -> PID -> steering angle, speed -> env action.
// The median test needs no sort: d > 9*median  <=>  at least w/2+1 window elements x satisfy 9*x < d.
// The percentile needs two order statistics of ~721 non-negative doubles: bisection on their bit patterns (monotone
// for non-negative IEEE doubles) with warp-wide counts.
// All arithmetic is float64 in the reference's operation order (translation unit compiled with -fmad=false); the only
// differences to NumPy are the summation order of the two means and acos() vs libm (<= 2 ulp).
#pragma once
#include "rd_common.cuh"

struct PolicyState {
  double* f64;    // [4][n]: PID previous input (NaN = none), PID integral, steering_angle, vehicle_speed
  int32_t* i32;   // [2][n]: scans seen, headings seen (the two "first message" gates of the node)
};
enum { RD_P_PREV = 0, RD_P_INT, RD_P_STEER, RD_P_SPEED, RD_NP_F64 };
enum { RD_P_SCANS = 0, RD_P_HEADINGS, RD_NP_I32 };

__device__ __forceinline__ void rd_policy_clear(const PolicyState& ps, int n, int e) {
  ps.f64[(size_t)RD_P_PREV * n + e] = __longlong_as_double(0x7ff8000000000000ll);
  ps.f64[(size_t)RD_P_INT * n + e] = 0.0;
  ps.f64[(size_t)RD_P_STEER * n + e] = 0.0;
  ps.f64[(size_t)RD_P_SPEED * n + e] = 0.0;
  ps.i32[(size_t)RD_P_SCANS * n + e] = 0;
  ps.i32[(size_t)RD_P_HEADINGS * n + e] = 0;
}

struct GapArgs {
  rd_gap_follower g;
  PolicyState ps;
  const float* lidar;     // [n][n_beams], metres, index 0 = left
  const float* speed;     // [n] or null
  const double* state_v;  // env state row RD_S_V (used when speed is null)
  float* actions;         // [n][2]
  double* debug;          // [n][4] or null
  int n, n_beams, m_pad;  // m_pad: doubles per shared-memory row
  int rescale;
  double low[2], high[2];
  double a_drive, c_drag, steer_scale;   // steer_scale = steer_gain * steer_max
};

__device__ __forceinline__ double rd_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// numpy float64 -> int64 cast followed by np.clip(., 0, hi): truncation toward zero; NaN and out-of-range values
// become INT64_MIN on x86 and therefore clip to 0.
__device__ __forceinline__ int rd_trunc_clip(double x, int hi) {
  if (!(x > -9.0e18 && x < 9.0e18)) return 0;
  long long k = (long long)x;
  return k < 0 ? 0 : (k > hi ? hi : (int)k);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_gap_follower(GapArgs A) {
  extern __shared__ double sm_pol[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WARPS + warp;
  if (e >= A.n) return;   // warps are independent: no block-wide barrier below
  const rd_gap_follower& g = A.g;
  const int n = A.n;
  double* rng = sm_pol + (size_t)warp * 2 * A.m_pad;
  double* adj = rng + A.m_pad;
  const int M = g.arc_last - g.arc_first + 1;
  int scans = A.ps.i32[(size_t)RD_P_SCANS * n + e] + 1;
  int headings = A.ps.i32[(size_t)RD_P_HEADINGS * n + e];
  double steering_angle = A.ps.f64[(size_t)RD_P_STEER * n + e];
  double vehicle_speed = A.ps.f64[(size_t)RD_P_SPEED * n + e];
  double heading = 0.0, heading_dist = 0.0;
  if (scans >= 2) {   // the first scan after a reset only arms the node's timestamp [REF agent.py:132-134]
    const float* row = A.lidar + (size_t)e * A.n_beams;
    for (int j = lane; j < M; j += 32) {
      double r = (double)row[(A.n_beams - 1) - (g.arc_first + j)];
      r = r < 0.0 ? 0.0 : r;
      r = r > g.lookahead ? g.lookahead : r;   // np.clip(ranges, 0, lookahead_distance)
      rng[j] = r;
      adj[j] = r;
    }
    __syncwarp();
    const int nd = M - 1, w = g.filter_width, half = w >> 1;
    const double inc = g.angle_increment, amin = g.angle_min;
    const double ang0 = (double)g.arc_first * inc + amin;
    const double w2 = g.vehicle_width * g.vehicle_width;
    for (int base = 0; base < nd; base += 32) {
      const int i = base + lane;
      bool cand = false;
      if (i < nd) {
        const double d = fabs(rng[i + 1] - rng[i]);
        if (d > g.minimum_gap_length) {
          bool ismax = true;
          int below = 0;
          for (int k = 0; k < w; ++k) {
            const int j = i - half + k;
            const int jr = j < 0 ? -j - 1 : (j >= nd ? 2 * nd - j - 1 : j);   // maximum_filter1d: mode 'reflect'
            const int jc = j < 0 ? 0 : (j >= nd ? nd - 1 : j);                // median_filter: mode 'nearest'
            const double dr = fabs(rng[jr + 1] - rng[jr]);
            const double dc = (jc == jr) ? dr : fabs(rng[jc + 1] - rng[jc]);
            if (dr > d) ismax = false;
            if (dc * g.median_dev_threshold < d) ++below;
          }
          cand = ismax && below >= half + 1;
        }
      }
      unsigned bal = __ballot_sync(0xffffffffu, cand);
      while (bal) {   // every lane handles the same candidate: its chord lowers a run of adjusted ranges
        const int ii = base + (__ffs(bal) - 1);
        bal &= bal - 1;
        double L = rng[ii];
        if (ii > 0) L = fmin(L, rng[ii - 1]);   // the reference's slice is empty (and raises) for ii == 0
        L = fmin(L, rng[ii + 1]);
        const double theta = (double)(g.arc_first + ii) * inc + amin;
        const double L2 = L * L;
        const double beta = acos((2.0 * L2 - w2) / (2.0 * L2));
        const int k0 = rd_trunc_clip(((theta - beta) - ang0) / inc, M - 1);
        const int k1 = rd_trunc_clip(((theta + beta) - ang0) / inc, M - 1);
        for (int k = k0 + lane; k <= k1; k += 32) adj[k] = fmin(adj[k], L);
      }
    }
    __syncwarp();
    // two order statistics of adj[0..M) by bisection on the bit patterns
    long long lo = 0x7fffffffffffffffll, hi = 0;
    for (int j = lane; j < M; j += 32) {
      const long long b = __double_as_longlong(adj[j]);
      lo = b < lo ? b : lo;
      hi = b > hi ? b : hi;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const long long l2 = __shfl_xor_sync(0xffffffffu, lo, off), h2 = __shfl_xor_sync(0xffffffffu, hi, off);
      lo = l2 < lo ? l2 : lo;
      hi = h2 > hi ? h2 : hi;
    }
    while (lo < hi) {
      const long long mid = lo + ((hi - lo) >> 1);
      unsigned c = 0;
      for (int j = lane; j < M; j += 32) c += (__double_as_longlong(adj[j]) <= mid) ? 1u : 0u;
      c = __reduce_add_sync(0xffffffffu, c);
      if ((int)c >= g.pct_lo + 1) hi = mid; else lo = mid + 1;
    }
    const double x = __longlong_as_double(lo);
    double y = x;
    if (g.pct_hi > g.pct_lo) {
      unsigned c = 0;
      long long nxt = 0x7fffffffffffffffll;
      for (int j = lane; j < M; j += 32) {
        const long long b = __double_as_longlong(adj[j]);
        c += (b <= lo) ? 1u : 0u;
        if (b > lo && b < nxt) nxt = b;
      }
      c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const long long o2 = __shfl_xor_sync(0xffffffffu, nxt, off);
        nxt = o2 < nxt ? o2 : nxt;
      }
      if ((int)c < g.pct_hi + 1) y = __longlong_as_double(nxt);
    }
    const double dxy = y - x;   // numpy _lerp
    const double pct = (g.pct_gamma >= 0.5) ? (y - dxy * (1.0 - g.pct_gamma)) : (x + dxy * g.pct_gamma);
    double sa = 0.0, sr = 0.0, cnt = 0.0;
    for (int j = lane; j < M; j += 32) {
      const double a = adj[j];
      if (a >= pct && a < g.range_max) {   // np.digitize(adjusted, [0, pct, range_max]) == 2
        sa += (double)(g.arc_first + j) * inc + amin;
        sr += rng[j];
        cnt += 1.0;
      }
    }
    sa = rd_warp_sum(sa); sr = rd_warp_sum(sr); cnt = rd_warp_sum(cnt);
    heading = sa / cnt;
    heading_dist = sr / cnt;
    headings += 1;
    if (headings >= 2) {   // the first heading only arms the PID clock [REF agent.py:206-208]
      const double prev = A.ps.f64[(size_t)RD_P_PREV * n + e];
      double integ = A.ps.f64[(size_t)RD_P_INT * n + e];
      const double err = 0.0 - heading;   // PID.calculate [REF agent.py:45-55]
      const double P = g.kp * err;
      integ = integ + g.ki * err * g.scan_dt;
      const double D = (prev != prev) ? 0.0 : (g.kd * (prev - heading) / g.scan_dt);
      const double control = (g.kp + g.ki + g.kd > 0.0) ? ((P + integ) + D) : heading;
      double ang = -control;
      const double lim = fabs(g.max_steering_angle);
      ang = ang < -lim ? -lim : (ang > lim ? lim : ang);
      const double aabs = fabs(ang);
      double spd = g.max_vehicle_speed;
      if (aabs > g.speed_limit_angle) spd = g.max_vehicle_speed - (aabs / g.max_steering_angle) * (g.max_vehicle_speed * 0.30);
      if (heading_dist < 5.0) spd = fmin(spd, heading_dist / 5.0 * 4.0);
      spd = fmax(spd, 1.5);
      steering_angle = ang;
      vehicle_speed = spd;
      if (lane == 0) {
        A.ps.f64[(size_t)RD_P_PREV * n + e] = heading;
        A.ps.f64[(size_t)RD_P_INT * n + e] = integ;
        A.ps.f64[(size_t)RD_P_STEER * n + e] = ang;
        A.ps.f64[(size_t)RD_P_SPEED * n + e] = spd;
      }
    }
  }
  if (lane == 0) {
    A.ps.i32[(size_t)RD_P_SCANS * n + e] = scans;
    A.ps.i32[(size_t)RD_P_HEADINGS * n + e] = headings;
    // drive command -> env action
    const double v = A.speed ? (double)A.speed[e] : A.state_v[e];
    const double vt = vehicle_speed * g.speed_scale;
    double motor = vt * A.c_drag / A.a_drive + g.speed_gain * (vt - v);
    double steer = steering_angle / A.steer_scale;
    motor = motor < A.low[0] ? A.low[0] : (motor > A.high[0] ? A.high[0] : motor);
    steer = steer < A.low[1] ? A.low[1] : (steer > A.high[1] ? A.high[1] : steer);
    if (A.rescale) {   // inverse of ReduceActionSpace._normalize [REF dreamer/wrappers.py:129-134]
      motor = (motor - A.low[0]) / (A.high[0] - A.low[0]) * 2.0 - 1.0;
      steer = (steer - A.low[1]) / (A.high[1] - A.low[1]) * 2.0 - 1.0;
    }
    A.actions[2 * (size_t)e] = (float)motor;
    A.actions[2 * (size_t)e + 1] = (float)steer;
    if (A.debug) {
      double* d = A.debug + 4 * (size_t)e;
      d[0] = steering_angle; d[1] = vehicle_speed; d[2] = heading; d[3] = heading_dist;
    }
  }
}
