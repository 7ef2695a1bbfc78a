// rd_dynamics.cuh -- K2: action transform + single-track dynamics + progress/lap/collision + reward/done
// + TimeLimit + auto-reset, one thread per env.
//
// Replaces, fused into one pass: ReduceActionSpace._normalize [REF dreamer/wrappers.py:129-134] (and the
// baselines clip [REF baselines/racing/environment/single_agent.py:55-56]), ActionRepeat.step
// [REF dreamer/wrappers.py:107-116], MultiAgentRaceEnv.step -> pybullet.stepSimulation + task reward/done
// (racecar_gym, not in tree; call site [REF dreamer/wrappers.py:63-64]), RaceCarWrapper's speed obs
// [REF dreamer/wrappers.py:66] and TimeLimit.step [REF dreamer/wrappers.py:147-154].
//
// State is SoA float64 (x, y, steer, v, yaw, yaw_rate, slip, ...) so that a warp's loads of one field are
// one coalesced 256-B request; all arithmetic is float64 in a fixed operation order (the translation unit is
// compiled with -fmad=false) so the result matches the CPU oracle to the rounding of sin/cos/tan.
#pragma once
#include "rd_common.cuh"
#include "rd_lidar.cuh"
#include "rd_policy.cuh"

struct StepParams {
  rd_config cfg;
  double* f64;      // [RD_NF64][n]
  int32_t* i32;     // [RD_NI32][n]
  double* stats;    // [8] accumulators (rd_stats order)
  OriginRec* recs;  // [n]
  const DevMap* maps;
  PolicyState pol;  // on-device controller state, cleared with the env (null pointers: no policy attached)
  double* hist;     // [n_step_progress][n] ring of lap + progress per sim tick (n_step_progress task), or null
  int n;
};

struct OutPtrs {
  float* pose; float* velocity; float* speed; float* reward; uint8_t* done; float* progress; int32_t* lap;
  float* time; uint8_t* flags; uint8_t* occupancy; int32_t* rank; uint8_t* opponents;
};

__device__ __forceinline__ void st_rhs(const rd_vehicle& p, const double (&q)[7], double sv, double acc, double (&f)[7]) {
  const double g = 9.81;
  const double steer = q[2], v = q[3], yaw = q[4], yr = q[5], slip = q[6];
  double svc;
  if ((steer <= p.steer_min && sv <= 0.0) || (steer >= p.steer_max && sv >= 0.0)) svc = 0.0;
  else if (sv <= -p.steer_vel_max) svc = -p.steer_vel_max;
  else if (sv >= p.steer_vel_max) svc = p.steer_vel_max;
  else svc = sv;
  // the division only exists for vehicles that can exceed v_switch (uniform test: no work for the default car)
  double pos_limit = p.a_max;
  if (p.v_max + 1.0 > p.v_switch) { if (v > p.v_switch) pos_limit = p.a_max * p.v_switch / v; }
  double ac;
  if ((v <= p.v_min && acc <= 0.0) || (v >= p.v_max && acc >= 0.0)) ac = 0.0;
  else if (acc <= -p.a_max) ac = -p.a_max;
  else if (acc >= pos_limit) ac = pos_limit;
  else ac = acc;
  const double lwb = p.lf + p.lr;
  const double rl = 1.0 / lwb;
  // one sincos serves both regimes (heading of the velocity vector); warps whose lanes sit in different regimes
  // share it instead of paying for both trig sets
  const bool kin = fabs(v) < p.v_kinematic;
  const double ang = kin ? yaw : (slip + yaw);
  double sn, cn;
  rd_sincos(ang, &sn, &cn);
  f[0] = v * cn;
  f[1] = v * sn;
  f[2] = svc;
  f[3] = ac;
  if (kin) {
    double ss, cs;   // |steer| <= steer_max (+ an RK4 stage's overshoot) < pi/4: no range reduction needed
    if (fabs(steer) < 0.78) rd_sincos_kernel(steer, ss, cs); else rd_sincos(steer, &ss, &cs);
    const double rc = rd_rcp(cs);
    const double tn = ss * rc;
    f[4] = (v * rl) * tn;
    f[5] = (ac * rl) * tn + ((v * rl) * (rc * rc)) * svc;
    f[6] = 0.0;
  } else {
    const double rv = rd_rcp(v);   // |v| >= v_kinematic here
    const double c1 = p.mu * p.mass / (p.inertia * lwb);
    const double c2 = p.mu * rl;
    double rear = g * p.lf + ac * p.h_cg;
    double front = g * p.lr - ac * p.h_cg;
    double k_yr = (-c1 * rv) * (p.lf * p.lf * p.c_sf * front + p.lr * p.lr * p.c_sr * rear);
    double k_sl = c1 * (p.lr * p.c_sr * rear - p.lf * p.c_sf * front);
    double k_st = c1 * (p.lf * p.c_sf * front);
    double b_yr = (c2 * (rv * rv)) * (p.c_sr * rear * p.lr - p.c_sf * front * p.lf) - 1.0;
    double b_sl = (c2 * rv) * (p.c_sr * rear + p.c_sf * front);
    double b_st = (c2 * rv) * (p.c_sf * front);
    f[4] = yr;
    f[5] = (k_yr * yr + k_sl * slip) + k_st * steer;
    f[6] = (b_yr * yr - b_sl * slip) + b_st * steer;
  }
}

#ifndef RD_STEP_MAP_LOCAL
#define RD_STEP_MAP_LOCAL 1    // 1: the track descriptor is copied into registers once per step instead of being re-read
#endif                         //    from global memory by every probe of the tick loop (long-scoreboard stalls)
#if RD_STEP_MAP_LOCAL
#define RD_STEP_MAP_T const DevMap
#else
#define RD_STEP_MAP_T const DevMap&
#endif
#ifndef RD_STEP_STAGE_LOOP
#define RD_STEP_STAGE_LOOP 1   // 1: the four RK4 stages share ONE copy of the RHS code (rolled loop); 0: four inlined copies
#endif
__device__ __forceinline__ void st_tick(const rd_config& cfg, double (&q)[7], double motor, double steering, double inv_dt) {
  const rd_vehicle& p = cfg.vehicle;
  const double dt = cfg.dt;
  double target = steering * p.steer_gain * p.steer_max;
  double sv = (target - q[2]) * inv_dt;
  double acc = (motor >= 0.0) ? (motor * p.a_drive - p.c_drag * q[3]) : (motor * p.a_brake - p.c_drag * q[3]);
  const double h2 = 0.5 * dt, h6 = dt / 6.0;
#if RD_STEP_STAGE_LOOP
  // One RHS body executed four times instead of four inlined copies: the tick loop's code shrinks ~3.5x and stays in
  // the instruction cache (k_step runs one warp per SM sub-partition; ncu showed 9 % no-instruction stalls).  Same
  // operations in the same order as the unrolled form: stage input q + c*k with c = (0, h/2, h/2, h) and k = 0 before
  // the first stage (q + 0*0 == q), sum ((1*k1 + 2*k2) + 2*k3) + 1*k4 (0 + 1*k1 == k1, products by 1 and 2 are exact).
  double k[7] = {0, 0, 0, 0, 0, 0, 0}, sum[7] = {0, 0, 0, 0, 0, 0, 0}, t[7];
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    const double c = st == 0 ? 0.0 : (st == 3 ? dt : h2);
    const double w = (st == 0 || st == 3) ? 1.0 : 2.0;
#pragma unroll
    for (int i = 0; i < 7; ++i) t[i] = q[i] + c * k[i];
    st_rhs(p, t, sv, acc, k);
#pragma unroll
    for (int i = 0; i < 7; ++i) sum[i] = sum[i] + w * k[i];
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) q[i] = q[i] + h6 * sum[i];
#else
  double k1[7], k2[7], k3[7], k4[7], t[7];
  st_rhs(p, q, sv, acc, k1);
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = q[i] + h2 * k1[i];
  st_rhs(p, t, sv, acc, k2);
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = q[i] + h2 * k2[i];
  st_rhs(p, t, sv, acc, k3);
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = q[i] + dt * k3[i];
  st_rhs(p, t, sv, acc, k4);
#pragma unroll
  for (int i = 0; i < 7; ++i) q[i] = q[i] + h6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
#endif
}

__device__ __forceinline__ int rd_checkpoint_of(const rd_config& cfg, double p) {
  int c = (int)(p * (double)cfg.n_checkpoints);
  return c > cfg.n_checkpoints - 1 ? cfg.n_checkpoints - 1 : c;
}
__device__ __forceinline__ bool rd_progress_at(const DevMap& m, double x, double y, double& p) {
  int cx, cy;
  if (!rd_cell_of(m, x, y, cx, cy)) return false;
  if (!rd_drivable_at(m, cx, cy)) return false;
  p = (double)__ldg(m.dist + (size_t)cy * m.w + cx) / (double)m.dmax;
  return true;
}
__device__ __forceinline__ bool rd_collides(const rd_config& cfg, const DevMap& m, double x, double y, double yaw) {
  int cx, cy;
  if (!rd_cell_of(m, x, y, cx, cy) || !rd_drivable_at(m, cx, cy)) return true;
  double c, s;
  sincos(yaw, &s, &c);
  double hl = 0.5 * cfg.vehicle.body_length, hw = 0.5 * cfg.vehicle.body_width;
  double ax = hl * c, ay = hl * s, bx = hw * s, by = hw * c;
  double px[4] = {(x + ax) - bx, (x + ax) + bx, (x - ax) - bx, (x - ax) + bx};
  double py[4] = {(y + ay) + by, (y + ay) - by, (y - ay) + by, (y - ay) - by};
  bool col = false;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (!rd_cell_of(m, px[k], py[k], cx, cy) || !rd_drivable_at(m, cx, cy)) col = true;
  return col;
}

// Footprint + progress probe of one pose: the centre and the four body corners are looked up with five independent
// loads (plus the centre's wavefront distance) issued together, instead of five dependent round trips.
// Same results as rd_collides() + rd_cell_of() + rd_progress_at().
__device__ __forceinline__ void rd_probe(const rd_config& cfg, const DevMap& m, double x, double y, double yaw,
                                         bool& col, bool& inside, double& p, double* c_out = nullptr,
                                         double* s_out = nullptr) {
  double c, s;
  rd_sincos(yaw, &s, &c);
  if (c_out) { *c_out = c; *s_out = s; }
  const double hl = 0.5 * cfg.vehicle.body_length, hw = 0.5 * cfg.vehicle.body_width;
  const double ax = hl * c, ay = hl * s, bx = hw * s, by = hw * c;
  const double px[5] = {x, (x + ax) - bx, (x + ax) + bx, (x - ax) - bx, (x - ax) + bx};
  const double py[5] = {y, (y + ay) + by, (y + ay) - by, (y - ay) + by, (y - ay) - by};
  int cx[5], cy[5];
  bool in[5];
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    in[k] = rd_cell_of(m, px[k], py[k], cx[k], cy[k]);
    w[k] = in[k] ? __ldg(m.bits + (size_t)cy[k] * m.rw + (cx[k] >> 5)) : 0u;
  }
  const uint32_t dval = in[0] ? (uint32_t)__ldg(m.dist + (size_t)cy[0] * m.w + cx[0]) : 0u;
  bool free_all = true;
#pragma unroll
  for (int k = 0; k < 5; ++k) free_all = free_all && in[k] && ((w[k] >> (cx[k] & 31)) & 1u);
  col = !free_all;
  inside = in[0];
  if (in[0] && ((w[0] >> (cx[0] & 31)) & 1u)) p = (double)dval / (double)m.dmax;
}

// reset of one env: pose from the map's tables (grid slot 0 or a Philox-sampled candidate)
// [REF dreamer/wrappers.py:91-92 reset(mode=...); sampler itself is racecar_gym -> NEW-SPEC]
__device__ __forceinline__ void rd_reset_one(const StepParams& P, int e, int mode) {
  const rd_config& cfg = P.cfg;
  const int n = P.n;
  const DevMap& m = P.maps[P.i32[(size_t)RD_I_MAP * n + e]];
  const uint32_t episode = (uint32_t)P.i32[(size_t)RD_I_EPISODE * n + e];
  // multi-agent worlds: the cars of a world draw ONE anchor (counter = global id of the world's agent 0) and line up
  // along the ball_next chain (cfg.ball_spacing metres of track apart); 'grid' hands out the staggered start slots
  const int A = cfg.agents_per_world > 1 ? cfg.agents_per_world : 1;
  const int a = e % A;
  const uint64_t gid = (uint64_t)(cfg.env_id_offset + (e - a));
  double x, y, yaw;
  if (mode == RD_RESET_GRID || m.n_reset <= 0) {
    const int slot = a < m.n_start ? a : m.n_start - 1;
    x = m.start[3 * slot]; y = m.start[3 * slot + 1]; yaw = m.start[3 * slot + 2];
  } else {
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), episode, 0u};
    philox4x32_10(c, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32) ^ RD_STREAM_RESET);
    uint32_t idx = __umulhi(c[0], (uint32_t)m.n_reset);
    for (int k = 0; k < a && m.ball_next; ++k) idx = (uint32_t)__ldg(m.ball_next + idx);
    x = m.reset[3 * idx]; y = m.reset[3 * idx + 1]; yaw = m.reset[3 * idx + 2];
    if (mode == RD_RESET_RANDOM_BIDIRECTIONAL && (c[1] & 1u)) yaw = yaw + 3.14159265358979323846;
  }
  double* f = P.f64;
  f[(size_t)RD_S_X * n + e] = x; f[(size_t)RD_S_Y * n + e] = y; f[(size_t)RD_S_STEER * n + e] = 0.0;
  f[(size_t)RD_S_V * n + e] = 0.0; f[(size_t)RD_S_YAW * n + e] = yaw; f[(size_t)RD_S_YAWRATE * n + e] = 0.0;
  f[(size_t)RD_S_SLIP * n + e] = 0.0; f[(size_t)RD_S_TIME * n + e] = 0.0;
  double p = 0.0;
  rd_progress_at(m, x, y, p);
  f[(size_t)RD_S_PROGRESS * n + e] = p;
  f[(size_t)RD_S_LAST * n + e] = 1.0 + p;
  f[(size_t)RD_S_START * n + e] = 1.0 + p;
  f[(size_t)RD_S_RETURN * n + e] = 0.0;
  int32_t* I = P.i32;
  I[(size_t)RD_I_LAP * n + e] = 1;
  I[(size_t)RD_I_CHECKPOINT * n + e] = rd_checkpoint_of(cfg, p);
  I[(size_t)RD_I_FLAGS * n + e] = 0;
  I[(size_t)RD_I_AGENT_STEP * n + e] = 0;
  I[(size_t)RD_I_EPISODE * n + e] = (int32_t)(episode + 1u);
  if (P.hist) for (int k = 0; k < cfg.n_step_progress; ++k) P.hist[(size_t)k * n + e] = 1.0 + p;
  if (P.pol.i32 || P.pol.dr_feat) rd_policy_clear(P.pol, n, e);
}

// observation scalars + the origin record for the LiDAR / occupancy kernels, from the committed state
__device__ __forceinline__ void rd_write_obs(const StepParams& P, const OutPtrs& o, int e, int was_reset) {
  const int n = P.n;
  const double* f = P.f64;
  const int32_t* I = P.i32;
  const int mid = I[(size_t)RD_I_MAP * n + e];
  const double x = f[(size_t)RD_S_X * n + e], y = f[(size_t)RD_S_Y * n + e], yaw = f[(size_t)RD_S_YAW * n + e];
  const double v = f[(size_t)RD_S_V * n + e], slip = f[(size_t)RD_S_SLIP * n + e];
  OriginRec rec;
  rd_make_origin(P.maps[mid], x, y, yaw, P.cfg.lidar_offset, rec);
  rec.gid = (uint32_t)(P.cfg.env_id_offset + e);
  rec.episode = (uint32_t)I[(size_t)RD_I_EPISODE * n + e];
  rec.step = (uint32_t)I[(size_t)RD_I_AGENT_STEP * n + e];
  rec.was_reset = was_reset;
  rec.pad = mid;
  P.recs[e] = rec;
  const double two_pi = 6.283185307179586;
  const double wy = yaw - rint(yaw / two_pi) * two_pi;
  const double vx = v * cos(slip), vy = v * sin(slip);
  if (o.pose) {
    float* p = o.pose + (size_t)e * 6;
    p[0] = (float)x; p[1] = (float)y; p[2] = 0.f; p[3] = 0.f; p[4] = 0.f; p[5] = (float)wy;
  }
  if (o.velocity) {
    float* q = o.velocity + (size_t)e * 6;
    q[0] = (float)vx; q[1] = (float)vy; q[2] = 0.f; q[3] = 0.f; q[4] = 0.f;
    q[5] = (float)f[(size_t)RD_S_YAWRATE * n + e];
  }
  if (o.speed) o.speed[e] = (float)sqrt(vx * vx + vy * vy);  // [REF dreamer/wrappers.py:66]
}

__global__ void __launch_bounds__(128) k_reset(StepParams P, OutPtrs o, const uint8_t* __restrict__ mask, int mode) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // k_lidar may start its set-up (see launch_lidar_t)
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  const int n = P.n;
  bool sel = (!mask || mask[e]);
  if (mask && P.cfg.agents_per_world > 1) {  // a world resets as a whole [REF dreamer/tools.py:178-179]
    const int A = P.cfg.agents_per_world, b = e - e % A;
    for (int j = 0; j < A; ++j) sel = sel || mask[b + j];
  }
  if (sel) rd_reset_one(P, e, mode);
  rd_write_obs(P, o, e, sel ? 1 : 3);  // 3: not reset here -> occupancy output left untouched
  if (sel) {
    if (o.reward) o.reward[e] = 0.f;
    if (o.done) o.done[e] = 0;
    if (o.progress) o.progress[e] = (float)P.f64[(size_t)RD_S_PROGRESS * n + e];
    if (o.lap) o.lap[e] = P.i32[(size_t)RD_I_LAP * n + e];
    if (o.time) o.time[e] = 0.f;
    if (o.flags) o.flags[e] = 0;
    if (o.rank) o.rank[e] = 1 + (P.cfg.agents_per_world > 1 ? e % P.cfg.agents_per_world : 0);  // refined by the first step
    if (o.opponents) o.opponents[e] = 0;
  }
}

// envs [e0, e1) of the batch (the host-facing path steps the batch in chunks on several streams)
__global__ void __launch_bounds__(128) k_step(StepParams P, OutPtrs o, const float* __restrict__ actions, int e0, int e1) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // k_lidar may start its set-up (see launch_lidar_t)
  const int e = e0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int n = P.n;
  const rd_config& cfg = P.cfg;
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // rd_stats contributions of this env
  if (e < e1) {
    double* f = P.f64;
    int32_t* I = P.i32;
    // Every load of this env's state is issued up front, before the first branch: after a cold L2 (another kernel ran in
    // between) each dependent round trip to HBM costs ~1 us of a ~40 us kernel -- the loads that used to sit behind the
    // flags test and behind the tick loop (return, episode start, step counter) showed up as its top stall lines.
    int flags = I[(size_t)RD_I_FLAGS * n + e];
    const int map_id = I[(size_t)RD_I_MAP * n + e];
    const float act0 = actions[2 * e], act1 = actions[2 * e + 1];
    double q[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) q[k] = f[(size_t)k * n + e];
    double time = f[(size_t)RD_S_TIME * n + e], p = f[(size_t)RD_S_PROGRESS * n + e];
    double last = f[(size_t)RD_S_LAST * n + e];
    const double ret0 = f[(size_t)RD_S_RETURN * n + e], start0 = f[(size_t)RD_S_START * n + e];
    int lap = I[(size_t)RD_I_LAP * n + e], cp = I[(size_t)RD_I_CHECKPOINT * n + e];
    const int agent_step0 = I[(size_t)RD_I_AGENT_STEP * n + e];
    if (flags & RD_F_NEEDS_RESET) {  // frozen until reset [REF dreamer/wrappers.py:148]
      if (o.reward) o.reward[e] = 0.f;
      if (o.done) o.done[e] = 1;
      if (o.progress) o.progress[e] = (float)p;
      if (o.lap) o.lap[e] = lap;
      if (o.time) o.time[e] = (float)time;
      if (o.flags) o.flags[e] = (uint8_t)flags;
      P.recs[e].was_reset = 2;
    } else {
      RD_STEP_MAP_T m = P.maps[map_id];
      // a4 [REF dreamer/wrappers.py:129-134; baselines single_agent.py:55-56]: numpy keeps (action+1)/2 of a
      // float32 policy output in float32 and promotes to float64 at `* (high-low)` (float64 arrays).
      double a[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float af = k == 0 ? act0 : act1;
        if (cfg.clip_actions) af = af < -1.0f ? -1.0f : (af > 1.0f ? 1.0f : af);
        if (cfg.rescale_actions) {
          const float t = __fdiv_rn(__fadd_rn(af, 1.0f), 2.0f);
          a[k] = (double)t * (cfg.action_high[k] - cfg.action_low[k]) + cfg.action_low[k];
        } else {
          a[k] = (double)af;
        }
      }
      double total = 0.0;
      int done = 0;
      const int ncp = cfg.n_checkpoints;
      const double inv_dt = 1.0 / cfg.dt;
      for (int t = 0; t < cfg.action_repeat; ++t) {  // ActionRepeat [REF dreamer/wrappers.py:107-116]
        st_tick(cfg, q, a[0], a[1], inv_dt);
        time = time + cfg.dt;
        bool col, inside;
        rd_probe(cfg, m, q[0], q[1], q[4], col, inside, p);
        flags &= ~(RD_F_COLLISION | RD_F_LEFT_MAP);
        if (col) flags |= RD_F_COLLISION;
        if (!inside) flags |= RD_F_LEFT_MAP;
        if (!(q[0] == q[0] && q[1] == q[1] && q[3] == q[3] && q[4] == q[4])) flags |= RD_F_NAN;
        const int cn = rd_checkpoint_of(cfg, p);
        if (cn == cp + 1) { cp = cn; flags &= ~RD_F_WRONG_WAY; }
        else if (cp == ncp - 1 && cn == 0 && ncp > 1) { lap += 1; cp = 0; flags &= ~RD_F_WRONG_WAY; }
        else if (cn == cp - 1 || (cp == 0 && cn == ncp - 1 && ncp > 1)) { flags |= RD_F_WRONG_WAY; }
        const double cur = (double)lap + p;
        double r;
        bool d;
        if (cfg.task == RD_TASK_MAX_SPEED) {  // [REF baselines/racing/environment/tasks.py:6-18]
          r = col ? -1.0 : -exp(fabs(a[1]) - q[3] * cos(q[6]));
          d = false;
        } else {  // maximize_progress [REF dreamer/scenarios/max_progress/austria.yml:8-10]
          double delta = cur - last;
          if (delta > 0.5) delta = delta - 1.0;
          if (delta < -0.5) delta = delta + 1.0;
          if (cfg.progress_abs) delta = fabs(delta);
          r = cfg.frame_reward + cfg.progress_reward * delta;
          if (col) r = r + cfg.collision_reward;
          d = (cfg.terminate_on_collision && col) || (lap > cfg.laps) || (time > cfg.time_limit);
        }
        last = cur;
        total = total + r;
        if (d && !(cfg.repeat_semantics == RD_REPEAT_BASELINES && t == 0 && cfg.action_repeat > 1)) { done = 1; break; }
      }
      const int agent_step = agent_step0 + 1;  // TimeLimit [REF dreamer/wrappers.py:147-154]
      int timeout = 0;
      if (cfg.time_limit_steps > 0 && agent_step >= cfg.time_limit_steps) { timeout = !done; done = 1; }
      const double ret = ret0 + total;
#pragma unroll
      for (int k = 0; k < 7; ++k) f[(size_t)k * n + e] = q[k];
      f[(size_t)RD_S_TIME * n + e] = time; f[(size_t)RD_S_PROGRESS * n + e] = p;
      f[(size_t)RD_S_LAST * n + e] = last; f[(size_t)RD_S_RETURN * n + e] = ret;
      I[(size_t)RD_I_LAP * n + e] = lap; I[(size_t)RD_I_CHECKPOINT * n + e] = cp;
      I[(size_t)RD_I_AGENT_STEP * n + e] = agent_step;
      if (done && !cfg.auto_reset) flags |= RD_F_NEEDS_RESET;
      I[(size_t)RD_I_FLAGS * n + e] = flags;
      if (o.reward) o.reward[e] = (float)total;
      if (o.done) o.done[e] = (uint8_t)done;
      if (o.progress) o.progress[e] = (float)p;
      if (o.lap) o.lap[e] = lap;
      if (o.time) o.time[e] = (float)time;
      if (o.flags) o.flags[e] = (uint8_t)flags;
      st[6] = 1.0;
      if (done) {
        st[0] = 1.0; st[1] = ret; st[2] = ((double)lap + p) - start0;
        st[3] = (double)agent_step; st[4] = (flags & RD_F_COLLISION) ? 1.0 : 0.0; st[5] = (double)(lap - 1);
        st[7] = timeout ? 1.0 : 0.0;
      }
      int was_reset = 0;
      if (done && cfg.auto_reset) { rd_reset_one(P, e, cfg.reset_mode); was_reset = 1; }
      rd_write_obs(P, o, e, was_reset);
    }
  }
  // K5 episode statistics: warp reduce, one atomic per warp and counter
  // [REF dreamer/tools.py:159-206 simulate(): per-episode return / progress lists]
  const unsigned any_done = __ballot_sync(0xffffffffu, st[0] != 0.0);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k != 6 && !any_done) continue;
    double v = st[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(P.stats + k, v);
  }
}

// ---- multi-agent worlds -----------------------------------------------------------------------
// Overlap of two body boxes (half extents hl x hw, centres d apart, headings (c1,s1), (c2,s2)): separating-axis test on
// the four box axes, float64, one IEEE operation per step in the order written (the oracle's rect_overlap).
// [NEW-SPEC: racecar_gym reports Bullet contacts between racecars as info['opponent_collisions']]
__device__ __forceinline__ bool rd_rect_overlap(double hl, double hw, double dx, double dy, double c1, double s1,
                                                double c2, double s2) {
  const double cr = fabs(c1 * c2 + s1 * s2), sr = fabs(c1 * s2 - s1 * c2);
  const double ex = hl + (hl * cr + hw * sr), ey = hw + (hl * sr + hw * cr);
  if (fabs(dx * c1 + dy * s1) > ex) return false;
  if (fabs(dy * c1 - dx * s1) > ey) return false;
  if (fabs(dx * c2 + dy * s2) > ex) return false;
  if (fabs(dy * c2 - dx * s2) > ey) return false;
  return true;
}

// k_step for worlds of A = cfg.agents_per_world cars (A = 1 allowed: single cars on the n_step_progress task).
// One thread per car; the cars of a world are adjacent threads of one CTA (a CTA holds blockDim.x / A whole worlds) and
// exchange poses / done flags through shared memory every tick:
//   tick:  every car integrates its own dynamics and probes the walls            -> pose to shared memory
//          body-box overlap with the other cars of the world (opponent mask), progress/lap machine, the car's own task
//          reward and done                                                        -> done flag to shared memory
//          dreamer ActionRepeat: the WORLD stops repeating at the first tick in which any car is done
//          [REF dreamer/wrappers.py:112]; baselines multi-agent ActionRepeat: all ticks run, dones are OR-ed
//          [REF baselines/racing/environment/multi_agent.py:72-79]
//   then:  TimeLimit (one counter per world, all dones True [REF dreamer/wrappers.py:151-153]), rank, commit, auto-reset
//          of the whole world when any car is done [REF dreamer/tools.py:178-179].
// Episode statistics follow tools.simulate: the world's first agent only [REF dreamer/tools.py:162-165 main_id].
__global__ void __launch_bounds__(128) k_step_ma(StepParams P, OutPtrs o, const float* __restrict__ actions) {
  __shared__ double sh_pose[128 * 4];
  __shared__ double sh_prog[128];
  __shared__ int sh_flag[128];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // k_lidar may start its set-up (see launch_lidar_t)
  const int n = P.n;
  const rd_config& cfg = P.cfg;
  const int A = cfg.agents_per_world > 1 ? cfg.agents_per_world : 1;
  const int wpc = (int)blockDim.x / A;             // worlds per CTA
  const int t = (int)threadIdx.x;
  const int e = (int)blockIdx.x * wpc * A + t;
  const bool live = t < wpc * A && e < n;
  const int a = t % A, base = t - a;               // agent index, first thread of my world
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double* f = P.f64;
  int32_t* I = P.i32;
  int flags = live ? I[(size_t)RD_I_FLAGS * n + e] : 0;
  const bool frozen = live && (flags & RD_F_NEEDS_RESET);
  const bool run = live && !frozen;
  const int task = A > 1 ? cfg.agent_task[a] : cfg.task;
  const int ncp = cfg.n_checkpoints;
  const double inv_dt = 1.0 / cfg.dt;
  const double hl = 0.5 * cfg.vehicle.body_length, hw = 0.5 * cfg.vehicle.body_width;
  RD_STEP_MAP_T m = P.maps[run ? I[(size_t)RD_I_MAP * n + e] : 0];
  double act[2] = {0.0, 0.0};
  double q[7] = {0, 0, 0, 0, 0, 0, 0};
  double time = 0.0, p = 0.0, last = 0.0, total = 0.0, ret0 = 0.0, start0 = 0.0;
  int lap = 1, cp = 0, agent_step0 = 0, opp = 0, done = 0;
  if (run) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // a4, as in k_step
      float af = actions[2 * e + k];
      if (cfg.clip_actions) af = af < -1.0f ? -1.0f : (af > 1.0f ? 1.0f : af);
      if (cfg.rescale_actions) {
        const float h = __fdiv_rn(__fadd_rn(af, 1.0f), 2.0f);
        act[k] = (double)h * (cfg.action_high[k] - cfg.action_low[k]) + cfg.action_low[k];
      } else {
        act[k] = (double)af;
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) q[k] = f[(size_t)k * n + e];
    time = f[(size_t)RD_S_TIME * n + e]; p = f[(size_t)RD_S_PROGRESS * n + e]; last = f[(size_t)RD_S_LAST * n + e];
    lap = I[(size_t)RD_I_LAP * n + e]; cp = I[(size_t)RD_I_CHECKPOINT * n + e];
    agent_step0 = I[(size_t)RD_I_AGENT_STEP * n + e];
    ret0 = f[(size_t)RD_S_RETURN * n + e]; start0 = f[(size_t)RD_S_START * n + e];
  }
  bool ticking = run;
  for (int tk = 0; tk < cfg.action_repeat; ++tk) {
    bool col = false, inside = true;
    if (ticking) {
      st_tick(cfg, q, act[0], act[1], inv_dt);
      time = time + cfg.dt;
      double c, s;
      rd_probe(cfg, m, q[0], q[1], q[4], col, inside, p, &c, &s);
      sh_pose[4 * t] = q[0]; sh_pose[4 * t + 1] = q[1]; sh_pose[4 * t + 2] = c; sh_pose[4 * t + 3] = s;
    }
    __syncthreads();
    int d = 0;
    if (ticking) {
      opp = 0;
      for (int j = 0; j < A; ++j) {
        if (j == a) continue;
        const double* w = sh_pose + 4 * (base + j);
        if (rd_rect_overlap(hl, hw, w[0] - q[0], w[1] - q[1], sh_pose[4 * t + 2], sh_pose[4 * t + 3], w[2], w[3])) opp |= 1 << j;
      }
      flags &= ~(RD_F_COLLISION | RD_F_LEFT_MAP | RD_F_OPPONENT);
      if (col) flags |= RD_F_COLLISION;
      if (opp) flags |= RD_F_OPPONENT;
      if (!inside) flags |= RD_F_LEFT_MAP;
      if (!(q[0] == q[0] && q[1] == q[1] && q[3] == q[3] && q[4] == q[4])) flags |= RD_F_NAN;
      const bool hitc = col || opp != 0;   // the task's collision: walls or other cars
      const int cn = rd_checkpoint_of(cfg, p);
      if (cn == cp + 1) { cp = cn; flags &= ~RD_F_WRONG_WAY; }
      else if (cp == ncp - 1 && cn == 0 && ncp > 1) { lap += 1; cp = 0; flags &= ~RD_F_WRONG_WAY; }
      else if (cn == cp - 1 || (cp == 0 && cn == ncp - 1 && ncp > 1)) { flags |= RD_F_WRONG_WAY; }
      const double cur = (double)lap + p;
      double r;
      if (task == RD_TASK_MAX_SPEED) {
        r = hitc ? -1.0 : -exp(fabs(act[1]) - q[3] * cos(q[6]));
        d = 0;
      } else {
        double ref = last;
        if (task == RD_TASK_N_STEP_PROGRESS && P.hist) {  // progress over the last n ticks
          const int slot = (agent_step0 * cfg.action_repeat + tk) % cfg.n_step_progress;
          ref = P.hist[(size_t)slot * n + e];
          P.hist[(size_t)slot * n + e] = cur;
        }
        double delta = cur - ref;
        if (delta > 0.5) delta = delta - 1.0;
        if (delta < -0.5) delta = delta + 1.0;
        if (cfg.progress_abs) delta = fabs(delta);
        r = cfg.frame_reward + cfg.progress_reward * delta;
        if (hitc) r = r + cfg.collision_reward;
        d = ((cfg.terminate_on_collision && hitc) || (lap > cfg.laps) || (time > cfg.time_limit)) ? 1 : 0;
      }
      last = cur;
      total = total + r;
      if (cfg.repeat_semantics == RD_REPEAT_BASELINES) done |= d; else done = d;
      sh_flag[t] = d;
    }
    __syncthreads();
    if (ticking && cfg.repeat_semantics != RD_REPEAT_BASELINES) {
      int any = 0;
      for (int j = 0; j < A; ++j) any |= sh_flag[base + j];
      if (any) ticking = false;
    }
  }
  // world-level done, rank
  __syncthreads();
  if (run) { sh_flag[t] = done; sh_prog[t] = (double)lap + p; }
  __syncthreads();
  if (frozen) {  // frozen until reset [REF dreamer/wrappers.py:148]
    if (o.reward) o.reward[e] = 0.f;
    if (o.done) o.done[e] = 1;
    if (o.progress) o.progress[e] = (float)f[(size_t)RD_S_PROGRESS * n + e];
    if (o.lap) o.lap[e] = I[(size_t)RD_I_LAP * n + e];
    if (o.time) o.time[e] = (float)f[(size_t)RD_S_TIME * n + e];
    if (o.flags) o.flags[e] = (uint8_t)flags;
    P.recs[e].was_reset = 2;
  } else if (run) {
    int wdone = 0, rank = 1;
    const double mine = sh_prog[t];
    for (int j = 0; j < A; ++j) {
      wdone |= sh_flag[base + j];
      if (j != a) { const double other = sh_prog[base + j]; if (other > mine || (other == mine && j < a)) rank += 1; }
    }
    const int agent_step = agent_step0 + 1;  // TimeLimit [REF dreamer/wrappers.py:147-154]
    int timeout = 0;
    if (cfg.time_limit_steps > 0 && agent_step >= cfg.time_limit_steps) { timeout = !wdone; done = 1; wdone = 1; }
    const double ret = ret0 + total;
#pragma unroll
    for (int k = 0; k < 7; ++k) f[(size_t)k * n + e] = q[k];
    f[(size_t)RD_S_TIME * n + e] = time; f[(size_t)RD_S_PROGRESS * n + e] = p;
    f[(size_t)RD_S_LAST * n + e] = last; f[(size_t)RD_S_RETURN * n + e] = ret;
    I[(size_t)RD_I_LAP * n + e] = lap; I[(size_t)RD_I_CHECKPOINT * n + e] = cp;
    I[(size_t)RD_I_AGENT_STEP * n + e] = agent_step;
    if (wdone && !cfg.auto_reset) flags |= RD_F_NEEDS_RESET;
    I[(size_t)RD_I_FLAGS * n + e] = flags;
    if (o.reward) o.reward[e] = (float)total;
    if (o.done) o.done[e] = (uint8_t)done;
    if (o.progress) o.progress[e] = (float)p;
    if (o.lap) o.lap[e] = lap;
    if (o.time) o.time[e] = (float)time;
    if (o.flags) o.flags[e] = (uint8_t)flags;
    if (o.rank) o.rank[e] = rank;
    if (o.opponents) o.opponents[e] = (uint8_t)opp;
    st[6] = 1.0;
    if (wdone && a == 0) {
      st[0] = 1.0; st[1] = ret; st[2] = ((double)lap + p) - start0;
      st[3] = (double)agent_step; st[4] = (flags & (RD_F_COLLISION | RD_F_OPPONENT)) ? 1.0 : 0.0; st[5] = (double)(lap - 1);
      st[7] = timeout ? 1.0 : 0.0;
    }
    int was_reset = 0;
    if (wdone && cfg.auto_reset) { rd_reset_one(P, e, cfg.reset_mode); was_reset = 1; }
    rd_write_obs(P, o, e, was_reset);
  }
  const unsigned any_done = __ballot_sync(0xffffffffu, st[0] != 0.0);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k != 6 && !any_done) continue;
    double v = st[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(P.stats + k, v);
  }
}

// a1 stage entry: dynamics only, state [7][n] SoA, commands [n][2] sim-facing (rd_dynamics)
__global__ void __launch_bounds__(128) k_dynamics(rd_config cfg, double* __restrict__ state,
                                                   const double* __restrict__ commands, int n, int n_ticks) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double q[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) q[k] = state[(size_t)k * n + e];
  const double motor = commands[2 * e], steering = commands[2 * e + 1];
  const double inv_dt = 1.0 / cfg.dt;
  for (int t = 0; t < n_ticks; ++t) st_tick(cfg, q, motor, steering, inv_dt);
#pragma unroll
  for (int k = 0; k < 7; ++k) state[(size_t)k * n + e] = q[k];
}
