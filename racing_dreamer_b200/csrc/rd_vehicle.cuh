// rd_vehicle.cuh -- the single-track vehicle model of K2, written for the latency of ONE
// env's dependency chain: k_step runs one warp per SM sub-partition, so its time is (instructions per env) x (cycles
// per dependent instruction), not throughput.
//
// Shared by the kernels (rd_dynamics.cuh) and by a host-compiled unit check (tests/native/vehicle_check.cpp, g++), which
// holds these functions against the CPU oracle and the independent NumPy rendition of the model.
//
// Replaces MultiAgentRaceEnv.step -> pybullet.stepSimulation (racecar_gym, not in tree; call site
// [REF dreamer/wrappers.py:63-64]) with the [NEW-SPEC] model of Appendix C: RK4, dt = 0.01 s, float64.
//
// Arithmetic contract: the equations and the RK4 scheme are those of the CPU oracle (its st_rhs / st_tick) and of
// tests/np_single_track.py; the operation ORDER is not (fused multiply-adds, hoisted sub-expressions, own sincos and
// reciprocal).  Results agree with both to ~1e-13 relative after hundreds of ticks (tests: < 1e-9; north_star's bar is
// 1e-5).  The checkers never follow this file; this file must stay within their tolerance.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include "../../include/rd_env.h"

#ifdef __CUDACC__
#define RDV_HD __host__ __device__ __forceinline__
#else
#define RDV_HD static inline
#endif

// ---- float64 helpers: short, branch-free, no slow paths ----
// sin/cos on [-pi/4, pi/4] (fdlibm __kernel_sin / __kernel_cos minimax polynomials, < 1 ulp on an exact argument).  The coefficients live in
// ONE table so that device code reads them from the constant bank (one uniform load per pair) instead of building every
// 64-bit immediate with two moves.
struct RdTrigTab {
  double s[6];   // S1..S6
  double c[6];   // C1..C6
  double two_over_pi, magic, pio2_1, pio2_2, pio2_3;
};
#define RD_TRIG_INIT                                                                                              \
  {{-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,                         \
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10},                         \
   {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,                          \
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11},                        \
   6.36619772367581382433e-01, 6755399441055744.0, 1.57079632679489655800e+00, 6.12323399573676603587e-17,        \
   -1.49738490485916983294e-33}
static const RdTrigTab rd_trig_host = RD_TRIG_INIT;
#ifdef __CUDACC__
static __constant__ RdTrigTab rd_trig_dev = RD_TRIG_INIT;
#endif
#ifdef __CUDA_ARCH__
#define RD_TRIG rd_trig_dev
#else
#define RD_TRIG rd_trig_host
#endif

RDV_HD void rdv_sincos_kernel(double r, double& s, double& c) {
  const RdTrigTab& T = RD_TRIG;
  const double z = r * r;
  double ps = fma(z, T.s[5], T.s[4]);
  double pc = fma(z, T.c[5], T.c[4]);
  ps = fma(z, ps, T.s[3]);
  pc = fma(z, pc, T.c[3]);
  ps = fma(z, ps, T.s[2]);
  pc = fma(z, pc, T.c[2]);
  ps = fma(z, ps, T.s[1]);
  pc = fma(z, pc, T.c[1]);
  ps = fma(z, ps, T.s[0]);
  pc = fma(z, pc, T.c[0]);
  s = fma(r * z, ps, r);
  c = fma(z * z, pc, fma(z, -0.5, 1.0));
}

// sincos for |x| < 2^30 * pi/2 (callers guard with |x| < 1e5: three-term Cody-Waite reduction by pi/2 into ONE word:
// <= 2 ulp there).  The quadrant comes from the low word of x * 2/pi + 1.5 * 2^52 (round to nearest even by the adder), so there
// is no conversion-pipe instruction (FRND / F2I) on the chain.
RDV_HD void rdv_sincos_reduced(double x, double& sn, double& cs) {
  const RdTrigTab& T = RD_TRIG;
  const double t = fma(x, T.two_over_pi, T.magic);
  const double k = t - T.magic;
#ifdef __CUDA_ARCH__
  const int q = __double2loint(t);
#else
  int64_t bits;
  memcpy(&bits, &t, 8);
  const int q = (int)(uint32_t)bits;
#endif
  double r = fma(-k, T.pio2_1, x);
  r = fma(-k, T.pio2_2, r);
  r = fma(-k, T.pio2_3, r);
  double s, c;
  rdv_sincos_kernel(r, s, c);
  const double s1 = (q & 1) ? c : s, c1 = (q & 1) ? s : c;
  sn = (q & 2) ? -s1 : s1;
  cs = ((q + 1) & 2) ? -c1 : c1;
}
#ifdef __CUDACC__
#define RDV_NOINLINE __host__ __device__ __noinline__
#else
#define RDV_NOINLINE static __attribute__((noinline))
#endif
// the rarely taken pieces live out of line so that the hot code stays small (k_step runs one warp per SM sub-partition:
// every instruction-cache miss is paid in full)
RDV_NOINLINE void rdv_sincos_slow(double x, double* sn, double* cs) {   // library path: any argument, NaN, inf
#ifdef __CUDA_ARCH__
  sincos(x, sn, cs);
#else
  *sn = sin(x); *cs = cos(x);
#endif
}
RDV_HD void rdv_sincos(double x, double& sn, double& cs) {
  if (fabs(x) < 1.0e5) rdv_sincos_reduced(x, sn, cs);
  else rdv_sincos_slow(x, &sn, &cs);   // a car that has spun thousands of times (or NaN)
}

// 1/x for normal, finite x (speeds above v_kinematic, cosines of small angles): hardware seed (>= 20 bits) + Newton
// steps, <= 1 ulp
RDV_HD double rdv_rcp(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// a / b correctly rounded from y = RN(1 / b) (Markstein): q0 = a*y is faithful, the residual a - q0*b is exact in one
// fma, and q0 + r*y rounds to RN(a / b).  Used for progress = dist / dmax (small integers; the quotient feeds the integer
// checkpoint index, which must be bit-exact against the oracle's true division); tests/native/vehicle_check.cpp checks
// every (dist, dmax) pair the maps can produce.
RDV_HD double rdv_div_by(double a, double b, double rcp_b) {
  const double q0 = a * rcp_b;
  const double r = fma(-q0, b, a);
  return fma(r, rcp_b, q0);
}

// ---- the model ----
// Everything that depends only on the vehicle parameters and dt, computed once per handle on the host.
struct VehConst {
  double steer_min, steer_max, sv_max, a_max, v_min, v_max, v_switch, v_kin;
  double a_drive, a_brake, c_drag, steer_gain;
  double rl;                 // 1 / (lf + lr)
  double c1, c2;             // mu*m / (I*lwb), mu / lwb
  double A, B;               // lf^2 * c_sf, lr^2 * c_sr
  double glf, glr, h;        // g*lf, g*lr, h_cg
  double lr_csr, lf_csf, c_sr, c_sf, lr, lf;
  double dt, inv_dt, h2, h6;
  double sw_num;             // a_max * v_switch
  int has_switch;            // the vehicle can exceed v_switch (v_max + 1 > v_switch): power-limited acceleration branch
  int pad;
};

// Coefficients of the dynamic regime that depend only on the (constrained) acceleration `ac`
struct VehAcc {
  double S1;                 // A*front + B*rear          (yaw-rate damping, times -c1/v)
  double k_sl, k_st;         // slip and steer gains of the yaw acceleration
  double S2, S3, Sf;         // slip-rate terms: (c_sr*rear*lr - c_sf*front*lf), (c_sr*rear + c_sf*front), c_sf*front
};

static inline VehConst rdv_make_const(const rd_vehicle& p, double dt) {
  VehConst k{};
  const double g = 9.81;
  k.steer_min = p.steer_min; k.steer_max = p.steer_max; k.sv_max = p.steer_vel_max; k.a_max = p.a_max;
  k.v_min = p.v_min; k.v_max = p.v_max; k.v_switch = p.v_switch; k.v_kin = p.v_kinematic;
  k.a_drive = p.a_drive; k.a_brake = p.a_brake; k.c_drag = p.c_drag; k.steer_gain = p.steer_gain;
  const double lwb = p.lf + p.lr;
  k.rl = 1.0 / lwb;
  k.c1 = p.mu * p.mass / (p.inertia * lwb);
  k.c2 = p.mu * k.rl;
  k.A = p.lf * p.lf * p.c_sf;
  k.B = p.lr * p.lr * p.c_sr;
  k.glf = g * p.lf; k.glr = g * p.lr; k.h = p.h_cg;
  k.lr_csr = p.lr * p.c_sr; k.lf_csf = p.lf * p.c_sf; k.c_sr = p.c_sr; k.c_sf = p.c_sf; k.lr = p.lr; k.lf = p.lf;
  k.dt = dt; k.inv_dt = 1.0 / dt; k.h2 = 0.5 * dt; k.h6 = dt / 6.0;
  k.sw_num = p.a_max * p.v_switch;
  k.has_switch = (p.v_max + 1.0 > p.v_switch) ? 1 : 0;
  return k;
}

RDV_HD VehAcc rdv_acc_set(const VehConst& k, double ac) {
  VehAcc s;
  const double rear = fma(ac, k.h, k.glf);      // g*lf + a*h
  const double front = fma(-ac, k.h, k.glr);    // g*lr - a*h
  s.S1 = fma(k.A, front, k.B * rear);
  const double lf_f = k.lf_csf * front;
  s.k_sl = k.c1 * fma(k.lr_csr, rear, -lf_f);
  s.k_st = k.c1 * lf_f;
  const double Sr = k.c_sr * rear;
  s.Sf = k.c_sf * front;
  s.S2 = fma(Sr, k.lr, -(s.Sf * k.lf));
  s.S3 = Sr + s.Sf;
  return s;
}

// Per-tick command: the steering-rate and acceleration requests are constant over the RK4 step; their clipped values and
// the sign tests of the two CommonRoad constraints are therefore per-tick work, only the state-dependent half of each
// constraint (at a stop / at a speed limit) is evaluated per stage.
struct VehCmd {
  double sv_clip;            // steering rate clipped to +-sv_max
  double acc;                // requested acceleration
  double acc_lo;             // max(acc, -a_max)
  bool sv_neg, sv_pos;       // sv <= 0, sv >= 0
  bool acc_neg, acc_pos;     // acc <= 0, acc >= 0
  VehAcc on;                 // coefficient set for ac = clip(acc, -a_max, a_max)
  double ac_on;
};

RDV_HD VehCmd rdv_command(const VehConst& k, const double (&q)[7], double motor, double steering) {
  VehCmd c;
  const double target = steering * k.steer_gain * k.steer_max;
  const double sv = (target - q[2]) * k.inv_dt;
  c.sv_clip = sv <= -k.sv_max ? -k.sv_max : (sv >= k.sv_max ? k.sv_max : sv);
  c.sv_neg = sv <= 0.0; c.sv_pos = sv >= 0.0;
  const double drive = motor >= 0.0 ? motor * k.a_drive : motor * k.a_brake;
  c.acc = fma(-k.c_drag, q[3], drive);
  c.acc_neg = c.acc <= 0.0; c.acc_pos = c.acc >= 0.0;
  c.acc_lo = c.acc <= -k.a_max ? -k.a_max : c.acc;
  c.ac_on = c.acc_lo >= k.a_max ? k.a_max : c.acc_lo;
  c.on = rdv_acc_set(k, c.ac_on);
  return c;
}

RDV_NOINLINE void rdv_acc_set_slow(const VehConst* k, double ac, VehAcc* out) { *out = rdv_acc_set(*k, ac); }

// dq/dt at state t (x, y, steer, v, yaw, yaw_rate, slip), split in two: the position (x, y) enters no derivative, so the
// recurrence between the RK4 stages runs over (steer, v, yaw, yaw_rate, slip) only.  rdv_rhs_core returns those five
// derivatives plus the heading `ang` of the velocity vector; the position derivative (v cos ang, v sin ang) -- the one
// sine / cosine pair both regimes need -- is formed afterwards (rdv_tick_t), off the dependency chain between stages.
// `off` = the coefficient set for ac = 0 (per handle).
// FAST: the caller has checked, for the whole tick, that the angles stay where the short trigonometric paths are valid
// (|heading| < 9e4 rad, |steer| < 0.7 rad) and that the vehicle cannot reach v_switch; then there are no range guards
// and no power-limit branch.  Same arithmetic either way.
template <bool FAST>
RDV_HD void rdv_rhs_core(const VehConst& k, const VehCmd& c, const VehAcc& off, const double (&t)[7], double (&f)[7],
                         double& ang) {
  const double steer = t[2], v = t[3], yaw = t[4], yr = t[5], slip = t[6];
  const bool s_blocked = (steer <= k.steer_min && c.sv_neg) || (steer >= k.steer_max && c.sv_pos);
  const double svc = s_blocked ? 0.0 : c.sv_clip;
  const bool a_blocked = (v <= k.v_min && c.acc_neg) || (v >= k.v_max && c.acc_pos);
  double ac = a_blocked ? 0.0 : c.ac_on;
  bool general = false;      // power-limited branch: the positive limit depends on v (vehicles faster than v_switch only)
  if (!FAST && k.has_switch) {
    if (v > k.v_switch && !a_blocked) {
      const double lim = k.sw_num / v;
      ac = c.acc_lo >= lim ? lim : c.acc_lo;
      general = true;
    }
  }
  const bool kin = fabs(v) < k.v_kin;
  ang = kin ? yaw : (slip + yaw);
  f[2] = svc;
  f[3] = ac;
  if (kin) {
    double ss, cs;   // |steer| <= steer_max (+ an RK4 stage's overshoot) < pi/4: no range reduction needed
    if (FAST || fabs(steer) < 0.78) rdv_sincos_kernel(steer, ss, cs); else rdv_sincos_slow(steer, &ss, &cs);
    const double rc = rdv_rcp(cs);
    const double tn = ss * rc;
    const double vl = v * k.rl;
    f[4] = vl * tn;
    f[5] = fma(ac * k.rl, tn, (vl * (rc * rc)) * svc);
    f[6] = 0.0;
  } else {
    VehAcc s;
    if (!FAST && general) rdv_acc_set_slow(&k, ac, &s);
    else {
      s.S1 = a_blocked ? off.S1 : c.on.S1; s.k_sl = a_blocked ? off.k_sl : c.on.k_sl; s.k_st = a_blocked ? off.k_st : c.on.k_st;
      s.S2 = a_blocked ? off.S2 : c.on.S2; s.S3 = a_blocked ? off.S3 : c.on.S3; s.Sf = a_blocked ? off.Sf : c.on.Sf;
    }
    const double rv = rdv_rcp(v);   // |v| >= v_kinematic here
    const double c2rv = k.c2 * rv;
    const double k_yr = (-k.c1 * rv) * s.S1;
    const double b_yr = fma(c2rv * rv, s.S2, -1.0);
    f[4] = yr;
    f[5] = fma(s.k_st, steer, fma(s.k_sl, slip, k_yr * yr));
    f[6] = fma(c2rv * s.Sf, steer, fma(-(c2rv * s.S3), slip, b_yr * yr));
  }
}
template <bool FAST> RDV_HD void rdv_heading(double ang, double& sn, double& cn);
template <bool FAST>
RDV_HD void rdv_rhs(const VehConst& k, const VehCmd& c, const VehAcc& off, const double (&t)[7], double (&f)[7]) {
  double ang, sn, cn;
  rdv_rhs_core<FAST>(k, c, off, t, f, ang);
  rdv_heading<FAST>(ang, sn, cn);
  f[0] = t[3] * cn;
  f[1] = t[3] * sn;
}
template <bool FAST>
RDV_HD void rdv_heading(double ang, double& sn, double& cn) {
  if (FAST) rdv_sincos_reduced(ang, sn, cn);
  else if (fabs(ang) < 1.0e5) rdv_sincos_reduced(ang, sn, cn);
  else rdv_sincos_slow(ang, &sn, &cn);
}

// One tick of the five coupled states (steer, v, yaw, yaw_rate, slip), stage after stage; q[0], q[1] are not touched.
// The four stage headings `ang` and speeds `vs` it leaves behind are all the position update needs (rdv_tick_position):
// the position never feeds back, so the values are those of the textbook stage-by-stage form, bit for bit -- and the
// two halves can run in different warps (k_step_split, rd_dynamics.cuh).
template <bool FAST>
RDV_HD void rdv_tick_core_t(const VehConst& k, const VehAcc& off, double (&q)[7], double motor, double steering,
                            double (&ang)[4], double (&vs)[4]) {
  const VehCmd c = rdv_command(k, q, motor, steering);
  double kk[7], acc[7], t[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = q[i];
  vs[0] = t[3];
  rdv_rhs_core<FAST>(k, c, off, t, kk, ang[0]);
#pragma unroll
  for (int i = 2; i < 7; ++i) { t[i] = fma(k.h2, kk[i], q[i]); acc[i] = kk[i]; }
  vs[1] = t[3];
  rdv_rhs_core<FAST>(k, c, off, t, kk, ang[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) { t[i] = fma(k.h2, kk[i], q[i]); acc[i] = fma(2.0, kk[i], acc[i]); }
  vs[2] = t[3];
  rdv_rhs_core<FAST>(k, c, off, t, kk, ang[2]);
#pragma unroll
  for (int i = 2; i < 7; ++i) { t[i] = fma(k.dt, kk[i], q[i]); acc[i] = fma(2.0, kk[i], acc[i]); }
  vs[3] = t[3];
  rdv_rhs_core<FAST>(k, c, off, t, kk, ang[3]);
#pragma unroll
  for (int i = 2; i < 7; ++i) q[i] = fma(k.h6, acc[i] + kk[i], q[i]);
}
// x, y: k_s = v_s (cos, sin)(ang_s);  q += h/6 (((k1 + 2 k2) + 2 k3) + k4).  Four INDEPENDENT sine / cosine evaluations
// (one straight-line block the scheduler can interleave).  FAST = false guards each argument (|ang| < 1e5: the same
// short path, same bits; beyond: the library).
template <bool FAST>
RDV_HD void rdv_tick_position(const VehConst& k, const double (&ang)[4], const double (&vs)[4], double& x, double& y) {
  double sn[4], cn[4];
#pragma unroll
  for (int st = 0; st < 4; ++st) rdv_heading<FAST>(ang[st], sn[st], cn[st]);
  const double ax = fma(2.0, vs[2] * cn[2], fma(2.0, vs[1] * cn[1], vs[0] * cn[0]));
  const double ay = fma(2.0, vs[2] * sn[2], fma(2.0, vs[1] * sn[1], vs[0] * sn[0]));
  x = fma(k.h6, ax + vs[3] * cn[3], x);
  y = fma(k.h6, ay + vs[3] * sn[3], y);
}

// one 10 ms tick under the sim-facing command (motor, steering): classical RK4.
// RD_TICK_ROLLED = 1: the four stages share ONE copy of the RHS code (stage input q + c*k with c = (0, h/2, h/2, h), sum
// of w*k with w = (1, 2, 2, 1): the same values, products by 0, 1 and 2 being exact) -- a quarter of the code for the
// instruction cache at the price of a few selects per stage; 0: four inlined copies.
#ifndef RD_TICK_ROLLED
#define RD_TICK_ROLLED 0
#endif
template <bool FAST>
RDV_HD void rdv_tick_t(const VehConst& k, const VehAcc& off, double (&q)[7], double motor, double steering) {
#if RD_TICK_ROLLED
  const VehCmd c = rdv_command(k, q, motor, steering);
  double kk[7] = {0, 0, 0, 0, 0, 0, 0}, acc[7] = {0, 0, 0, 0, 0, 0, 0}, t[7];
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    const double cs = st == 0 ? 0.0 : (st == 3 ? k.dt : k.h2);
    const double w = (st == 0 || st == 3) ? 1.0 : 2.0;
#pragma unroll
    for (int i = 0; i < 7; ++i) t[i] = fma(cs, kk[i], q[i]);
    rdv_rhs<FAST>(k, c, off, t, kk);
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[i] = fma(w, kk[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) q[i] = fma(k.h6, acc[i], q[i]);
#else
  double ang[4], vs[4];
  rdv_tick_core_t<FAST>(k, off, q, motor, steering, ang, vs);
  rdv_tick_position<FAST>(k, ang, vs, q[0], q[1]);
#endif
}
// the general tick, out of line: any heading, any steering angle, vehicles that reach v_switch
RDV_NOINLINE void rdv_tick_slow(const VehConst* k, const VehAcc* off, double* q, double motor, double steering) {
  double r[7];
  for (int i = 0; i < 7; ++i) r[i] = q[i];
  rdv_tick_t<false>(*k, *off, r, motor, steering);
  for (int i = 0; i < 7; ++i) q[i] = r[i];
}
// One tick: the short path when the whole tick provably stays inside its validity range -- the heading moves by
// |yaw_rate| * dt plus the slip change per tick (a rad at the very most), the steering angle by sv_max * dt.
// kg / offg: the same constants behind plain pointers (global memory on the device) for the out-of-line general tick, so
// that taking their address does not force a local copy of the kernel-parameter structs.
RDV_HD void rdv_tick(const VehConst& k, const VehAcc& off, const VehConst* kg, const VehAcc* offg, double (&q)[7],
                     double motor, double steering) {
  const bool fast = !k.has_switch && fabs(q[4]) < 9.0e4 && fabs(q[6]) < 1.0e3 && fabs(q[5]) < 1.0e3 && fabs(q[2]) < 0.6;
  if (fast) rdv_tick_t<true>(k, off, q, motor, steering);
  else rdv_tick_slow(kg, offg, q, motor, steering);
}
// The same tick without the position: the coupled states advance, the stage headings / speeds are handed out.
RDV_NOINLINE void rdv_tick_core_slow(const VehConst* k, const VehAcc* off, double* q, double motor, double steering,
                                     double* ang, double* vs) {
  double r[7], a[4], v[4];
  for (int i = 0; i < 7; ++i) r[i] = q[i];
  rdv_tick_core_t<false>(*k, *off, r, motor, steering, a, v);
  for (int i = 2; i < 7; ++i) q[i] = r[i];
  for (int i = 0; i < 4; ++i) { ang[i] = a[i]; vs[i] = v[i]; }
}
RDV_HD void rdv_tick_core(const VehConst& k, const VehAcc& off, const VehConst* kg, const VehAcc* offg, double (&q)[7],
                          double motor, double steering, double (&ang)[4], double (&vs)[4]) {
  const bool fast = !k.has_switch && fabs(q[4]) < 9.0e4 && fabs(q[6]) < 1.0e3 && fabs(q[5]) < 1.0e3 && fabs(q[2]) < 0.6;
  if (fast) rdv_tick_core_t<true>(k, off, q, motor, steering, ang, vs);
  else rdv_tick_core_slow(kg, offg, q, motor, steering, ang, vs);
}
