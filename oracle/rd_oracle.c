/* rd_oracle.c -- CPU restatement (float64 / integer, plain C) of the racing-environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in racing_dreamer_b200/ may include, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or the CPU baseline.
 *
 * PARITY STATUS
 *   - a3/a4/a9/a10/a11 (wrapper arithmetic) and a5 (OccupancyMapObs) are PINNED: tests/golden holds
 *     vectors produced by the unmodified reference classes of dreamer/wrappers.py (imported under
 *     sys.modules stubs, scipy + Pillow of this image) -- see tests/golden/make_golden.py.
 *   - a1 dynamics, a2 LiDAR, a7 progress/lap, a8 reward/done, collision and reset sampling are
 *     **PARITY UNPINNED**: in the reference that arithmetic lives in the un-vendored third-party
 *     packages racecar_gym (@icra22 / @gym-api, commits a9e6f5f..., e117432...) on pybullet 3.0.8/3.2.1
 *     [REF dreamer/requirements.txt:6; baselines/docker/requirements_acme.txt:85,100], absent from
 *     /root/reference and not installable offline; the reference has no tests or golden vectors.
 *     For those stages this file IS the specification (north_star: "a float64 ... rendition of the
 *     same model"), anchored on the reference's call sites and constants cited per function.
 *   - multi-agent worlds: the dict-of-agents semantics (ActionRepeat stops when ANY agent is
 *     done and sums per agent, TimeLimit sets every done, reset when any agent is done) are PINNED against the
 *     unmodified reference wrappers (tests/golden/multi_agent_stack_golden.npz); car-car contact, the other
 *     cars in the scans, rank, the n_step_progress task and the random_ball reset are racecar_gym arithmetic
 *     -> **PARITY UNPINNED**, [NEW-SPEC] here.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  No FMA contraction:
 * every floating-point operation below is one IEEE-754 operation in the order written.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rd_env.h"

#define ORC_API __attribute__((visibility("default")))

/* One compiled track (racing_dreamer_b200.maps.TrackMap), unpacked, y-up rows.
 * [REF docs/maps/costmaps/generate-costmap.py:405-420 npz keys drivable_area / norm_distance_from_start] */
typedef struct orc_map {
  int32_t h, w;
  int32_t col0, row0;        /* full-image cell index (x, y-up) of crop cell (0,0) */
  int32_t full_h;
  int32_t dmax;
  double resolution, origin_x, origin_y;
  const uint8_t* drivable;   /* [h][w] y-up, 1 = drivable */
  const uint16_t* dist;      /* [h][w] y-up wavefront distance */
  const double* start_poses; /* [n_start][3] */
  const double* reset_poses; /* [n_reset][3] */
  int32_t n_start, n_reset;
  const int32_t* ball_next;  /* [n_reset] orc_ball_next(): the reset pose ball_spacing metres further along the lap */
} orc_map;

/* ------------------------------------------------------------------------------------------------ */
/* defaults: dreamer training setup [REF dreamer/dream.py:55-58,138; scenarios/max_progress/austria.yml] */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_default_config(rd_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = RD_ABI_VERSION;
  c->n_envs = 1;
  c->n_beams = 1080;
  c->action_repeat = 4;
  c->repeat_semantics = RD_REPEAT_DREAMER;
  c->obs_flags = RD_OBS_LIDAR;
  c->task = RD_TASK_MAX_PROGRESS;
  c->laps = 10;
  c->terminate_on_collision = 1;
  c->n_checkpoints = 20;
  c->time_limit_steps = 0;
  c->auto_reset = 0;
  c->reset_mode = RD_RESET_GRID;
  c->rescale_actions = 1;
  c->clip_actions = 0;
  c->progress_abs = 0;
  c->env_id_offset = 0;
  c->seed = 0;
  c->dt = 0.01;
  c->time_limit = 180.0;
  c->collision_reward = -1.0;
  c->progress_reward = 100.0;
  c->frame_reward = 0.0;
  c->action_low[0] = 0.005; c->action_low[1] = -1.0;
  c->action_high[0] = 1.0;  c->action_high[1] = 1.0;
  c->lidar_fov = 270.0 * (3.14159265358979323846 / 180.0);
  c->lidar_range_min = 0.25;
  c->lidar_range_max = 15.0;
  c->lidar_offset = 0.0;
  c->lidar_noise = 0.0f;
  c->agents_per_world = 1;
  for (int a = 0; a < RD_MAX_AGENTS; ++a) c->agent_task[a] = RD_TASK_MAX_PROGRESS;
  c->n_step_progress = 10; /* [REF baselines/scenarios/max_progress/austria.yml:18] */
  c->ball_spacing = 1.5;
  c->time_limit_ticks = 0;
  /* Box bounds of the observation space [NEW-SPEC: racecar_gym's sensor spaces are not in tree] */
  c->obs_low[RD_NORM_LIDAR] = 0.0;      c->obs_high[RD_NORM_LIDAR] = 15.0;
  c->obs_low[RD_NORM_POSE] = -100.0;    c->obs_high[RD_NORM_POSE] = 100.0;
  c->obs_low[RD_NORM_VELOCITY] = -10.0; c->obs_high[RD_NORM_VELOCITY] = 10.0;
  rd_vehicle* v = &c->vehicle;
  v->mu = 1.0489; v->c_sf = 4.718; v->c_sr = 5.4562; v->lf = 0.15875; v->lr = 0.17145; v->h_cg = 0.074;
  v->mass = 3.74; v->inertia = 0.04712;
  v->steer_min = -0.42; v->steer_max = 0.42; v->steer_vel_max = 3.2;
  v->v_switch = 7.319; v->a_max = 9.51; v->v_min = 0.0; v->v_max = 5.0;
  v->v_kinematic = 0.5;
  v->a_drive = 6.0; v->a_brake = 8.26; v->c_drag = 1.0;
  v->steer_gain = -1.0;  /* positive steering action = right turn [REF ros_agent/agents/dreamer/src/agent.py:111] */
  v->body_length = 0.50; v->body_width = 0.27;
}

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), counter-based RNG for reset sampling and LiDAR noise          */
/* ------------------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
enum { STREAM_RESET = 0x52455345, STREAM_LIDAR = 0x4c494441 };

/* ------------------------------------------------------------------------------------------------ */
/* map lookups.  Cell of a world point: col = floor((x-ox)*inv_res), row_up = floor((y-oy)*inv_res)   */
/* [REF docs/maps/costmaps/generate-costmap.py:49-52 grid = (world-origin)/resolution, y flipped]     */
/* ------------------------------------------------------------------------------------------------ */
static inline int cell_of(const orc_map* m, double x, double y, int* cx, int* cy) {
  double inv_res = 1.0 / m->resolution;
  double u = (x - m->origin_x) * inv_res;
  double v = (y - m->origin_y) * inv_res;
  double fu = floor(u), fv = floor(v);
  /* far outside: avoid int overflow */
  if (!(fu > -1.0e9 && fu < 1.0e9 && fv > -1.0e9 && fv < 1.0e9)) { *cx = -1; *cy = -1; return 0; }
  *cx = (int)fu - m->col0;
  *cy = (int)fv - m->row0;
  return (*cx >= 0 && *cx < m->w && *cy >= 0 && *cy < m->h);
}
static inline int drivable_at(const orc_map* m, int cx, int cy) {
  if (cx < 0 || cx >= m->w || cy < 0 || cy >= m->h) return 0;
  return m->drivable[(size_t)cy * m->w + cx];
}

/* ------------------------------------------------------------------------------------------------ */
/* a2 LiDAR.  Reference: sensor 'lidar' [REF dreamer/scenarios/max_progress/austria.yml:7] served by  */
/* pybullet.rayTestBatch inside racecar_gym (not in tree).  In-tree pins: 1080 beams                  */
/* [REF dreamer/dream.py:66], 15 m [REF dreamer/tools.py:274], 270 deg, beam 0 = +135 deg (left),     */
/* last = -135 deg, endpoints included [REF dreamer/tools.py:84-86 linspace(...)[::-1]].              */
/* [NEW-SPEC] exact integer grid traversal: origin quantised to 2^-12 cell, direction to 2^-18;        */
/* hit = first non-drivable cell entered; range = distance to the crossing into that cell.           */
/* ------------------------------------------------------------------------------------------------ */
#define SUB_BITS 12
#define SUB (1 << SUB_BITS)
#define DIR_BITS 18

ORC_API void orc_beam_table(const rd_config* cfg, double* ca, double* sa) {
  int n = cfg->n_beams;
  for (int i = 0; i < n; ++i) {
    double a = (n > 1) ? (0.5 * cfg->lidar_fov - (double)i * (cfg->lidar_fov / (double)(n - 1))) : 0.0;
    ca[i] = cos(a);
    sa[i] = sin(a);
  }
}

static float finish_range(const rd_config* cfg, float r, uint32_t gid, uint32_t episode, uint32_t step,
                          uint32_t beam) {
  float rmin = (float)cfg->lidar_range_min, rmax = (float)cfg->lidar_range_max;
  if (cfg->lidar_noise > 0.0f) {
    uint32_t c[4] = {gid, episode, step, beam};
    philox4x32_10(c, (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32) ^ (uint32_t)STREAM_LIDAR);
    float u = (float)(c[0] >> 8) * (1.0f / 16777216.0f); /* [0,1) exact */
    float t = u * 2.0f - 1.0f;
    float f = 1.0f + cfg->lidar_noise * t;
    r = r * f;
  }
  r = r < rmin ? rmin : r;
  r = r > rmax ? rmax : r;
  if (cfg->obs_flags & RD_OBS_LIDAR_NORM) r = r / rmax - 0.5f; /* [REF dreamer/tools.py:274] */
  /* NormalizeObservations: (observation - low) * scaler, scaler = 1.0 / (high - low), float64 arrays
   * [REF baselines/racing/environment/single_agent.py:66-99]; SinglePrecisionWrapper rounds to float32 afterwards */
  else if (cfg->obs_flags & RD_OBS_NORM_BASELINES)
    r = (float)(((double)r - cfg->obs_low[RD_NORM_LIDAR]) * (1.0 / (cfg->obs_high[RD_NORM_LIDAR] - cfg->obs_low[RD_NORM_LIDAR])));
  return r;
}

/* Another car of the same world as the scan sees it (multi-agent worlds): ITS sensor origin in
 * 2^-12 cells and heading.  [NEW-SPEC: racecar_gym's rayTestBatch also hits the other racecars' collision shapes] */
typedef struct orc_car { int32_t px, py; int has_pos; double c, s; } orc_car;

/* sensor origin of a pose, quantised to 2^-12 cells relative to the crop; returns 1 when representable */
static int sensor_origin(const rd_config* cfg, const orc_map* m, double x, double y, double yaw, int64_t* PX,
                         int64_t* PY, double* c_out, double* s_out) {
  const double inv_res = 1.0 / m->resolution;
  const double c = cos(yaw), s = sin(yaw);
  const double sx = x + cfg->lidar_offset * c;
  const double sy = y + cfg->lidar_offset * s;
  const double u = (sx - m->origin_x) * inv_res;
  const double v = (sy - m->origin_y) * inv_res;
  const double pu = floor(u * (double)SUB), pv = floor(v * (double)SUB);
  *c_out = c; *s_out = s;
  *PX = 0; *PY = 0;
  if (!(pu > -1.0e12 && pu < 1.0e12 && pv > -1.0e12 && pv < 1.0e12)) return 0;
  *PX = (int64_t)pu - (int64_t)m->col0 * SUB;
  *PY = (int64_t)pv - (int64_t)m->row0 * SUB;
  return (*PX > -((int64_t)1 << 30) && *PX < ((int64_t)1 << 30) && *PY > -((int64_t)1 << 30) && *PY < ((int64_t)1 << 30));
}

static orc_car car_of(const rd_config* cfg, const orc_map* m, double x, double y, double yaw) {
  orc_car k;
  int64_t PX, PY;
  k.has_pos = sensor_origin(cfg, m, x, y, yaw, &PX, &PY, &k.c, &k.s);
  k.px = k.has_pos ? (int32_t)PX : 0;
  k.py = k.has_pos ? (int32_t)PY : 0;
  return k;
}

/* Range (metres) at which the beam (DX, DY) * 2^-18 from (px, py) enters the body box of car q, or +inf: float32 slab
 * test in q's sensor frame, one IEEE operation per step in this order (the kernel's rd_car_hit). */
static float car_hit(const rd_config* cfg, const orc_map* m, int32_t px, int32_t py, int32_t DX, int32_t DY,
                     const orc_car* q) {
  const double inv_res = 1.0 / m->resolution;
  const double hl = 0.5 * cfg->vehicle.body_length * inv_res, hw = 0.5 * cfg->vehicle.body_width * inv_res;
  const double off = cfg->lidar_offset * inv_res;
  const float ulo = (float)(-hl - off), uhi = (float)(hl - off), vhw = (float)hw, res = (float)m->resolution;
  const int reach = (int)ceil((cfg->lidar_range_max * inv_res + 2.0 * (hl + hw + fabs(off)) + 2.0) * (double)SUB);
  const int32_t rx = px - q->px, ry = py - q->py;
  const int32_t arx = rx < 0 ? -rx : rx, ary = ry < 0 ? -ry : ry;
  if (!q->has_pos || arx > reach || ary > reach) return INFINITY;
  const float ox = (float)rx * (1.0f / (float)SUB), oy = (float)ry * (1.0f / (float)SUB);
  const float c = (float)q->c, s = (float)q->s;
  const float dxf = (float)DX * (1.0f / (float)(1 << DIR_BITS)), dyf = (float)DY * (1.0f / (float)(1 << DIR_BITS));
  const float u0 = ox * c + oy * s;
  const float v0 = oy * c - ox * s;
  const float du = dxf * c + dyf * s;
  const float dv = dyf * c - dxf * s;
  float tmin = 0.0f, tmax = INFINITY;
  if (du != 0.0f) {
    float t1 = (ulo - u0) / du, t2 = (uhi - u0) / du;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  } else if (u0 < ulo || u0 > uhi) {
    return INFINITY;
  }
  if (dv != 0.0f) {
    float t1 = (-vhw - v0) / dv, t2 = (vhw - v0) / dv;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  } else if (v0 < -vhw || v0 > vhw) {
    return INFINITY;
  }
  return tmin <= tmax ? tmin * res : INFINITY;
}

static void lidar_one(const rd_config* cfg, const orc_map* m, const double* ca, const double* sa, double x,
                      double y, double yaw, uint32_t gid, uint32_t episode, uint32_t step, float* out,
                      const orc_car* cars, int n_cars) {
  const int nb = cfg->n_beams;
  const double inv_res = 1.0 / m->resolution;
  const double c = cos(yaw), s = sin(yaw);
  const double sx = x + cfg->lidar_offset * c;
  const double sy = y + cfg->lidar_offset * s;
  const double u = (sx - m->origin_x) * inv_res;
  const double v = (sy - m->origin_y) * inv_res;
  const double pu = floor(u * (double)SUB), pv = floor(v * (double)SUB);
  const float scale = (float)((double)(1 << (DIR_BITS - SUB_BITS)) * m->resolution);
  const int64_t rsub = (int64_t)rint(cfg->lidar_range_max * inv_res * (double)SUB);
  int ok = (pu > -1.0e12 && pu < 1.0e12 && pv > -1.0e12 && pv < 1.0e12);
  int64_t PX = 0, PY = 0;
  int ix0 = -1, iy0 = -1;
  if (ok) {
    PX = (int64_t)pu - (int64_t)m->col0 * SUB;
    PY = (int64_t)pv - (int64_t)m->row0 * SUB;
    ok = (PX >= 0 && PY >= 0 && (PX >> SUB_BITS) < m->w && (PY >> SUB_BITS) < m->h);
  }
  if (ok) {
    ix0 = (int)(PX >> SUB_BITS);
    iy0 = (int)(PY >> SUB_BITS);
    ok = drivable_at(m, ix0, iy0);
  }
  if (!ok) { /* sensor outside the drivable area: every beam reads 0 (-> range_min) */
    for (int i = 0; i < nb; ++i) out[i] = finish_range(cfg, 0.0f, gid, episode, step, (uint32_t)i);
    return;
  }
  const int32_t fx = (int32_t)(PX & (SUB - 1)), fy = (int32_t)(PY & (SUB - 1));
  for (int i = 0; i < nb; ++i) {
    double dx = c * ca[i] - s * sa[i];
    double dy = s * ca[i] + c * sa[i];
    int32_t DX = (int32_t)rint(dx * (double)(1 << DIR_BITS));
    int32_t DY = (int32_t)rint(dy * (double)(1 << DIR_BITS));
    int32_t adx = DX < 0 ? -DX : DX, ady = DY < 0 ? -DY : DY;
    int stepx = DX > 0 ? 1 : -1, stepy = DY > 0 ? 1 : -1;
    int32_t bx = DX > 0 ? SUB - fx : fx;
    int32_t by = DY > 0 ? SUB - fy : fy;
    int32_t e = (int32_t)((int64_t)bx * ady - (int64_t)by * adx);
    if (ady == 0) e = -1;
    /* crossings no farther than range_max along the ray */
    int64_t lx = (rsub * adx) >> DIR_BITS, ly = (rsub * ady) >> DIR_BITS;
    int nx = (adx != 0 && lx >= bx) ? (int)((lx - bx) >> SUB_BITS) + 1 : 0;
    int ny = (ady != 0 && ly >= by) ? (int)((ly - by) >> SUB_BITS) + 1 : 0;
    int nmax = nx + ny;
    int ix = ix0, iy = iy0, hit = 0, lastx = 0;
    const int32_t ex = ady << SUB_BITS, ey = adx << SUB_BITS;
    for (int n = 0; n < nmax; ++n) {
      if (e < 0) { ix += stepx; e += ex; lastx = 1; }
      else       { iy += stepy; e -= ey; lastx = 0; }
      if (!drivable_at(m, ix, iy)) { hit = 1; break; }
    }
    float r;
    if (hit) {
      int32_t num, den;
      if (lastx) { int k = (ix > ix0 ? ix - ix0 : ix0 - ix) - 1; num = bx + k * SUB; den = adx; }
      else       { int k = (iy > iy0 ? iy - iy0 : iy0 - iy) - 1; num = by + k * SUB; den = ady; }
      r = (float)num / (float)den;
      r = r * scale;
    } else {
      r = (float)cfg->lidar_range_max;
    }
    for (int k = 0; k < n_cars; ++k) r = fminf(r, car_hit(cfg, m, (int32_t)PX, (int32_t)PY, DX, DY, &cars[k]));
    out[i] = finish_range(cfg, r, gid, episode, step, (uint32_t)i);
  }
}

/* poses [n][3] = (x, y, yaw); map_ids NULL = map 0; ranges [n][n_beams] */
ORC_API void orc_lidar_cast(const rd_config* cfg, const orc_map* maps, const double* poses,
                            const int32_t* map_ids, int n, float* ranges, int n_threads) {
  int nb = cfg->n_beams;
  double* ca = (double*)malloc(sizeof(double) * 2 * nb);
  double* sa = ca + nb;
  orc_beam_table(cfg, ca, sa);
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
  for (int e = 0; e < n; ++e) {
    const orc_map* m = &maps[map_ids ? map_ids[e] : 0];
    orc_car cars[RD_MAX_AGENTS];
    int n_cars = 0;
    const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
    for (int j = e - e % A; j < e - e % A + A && j < n && A > 1; ++j)
      if (j != e) cars[n_cars++] = car_of(cfg, m, poses[3 * j], poses[3 * j + 1], poses[3 * j + 2]);
    lidar_one(cfg, m, ca, sa, poses[3 * e], poses[3 * e + 1], poses[3 * e + 2], (uint32_t)e, 0u, 0u,
              ranges + (size_t)e * nb, cars, n_cars);
  }
  free(ca);
}

/* ------------------------------------------------------------------------------------------------ */
/* a1 vehicle dynamics.  Reference: MultiAgentRaceEnv.step -> pybullet.stepSimulation (not in tree;   */
/* call site [REF dreamer/wrappers.py:63-64]).  [NEW-SPEC] single-track model of SURVEY.md Appendix C */
/* (CommonRoad/f1tenth formulation), RK4 at dt = 0.01 s, float64.                                     */
/* state q = (x, y, steer, v, yaw, yaw_rate, slip)                                                    */
/* ------------------------------------------------------------------------------------------------ */
static void st_rhs(const rd_vehicle* p, const double q[7], double sv, double acc, double f[7]) {
  const double g = 9.81;
  const double steer = q[2], v = q[3], yaw = q[4], yr = q[5], slip = q[6];
  /* steering constraint */
  double svc;
  if ((steer <= p->steer_min && sv <= 0.0) || (steer >= p->steer_max && sv >= 0.0)) svc = 0.0;
  else if (sv <= -p->steer_vel_max) svc = -p->steer_vel_max;
  else if (sv >= p->steer_vel_max) svc = p->steer_vel_max;
  else svc = sv;
  /* acceleration constraint */
  double pos_limit = (v > p->v_switch) ? (p->a_max * p->v_switch / v) : p->a_max;
  double ac;
  if ((v <= p->v_min && acc <= 0.0) || (v >= p->v_max && acc >= 0.0)) ac = 0.0;
  else if (acc <= -p->a_max) ac = -p->a_max;
  else if (acc >= pos_limit) ac = pos_limit;
  else ac = acc;
  const double lwb = p->lf + p->lr;
  const double rl = 1.0 / lwb;
  /* one sine/cosine pair serves both regimes: heading of the velocity vector */
  const int kin = fabs(v) < p->v_kinematic;
  const double ang = kin ? yaw : (slip + yaw);
  const double sn = sin(ang), cn = cos(ang);
  f[0] = v * cn;
  f[1] = v * sn;
  f[2] = svc;
  f[3] = ac;
  if (kin) {
    const double ss = sin(steer), cs = cos(steer);
    const double rc = 1.0 / cs;
    const double tn = ss * rc;
    f[4] = (v * rl) * tn;
    f[5] = (ac * rl) * tn + ((v * rl) * (rc * rc)) * svc;
    f[6] = 0.0;
  } else {
    const double rv = 1.0 / v;
    const double c1 = p->mu * p->mass / (p->inertia * lwb);
    const double c2 = p->mu * rl;
    double rear = g * p->lf + ac * p->h_cg;   /* g*lf + a*h */
    double front = g * p->lr - ac * p->h_cg;  /* g*lr - a*h */
    double k_yr = (-c1 * rv) * (p->lf * p->lf * p->c_sf * front + p->lr * p->lr * p->c_sr * rear);
    double k_sl = c1 * (p->lr * p->c_sr * rear - p->lf * p->c_sf * front);
    double k_st = c1 * (p->lf * p->c_sf * front);
    double b_yr = (c2 * (rv * rv)) * (p->c_sr * rear * p->lr - p->c_sf * front * p->lf) - 1.0;
    double b_sl = (c2 * rv) * (p->c_sr * rear + p->c_sf * front);
    double b_st = (c2 * rv) * (p->c_sf * front);
    f[4] = yr;
    f[5] = (k_yr * yr + k_sl * slip) + k_st * steer;
    f[6] = (b_yr * yr - b_sl * slip) + b_st * steer;
  }
}

/* one 10 ms tick under the sim-facing command (motor, steering) in [-1,1]
 * [REF dreamer/dream.py:138: motor in [0.005,1], steering in [-1,1] after ReduceActionSpace] */
static void st_tick(const rd_config* cfg, double q[7], double motor, double steering) {
  const rd_vehicle* p = &cfg->vehicle;
  const double dt = cfg->dt;
  double target = steering * p->steer_gain * p->steer_max;
  double sv = (target - q[2]) / dt;
  double acc = (motor >= 0.0) ? (motor * p->a_drive - p->c_drag * q[3]) : (motor * p->a_brake - p->c_drag * q[3]);
  double k1[7], k2[7], k3[7], k4[7], t[7];
  const double h2 = 0.5 * dt, h6 = dt / 6.0;
  st_rhs(p, q, sv, acc, k1);
  for (int i = 0; i < 7; ++i) t[i] = q[i] + h2 * k1[i];
  st_rhs(p, t, sv, acc, k2);
  for (int i = 0; i < 7; ++i) t[i] = q[i] + h2 * k2[i];
  st_rhs(p, t, sv, acc, k3);
  for (int i = 0; i < 7; ++i) t[i] = q[i] + dt * k3[i];
  st_rhs(p, t, sv, acc, k4);
  for (int i = 0; i < 7; ++i) q[i] = q[i] + h6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
}

/* state [7][n] SoA in/out; commands [n][2] sim-facing */
ORC_API void orc_dynamics(const rd_config* cfg, double* state, const double* commands, int n, int n_ticks) {
  for (int e = 0; e < n; ++e) {
    double q[7];
    for (int i = 0; i < 7; ++i) q[i] = state[(size_t)i * n + e];
    for (int t = 0; t < n_ticks; ++t) st_tick(cfg, q, commands[2 * e], commands[2 * e + 1]);
    for (int i = 0; i < 7; ++i) state[(size_t)i * n + e] = q[i];
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* a7 progress / lap / wrong-way, a8 reward / done, collision, reset.  Consumers in tree:             */
/* [REF dreamer/wrappers.py:218-219 lap+progress-1; dreamer/tools.py:195; tasks.py:8-11];             */
/* task parameters [REF dreamer/scenarios/max_progress/austria.yml:8-10].  Implementation in          */
/* racecar_gym (not in tree) -> [NEW-SPEC], SURVEY.md §8 a7/a8.                                       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct orc_view { double* f[RD_NF64]; int32_t* i[RD_NI32]; double* hist; int n; } orc_view;
static orc_view view_of(double* f64, int32_t* i32, int n) {
  orc_view v;
  for (int k = 0; k < RD_NF64; ++k) v.f[k] = f64 + (size_t)k * n;
  for (int k = 0; k < RD_NI32; ++k) v.i[k] = i32 + (size_t)k * n;
  v.hist = NULL; v.n = n;
  return v;
}

static int checkpoint_of(const rd_config* cfg, double p) {
  int c = (int)(p * (double)cfg->n_checkpoints);
  return c > cfg->n_checkpoints - 1 ? cfg->n_checkpoints - 1 : c;
}

/* progress at (x,y): wavefront distance / dmax; returns 0 and leaves *p when the cell is not drivable */
static int progress_at(const orc_map* m, double x, double y, double* p) {
  int cx, cy;
  if (!cell_of(m, x, y, &cx, &cy)) return 0;
  if (!m->drivable[(size_t)cy * m->w + cx]) return 0;
  *p = (double)m->dist[(size_t)cy * m->w + cx] / (double)m->dmax;
  return 1;
}

static int collides(const rd_config* cfg, const orc_map* m, double x, double y, double yaw) {
  int cx, cy;
  if (!cell_of(m, x, y, &cx, &cy) || !drivable_at(m, cx, cy)) return 1;
  double c = cos(yaw), s = sin(yaw);
  double hl = 0.5 * cfg->vehicle.body_length, hw = 0.5 * cfg->vehicle.body_width;
  double ax = hl * c, ay = hl * s, bx = hw * s, by = hw * c;
  /* front-left, front-right, rear-left, rear-right */
  double px[4] = {(x + ax) - bx, (x + ax) + bx, (x - ax) - bx, (x - ax) + bx};
  double py[4] = {(y + ay) + by, (y + ay) - by, (y - ay) + by, (y - ay) - by};
  for (int k = 0; k < 4; ++k) {
    if (!cell_of(m, px[k], py[k], &cx, &cy) || !drivable_at(m, cx, cy)) return 1;
  }
  return 0;
}

/* random_ball / multi-agent random reset [REF dreamer/dream.py:105-106; sampler in racecar_gym -> NEW-SPEC]: the cars of
 * a world line up behind one random anchor pose.  next[i] = the reset pose with the smallest wavefront distance that is
 * at least ceil(ball_spacing / resolution) cells of progress beyond pose i (wrapping over the finish line); the
 * wavefront distance is a Chebyshev path length, so consecutive cars stand at least ball_spacing metres apart. */
typedef struct ball_key { int32_t d, i; } ball_key;
static int ball_key_cmp(const void* a, const void* b) {
  const ball_key* x = (const ball_key*)a; const ball_key* y = (const ball_key*)b;
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
static int ball_lower_bound(const ball_key* k, int n, int32_t d) { /* first entry with k.d >= d, or n */
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) / 2; if (k[mid].d < d) lo = mid + 1; else hi = mid; }
  return lo;
}
ORC_API void orc_ball_next(const rd_config* cfg, const orc_map* m, int32_t* next) {
  const int n = m->n_reset;
  if (n <= 0) return;
  ball_key* k = (ball_key*)malloc(sizeof(ball_key) * (size_t)n);
  int32_t* d = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    int cx, cy;
    d[i] = cell_of(m, m->reset_poses[3 * i], m->reset_poses[3 * i + 1], &cx, &cy) ? (int32_t)m->dist[(size_t)cy * m->w + cx] : 0;
    k[i].d = d[i]; k[i].i = i;
  }
  qsort(k, (size_t)n, sizeof(ball_key), ball_key_cmp);
  const int32_t sc = (int32_t)ceil(cfg->ball_spacing / m->resolution);
  for (int i = 0; i < n; ++i) {
    int j = ball_lower_bound(k, n, d[i] + sc);
    if (j == n) j = ball_lower_bound(k, n, d[i] + sc - m->dmax);
    next[i] = j < n ? k[j].i : i;
  }
  free(d);
  free(k);
}

static void reset_one(const rd_config* cfg, const orc_map* maps, orc_view* s, int e, int mode) {
  const orc_map* m = &maps[s->i[RD_I_MAP][e]];
  uint32_t episode = (uint32_t)s->i[RD_I_EPISODE][e];
  /* multi-agent worlds: one anchor per world (counter = global id of its agent 0), the other cars follow the
   * ball_next chain (cfg->ball_spacing metres of track apart); 'grid' = the staggered start slots */
  const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
  const int a = e % A;
  uint64_t gid = (uint64_t)(cfg->env_id_offset + (e - a));
  double x, y, yaw;
  if (mode == RD_RESET_GRID || m->n_reset <= 0) {
    int slot = a < m->n_start ? a : m->n_start - 1;
    x = m->start_poses[3 * slot]; y = m->start_poses[3 * slot + 1]; yaw = m->start_poses[3 * slot + 2];
  } else {
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), episode, 0u};
    philox4x32_10(c, (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32) ^ (uint32_t)STREAM_RESET);
    uint32_t idx = (uint32_t)(((uint64_t)c[0] * (uint64_t)m->n_reset) >> 32);
    for (int k = 0; k < a && m->ball_next; ++k) idx = (uint32_t)m->ball_next[idx];
    x = m->reset_poses[3 * idx]; y = m->reset_poses[3 * idx + 1]; yaw = m->reset_poses[3 * idx + 2];
    if (mode == RD_RESET_RANDOM_BIDIRECTIONAL && (c[1] & 1u)) yaw = yaw + 3.14159265358979323846;
  }
  s->f[RD_S_X][e] = x; s->f[RD_S_Y][e] = y; s->f[RD_S_STEER][e] = 0.0; s->f[RD_S_V][e] = 0.0;
  s->f[RD_S_YAW][e] = yaw; s->f[RD_S_YAWRATE][e] = 0.0; s->f[RD_S_SLIP][e] = 0.0;
  s->f[RD_S_TIME][e] = 0.0;
  double p = 0.0;
  progress_at(m, x, y, &p);
  s->f[RD_S_PROGRESS][e] = p;
  s->f[RD_S_LAST][e] = 1.0 + p;
  s->f[RD_S_START][e] = 1.0 + p;
  s->f[RD_S_RETURN][e] = 0.0;
  s->f[RD_S_MAXPROG][e] = -1.0; /* no step yet: lap + progress - 1 is never negative */
  s->i[RD_I_LAP][e] = 1;
  s->i[RD_I_CHECKPOINT][e] = checkpoint_of(cfg, p);
  s->i[RD_I_FLAGS][e] = 0;
  s->i[RD_I_AGENT_STEP][e] = 0;
  s->i[RD_I_EPISODE][e] = (int32_t)(episode + 1u);
  if (s->hist) for (int k = 0; k < cfg->n_step_progress; ++k) s->hist[(size_t)k * s->n + e] = 1.0 + p;
}

/* outputs of one step, host arrays (any may be NULL) -- mirrors rd_outputs */
typedef struct orc_outputs {
  float* lidar; uint8_t* occupancy; float* pose; float* velocity; float* speed; float* reward;
  uint8_t* done; float* progress; int32_t* lap; float* time; uint8_t* flags;
  double* reward64; /* un-rounded step reward (the reference sums python floats [REF dreamer/wrappers.py:114]) */
  int32_t* rank; uint8_t* opponents; /* multi-agent worlds */
} orc_outputs;

static void orc_occupancy_one(const orc_map* m, double x, double y, double yaw, uint8_t* out);

static void write_obs(const rd_config* cfg, const orc_map* maps, const double* ca, const double* sa,
                      orc_view* s, int e, const orc_outputs* o, int zero_occupancy) {
  const int nb = cfg->n_beams;
  const orc_map* m = &maps[s->i[RD_I_MAP][e]];
  double x = s->f[RD_S_X][e], y = s->f[RD_S_Y][e], yaw = s->f[RD_S_YAW][e];
  double v = s->f[RD_S_V][e], slip = s->f[RD_S_SLIP][e];
  if (o->lidar) {
    orc_car cars[RD_MAX_AGENTS];
    int n_cars = 0;
    const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
    for (int j = e - e % A; j < e - e % A + A && A > 1; ++j)
      if (j != e) cars[n_cars++] = car_of(cfg, m, s->f[RD_S_X][j], s->f[RD_S_Y][j], s->f[RD_S_YAW][j]);
    lidar_one(cfg, m, ca, sa, x, y, yaw, (uint32_t)(cfg->env_id_offset + e), (uint32_t)s->i[RD_I_EPISODE][e],
              (uint32_t)s->i[RD_I_AGENT_STEP][e], o->lidar + (size_t)e * nb, cars, n_cars);
  }
  if (o->occupancy) {
    if (zero_occupancy) memset(o->occupancy + (size_t)e * 4096, 0, 4096); /* [REF wrappers.py:410-414] */
    else orc_occupancy_one(m, x, y, yaw, o->occupancy + (size_t)e * 4096);
  }
  const double two_pi = 6.283185307179586;
  double wy = yaw - rint(yaw / two_pi) * two_pi;
  double vx = v * cos(slip), vy = v * sin(slip);
  double pose[6] = {x, y, 0.0, 0.0, 0.0, wy};
  double vel[6] = {vx, vy, 0.0, 0.0, 0.0, s->f[RD_S_YAWRATE][e]};
  if (cfg->obs_flags & RD_OBS_NORM_BASELINES) { /* NormalizeObservations [REF baselines single_agent.py:92-99] */
    const double pl = cfg->obs_low[RD_NORM_POSE], ps = 1.0 / (cfg->obs_high[RD_NORM_POSE] - cfg->obs_low[RD_NORM_POSE]);
    const double vl = cfg->obs_low[RD_NORM_VELOCITY], vs = 1.0 / (cfg->obs_high[RD_NORM_VELOCITY] - cfg->obs_low[RD_NORM_VELOCITY]);
    for (int k = 0; k < 6; ++k) { pose[k] = (pose[k] - pl) * ps; vel[k] = (vel[k] - vl) * vs; }
  }
  if (o->pose) for (int k = 0; k < 6; ++k) o->pose[(size_t)e * 6 + k] = (float)pose[k];
  if (o->velocity) for (int k = 0; k < 6; ++k) o->velocity[(size_t)e * 6 + k] = (float)vel[k];
  if (o->speed) o->speed[e] = (float)sqrt(vx * vx + vy * vy); /* [REF dreamer/wrappers.py:66] */
}

/* env.reset(mode) [REF dreamer/wrappers.py:71-77,91-92,156-158] */
ORC_API void orc_reset(const rd_config* cfg, const orc_map* maps, double* f64, int32_t* i32, double* hist,
                       const uint8_t* mask, int mode, const orc_outputs* out) {
  int n = cfg->n_envs, nb = cfg->n_beams;
  orc_view s = view_of(f64, i32, n);
  s.hist = hist;
  double* ca = (double*)malloc(sizeof(double) * 2 * nb);
  double* sa = ca + nb;
  orc_beam_table(cfg, ca, sa);
  const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
  uint8_t* sel = (uint8_t*)malloc((size_t)n);
  for (int e = 0; e < n; ++e) { /* a world resets as a whole [REF dreamer/tools.py:178-179] */
    int v = (!mask || mask[e]);
    for (int j = e - e % A; j < e - e % A + A && mask && A > 1; ++j) v = v || mask[j];
    sel[e] = (uint8_t)v;
  }
  for (int e = 0; e < n; ++e) if (sel[e]) reset_one(cfg, maps, &s, e, mode);
  for (int e = 0; e < n && out; ++e) { /* observations after every car has its new pose (scans see the other cars) */
    orc_outputs o = *out;
    if (!sel[e]) o.occupancy = NULL; /* untouched for envs that were not reset */
    write_obs(cfg, maps, ca, sa, &s, e, &o, 1);
    if (sel[e]) {
      if (out->reward) out->reward[e] = 0.f;
      if (out->done) out->done[e] = 0;
      if (out->progress) out->progress[e] = (float)s.f[RD_S_PROGRESS][e];
      if (out->lap) out->lap[e] = s.i[RD_I_LAP][e];
      if (out->time) out->time[e] = 0.f;
      if (out->flags) out->flags[e] = 0;
      if (out->rank) out->rank[e] = 1 + e % A;
      if (out->opponents) out->opponents[e] = 0;
    }
  }
  free(sel);
  free(ca);
}

/* One agent step of one env: ReduceActionSpace -> ActionRepeat{tick: dynamics, maps, lap logic, reward,
 * done} -> TimeLimit -> (auto-reset) [REF dreamer/wrappers.py:129-134,107-116,147-154]. */
static void step_one(const rd_config* cfg, const orc_map* maps, orc_view* s, int e, const float* actions,
                     const double* commands, const orc_outputs* o, rd_stats* st, int* was_reset) {
  const orc_map* m = &maps[s->i[RD_I_MAP][e]];
  *was_reset = 0;
  if (s->i[RD_I_FLAGS][e] & RD_F_NEEDS_RESET) { /* frozen until reset [REF wrappers.py:148 'Must reset'] */
    if (o->reward) o->reward[e] = 0.f;
    if (o->done) o->done[e] = 1;
    if (o->progress) o->progress[e] = (float)s->f[RD_S_PROGRESS][e];
    if (o->lap) o->lap[e] = s->i[RD_I_LAP][e];
    if (o->time) o->time[e] = (float)s->f[RD_S_TIME][e];
    if (o->flags) o->flags[e] = (uint8_t)s->i[RD_I_FLAGS][e];
    return;
  }
  /* a4 action transform [REF dreamer/wrappers.py:129-134; baselines single_agent.py:55-56].  The policy hands
   * over a float32 array; numpy keeps `(action + 1) / 2` in float32 (python scalars are weak) and promotes to
   * float64 only at `* (high - low)` because low/high are float64 arrays.  `commands` (float64, sim-facing)
   * bypasses the transform: that is racecar_gym's own step({'motor','steering'}) entry. */
  double a[2];
  if (commands) { a[0] = commands[2 * e]; a[1] = commands[2 * e + 1]; }
  else {
    for (int k = 0; k < 2; ++k) {
      float af = actions[2 * e + k];
      if (cfg->clip_actions) af = af < -1.0f ? -1.0f : (af > 1.0f ? 1.0f : af);
      if (cfg->rescale_actions) {
        float t = (af + 1.0f) / 2.0f;
        a[k] = (double)t * (cfg->action_high[k] - cfg->action_low[k]) + cfg->action_low[k];
      } else {
        a[k] = (double)af;
      }
    }
  }
  double q[7];
  for (int k = 0; k < 7; ++k) q[k] = s->f[k][e];
  double time = s->f[RD_S_TIME][e], p = s->f[RD_S_PROGRESS][e], last = s->f[RD_S_LAST][e];
  int lap = s->i[RD_I_LAP][e], cp = s->i[RD_I_CHECKPOINT][e], flags = s->i[RD_I_FLAGS][e];
  double total = 0.0;
  int done = 0, tick_timeout = 0;
  const int ncp = cfg->n_checkpoints;
  for (int t = 0; t < cfg->action_repeat; ++t) {
    st_tick(cfg, q, a[0], a[1]);
    time = time + cfg->dt;
    int col = collides(cfg, m, q[0], q[1], q[4]);
    int cx, cy;
    int inside = cell_of(m, q[0], q[1], &cx, &cy);
    flags &= ~(RD_F_COLLISION | RD_F_LEFT_MAP);
    if (col) flags |= RD_F_COLLISION;
    if (!inside) flags |= RD_F_LEFT_MAP;
    if (!(q[0] == q[0] && q[1] == q[1] && q[3] == q[3] && q[4] == q[4])) flags |= RD_F_NAN;
    progress_at(m, q[0], q[1], &p);
    int cn = checkpoint_of(cfg, p);
    if (cn == cp + 1) { cp = cn; flags &= ~RD_F_WRONG_WAY; }
    else if (cp == ncp - 1 && cn == 0 && ncp > 1) { lap += 1; cp = 0; flags &= ~RD_F_WRONG_WAY; }
    else if (cn == cp - 1 || (cp == 0 && cn == ncp - 1 && ncp > 1)) { flags |= RD_F_WRONG_WAY; }
    double cur = (double)lap + p;
    double r;
    int d;
    if (cfg->task == RD_TASK_MAX_SPEED) { /* [REF baselines/racing/environment/tasks.py:6-18] */
      r = col ? -1.0 : -exp(fabs(a[1]) - q[3] * cos(q[6]));
      d = 0;
    } else {
      double delta = cur - last;
      if (delta > 0.5) delta = delta - 1.0;
      if (delta < -0.5) delta = delta + 1.0;
      if (cfg->progress_abs) delta = fabs(delta);
      r = cfg->frame_reward + cfg->progress_reward * delta;
      if (col) r = r + cfg->collision_reward;
      d = (cfg->terminate_on_collision && col) || (lap > cfg->laps) || (time > cfg->time_limit);
    }
    /* gym TimeLimit inside ActionRepeat (the model-free chain): the tick that brings the episode's tick count to the
     * limit is done [REF baselines/racing/experiments/acme/experiment.py:66-72; gym 0.18.0 wrappers/time_limit.py] */
    tick_timeout = 0;
    if (cfg->time_limit_ticks > 0 && s->i[RD_I_AGENT_STEP][e] * cfg->action_repeat + t + 1 >= cfg->time_limit_ticks) {
      tick_timeout = !d; d = 1;
    }
    last = cur;
    total = total + r;
    /* dreamer: stop at the first done [REF dreamer/wrappers.py:112]; baselines: the done of tick 0 is not
     * tested when more ticks follow [REF baselines/racing/environment/single_agent.py:32-38] */
    if (d && !(cfg->repeat_semantics == RD_REPEAT_BASELINES && t == 0 && cfg->action_repeat > 1)) { done = 1; break; }
  }
  int agent_step = s->i[RD_I_AGENT_STEP][e] + 1;
  int timeout = done ? tick_timeout : 0;
  if (cfg->time_limit_steps > 0 && agent_step >= cfg->time_limit_steps) { timeout = timeout || !done; done = 1; }
  double ret = s->f[RD_S_RETURN][e] + total;
  /* tools.simulate's per-episode statistic: max over the agent steps of lap + progress - 1 [REF dreamer/tools.py:181,195] */
  double epi = ((double)lap + p) - 1.0;
  double mp = epi > s->f[RD_S_MAXPROG][e] ? epi : s->f[RD_S_MAXPROG][e];
  s->f[RD_S_MAXPROG][e] = mp;
  /* commit */
  for (int k = 0; k < 7; ++k) s->f[k][e] = q[k];
  s->f[RD_S_TIME][e] = time; s->f[RD_S_PROGRESS][e] = p; s->f[RD_S_LAST][e] = last; s->f[RD_S_RETURN][e] = ret;
  s->i[RD_I_LAP][e] = lap; s->i[RD_I_CHECKPOINT][e] = cp; s->i[RD_I_AGENT_STEP][e] = agent_step;
  if (done && !cfg->auto_reset) flags |= RD_F_NEEDS_RESET;
  s->i[RD_I_FLAGS][e] = flags;
  if (o->reward) o->reward[e] = (float)total;
  if (o->reward64) o->reward64[e] = total;
  if (o->done) o->done[e] = (uint8_t)done;
  if (o->progress) o->progress[e] = (float)p;
  if (o->lap) o->lap[e] = lap;
  if (o->time) o->time[e] = (float)time;
  if (o->flags) o->flags[e] = (uint8_t)flags;
  if (st) {
    st->env_steps += 1.0;
    if (done) {
      st->episodes += 1.0;
      st->return_sum += ret;
      st->progress_sum += ((double)lap + p) - s->f[RD_S_START][e];
      st->length_sum += (double)agent_step;
      st->collisions += (flags & RD_F_COLLISION) ? 1.0 : 0.0;
      st->laps_completed += (double)(lap - 1);
      st->timeouts += timeout ? 1.0 : 0.0;
      st->max_progress_sum += mp;
    }
  }
  if (done && cfg->auto_reset) { reset_one(cfg, maps, s, e, cfg->reset_mode); *was_reset = 1; }
}

/* ---- multi-agent worlds --------------------------------------------------------------------- */
/* Body-box overlap of two cars: separating-axis test on the four box axes [NEW-SPEC: racecar_gym reports Bullet
 * contacts between racecars as info['opponent_collisions']]. */
static int rect_overlap(double hl, double hw, double dx, double dy, double c1, double s1, double c2, double s2) {
  const double cr = fabs(c1 * c2 + s1 * s2), sr = fabs(c1 * s2 - s1 * c2);
  const double ex = hl + (hl * cr + hw * sr), ey = hw + (hl * sr + hw * cr);
  if (fabs(dx * c1 + dy * s1) > ex) return 0;
  if (fabs(dy * c1 - dx * s1) > ey) return 0;
  if (fabs(dx * c2 + dy * s2) > ex) return 0;
  if (fabs(dy * c2 - dx * s2) > ey) return 0;
  return 1;
}

/* One agent step of one WORLD of A cars (env indices w*A .. w*A+A-1), the dict-of-agents semantics of the reference's
 * wrappers: ActionRepeat stops at the first tick in which ANY car is done [REF dreamer/wrappers.py:112] (baselines
 * multi-agent variant: every tick runs, dones are OR-ed [REF baselines/racing/environment/multi_agent.py:72-79]);
 * TimeLimit sets every done [REF dreamer/wrappers.py:151-153]; the world is reset when any car is done
 * [REF dreamer/tools.py:178-179]; statistics follow the first agent [REF dreamer/tools.py:162-165].  Tasks per car:
 * cfg->agent_task[a] (A > 1) or cfg->task; n_step_progress rewards progress over the last n ticks [NEW-SPEC]. */
static void step_world(const rd_config* cfg, const orc_map* maps, orc_view* s, int w, const float* actions,
                       const double* commands, const orc_outputs* o, rd_stats* st, int* was_reset) {
  const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
  const int e0 = w * A, n = s->n;
  const orc_map* m = &maps[s->i[RD_I_MAP][e0]];
  *was_reset = 0;
  if (s->i[RD_I_FLAGS][e0] & RD_F_NEEDS_RESET) {
    for (int e = e0; e < e0 + A; ++e) {
      if (o->reward) o->reward[e] = 0.f;
      if (o->done) o->done[e] = 1;
      if (o->progress) o->progress[e] = (float)s->f[RD_S_PROGRESS][e];
      if (o->lap) o->lap[e] = s->i[RD_I_LAP][e];
      if (o->time) o->time[e] = (float)s->f[RD_S_TIME][e];
      if (o->flags) o->flags[e] = (uint8_t)s->i[RD_I_FLAGS][e];
    }
    return;
  }
  double act[RD_MAX_AGENTS][2], q[RD_MAX_AGENTS][7], time[RD_MAX_AGENTS], p[RD_MAX_AGENTS], last[RD_MAX_AGENTS];
  double total[RD_MAX_AGENTS], cs[RD_MAX_AGENTS][2];
  int lap[RD_MAX_AGENTS], cp[RD_MAX_AGENTS], flags[RD_MAX_AGENTS], opp[RD_MAX_AGENTS], done[RD_MAX_AGENTS];
  int tick_timeout[RD_MAX_AGENTS];
  /* baselines semantics: several cars -> multi_agent.py (every tick runs, dones are OR-ed); one car -> single_agent.py
   * (the first tick's done is not tested, a later one stops the repeat), exactly as step_one */
  const int ma_or = cfg->repeat_semantics == RD_REPEAT_BASELINES && A > 1;
  const int sa_skip = cfg->repeat_semantics == RD_REPEAT_BASELINES && A == 1;
  int col[RD_MAX_AGENTS], inside[RD_MAX_AGENTS];
  const int ncp = cfg->n_checkpoints;
  const double hl = 0.5 * cfg->vehicle.body_length, hw = 0.5 * cfg->vehicle.body_width;
  for (int a = 0; a < A; ++a) {
    const int e = e0 + a;
    for (int k = 0; k < 2; ++k) {
      if (commands) { act[a][k] = commands[2 * e + k]; continue; } /* sim-facing float64: racecar_gym's own step() */
      float af = actions[2 * e + k];
      if (cfg->clip_actions) af = af < -1.0f ? -1.0f : (af > 1.0f ? 1.0f : af);
      if (cfg->rescale_actions) {
        float t = (af + 1.0f) / 2.0f;
        act[a][k] = (double)t * (cfg->action_high[k] - cfg->action_low[k]) + cfg->action_low[k];
      } else {
        act[a][k] = (double)af;
      }
    }
    for (int k = 0; k < 7; ++k) q[a][k] = s->f[k][e];
    time[a] = s->f[RD_S_TIME][e]; p[a] = s->f[RD_S_PROGRESS][e]; last[a] = s->f[RD_S_LAST][e];
    lap[a] = s->i[RD_I_LAP][e]; cp[a] = s->i[RD_I_CHECKPOINT][e]; flags[a] = s->i[RD_I_FLAGS][e];
    total[a] = 0.0; opp[a] = 0; done[a] = 0; tick_timeout[a] = 0;
  }
  const int agent_step0 = s->i[RD_I_AGENT_STEP][e0];
  for (int tk = 0; tk < cfg->action_repeat; ++tk) {
    for (int a = 0; a < A; ++a) {
      st_tick(cfg, q[a], act[a][0], act[a][1]);
      time[a] = time[a] + cfg->dt;
      col[a] = collides(cfg, m, q[a][0], q[a][1], q[a][4]);
      int cx, cy;
      inside[a] = cell_of(m, q[a][0], q[a][1], &cx, &cy);
      progress_at(m, q[a][0], q[a][1], &p[a]);
      cs[a][0] = cos(q[a][4]); cs[a][1] = sin(q[a][4]);
    }
    int any = 0;
    for (int a = 0; a < A; ++a) {
      const int e = e0 + a;
      opp[a] = 0;
      for (int j = 0; j < A; ++j)
        if (j != a && rect_overlap(hl, hw, q[j][0] - q[a][0], q[j][1] - q[a][1], cs[a][0], cs[a][1], cs[j][0], cs[j][1]))
          opp[a] |= 1 << j;
      flags[a] &= ~(RD_F_COLLISION | RD_F_LEFT_MAP | RD_F_OPPONENT);
      if (col[a]) flags[a] |= RD_F_COLLISION;
      if (opp[a]) flags[a] |= RD_F_OPPONENT;
      if (!inside[a]) flags[a] |= RD_F_LEFT_MAP;
      if (!(q[a][0] == q[a][0] && q[a][1] == q[a][1] && q[a][3] == q[a][3] && q[a][4] == q[a][4])) flags[a] |= RD_F_NAN;
      const int hitc = col[a] || opp[a] != 0;
      int cn = checkpoint_of(cfg, p[a]);
      if (cn == cp[a] + 1) { cp[a] = cn; flags[a] &= ~RD_F_WRONG_WAY; }
      else if (cp[a] == ncp - 1 && cn == 0 && ncp > 1) { lap[a] += 1; cp[a] = 0; flags[a] &= ~RD_F_WRONG_WAY; }
      else if (cn == cp[a] - 1 || (cp[a] == 0 && cn == ncp - 1 && ncp > 1)) { flags[a] |= RD_F_WRONG_WAY; }
      const double cur = (double)lap[a] + p[a];
      const int task = A > 1 ? cfg->agent_task[a] : cfg->task;
      double r;
      int d;
      if (task == RD_TASK_MAX_SPEED) {
        r = hitc ? -1.0 : -exp(fabs(act[a][1]) - q[a][3] * cos(q[a][6]));
        d = 0;
      } else {
        double ref = last[a];
        if (task == RD_TASK_N_STEP_PROGRESS && s->hist) {
          int slot = (agent_step0 * cfg->action_repeat + tk) % cfg->n_step_progress;
          ref = s->hist[(size_t)slot * n + e];
          s->hist[(size_t)slot * n + e] = cur;
        }
        double delta = cur - ref;
        if (delta > 0.5) delta = delta - 1.0;
        if (delta < -0.5) delta = delta + 1.0;
        if (cfg->progress_abs) delta = fabs(delta);
        r = cfg->frame_reward + cfg->progress_reward * delta;
        if (hitc) r = r + cfg->collision_reward;
        d = (cfg->terminate_on_collision && hitc) || (lap[a] > cfg->laps) || (time[a] > cfg->time_limit);
      }
      int tto = 0;
      if (cfg->time_limit_ticks > 0 && agent_step0 * cfg->action_repeat + tk + 1 >= cfg->time_limit_ticks) { tto = !d; d = 1; }
      last[a] = cur;
      total[a] = total[a] + r;
      if (ma_or) { done[a] |= d; tick_timeout[a] |= tto; }
      else {
        d = d && !(sa_skip && tk == 0 && cfg->action_repeat > 1);
        done[a] = d; tick_timeout[a] = d ? tto : 0;
        any |= d;
      }
    }
    if (any) break;
  }
  int wdone = 0;
  for (int a = 0; a < A; ++a) wdone |= done[a];
  const int agent_step = agent_step0 + 1;
  int timeout = done[0] ? tick_timeout[0] : 0;
  if (cfg->time_limit_steps > 0 && agent_step >= cfg->time_limit_steps) {
    timeout = timeout || !wdone; wdone = 1;
    for (int a = 0; a < A; ++a) done[a] = 1;
  }
  for (int a = 0; a < A; ++a) {
    const int e = e0 + a;
    int rank = 1;
    const double mine = (double)lap[a] + p[a];
    for (int j = 0; j < A; ++j) {
      if (j == a) continue;
      const double other = (double)lap[j] + p[j];
      if (other > mine || (other == mine && j < a)) rank += 1;
    }
    const double ret = s->f[RD_S_RETURN][e] + total[a];
    const double epi = ((double)lap[a] + p[a]) - 1.0;
    const double mp = epi > s->f[RD_S_MAXPROG][e] ? epi : s->f[RD_S_MAXPROG][e];
    s->f[RD_S_MAXPROG][e] = mp;
    for (int k = 0; k < 7; ++k) s->f[k][e] = q[a][k];
    s->f[RD_S_TIME][e] = time[a]; s->f[RD_S_PROGRESS][e] = p[a]; s->f[RD_S_LAST][e] = last[a];
    s->f[RD_S_RETURN][e] = ret;
    s->i[RD_I_LAP][e] = lap[a]; s->i[RD_I_CHECKPOINT][e] = cp[a]; s->i[RD_I_AGENT_STEP][e] = agent_step;
    if (wdone && !cfg->auto_reset) flags[a] |= RD_F_NEEDS_RESET;
    s->i[RD_I_FLAGS][e] = flags[a];
    if (o->reward) o->reward[e] = (float)total[a];
    if (o->reward64) o->reward64[e] = total[a];
    if (o->done) o->done[e] = (uint8_t)done[a];
    if (o->progress) o->progress[e] = (float)p[a];
    if (o->lap) o->lap[e] = lap[a];
    if (o->time) o->time[e] = (float)time[a];
    if (o->flags) o->flags[e] = (uint8_t)flags[a];
    if (o->rank) o->rank[e] = rank;
    if (o->opponents) o->opponents[e] = (uint8_t)opp[a];
    if (st) {
      st->env_steps += 1.0;
      if (wdone && a == 0) {
        st->episodes += 1.0;
        st->return_sum += ret;
        st->progress_sum += ((double)lap[a] + p[a]) - s->f[RD_S_START][e];
        st->length_sum += (double)agent_step;
        st->collisions += (flags[a] & (RD_F_COLLISION | RD_F_OPPONENT)) ? 1.0 : 0.0;
        st->laps_completed += (double)(lap[a] - 1);
        st->timeouts += timeout ? 1.0 : 0.0;
        st->max_progress_sum += mp;
      }
    }
  }
  if (wdone && cfg->auto_reset) {
    for (int e = e0; e < e0 + A; ++e) reset_one(cfg, maps, s, e, cfg->reset_mode);
    *was_reset = 1;
  }
}

static int uses_worlds(const rd_config* cfg) {
  if (cfg->agents_per_world > 1) return 1;
  return cfg->task == RD_TASK_N_STEP_PROGRESS;
}

/* env.step(actions) for the whole batch.  stats may be NULL.  n_threads >= 1.  hist: [n_step_progress][n] ring of the
 * n_step_progress task or NULL. */
ORC_API void orc_step(const rd_config* cfg, const orc_map* maps, double* f64, int32_t* i32, double* hist,
                      const float* actions, const double* commands, const orc_outputs* out, rd_stats* stats,
                      int n_threads) {
  int n = cfg->n_envs, nb = cfg->n_beams;
  orc_view s = view_of(f64, i32, n);
  s.hist = hist;
  double* ca = (double*)malloc(sizeof(double) * 2 * nb);
  double* sa = ca + nb;
  orc_beam_table(cfg, ca, sa);
  if (n_threads < 1) n_threads = 1;
  rd_stats acc;
  memset(&acc, 0, sizeof(acc));
  const int worlds = uses_worlds(cfg);
  const int A = cfg->agents_per_world > 1 ? cfg->agents_per_world : 1;
  uint8_t* mark = (uint8_t*)calloc((size_t)n, 1); /* bit0 frozen, bit1 was reset */
#pragma omp parallel num_threads(n_threads)
  {
    rd_stats loc;
    memset(&loc, 0, sizeof(loc));
    if (worlds) {
      /* phase 1: every world advances; phase 2: observations (scans see the other cars' committed poses) */
#pragma omp for schedule(dynamic, 8)
      for (int w = 0; w < n / A; ++w) {
        int was_reset = 0;
        int frozen = (s.i[RD_I_FLAGS][w * A] & RD_F_NEEDS_RESET) != 0;
        step_world(cfg, maps, &s, w, actions, commands, out, &loc, &was_reset);
        for (int e = w * A; e < w * A + A; ++e) mark[e] = (uint8_t)((frozen ? 1 : 0) | (was_reset ? 2 : 0));
      }
#pragma omp for schedule(dynamic, 8)
      for (int e = 0; e < n; ++e)
        if (!(mark[e] & 1)) write_obs(cfg, maps, ca, sa, &s, e, out, (mark[e] & 2) != 0);
    } else {
#pragma omp for schedule(dynamic, 8)
      for (int e = 0; e < n; ++e) {
        int was_reset = 0;
        int frozen = (s.i[RD_I_FLAGS][e] & RD_F_NEEDS_RESET) != 0;
        step_one(cfg, maps, &s, e, actions, commands, out, &loc, &was_reset);
        if (!frozen) write_obs(cfg, maps, ca, sa, &s, e, out, was_reset);
      }
    }
#pragma omp critical
    {
      acc.episodes += loc.episodes; acc.return_sum += loc.return_sum; acc.progress_sum += loc.progress_sum;
      acc.length_sum += loc.length_sum; acc.collisions += loc.collisions; acc.laps_completed += loc.laps_completed;
      acc.env_steps += loc.env_steps; acc.timeouts += loc.timeouts; acc.max_progress_sum += loc.max_progress_sum;
    }
  }
  if (stats) {
    stats->episodes += acc.episodes; stats->return_sum += acc.return_sum; stats->progress_sum += acc.progress_sum;
    stats->length_sum += acc.length_sum; stats->collisions += acc.collisions;
    stats->laps_completed += acc.laps_completed; stats->env_steps += acc.env_steps; stats->timeouts += acc.timeouts;
    stats->max_progress_sum += acc.max_progress_sum;
  }
  free(mark);
  free(ca);
}

/* a7/a8 stage entry (mirrors rd_reward_done): one sim tick of progress / lap / wrong-way bookkeeping + the task's reward
 * and done for teacher-forced poses.  kin [5][n] = (x, y, yaw, v, slip) AFTER the tick; steering [n] or NULL;
 * book_f64 [3][n] = (time, progress, last), book_i32 [3][n] = (lap, checkpoint, flags), both in/out. */
ORC_API void orc_reward_done(const rd_config* cfg, const orc_map* maps, const double* kin, const double* steering,
                             const int32_t* map_ids, int n, double* book_f64, int32_t* book_i32, double* reward,
                             uint8_t* done) {
  const int ncp = cfg->n_checkpoints;
  for (int e = 0; e < n; ++e) {
    const orc_map* m = &maps[map_ids ? map_ids[e] : 0];
    const double x = kin[e], y = kin[(size_t)n + e], yaw = kin[(size_t)2 * n + e], v = kin[(size_t)3 * n + e];
    const double slip = kin[(size_t)4 * n + e];
    double time = book_f64[e] + cfg->dt, p = book_f64[(size_t)n + e], last = book_f64[(size_t)2 * n + e];
    int lap = book_i32[e], cp = book_i32[(size_t)n + e], flags = book_i32[(size_t)2 * n + e];
    int col = collides(cfg, m, x, y, yaw);
    int cx, cy;
    int inside = cell_of(m, x, y, &cx, &cy);
    flags &= ~(RD_F_COLLISION | RD_F_LEFT_MAP);
    if (col) flags |= RD_F_COLLISION;
    if (!inside) flags |= RD_F_LEFT_MAP;
    if (!(x == x && y == y && v == v && yaw == yaw)) flags |= RD_F_NAN;
    progress_at(m, x, y, &p);
    int cn = checkpoint_of(cfg, p);
    if (cn == cp + 1) { cp = cn; flags &= ~RD_F_WRONG_WAY; }
    else if (cp == ncp - 1 && cn == 0 && ncp > 1) { lap += 1; cp = 0; flags &= ~RD_F_WRONG_WAY; }
    else if (cn == cp - 1 || (cp == 0 && cn == ncp - 1 && ncp > 1)) { flags |= RD_F_WRONG_WAY; }
    const double cur = (double)lap + p;
    double r;
    int d;
    if (cfg->task == RD_TASK_MAX_SPEED) { /* [REF baselines/racing/environment/tasks.py:6-18] */
      r = col ? -1.0 : -exp(fabs(steering ? steering[e] : 0.0) - v * cos(slip));
      d = 0;
    } else {
      double delta = cur - last;
      if (delta > 0.5) delta = delta - 1.0;
      if (delta < -0.5) delta = delta + 1.0;
      if (cfg->progress_abs) delta = fabs(delta);
      r = cfg->frame_reward + cfg->progress_reward * delta;
      if (col) r = r + cfg->collision_reward;
      d = (cfg->terminate_on_collision && col) || (lap > cfg->laps) || (time > cfg->time_limit);
    }
    book_f64[e] = time; book_f64[(size_t)n + e] = p; book_f64[(size_t)2 * n + e] = cur;
    book_i32[e] = lap; book_i32[(size_t)n + e] = cp; book_i32[(size_t)2 * n + e] = flags;
    reward[e] = r;
    done[e] = (uint8_t)(d ? 1 : 0);
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* a5 OccupancyMapObs.step [REF dreamer/wrappers.py:390-408]:                                         */
/*   (pr,pc) = to_pixel(pose); M[pr-110:pr+110, pc-110:pc+110] as uint8;                              */
/*   scipy.ndimage.rotate(., rad2deg(2*pi - yaw)) (order 3, reshape, mode constant, prefilter);       */
/*   centre crop 200x200; PIL.Image.fromarray(.).resize((64,64)) (bicubic); -> (64,64,1) uint8.       */
/* The two library stages are restated from their published algorithms (scipy ni_splines.c /          */
/* ni_interpolation.c, Pillow Resample.c) and pinned against the libraries in tests/golden.           */
/* ------------------------------------------------------------------------------------------------ */
#define OCC_NEIGH 100
#define OCC_IN (2 * (OCC_NEIGH + 10)) /* 220 */
#define OCC_MID (2 * OCC_NEIGH)       /* 200 */
#define OCC_OUT 64

/* cubic B-spline prefilter of one line, mirror boundary (scipy uses it for mode='constant') */
static void spline_prefilter_line(double* c, int n) {
  const double z = sqrt(3.0) - 2.0;
  const double gain = (1.0 - z) * (1.0 - 1.0 / z);
  for (int i = 0; i < n; ++i) c[i] *= gain;
  /* causal initialisation: exact mirror sum */
  double z_i = z;
  const double z_n_1 = pow(z, (double)(n - 1));
  c[0] = c[0] + z_n_1 * c[n - 1];
  for (int i = 1; i < n - 1; ++i) {
    c[0] += z_i * (c[i] + z_n_1 * c[n - 1 - i]);
    z_i *= z;
  }
  c[0] /= 1.0 - z_n_1 * z_n_1;
  for (int i = 1; i < n; ++i) c[i] += z * c[i - 1];
  c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0);
  for (int i = n - 2; i >= 0; --i) c[i] = z * (c[i + 1] - c[i]);
}

static inline int mirror_index(int idx, int len) {
  int s2 = 2 * len - 2;
  if (idx < 0) {
    idx = s2 * (int)(-idx / s2) + idx;
    idx = idx <= 1 - len ? idx + s2 : -idx;
  } else if (idx >= len) {
    idx -= s2 * (int)(idx / s2);
    if (idx >= len) idx = s2 - idx;
  }
  return idx;
}

static inline double bicubic_kernel(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* Pillow precompute_coeffs + normalize_coeffs_8bpc for in_size -> out_size, bicubic */
#define PIL_PRECISION_BITS (32 - 8 - 2)
static int pil_coeffs(int in_size, int out_size, int** bounds_out, int32_t** kk_out) {
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  double support = 2.0 * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  int* bounds = (int*)malloc(sizeof(int) * 2 * out_size);
  int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * out_size * ksize);
  double* k = (double*)malloc(sizeof(double) * ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale, ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) { double w = bicubic_kernel((x + xmin - center + 0.5) * ss); k[x] = w; ww += w; }
    for (x = 0; x < xmax; ++x) if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    for (x = 0; x < ksize; ++x)
      kk[xx * ksize + x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << PIL_PRECISION_BITS))
                                    : (int32_t)(0.5 + k[x] * (1 << PIL_PRECISION_BITS));
    bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
  }
  free(k);
  *bounds_out = bounds; *kk_out = kk;
  return ksize;
}
static inline uint8_t clip8(int32_t v) { v >>= PIL_PRECISION_BITS; return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static void orc_occupancy_one(const orc_map* m, double x, double y, double yaw, uint8_t* out) {
  /* to_pixel: full-image (row, col); the crop below is expressed on y-up cells */
  double inv_res = 1.0 / m->resolution;
  int col = (int)floor((x - m->origin_x) * inv_res);
  int rup = (int)floor((y - m->origin_y) * inv_res);
  int pr = m->full_h - 1 - rup, pc = col;
  static const int N = OCC_IN;
  double* coef = (double*)malloc(sizeof(double) * N * N);
  double* line = (double*)malloc(sizeof(double) * N);
  /* crop rows pr-110 .. pr+109 (image order), cols pc-110 .. pc+109 */
  for (int i = 0; i < N; ++i) {
    int r_img = pr - N / 2 + i;
    int cy = (m->full_h - 1 - r_img) - m->row0;
    for (int j = 0; j < N; ++j) coef[i * N + j] = (double)drivable_at(m, pc - N / 2 + j - m->col0, cy);
  }
  /* spline_filter: axis 0 then axis 1 */
  for (int j = 0; j < N; ++j) {
    for (int i = 0; i < N; ++i) line[i] = coef[i * N + j];
    spline_prefilter_line(line, N);
    for (int i = 0; i < N; ++i) coef[i * N + j] = line[i];
  }
  for (int i = 0; i < N; ++i) spline_prefilter_line(coef + i * N, N);
  /* rotation matrix: angle = rad2deg(2*pi - yaw); c,s = cos,sin of that angle */
  double ang = 2.0 * 3.141592653589793 - yaw;
  double c = cos(ang), s = sin(ang);
  /* output plane shape = int(ptp(rot @ corners) + 0.5) */
  double b0[4] = {0.0, s * N, c * N, c * N + s * N};        /* row 0 of rot @ [[0,0,iy,iy],[0,ix,0,ix]] */
  double b1[4] = {0.0, c * N, -s * N, -s * N + c * N};
  double mn0 = b0[0], mx0 = b0[0], mn1 = b1[0], mx1 = b1[0];
  for (int k = 1; k < 4; ++k) {
    if (b0[k] < mn0) mn0 = b0[k]; if (b0[k] > mx0) mx0 = b0[k];
    if (b1[k] < mn1) mn1 = b1[k]; if (b1[k] > mx1) mx1 = b1[k];
  }
  int oh = (int)((mx0 - mn0) + 0.5), ow = (int)((mx1 - mn1) + 0.5);
  double oc0 = (oh - 1) / 2.0, oc1 = (ow - 1) / 2.0, ic = (N - 1) / 2.0;
  double off0 = ic - (c * oc0 + s * oc1);
  double off1 = ic - (-s * oc0 + c * oc1);
  int cr = oh / 2, cc = ow / 2;
  uint8_t* mid = (uint8_t*)malloc(OCC_MID * OCC_MID);
  for (int a = 0; a < OCC_MID; ++a) {
    int o0 = cr - OCC_NEIGH + a;
    for (int b = 0; b < OCC_MID; ++b) {
      int o1 = cc - OCC_NEIGH + b;
      double t = 0.0;
      if (o0 >= 0 && o0 < oh && o1 >= 0 && o1 < ow) {
        /* scipy accumulates shift first, then one term per output axis */
        double c0 = off0; c0 += (double)o0 * c;  c0 += (double)o1 * s;
        double c1 = off1; c1 += (double)o0 * -s; c1 += (double)o1 * c;
        if (!(c0 < 0.0 || c0 > N - 1 || c1 < 0.0 || c1 > N - 1)) {
          int s0 = (int)floor(c0) - 1, s1 = (int)floor(c1) - 1;
          double w0[4], w1[4];
          {
            double yv = c0 - floor(c0), zv = 1.0 - yv;
            w0[1] = (yv * yv * (yv - 2.0) * 3.0 + 4.0) / 6.0;
            w0[2] = (zv * zv * (zv - 2.0) * 3.0 + 4.0) / 6.0;
            w0[0] = zv * zv * zv / 6.0;
            w0[3] = 1.0 - w0[0] - w0[1] - w0[2];
          }
          {
            double yv = c1 - floor(c1), zv = 1.0 - yv;
            w1[1] = (yv * yv * (yv - 2.0) * 3.0 + 4.0) / 6.0;
            w1[2] = (zv * zv * (zv - 2.0) * 3.0 + 4.0) / 6.0;
            w1[0] = zv * zv * zv / 6.0;
            w1[3] = 1.0 - w1[0] - w1[1] - w1[2];
          }
          for (int p = 0; p < 4; ++p) {
            int ii = mirror_index(s0 + p, N);
            for (int q = 0; q < 4; ++q) {
              int jj = mirror_index(s1 + q, N);
              double cf = coef[ii * N + jj];
              cf *= w0[p];
              cf *= w1[q];
              t += cf;
            }
          }
        }
      } else {
        /* outside the rotated image: numpy slicing would have shortened the crop; the reference never
         * gets here because out >= 220 > 200 */
      }
      double tv = t > 0 ? t + 0.5 : 0.0;
      tv = tv > 255.0 ? 255.0 : tv;
      mid[a * OCC_MID + b] = (uint8_t)tv;
    }
  }
  /* Pillow resize 200x200 -> 64x64 bicubic: horizontal pass to a uint8 image, then vertical */
  int *bounds; int32_t* kk;
  int ksize = pil_coeffs(OCC_MID, OCC_OUT, &bounds, &kk);
  uint8_t* tmp = (uint8_t*)malloc(OCC_MID * OCC_OUT);
  for (int yy = 0; yy < OCC_MID; ++yy)
    for (int xx = 0; xx < OCC_OUT; ++xx) {
      int32_t ss0 = 1 << (PIL_PRECISION_BITS - 1);
      int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
      for (int k = 0; k < xmax; ++k) ss0 += (int32_t)mid[yy * OCC_MID + xmin + k] * kk[xx * ksize + k];
      tmp[yy * OCC_OUT + xx] = clip8(ss0);
    }
  for (int yy = 0; yy < OCC_OUT; ++yy) {
    int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    for (int xx = 0; xx < OCC_OUT; ++xx) {
      int32_t ss0 = 1 << (PIL_PRECISION_BITS - 1);
      for (int k = 0; k < ymax; ++k) ss0 += (int32_t)tmp[(ymin + k) * OCC_OUT + xx] * kk[yy * ksize + k];
      out[yy * OCC_OUT + xx] = clip8(ss0);
    }
  }
  free(tmp); free(bounds); free(kk); free(mid); free(line); free(coef);
}

/* poses [n][3]; out [n][64*64] */
ORC_API void orc_occupancy_obs(const orc_map* maps, const double* poses, const int32_t* map_ids, int n,
                               uint8_t* out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int e = 0; e < n; ++e)
    orc_occupancy_one(&maps[map_ids ? map_ids[e] : 0], poses[3 * e], poses[3 * e + 1], poses[3 * e + 2],
                      out + (size_t)e * 4096);
}

/* stage pieces exposed for pinning tests against scipy / Pillow */
ORC_API void orc_spline_prefilter_2d(double* a, int n) {
  double* line = (double*)malloc(sizeof(double) * n);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) line[i] = a[i * n + j];
    spline_prefilter_line(line, n);
    for (int i = 0; i < n; ++i) a[i * n + j] = line[i];
  }
  for (int i = 0; i < n; ++i) spline_prefilter_line(a + i * n, n);
  free(line);
}
ORC_API void orc_pil_resize_200_to_64(const uint8_t* in, uint8_t* out) {
  int *bounds; int32_t* kk;
  int ksize = pil_coeffs(OCC_MID, OCC_OUT, &bounds, &kk);
  uint8_t* tmp = (uint8_t*)malloc(OCC_MID * OCC_OUT);
  for (int yy = 0; yy < OCC_MID; ++yy)
    for (int xx = 0; xx < OCC_OUT; ++xx) {
      int32_t ss0 = 1 << (PIL_PRECISION_BITS - 1);
      for (int k = 0; k < bounds[2 * xx + 1]; ++k) ss0 += (int32_t)in[yy * OCC_MID + bounds[2 * xx] + k] * kk[xx * ksize + k];
      tmp[yy * OCC_OUT + xx] = clip8(ss0);
    }
  for (int yy = 0; yy < OCC_OUT; ++yy)
    for (int xx = 0; xx < OCC_OUT; ++xx) {
      int32_t ss0 = 1 << (PIL_PRECISION_BITS - 1);
      for (int k = 0; k < bounds[2 * yy + 1]; ++k) ss0 += (int32_t)tmp[(bounds[2 * yy] + k) * OCC_OUT + xx] * kk[yy * ksize + k];
      out[yy * OCC_OUT + xx] = clip8(ss0);
    }
  free(tmp); free(bounds); free(kk);
}
ORC_API int orc_sizeof_config(void) { return (int)sizeof(rd_config); }
ORC_API int orc_sizeof_map(void) { return (int)sizeof(orc_map); }
