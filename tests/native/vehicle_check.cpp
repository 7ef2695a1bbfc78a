// vehicle_check.cpp -- host-compiled check of the kernel's vehicle model (csrc/rd_vehicle.cuh) against the CPU oracle and
// the NumPy rendition (driven from tests/test_cpu_vehicle_model.py).  TEST INFRASTRUCTURE: the same functions k_step
// calls, compiled by g++ (the device-only pieces -- the hardware reciprocal seed -- fall back to 1.0 / x here); it is
// not a CPU path of the product.
#include <cstdint>
#include <cstdio>

#include "../../racing_dreamer_b200/csrc/rd_vehicle.cuh"

extern "C" {

// state [7][n] SoA in/out, commands [n][2] sim-facing, n_ticks ticks of cfg->dt
__attribute__((visibility("default")))
void vehicle_ticks(const rd_config* cfg, double* state, const double* commands, int n, int n_ticks) {
  const VehConst k = rdv_make_const(cfg->vehicle, cfg->dt);
  const VehAcc off = rdv_acc_set(k, 0.0);
  for (int e = 0; e < n; ++e) {
    double q[7];
    for (int i = 0; i < 7; ++i) q[i] = state[(size_t)i * n + e];
    for (int t = 0; t < n_ticks; ++t) rdv_tick(k, off, &k, &off, q, commands[2 * e], commands[2 * e + 1]);
    for (int i = 0; i < 7; ++i) state[(size_t)i * n + e] = q[i];
  }
}

// out [n][2] = sin, cos by the kernel's reduction + polynomials
__attribute__((visibility("default")))
void vehicle_sincos(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) rdv_sincos(x[i], out[2 * i], out[2 * i + 1]);
}

// number of d in [0, dmax] for which the Markstein quotient differs from the correctly rounded d / dmax
__attribute__((visibility("default")))
long long vehicle_div_mismatches(int dmax) {
  const double b = (double)dmax, y = 1.0 / b;
  long long bad = 0;
  for (int d = 0; d <= dmax; ++d)
    if (rdv_div_by((double)d, b, y) != (double)d / b) ++bad;
  return bad;
}

__attribute__((visibility("default")))
int vehicle_sizeof_config(void) { return (int)sizeof(rd_config); }

}  // extern "C"
