// probe_fp64.cu -- dependent-issue latency of the float64 operations k_step's chains are made of (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/probe_fp64 tools/probe_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double v = threadIdx.x * 1e-3 + 1.0, w = 0.5;
  int q = 3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) v = fma(v, a, b);
    else if (OP == 1) v = v + a;
    else if (OP == 2) v = v * a;
    else if (OP == 3) v = (v > b) ? v * a : v + a;             // compare + select + op
    else if (OP == 4) { q = __double2int_rn(v); v = (double)q * a; }   // F2I + I2F + mul
    else if (OP == 5) v = floor(v * a);
    else if (OP == 6) { float f = (float)v; f = fmaf(f, 0.999f, 0.001f); v = (double)f; }
    else if (OP == 7) { v = fma(v, a, b); w = fma(w, a, b); }   // two independent chains
    else if (OP == 8) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v)); v = y + a; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = v + w + q;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name) {
  double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP><<<1, 32>>>(out, cyc, iters, 0.999, 0.001);
  k<OP><<<1, 32>>>(out, cyc, iters, 0.999, 0.001);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  std::printf("%-44s %.1f cycles per iteration\n", name, (double)c / iters);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("DFMA dependent");
  run<1>("DADD dependent");
  run<2>("DMUL dependent");
  run<3>("DSETP + select + DMUL/DADD");
  run<4>("F2I.F64 + I2F.F64 + DMUL");
  run<5>("DMUL + floor");
  run<6>("F2F 64->32, FFMA, F2F 32->64");
  run<7>("two independent DFMA chains");
  run<8>("rcp.approx.ftz.f64 + DADD");
  return 0;
}
