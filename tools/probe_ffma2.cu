// probe_ffma2.cu -- issue rate and dependent latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/probe_ffma2 tools/probe_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float2 v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) v[i] = __ffma2_rn(v[i], A, B);
      else { v[i].x = __fmaf_rn(v[i].x, a, b); v[i].y = __fmaf_rn(v[i].y, a, b); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, bool PACKED>
void run(const char* name, int threads) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, PACKED><<<148, threads>>>(out, iters, 0.999f, 0.001f);
  cudaEventRecord(e0);
  k<ILP, PACKED><<<148, threads>>>(out, iters, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)148 * threads * iters * ILP * 2;
  std::printf("%-28s threads/SM %4d ILP %d: %.3f ms, %.1f Gfma/s, %.2f ns per iteration\n", name, threads, ILP, ms, fma / ms * 1e-6, ms * 1e6 / iters);
  cudaFree(out);
}

int main() {
  run<1, false>("FFMA  dependent chain", 32);
  run<1, true>("FFMA2 dependent chain", 32);
  run<8, false>("FFMA  8 chains", 1024);
  run<8, true>("FFMA2 8 chains", 1024);
  run<8, false>("FFMA  8 chains", 512);
  run<8, true>("FFMA2 8 chains", 512);
  run<2, false>("FFMA  2 chains", 512);
  run<2, true>("FFMA2 2 chains", 512);
  return 0;
}
