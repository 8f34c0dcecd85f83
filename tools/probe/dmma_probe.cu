// Micro-benchmark: fp64 FMA throughput of the vector pipe (DFMA) against the tensor pipe (mma.sync f64) on sm_100a.
// Decides whether the q >= 20 level kernels may put their 20x20 products on DMMA.  Build: see tools/probe/run_dmma.sh
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) { c0[k] = threadIdx.x + k; c1[k] = k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA with distinct operand registers per instruction (the level kernels' situation: 15 fragments x m-tiles), not one
// shared (a, b) pair: does operand collection cost pipe cycles?
template <int CH>
__global__ void dmma884_distinct_kernel(double* out, const double* in, int iters) {
  double c0[CH], c1[CH], a[CH], b[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) { c0[k] = threadIdx.x + k; c1[k] = k; a[k] = in[threadIdx.x + 32 * k]; b[k] = in[threadIdx.x + 32 * k + 7]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a[k]), "d"(b[(k + 3) % CH]));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent-issue latencies: one chain per warp, one warp per SMSP
__global__ void lat_kernel(double* out, int iters, double a, double b, long long* clk) {
  double x = threadIdx.x, c0 = threadIdx.x, c1 = 1.0, r = threadIdx.x + 1.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r));
  long long t3 = clock64();
  double s = x;
  for (int it = 0; it < iters; ++it) { s += __shfl_xor_sync(0xffffffffu, s, 1); }
  long long t4 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + c0 + c1 + r + s;
}
// DMMA and DFMA interleaved 1 : R in one instruction stream: do the two share a pipe?
template <int R>
__global__ void mixed_kernel(double* out, int iters, double a, double b) {
  double c0[4], c1[4], x[4 * R];
#pragma unroll
  for (int k = 0; k < 4; ++k) { c0[k] = threadIdx.x + k; c1[k] = k; }
#pragma unroll
  for (int k = 0; k < 4 * R; ++k) x[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k * R + r]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c0[k] + c1[k];
#pragma unroll
  for (int k = 0; k < 4 * R; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#ifdef BIG
template <int CH>
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[CH][4];
#pragma unroll
  for (int k = 0; k < CH; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                   : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif
template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  double* out; cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
  const int iters = 20000;
  {
    long long* clk; cudaMallocManaged(&clk, 64);
    lat_kernel<<<1, 32>>>(out, 4096, 1.0000001, 1e-9, clk); cudaDeviceSynchronize();
    printf("dependent latency (clk): DFMA %.1f  DMMA.8x8x4 %.1f  MUFU.RCP64H(+mov) %.1f  SHFL.64+DADD %.1f\n", clk[0] / 4096.0, clk[1] / 4096.0, clk[2] / 4096.0, clk[3] / 4096.0);
  }
  for (int wps : {4, 8, 16, 32}) {   // warps per SM
    const int threads = 128, blocks = sms * wps / 4;
    float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fma = (double)blocks * threads * 16.0 * iters;
    printf("DFMA      warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", wps, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
    ms = time_it([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fma = (double)blocks * (threads / 32) * 8.0 * iters * 256.0;
    printf("DMMA 884  warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", wps, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
    ms = time_it([&] { dmma884_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fma = (double)blocks * (threads / 32) * 2.0 * iters * 256.0;
    printf("DMMA 884 (2 chains) warps/SM %2d: %.3f ms  %.2f TFLOP/s  -> latency %.1f clk per dependent mma at 1 warp/SMSP\n", wps, ms, 2 * fma / ms * 1e-9, ms * 1e-3 * 1.965e9 / (2.0 * iters) );
    if (wps == 16) {
      ms = time_it([&] { dmma884_distinct_kernel<8><<<blocks, threads>>>(out, out + 4096, iters); });
      fma = (double)blocks * (threads / 32) * 8.0 * iters * 256.0;
      printf("DMMA 884 distinct operand registers, %d warps/SM: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", wps, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
      ms = time_it([&] { mixed_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double per = ms * 1e-3 * 1.965e9 / (4.0 * iters) / (wps / 4);   // clocks per (1 DMMA + 2 DFMA) group per SMSP
      printf("MIXED 1 DMMA : 2 DFMA, %d warps/SM: %.3f ms -> %.1f clk per group per SMSP (DMMA alone 16, 2 DFMA alone ~4.4: shared pipe ~20.4)\n", wps, ms, per);
      ms = time_it([&] { mixed_kernel<6><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      per = ms * 1e-3 * 1.965e9 / (4.0 * iters) / (wps / 4);
      printf("MIXED 1 DMMA : 6 DFMA, %d warps/SM: %.3f ms -> %.1f clk per group per SMSP (shared pipe ~29.2, separate pipes 16)\n", wps, ms, per);
    }
#ifdef BIG
    ms = time_it([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fma = (double)blocks * (threads / 32) * 4.0 * iters * 2048.0;
    printf("DMMA 16816 warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", wps, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
#endif
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
