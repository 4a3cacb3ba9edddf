// FP64 peak micro-benchmarks: DFMA (FP64 pipe) and DMMA (mma.sync.*.f64, the FP64 tensor path tcgen05 does not
// have).  Prints one JSON object; scripts/gpu_r2_*.sh stores it as profiles/r02_fp64_peaks.json, which bench.py
// reads as the FP64 roofline denominator (max of this file and its own live probe).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_dmma.bin scripts/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <string>

__global__ void dfma_tput(int iters, double* out) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x + i;
  const double m = 0.999999, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 42.0) out[0] = s;
}

// m8n8k4: A 8x4 (1 double / lane), B 4x8 (1 / lane), C 8x8 (2 / lane): 256 FMA per warp instruction
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// m16n8k8 (sm_90+): A 16x8 (4 / lane), B 8x8 (2 / lane), C 16x8 (4 / lane): 1024 FMA per warp instruction
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// m16n8k16 (sm_90+): A 16x16 (8 / lane), B 16x8 (4 / lane): 2048 FMA per warp instruction
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
               "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                 "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884_tput(int iters, double* out, long long* cyc) {
  double d0[ILP], d1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d0[i] = threadIdx.x * 1e-3 + i; d1[i] = 1.0 - i; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 0.25;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(d0[i], d1[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d0[i] + d1[i];
  if (s == 42.0) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0 && cyc) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void dmma1688_tput(int iters, double* out, long long* cyc) {
  double d[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-3 + i + j;
  const double a[4] = {1.0 + 1e-9 * threadIdx.x, 0.5, 0.25, 0.125}, b[2] = {0.25, 0.5};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(d[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 42.0) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0 && cyc) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void dmma16816_tput(int iters, double* out, long long* cyc) {
  double d[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-3 + i + j;
  const double a[8] = {1.0 + 1e-9 * threadIdx.x, 0.5, 0.25, 0.125, 0.3, 0.2, 0.1, 0.05}, b[4] = {0.25, 0.5, 0.1, 0.2};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(d[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 42.0) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0 && cyc) cyc[0] = t1 - t0;
}

template <class F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  const int iters = 20000;
  std::string js = "{";
  char buf[512];
  snprintf(buf, sizeof buf, "\"sms\": %d, \"sm_clock_mhz_max\": %.0f, ", sms, clk_khz / 1e3); js += buf;
  double best_dfma = 0, best_dmma = 0;
  js += "\"dfma_tflops\": {";
  for (int w : {4, 8, 16, 32}) {
    const int threads = 32 * w > 1024 ? 1024 : 32 * w, blocks = sms * (32 * w / threads);
    double ms = time_ms([&] { dfma_tput<<<blocks, threads>>>(iters, out); });
    double tf = (double)blocks * threads * 8.0 * iters * 2.0 / (ms * 1e-3) / 1e12;
    if (tf > best_dfma) best_dfma = tf;
    snprintf(buf, sizeof buf, "%s\"%d_warps_per_sm\": %.2f", w == 4 ? "" : ", ", w, tf); js += buf;
  }
  js += "}, ";
  // single-warp latency / issue interval of the three DMMA shapes
  long long h;
  js += "\"dmma_single_warp_cycles_per_instr\": {";
  dmma884_tput<1><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, "\"m8n8k4_dependent\": %.2f", (double)h / iters); js += buf;
  dmma884_tput<8><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, ", \"m8n8k4_ilp8\": %.2f", (double)h / iters / 8); js += buf;
  dmma1688_tput<1><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, ", \"m16n8k8_dependent\": %.2f", (double)h / iters); js += buf;
  dmma1688_tput<4><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, ", \"m16n8k8_ilp4\": %.2f", (double)h / iters / 4); js += buf;
  dmma16816_tput<1><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, ", \"m16n8k16_dependent\": %.2f", (double)h / iters); js += buf;
  dmma16816_tput<4><<<1, 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  snprintf(buf, sizeof buf, ", \"m16n8k16_ilp4\": %.2f", (double)h / iters / 4); js += buf;
  js += "}, ";
  auto sweep = [&](const char* name, double fma_per_instr, int ilp, auto launch) {
    js += std::string("\"") + name + "\": {";
    bool first = true;
    for (int w : {4, 8, 16, 32}) {
      const int threads = 32 * w > 1024 ? 1024 : 32 * w, blocks = sms * (32 * w / threads);
      double ms = time_ms([&] { launch(blocks, threads); });
      double tf = (double)blocks * (threads / 32) * ilp * fma_per_instr * 2.0 * iters / (ms * 1e-3) / 1e12;
      if (tf > best_dmma) best_dmma = tf;
      snprintf(buf, sizeof buf, "%s\"%d_warps_per_sm\": %.2f", first ? "" : ", ", w, tf); js += buf;
      first = false;
    }
    js += "}, ";
  };
  sweep("dmma_m8n8k4_tflops", 256, 8, [&](int b, int t) { dmma884_tput<8><<<b, t>>>(iters, out, nullptr); });
  sweep("dmma_m16n8k8_tflops", 1024, 4, [&](int b, int t) { dmma1688_tput<4><<<b, t>>>(iters, out, nullptr); });
  sweep("dmma_m16n8k16_tflops", 2048, 4, [&](int b, int t) { dmma16816_tput<4><<<b, t>>>(iters, out, nullptr); });
  cudaError_t e = cudaDeviceSynchronize();
  snprintf(buf, sizeof buf, "\"fp64_dfma_peak_tflops\": %.2f, \"fp64_dmma_peak_tflops\": %.2f, \"cuda_status\": \"%s\"}",
           best_dfma, best_dmma, cudaGetErrorString(e));
  js += buf;
  printf("%s\n", js.c_str());
  return 0;
}
