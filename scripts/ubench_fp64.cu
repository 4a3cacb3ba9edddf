// FP64 pipe micro-benchmarks on one warp (issue interval / dependent latency of DFMA, division, SHFL).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_fp64.bin scripts/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(int iters, double* out, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x + i;
  const double m = 0.999999, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 42.0) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
__global__ void dop_tput(int iters, double* out) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x + i;
  const double m = 0.999999, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], m, c);
      if (OP == 1) a[i] = a[i] * m;
      if (OP == 2) a[i] = a[i] + c;
      if (OP == 3) { a[i] = a[i] * m; a[i] = a[i] + c; }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 42.0) out[0] = s;
}

__global__ void ddiv_chain(int iters, double* out, long long* cyc) {
  double a = 1.0 + threadIdx.x, b = 3.0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { a = b / (a + 1.5); }
  long long t1 = clock64();
  if (a == 42.0) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ddiv_ilp4(int iters, double* out, long long* cyc) {
  double a0 = 1.0 + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b = 3.0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { a0 = b / (a0 + 1.5); a1 = b / (a1 + 1.5); a2 = b / (a2 + 1.5); a3 = b / (a3 + 1.5); }
  long long t1 = clock64();
  if (a0 + a1 + a2 + a3 == 42.0) out[0] = a0;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void shfl_chain(int iters, double* out, long long* cyc) {
  double a = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0; }
  long long t1 = clock64();
  if (a == 42.0) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput with many warps: blocks x warps DFMA ILP8
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

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 100000;
#define RUN(K, name, per)                                             \
  K<<<1, 32>>>(iters, out, cyc); cudaDeviceSynchronize();             \
  K<<<1, 32>>>(iters, out, cyc); cudaDeviceSynchronize();             \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                     \
  printf("%-28s %.2f cycles per op (%d ops/iter)\n", name, (double)h / iters / per, per);
  RUN(dfma_chain<1>, "DFMA dependent chain", 1)
  RUN(dfma_chain<2>, "DFMA ILP2", 2)
  RUN(dfma_chain<4>, "DFMA ILP4", 4)
  RUN(dfma_chain<8>, "DFMA ILP8", 8)
  RUN(dfma_chain<16>, "DFMA ILP16", 16)
  RUN(ddiv_chain, "DDIV(+DADD) dependent", 1)
  RUN(ddiv_ilp4, "DDIV(+DADD) ILP4", 4)
  RUN(shfl_chain, "SHFL.64+DADD dependent", 1)
  for (int wps = 1; wps <= 16; wps *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_tput<<<148, 32 * wps>>>(1000, out);
    cudaEventRecord(e0);
    dfma_tput<<<148, 32 * wps>>>(200000, out);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * 32 * wps * 8 * 200000 * 2;
    printf("DFMA tput %2d warps/SM: %.2f TFLOP/s\n", wps, fl / ms / 1e9);
  }
  const char* names[4] = {"DFMA", "DMUL", "DADD", "DMUL+DADD (2 ops)"};
  for (int op = 0; op < 4; ++op) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    auto launch = [&](int it) {
      if (op == 0) dop_tput<0><<<148, 512>>>(it, out);
      if (op == 1) dop_tput<1><<<148, 512>>>(it, out);
      if (op == 2) dop_tput<2><<<148, 512>>>(it, out);
      if (op == 3) dop_tput<3><<<148, 512>>>(it, out);
    };
    launch(1000);
    cudaEventRecord(e0); launch(100000); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 512 * 8 * 100000 * (op == 3 ? 2 : 1);
    printf("%-20s 16 warps/SM: %.2f T ops/s  (%.2f cycles per warp-instruction per SMSP)\n", names[op], ops / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (100000.0 * 8 * (op == 3 ? 2 : 1) * 4));
  }
  return 0;
}
