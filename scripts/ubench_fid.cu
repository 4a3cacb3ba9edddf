// Timing of qt_fidelity_batch on random full-rank 4-qubit pairs (dev tool for kernel variants selected with -D flags).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DFID_TRI_...] -o scripts/ubench_fid.bin \
//        scripts/ubench_fid.cu forest_benchmarking_b200/csrc/qt_distance.cu forest_benchmarking_b200/csrc/qt_api.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/qtomo.h"

__global__ void make_states(long long B, int D, double2* out, unsigned seed) {
  // rho = G G^dagger / tr, G a D x D matrix of LCG noise; one thread per state (setup only)
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  unsigned s = seed + (unsigned)b * 2654435761u;
  double2 G[32 * 32];  // up to d = 32 (local memory: setup only)
  for (int e = 0; e < D * D; ++e) {
    s = s * 1664525u + 1013904223u; double x = (double)(s >> 8) / (1 << 24) - 0.5;
    s = s * 1664525u + 1013904223u; double y = (double)(s >> 8) / (1 << 24) - 0.5;
    G[e] = make_double2(x, y);
  }
  double tr = 0;
  for (int i = 0; i < D; ++i)
    for (int k = 0; k < D; ++k) tr += G[i * D + k].x * G[i * D + k].x + G[i * D + k].y * G[i * D + k].y;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double re = 0, im = 0;
      for (int k = 0; k < D; ++k) {
        double2 a = G[i * D + k], c = G[j * D + k];
        re += a.x * c.x + a.y * c.y;
        im += a.y * c.x - a.x * c.y;
      }
      out[(b * D + i) * D + j] = make_double2(re / tr, im / tr);
    }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4, D = 1 << n;
  const long long B = argc > 2 ? atoll(argv[2]) : 262144;
  double2 *rho, *sig; double* out;
  cudaMalloc(&rho, B * D * D * 16); cudaMalloc(&sig, B * D * D * 16); cudaMalloc(&out, B * 8);
  make_states<<<(unsigned)((B + 63) / 64), 64>>>(B, D, rho, 1u);
  make_states<<<(unsigned)((B + 63) / 64), 64>>>(B, D, sig, 77u);
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) qt_fidelity_batch(n, B, rho, sig, out, nullptr);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) qt_fidelity_batch(n, B, rho, sig, out, nullptr);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> h(B);
  cudaMemcpy(h.data(), out, B * 8, cudaMemcpyDeviceToHost);
  double sum = 0; for (double v : h) sum += v;
  char msg[256] = "";
  qt_last_error(msg, 256);
  printf("n=%d B=%lld: %.3f ms per call, %.3e pairs/s, checksum %.12f (%s) %s\n", n, B, ms / reps, B / (ms / reps * 1e-3), sum / B,
         cudaGetErrorString(err), msg);
  return 0;
}
