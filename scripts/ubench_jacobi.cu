// Ablation timing of the 64x64 block Jacobi step loop (dev tool; results with a non-zero mask are garbage).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_jacobi.bin scripts/ubench_jacobi.cu
#include <cstdio>
#include <vector>
#include "../forest_benchmarking_b200/csrc/qt_eigh.cuh"
void qt_set_error(const char*, ...) {}

constexpr int M = 64, LD = 65;
template <int NT, int ABL>
__global__ void __launch_bounds__(NT) jac_kernel(const cplx* in, double* evout, int sweeps, long long* cyc) {
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* A = reinterpret_cast<cplx*>(raw);
  cplx* V = A + M * LD;
  double* ev = reinterpret_cast<double*>(V + M * LD);
  double* scr = ev + M;
  const int tid = threadIdx.x;
  for (int e = tid; e < M * M; e += NT) A[(e / M) * LD + e % M] = in[(size_t)blockIdx.x * M * M + e];
  __syncthreads();
  long long t0 = clock64();
  jacobi_eigh_ring<64, NT, LD, true, ABL>(A, V, ev, scr, tid, true, sweeps, 1e-300);
  long long t1 = clock64();
  if (tid < M) evout[blockIdx.x * M + tid] = ev[tid] + V[tid * LD + 3].x;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NT, int ABL>
void run(const cplx* d_in, double* d_ev, long long* d_cyc, int sweeps, const char* name) {
  size_t smem = sizeof(cplx) * 2 * M * LD + sizeof(double) * (M + 3 * M + 64);
  cudaFuncSetAttribute(jac_kernel<NT, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#ifdef QT_JACOBI_PLAIN_BARRIER
  const int grid = 2;
#else
  const int grid = 148;
#endif
  jac_kernel<NT, ABL><<<grid, NT, smem>>>(d_in, d_ev, sweeps, d_cyc);
  cudaDeviceSynchronize();
  jac_kernel<NT, ABL><<<grid, NT, smem>>>(d_in, d_ev, sweeps, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("NT=%d %-46s %8.0f cycles per step (%s)\n", NT, name, (double)h / (sweeps * 63.0), cudaGetErrorString(e));
}

int main() {
  std::vector<cplx> h((size_t)148 * M * M);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
  for (int b = 0; b < 148; ++b)
    for (int r = 0; r < M; ++r)
      for (int c = r; c < M; ++c) {
        cplx v = make_double2(rnd(), r == c ? 0.0 : rnd());
        h[((size_t)b * M + r) * M + c] = v;
        h[((size_t)b * M + c) * M + r] = make_double2(v.x, -v.y);
      }
  cplx* d_in; double* d_ev; long long* d_cyc;
  cudaMalloc(&d_in, h.size() * sizeof(cplx)); cudaMalloc(&d_ev, 148 * M * 8); cudaMalloc(&d_cyc, 8);
  cudaMemcpy(d_in, h.data(), h.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  const int sw = 8;
#ifdef QT_JACOBI_PLAIN_BARRIER
  run<512, 0>(d_in, d_ev, d_cyc, 2, "full step, plain barrier (racecheck build)");
  return 0;
#endif
  run<512, 0>(d_in, d_ev, d_cyc, sw, "full step");
  run<512, 1>(d_in, d_ev, d_cyc, sw, "no V update");
  run<512, 2>(d_in, d_ev, d_cyc, sw, "no block update");
  run<512, 4>(d_in, d_ev, d_cyc, sw, "no rotation chain");
  run<512, 3>(d_in, d_ev, d_cyc, sw, "no V, no block (rotation chain + barrier)");
  run<512, 5>(d_in, d_ev, d_cyc, sw, "no V, no chain (block update + barrier)");
  run<512, 6>(d_in, d_ev, d_cyc, sw, "no block, no chain (V update + barrier)");
  run<512, 7>(d_in, d_ev, d_cyc, sw, "nothing but loads + barrier");
  run<512, 8>(d_in, d_ev, d_cyc, sw, "full step, no barrier (racy)");
  run<512, 5 + 16>(d_in, d_ev, d_cyc, sw, "block update w/o parameter shuffles");
  run<512, 5 + 32>(d_in, d_ev, d_cyc, sw, "block update w/o stores");
  run<512, 5 + 48>(d_in, d_ev, d_cyc, sw, "block update w/o shuffles and stores");
  run<512, 16>(d_in, d_ev, d_cyc, sw, "full step w/o parameter shuffles");
  run<256, 0>(d_in, d_ev, d_cyc, sw, "full step");
  run<256, 1>(d_in, d_ev, d_cyc, sw, "no V update");
  run<256, 2>(d_in, d_ev, d_cyc, sw, "no block update");
  run<256, 4>(d_in, d_ev, d_cyc, sw, "no rotation chain");
  return 0;
}
