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

// role-split variant (A warps / V warps); same inputs, extra FIFO buffer after V
template <int ABL>
__global__ void __launch_bounds__(512) jac_split_abl_kernel(const cplx* in, double* evout, int sweeps, long long* cyc) {
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* A = reinterpret_cast<cplx*>(raw);
  cplx* V = A + M * LD;
  double* fifo = reinterpret_cast<double*>(V + M * LD);
  double* ev = fifo + M * LD * 2;
  double* scr = ev + M;
  const int tid = threadIdx.x;
  for (int e = tid; e < M * M; e += 512) A[(e / M) * LD + e % M] = in[(size_t)blockIdx.x * M * M + e];
  __syncthreads();
  long long t0 = clock64();
  jacobi_eigh_ring_split64<LD, ABL>(A, V, ev, scr, fifo, tid, true, sweeps, 1e-300, false);
  long long t1 = clock64();
  if (tid < M) evout[blockIdx.x * M + tid] = ev[tid] + V[tid * LD + 3].x;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ABL>
void run_split_abl(const cplx* d_in, double* d_ev, long long* d_cyc, int sweeps, const char* name) {
  size_t smem = sizeof(cplx) * 3 * M * LD + sizeof(double) * (M + 3 * M + 64);
  cudaFuncSetAttribute(jac_split_abl_kernel<ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  jac_split_abl_kernel<ABL><<<148, 512, smem>>>(d_in, d_ev, sweeps, d_cyc);
  cudaDeviceSynchronize();
  jac_split_abl_kernel<ABL><<<148, 512, smem>>>(d_in, d_ev, sweeps, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("SPLIT %-46s %8.0f cycles per step (%s)\n", name, (double)h / (sweeps * 63.0), cudaGetErrorString(e));
}

__global__ void __launch_bounds__(512) jac_split_kernel(const cplx* in, double* evout, cplx* vout, int sweeps, double rel2,
                                                        int split, long long* cyc, int* sweeps_out) {
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* A = reinterpret_cast<cplx*>(raw);
  cplx* V = A + M * LD;
  double* fifo = reinterpret_cast<double*>(V + M * LD);
  double* ev = fifo + M * LD * 2;
  double* scr = ev + M;
  const int tid = threadIdx.x;
  for (int e = tid; e < M * M; e += 512) A[(e / M) * LD + e % M] = in[(size_t)blockIdx.x * M * M + e];
  __syncthreads();
  long long t0 = clock64();
  const int sw = split ? jacobi_eigh_ring_split64<LD>(A, V, ev, scr, fifo, tid, true, sweeps, rel2, true)
                       : jacobi_eigh_ring<64, 512, LD, true, 0>(A, V, ev, scr, tid, true, sweeps, rel2, true);
  long long t1 = clock64();
  if (tid < M) evout[blockIdx.x * M + tid] = ev[tid];
  for (int e = tid; e < M * M; e += 512) {
    vout[((size_t)blockIdx.x * 2) * M * M + e] = V[(e / M) * LD + e % M];
    vout[((size_t)blockIdx.x * 2 + 1) * M * M + e] = A[(e / M) * LD + e % M];
  }
  if (tid == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; sweeps_out[0] = sw; }
}

static void run_split(const cplx* d_in, const std::vector<cplx>& h_in, int sweeps, double rel2) {
  const int grid = 148;
  size_t smem = sizeof(cplx) * 3 * M * LD + sizeof(double) * (M + 3 * M + 64);
  cudaFuncSetAttribute(jac_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double* d_ev; cplx* d_v; long long* d_cyc; int* d_sw;
  cudaMalloc(&d_ev, grid * M * 8); cudaMalloc(&d_v, (size_t)grid * 2 * M * M * sizeof(cplx)); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sw, 4);
  std::vector<double> ev[2]; std::vector<cplx> vv[2];
  for (int split = 0; split < 2; ++split) {
    jac_split_kernel<<<grid, 512, smem>>>(d_in, d_ev, d_v, sweeps, rel2, split, d_cyc, d_sw);
    cudaDeviceSynchronize();
    jac_split_kernel<<<grid, 512, smem>>>(d_in, d_ev, d_v, sweeps, rel2, split, d_cyc, d_sw);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; int sw;
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&sw, d_sw, 4, cudaMemcpyDeviceToHost);
    ev[split].resize(grid * M); vv[split].resize((size_t)grid * 2 * M * M);
    cudaMemcpy(ev[split].data(), d_ev, grid * M * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(vv[split].data(), d_v, vv[split].size() * sizeof(cplx), cudaMemcpyDeviceToHost);
    printf("%s rel2=%g: %d sweeps, %8.0f cycles per step (%s)\n", split ? "ROLE-SPLIT" : "plain     ", rel2, sw,
           (double)h / (sw * 63.0), cudaGetErrorString(e));
  }
  // the two variants execute the same arithmetic in the same order: results must agree to rounding
  double dev = 0, dv = 0, da = 0, recon = 0;
  for (size_t i = 0; i < ev[0].size(); ++i) dev = fmax(dev, fabs(ev[0][i] - ev[1][i]));
  for (int b = 0; b < grid; ++b)
    for (int e = 0; e < M * M; ++e) {
      const cplx a = vv[0][((size_t)b * 2) * M * M + e], c = vv[1][((size_t)b * 2) * M * M + e];
      dv = fmax(dv, hypot(a.x - c.x, a.y - c.y));
      const cplx a2 = vv[0][((size_t)b * 2 + 1) * M * M + e], c2 = vv[1][((size_t)b * 2 + 1) * M * M + e];
      da = fmax(da, hypot(a2.x - c2.x, a2.y - c2.y));
    }
  // split result on its own: || V (D + R) V^dagger - A0 ||_max for block 0
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < M; ++c) {
      double sx = 0, sy = 0;
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
          const cplx v1 = vv[1][r * M + i], aij = vv[1][(size_t)M * M + i * M + j], v2 = vv[1][c * M + j];
          const double tx = v1.x * aij.x - v1.y * aij.y, ty = v1.x * aij.y + v1.y * aij.x;
          sx += tx * v2.x + ty * v2.y; sy += ty * v2.x - tx * v2.y;
        }
      recon = fmax(recon, hypot(sx - h_in[r * M + c].x, sy - h_in[r * M + c].y));
    }
  printf("  split vs plain: max |d ev| %.2e  max |d V| %.2e  max |d A| %.2e;  split: max |V (D+R) V^H - A0| %.2e\n", dev, dv, da, recon);
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
#ifndef QT_JACOBI_PLAIN_BARRIER
  run_split(d_in, h, 30, 0.0);
  run_split(d_in, h, 30, 1e-10);
  run_split(d_in, h, 4, 1e-300);
  run_split_abl<0>(d_in, d_ev, d_cyc, 8, "full (unroll 1)");
  run_split_abl<1>(d_in, d_ev, d_cyc, 8, "V warps idle (A path alone)");
  run_split_abl<2>(d_in, d_ev, d_cyc, 8, "no block update (rotation + FIFO + V warps)");
  run_split_abl<3>(d_in, d_ev, d_cyc, 8, "V idle, no block update (rotation chain only)");
#endif
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
