// Throughput of SHFL.32 / LDS.128 / STS.128 / FSEL with 16 warps per SM (the Jacobi block configuration).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_shfl.bin scripts/ubench_shfl.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void shfl_tput(int iters, float* out, long long* cyc) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __shfl_down_sync(0xffffffffu, a[i], 1);
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 42.0f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_tput(int iters, float* out, long long* cyc) {
  __shared__ double2 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 v = buf[(idx + i * 256) & 2047];
      acc.x += v.x;
      acc.y += v.y;
    }
    idx = (idx + 32) & 2047;
  }
  __syncthreads();
  long long t1 = clock64();
  if (acc.x + acc.y == 42.0) out[0] = (float)acc.x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void sts_tput(int iters, float* out, long long* cyc) {
  __shared__ double2 buf[2048];
  int idx = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) buf[(idx + i * 256) & 2047] = make_double2(it, i);
    idx = (idx + 32) & 2047;
  }
  __syncthreads();
  long long t1 = clock64();
  if (buf[threadIdx.x].x == 42.5) out[0] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void fsel_tput(int iters, float* out, long long* cyc) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const bool f = (threadIdx.x & 1);
  double b = out[1];
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (f ^ ((it + i) & 1)) ? a[(i + 1) & 7] : b;
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 42.0) out[0] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 64); cudaMemset(out, 0, 64); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  for (int threads = 128; threads <= 512; threads *= 2) {
#define RUN(K, name, per, bytes)                                                     \
    K<<<148, threads>>>(iters, out, cyc); cudaDeviceSynchronize();                   \
    K<<<148, threads>>>(iters, out, cyc); cudaDeviceSynchronize();                   \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                  \
    printf("%-10s threads/SM=%3d: %.2f cycles per warp-instruction per SM\n", name, threads, \
           (double)h / ((double)iters * per * (threads / 32)));
    RUN(shfl_tput, "SHFL.32", 8, 4)
    RUN(lds_tput, "LDS.128", 8, 16)
    RUN(sts_tput, "STS.128", 8, 16)
    RUN(fsel_tput, "SEL.64", 8, 8)
  }
  return 0;
}
