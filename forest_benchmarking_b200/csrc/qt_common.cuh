// Shared device/host helpers for libqtomo (sm_100a only).
//
// Layout contract (include/qtomo.h): every matrix batch is row-major [B, rows, cols] interleaved
// complex128, i.e. element (b, r, c) lives at ((b*R + r)*C + c) * 16 bytes -- the C-order layout
// of a numpy complex128 array, so a reference user's arrays can be handed over unchanged.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef double2 cplx;  // .x = real, .y = imag

#define QT_OK 0
#define QT_ERR_ARG (-1)
#define QT_ERR_CUDA (-2)
#define QT_ERR_UNSUPPORTED (-3)
#define QT_ERR_WORKSPACE (-4)

void qt_set_error(const char* fmt, ...);
// eigh_rel_tol argument of the C ABI -> squared relative off-diagonal tolerance handed to the Jacobi solver
// (< 0: the default for n qubits; 0: the tight 1e-15 * 4^n; otherwise the value, which must be < 1e-3)
int qt_eigh_rel2_from_tol(double rel_tol, int n, double* rel2_out, const char* who);
int qt_num_sms();  // multiprocessor count of the current device (cached per device)
int qt_check_launch(const char* what);

#define QT_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      qt_set_error(__VA_ARGS__);             \
      return QT_ERR_ARG;                     \
    }                                        \
  } while (0)

#define QT_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      qt_set_error("%s failed: %s", #call, cudaGetErrorString(e_));          \
      return QT_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define QT_NUM_SMS qt_num_sms()  // 148 on B200; queried, not assumed

// ---------------------------------------------------------------------------------------------
// complex arithmetic on double2
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += a * conj(b)
__device__ __forceinline__ void cfma_conj(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
}
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// multiply by i^k, k in 0..3
__host__ __device__ __forceinline__ cplx cmul_ipow(cplx a, int k) {
  switch (k & 3) {
    case 0: return a;
    case 1: return make_double2(-a.y, a.x);
    case 2: return make_double2(-a.x, -a.y);
    default: return make_double2(a.y, -a.x);
  }
}

// 1/d to ~1 ulp: MUFU.RCP64H seed (20+ bits) + two Newton steps = 1 MUFU + 4 DFMA.  An IEEE division costs
// ~133 issue cycles per warp on B200 (profiles/r01_ubench_fp64.txt).  d must be a normal, finite number.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// 1/sqrt(d) to ~1 ulp, branch-free: MUFU.RSQ64H seed + two Newton steps.  CUDA's rsqrt() carries a slow-path
// CALL for special inputs, which splits the caller's basic block (see qt_eigh.cuh).  d must be normal, > 0.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-(d * r), r, 1.0);
  r = fma(0.5 * r, e, r);
  e = fma(-(d * r), r, 1.0);
  return fma(0.5 * r, e, r);
}

// ---------------------------------------------------------------------------------------------
// Pauli bookkeeping.  Canonical index = base-4 number, digits I=0 X=1 Y=2 Z=3, first qubit most
// significant (reference utils.py:146-156, 398-409).  A Pauli is i^{|x&z|} X^x Z^z with n-bit masks
// (qubit 0 = most significant bit of the row index):  P[r, c] = [r^c == x] * i^{|x&z|} * (-1)^{|z&c|}.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ constexpr int pauli_xmask(int idx, int n) {
  int m = 0;
  for (int q = 0; q < n; ++q) {
    int d = (idx >> (2 * (n - 1 - q))) & 3;
    m |= ((d == 1 || d == 2) ? 1 : 0) << (n - 1 - q);
  }
  return m;
}
__host__ __device__ __forceinline__ constexpr int pauli_zmask(int idx, int n) {
  int m = 0;
  for (int q = 0; q < n; ++q) {
    int d = (idx >> (2 * (n - 1 - q))) & 3;
    m |= ((d == 2 || d == 3) ? 1 : 0) << (n - 1 - q);
  }
  return m;
}
__host__ __device__ __forceinline__ constexpr int popc_c(int v) {
  int c = 0;
  for (; v; v &= v - 1) ++c;
  return c;
}
// inverse: (x, z) masks -> canonical index
__host__ __device__ __forceinline__ constexpr int pauli_from_masks(int x, int z, int n) {
  int idx = 0;
  for (int q = 0; q < n; ++q) {
    int xb = (x >> (n - 1 - q)) & 1, zb = (z >> (n - 1 - q)) & 1;
    int d = xb ? (zb ? 2 : 1) : (zb ? 3 : 0);
    idx |= d << (2 * (n - 1 - q));
  }
  return idx;
}

// Thread-per-item kernels stage 128 items through shared memory element-major: tile[element * QT_TS + item].  With a
// stride of exactly 128 complex numbers the coalesced side of the transposition (consecutive lanes = consecutive
// elements of one item) puts a whole quarter warp on the same banks (8-way conflict); 129 spreads it over all 32.
constexpr int QT_TS = 129;

// ---------------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
