// Batched state distance measures: one pair (rho_b, sigma_b) per warp.
//
//   trace_distance  distance_measures.py:100-114  0.5 * ||rho - sigma||_1 with NumPy's INDUCED 1-norm
//                   (max column abs-sum) -- the reference's actual behaviour, pinned by its own test
//                   tests/test_distance_measures.py:73-82.  Pure streaming, HBM-bound.
//   trace_distance_nuclear   the textbook 0.5 * sum |eig(rho - sigma)| (extra, not a reference function)
//   fidelity        distance_measures.py:64-84 + calculational.py:77-91
//                   (tr sqrt( sqrt(rho) sigma sqrt(rho) ))^2 = (sum_i sqrt(max(lambda_i, 0)))^2
//   purity          distance_measures.py:14-37    tr(rho^2)
#include "qt_common.cuh"
#include "qt_eigh.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

// Small states (d = 2, 4: fewer than 32 elements) are packed IPW = 32 / d^2 pairs per warp, one element per lane, so
// that every lane of every load is used; larger ones take one pair per warp with the element loop unrolled.
template <int D>
__global__ void trace_distance_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                      double* __restrict__ out) {
  constexpr int DD = D * D;
  constexpr int IPW = (DD < 32) ? 32 / DD : 1;   // pairs per warp
  constexpr int GL = (DD < 32) ? DD : 32;        // lanes per pair
  const int lane = threadIdx.x & 31, gl = lane % GL;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t b = warp * IPW + lane / GL;
  const bool live = b < B;
  const cplx* r = rho + (live ? b : 0) * DD;
  const cplx* s = sigma + (live ? b : 0) * DD;
  double colsum = 0.0;  // lane owns column gl % D, rows gl / D + k * (GL / D)
  if (D <= 32) {
#pragma unroll
    for (int e = gl; e < DD; e += GL) {
      const cplx a = r[e], c = s[e];
      colsum += sqrt(cabs2(csub(a, c)));
    }
#pragma unroll
    for (int o = GL / 2; o >= D; o >>= 1) colsum += __shfl_xor_sync(0xffffffffu, colsum, o);
    double m = colsum;
#pragma unroll
    for (int o = (D < 32 ? D : 32) / 2; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (live && gl == 0) out[b] = 0.5 * m;
  }
}

// tr(rho rho) = sum_ij rho_ij rho_ji (real part).  The transposed partner of an element is fetched without a second,
// uncoalesced global read: by shuffle when a state fits one warp pass (d <= 4), from a padded shared-memory copy
// of the state otherwise.
template <int D>
__global__ void purity_kernel(int64_t B, const cplx* __restrict__ rho, double* __restrict__ out) {
  constexpr int DD = D * D;
  constexpr int IPW = (DD < 32) ? 32 / DD : 1;
  constexpr int GL = (DD < 32) ? DD : 32;
  constexpr int WPB = (D >= 32) ? 2 : 8, LD = D + 1;
  const int lane = threadIdx.x & 31, gl = lane % GL, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * WPB + wib;
  const int64_t b = warp * IPW + lane / GL;
  const bool live = b < B;
  const cplx* r = rho + (live ? b : 0) * DD;
  double acc = 0.0;
  if constexpr (DD < 32) {
    const int i = gl / D, j = gl % D;
    const cplx a = r[gl];
    cplx c;
    c.x = __shfl_sync(0xffffffffu, a.x, (lane - gl) + j * D + i);
    c.y = __shfl_sync(0xffffffffu, a.y, (lane - gl) + j * D + i);
    acc = a.x * c.x - a.y * c.y;
  } else {
    __shared__ cplx tile[WPB][D * LD];
    cplx* t = tile[wib];
#pragma unroll
    for (int e = lane; e < DD; e += 32) t[(e / D) * LD + e % D] = r[e];
    __syncwarp();
#pragma unroll
    for (int e = lane; e < DD; e += 32) {
      const int i = e / D, j = e % D;
      const cplx a = t[i * LD + j], c = t[j * LD + i];
      acc += a.x * c.x - a.y * c.y;
    }
  }
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && gl == 0) out[b] = acc;
}

// shared memory per warp: 3 matrices (leading dimension LD = D + 1 for D >= 8: the Jacobi block updates and the
// products read columns, which an unpadded power-of-two row stride maps onto the same banks) + eigenvalues +
// Jacobi scratch
template <int D>
struct FidSmem {
  static constexpr int LD = (D >= 8) ? D + 1 : D;
  static constexpr int MP = D * LD;
  static constexpr size_t bytes =
      (sizeof(cplx) * MP * 3 + sizeof(double) * (D + JacobiScratch<D>::doubles) + 15) / 16 * 16;
};

// Fast path of the fidelity for a numerically positive-definite rho: rho = L L^dagger (Cholesky), and the eigenvalues of
// L^dagger sigma L are those of sqrt(rho) sigma sqrt(rho) (similar matrices), so ONE values-only eigendecomposition
// replaces eigh(rho) with vectors + sqrtm + the second eigh.  Two matrices of shared memory per warp instead of three
// and no eigenvector registers: 1.5x the resident warps of fidelity_kernel.  A pivot that is not safely positive
// (rank-deficient or non-PSD input, where the reference's clamping in sqrtm_psd matters) writes FID_FLAG and leaves the
// item to fidelity_kernel, which runs the reference's own sequence on the flagged items.
constexpr double FID_FLAG = -1.0;  // a fidelity is never negative
template <int D>
struct FidFastSmem {
  static constexpr int LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  static constexpr size_t bytes = (sizeof(cplx) * MP * 2 + sizeof(double) * (D + JacobiScratch<D>::doubles) + 15) / 16 * 16;
};

template <int D>
__global__ void __launch_bounds__(256, (D == 16) ? 3 : 1) fidelity_fast_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                     double* __restrict__ out) {
  constexpr int DD = D * D, LD = FidFastSmem<D>::LD, MP = FidFastSmem<D>::MP, PER = (DD + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidFastSmem<D>::bytes * wib);
  cplx* W = A + MP;
  double* ev = reinterpret_cast<double*>(W + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  const cplx* r = rho + b * DD;
  const cplx* s = sigma + b * DD;
  // scipy.linalg.eigh reads the lower triangle only: factor the Hermitian matrix it sees
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = r[e];
    if (i == j) v.y = 0.0;
    if (i >= j) A[i * LD + j] = v;
  }
  __syncwarp();
  double dmax = 0.0;
  for (int k = 0; k < D; ++k) dmax = fmax(dmax, A[k * LD + k].x);
  const double thr = 1e-10 * dmax;
  bool ok = dmax > 0.0;
  for (int j = 0; j < D && ok; ++j) {
    const double piv = A[j * LD + j].x;
    if (!(piv > thr)) {
      ok = false;
      break;
    }
    const double inv = rsqrt(piv);
    __syncwarp();
    for (int i = j + lane; i < D; i += 32) A[i * LD + j] = (i == j) ? cmake(piv * inv, 0.0) : cscale(A[i * LD + j], inv);
    __syncwarp();
    const int nt = D - j - 1;
    for (int e = lane; e < nt * nt; e += 32) {
      const int i = j + 1 + e / nt, k = j + 1 + e % nt;
      if (k <= i) {
        const cplx li = A[i * LD + j], lk = A[k * LD + j];
        cplx v = A[i * LD + k];
        v.x -= li.x * lk.x + li.y * lk.y;  // v -= li * conj(lk)
        v.y -= li.y * lk.x - li.x * lk.y;
        if (i == k) v.y = 0.0;
        A[i * LD + k] = v;
      }
    }
    __syncwarp();
  }
  if (!ok) {
    if (lane == 0) out[b] = FID_FLAG;
    return;
  }
  // W = sigma L  (L lower triangular in A)
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = j; k < D; ++k) cfma(acc, s[i * D + k], A[k * LD + j]);
    W[i * LD + j] = acc;
  }
  __syncwarp();
  // Y = L^dagger W: lower triangle (incl. diagonal) held in registers until every lane has finished reading L, then
  // written over A together with its conjugate mirror -- the Hermitian matrix an eigensolver reading the lower
  // triangle would see
  cplx y[PER];
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int e = lane + 32 * t, i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    if (e < DD && i >= j)
      for (int k = i; k < D; ++k) cfma(acc, cconj(A[k * LD + i]), W[k * LD + j]);
    y[t] = acc;
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int e = lane + 32 * t, i = e / D, j = e % D;
    if (e < DD && i >= j) {
      cplx v = y[t];
      if (i == j) v.y = 0.0;
      A[i * LD + j] = v;
      if (i > j) A[j * LD + i] = cconj(v);
    }
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, false, LD>(A, nullptr, ev, ev + D, lane);
  double acc = 0.0;
  for (int k = lane; k < D; k += 32) acc += sqrt(fmax(ev[k], 0.0));
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc * acc;
}

// MODE 0: fidelity.  MODE 1: nuclear-norm trace distance.
template <int D, int MODE>
__global__ void fidelity_kernel(int64_t B, const cplx* __restrict__ rho, const cplx* __restrict__ sigma,
                                double* __restrict__ out, int only_flagged) {
  constexpr int DD = D * D, LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidSmem<D>::bytes * wib);
  cplx* V = A + MP;
  cplx* W = V + MP;
  double* ev = reinterpret_cast<double*>(W + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  if (only_flagged && out[b] != FID_FLAG) return;  // fidelity_fast_kernel already produced this item
  const cplx* r = rho + b * DD;
  const cplx* s = sigma + b * DD;
  if constexpr (MODE == 1) {
    for (int e = lane; e < DD; e += 32) {
      // Hermitian part of rho - sigma
      const int i = e / D, j = e % D;
      const cplx d1 = csub(r[e], s[e]), d2 = csub(r[j * D + i], s[j * D + i]);
      A[i * LD + j] = cmake(0.5 * (d1.x + d2.x), 0.5 * (d1.y - d2.y));
    }
    __syncwarp();
    jacobi_eigh<D, 32, SyncWarp, false, LD>(A, nullptr, ev, ev + D, lane);
    double acc = 0.0;
    for (int k = lane; k < D; k += 32) acc += fabs(ev[k]);
    acc = warp_sum(acc);
    if (lane == 0) out[b] = 0.5 * acc;
    return;
  }
  // scipy.linalg.eigh reads the lower triangle only: build the Hermitian matrix it sees
  auto load_rho = [&]() {
    for (int e = lane; e < DD; e += 32) {  // coalesced read; the lower triangle is mirrored into the upper one
      const int i = e / D, j = e % D;
      cplx v = r[e];
      if (i == j) v.y = 0.0;
      if (i >= j) {
        A[i * LD + j] = v;
        if (i > j) A[j * LD + i] = cconj(v);
      }
    }
    __syncwarp();
  };
  load_rho();
  jacobi_eigh<D, 32, SyncWarp, true, LD>(A, V, ev, ev + D, lane);
  // S = V sqrt(max(ev,0)) V^dagger  -> A
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) {
      const double w = sqrt(fmax(ev[k], 0.0));
      cfma_conj(acc, cscale(V[i * LD + k], w), V[j * LD + k]);
    }
    A[i * LD + j] = acc;
  }
  __syncwarp();
  // W = S sigma
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma(acc, A[i * LD + k], s[k * D + j]);
    W[i * LD + j] = acc;
  }
  __syncwarp();
  // V = W S, then the Hermitian matrix eigh would see (lower triangle)
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx acc = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma(acc, W[i * LD + k], A[k * LD + j]);
    V[i * LD + j] = acc;
  }
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = (i >= j) ? V[i * LD + j] : cconj(V[j * LD + i]);
    if (i == j) v.y = 0.0;
    W[i * LD + j] = v;
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, false, LD>(W, nullptr, ev, ev + D, lane);
  double acc = 0.0;
  for (int k = lane; k < D; k += 32) acc += sqrt(fmax(ev[k], 0.0));
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc * acc;
}

// project_state_matrix_to_physical (operator_tools/project_state_matrix.py:6-52), one state per warp:
// rho / tr(rho) -> eigh (lower triangle, like scipy) -> if any eigenvalue is negative, zero the smallest ones and
// spread their mass evenly over the rest (water filling) -> V diag(l') V^dagger.  A matrix that is already PSD
// is returned as rho / tr(rho) unchanged, exactly like the reference.
template <int D>
__global__ void project_state_kernel(int64_t B, const cplx* __restrict__ rho, cplx* __restrict__ out) {
  constexpr int DD = D * D, LD = FidSmem<D>::LD, MP = FidSmem<D>::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  cplx* A = reinterpret_cast<cplx*>(smem_raw + FidSmem<D>::bytes * wib);
  cplx* V = A + MP;
  cplx* R = V + MP;  // rho / tr(rho), kept for the already-physical case
  double* ev = reinterpret_cast<double*>(R + MP);
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  const cplx* r = rho + b * DD;
  cplx tr = cmake(0.0, 0.0);
  for (int k = lane; k < D; k += 32) tr = cadd(tr, r[k * D + k]);
  tr.x = warp_sum(tr.x);
  tr.y = warp_sum(tr.y);
  const double n2 = cabs2(tr);
  const cplx itr = cmake(tr.x / n2, -tr.y / n2);
  for (int e = lane; e < DD; e += 32) R[(e / D) * LD + e % D] = cmul(r[e], itr);
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int i = e / D, j = e % D;
    cplx v = (i >= j) ? R[i * LD + j] : cconj(R[j * LD + i]);
    if (i == j) v.y = 0.0;
    A[i * LD + j] = v;
  }
  __syncwarp();
  jacobi_eigh<D, 32, SyncWarp, true, LD>(A, V, ev, ev + D, lane);
  double mn = 1e300;
  for (int k = 0; k < D; ++k) mn = fmin(mn, ev[k]);
  cplx* dst = out + b * DD;
  if (mn >= 0.0) {
    for (int e = lane; e < DD; e += 32) dst[e] = R[(e / D) * LD + e % D];
    return;
  }
  // rank of each eigenvalue in DESCENDING order (ties broken by index), then the water-filling scan
  double* desc = ev + D;      // the Jacobi scratch (3 D + 32 doubles) is free now
  double* wgt = ev + 2 * D;   // new eigenvalue of eigenpair k
  __syncwarp();
  for (int k = lane; k < D; k += 32) {
    int rk = 0;
    for (int j = 0; j < D; ++j) rk += (ev[j] > ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
    desc[rk] = ev[k];
  }
  __syncwarp();
  int i = D;
  double acc = 0.0;
  while (i > 1 && desc[i - 1] + acc / (double)i < 0.0) {
    acc += desc[i - 1];
    --i;
  }
  const double shift = acc / (double)i;  // the i largest eigenvalues are kept and shifted
  for (int k = lane; k < D; k += 32) {
    int rk = 0;
    for (int j = 0; j < D; ++j) rk += (ev[j] > ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
    wgt[k] = (rk < i) ? ev[k] + shift : 0.0;
  }
  __syncwarp();
  for (int e = lane; e < DD; e += 32) {
    const int a = e / D, c = e % D;
    cplx s = cmake(0.0, 0.0);
    for (int k = 0; k < D; ++k) cfma_conj(s, cscale(V[a * LD + k], wgt[k]), V[c * LD + k]);
    dst[e] = s;
  }
}

template <int D>
static int launch_project_state(int64_t B, const void* rho, void* out, cudaStream_t st) {
  const size_t per_warp = FidSmem<D>::bytes;
  int wpb = (int)max((size_t)1, min((size_t)8, (size_t)(112 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  QT_CUDA(cudaFuncSetAttribute(project_state_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_state_kernel<D><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho, (cplx*)out);
  return qt_check_launch("project_state_kernel");
}

// Hilbert-Schmidt inner product tr(A^dagger B) per pair (distance_measures.py:198-216); the kernel behind
// entanglement_fidelity / process_fidelity (:271-375).  Pure streaming: one warp per pair for small matrices, one
// block per pair otherwise.  out[b] is complex.
__global__ void hs_inner_kernel(int64_t elems, int64_t B, const cplx* __restrict__ a, const cplx* __restrict__ b,
                                cplx* __restrict__ out, int block_per_pair) {
  __shared__ double red[2][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  if (block_per_pair == 2) {
    // fewer than 32 elements per matrix (a power of two): 32 / elems pairs per warp, one element per lane
    const int gl_n = (int)elems, ipw = 32 / gl_n, gl = lane % gl_n;
    const int64_t p = ((int64_t)blockIdx.x * wpb + wib) * ipw + lane / gl_n;
    const bool live = p < B;
    cplx acc = cmake(0.0, 0.0);
    if (live) cfma(acc, cconj(a[p * elems + gl]), b[p * elems + gl]);
    for (int o = gl_n / 2; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (live && gl == 0) out[p] = acc;
    return;
  }
  if (!block_per_pair) {
    const int64_t p = (int64_t)blockIdx.x * wpb + wib;
    if (p >= B) return;
    cplx acc = cmake(0.0, 0.0);
    for (int64_t e = lane; e < elems; e += 32) cfma(acc, cconj(a[p * elems + e]), b[p * elems + e]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) out[p] = acc;
    return;
  }
  for (int64_t p = blockIdx.x; p < B; p += gridDim.x) {
    cplx acc = cmake(0.0, 0.0);
    for (int64_t e = threadIdx.x; e < elems; e += blockDim.x) cfma(acc, cconj(a[p * elems + e]), b[p * elems + e]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    __syncthreads();
    if (lane == 0) {
      red[0][wib] = acc.x;
      red[1][wib] = acc.y;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      cplx t = cmake(0.0, 0.0);
      for (int w = 0; w < wpb; ++w) t = cadd(t, cmake(red[0][w], red[1][w]));
      out[p] = t;
    }
  }
}

template <int D>
static int launch_td(int64_t B, const void* rho, const void* sigma, double* out, cudaStream_t st) {
  const int wpb = 8;
  constexpr int IPW_TD = (D * D < 32) ? 32 / (D * D) : 1;
  const int64_t warps_td = (B + IPW_TD - 1) / IPW_TD;
  trace_distance_kernel<D><<<(unsigned)((warps_td + wpb - 1) / wpb), 32 * wpb, 0, st>>>(B, (const cplx*)rho,
                                                                                 (const cplx*)sigma, out);
  return qt_check_launch("trace_distance_kernel");
}
template <int D>
static int launch_purity(int64_t B, const void* rho, double* out, cudaStream_t st) {
  constexpr int IPW_P = (D * D < 32) ? 32 / (D * D) : 1;
  constexpr int WPB_P = (D >= 32) ? 2 : 8;  // must match purity_kernel
  const int64_t warps_p = (B + IPW_P - 1) / IPW_P;
  purity_kernel<D><<<(unsigned)((warps_p + WPB_P - 1) / WPB_P), 32 * WPB_P, 0, st>>>(B, (const cplx*)rho, out);
  return qt_check_launch("purity_kernel");
}
template <int D, int MODE>
static int launch_fid(int64_t B, const void* rho, const void* sigma, double* out, cudaStream_t st) {
  const size_t per_warp = FidSmem<D>::bytes;
  int wpb = (int)max((size_t)1, min((size_t)8, (size_t)(112 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  int only_flagged = 0;
  if constexpr (MODE == 0 && D >= 4) {
    const size_t per_fast = FidFastSmem<D>::bytes;
    const int wf = (int)max((size_t)1, min((size_t)8, (size_t)(74 * 1024) / per_fast));
    QT_CUDA(cudaFuncSetAttribute(fidelity_fast_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_fast * wf)));
    fidelity_fast_kernel<D><<<(unsigned)((B + wf - 1) / wf), 32 * wf, per_fast * wf, st>>>(B, (const cplx*)rho,
                                                                                          (const cplx*)sigma, out);
    int rc = qt_check_launch("fidelity_fast_kernel");
    if (rc) return rc;
    only_flagged = 1;
  }
  QT_CUDA(cudaFuncSetAttribute(fidelity_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fidelity_kernel<D, MODE><<<(unsigned)((B + wpb - 1) / wpb), 32 * wpb, smem, st>>>(B, (const cplx*)rho,
                                                                                    (const cplx*)sigma, out, only_flagged);
  return qt_check_launch("fidelity_kernel");
}

#define DISPATCH_D(n, CALL)                      \
  switch (n) {                                   \
    case 1: return CALL(2);                      \
    case 2: return CALL(4);                      \
    case 3: return CALL(8);                      \
    case 4: return CALL(16);                     \
    case 5: return CALL(32);                     \
    default:                                     \
      qt_set_error("n=%d out of range 1..5", n); \
      return QT_ERR_ARG;                         \
  }

extern "C" int qt_trace_distance_batch(int n, int64_t B, const void* rho, const void* sigma, double* out,
                                       void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_trace_distance_batch: null argument");
#define CALL(D) launch_td<D>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_trace_distance_nuclear_batch(int n, int64_t B, const void* rho, const void* sigma, double* out,
                                               void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_trace_distance_nuclear_batch: null argument");
#define CALL(D) launch_fid<D, 1>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_fidelity_batch(int n, int64_t B, const void* rho, const void* sigma, double* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && sigma && out, "qt_fidelity_batch: null argument");
#define CALL(D) launch_fid<D, 0>(B, rho, sigma, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_purity_batch(int n, int64_t B, const void* rho, double* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && out, "qt_purity_batch: null argument");
#define CALL(D) launch_purity<D>(B, rho, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_project_state_batch(int n, int64_t B, const void* rho, void* out, void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rho && out, "qt_project_state_batch: null argument");
#define CALL(D) launch_project_state<D>(B, rho, out, (cudaStream_t)stream)
  DISPATCH_D(n, CALL)
#undef CALL
}

extern "C" int qt_hs_inner_batch(int64_t rows, int64_t cols, int64_t B, const void* a, const void* b, void* out,
                                 void* stream) {
  if (B == 0) return QT_OK;  // empty batch: nothing to check, nothing to do
  QT_REQUIRE(rows > 0 && cols > 0 && a && b && out, "qt_hs_inner_batch: bad arguments");
  const int64_t elems = rows * cols;
  const bool packed = elems < 32 && (elems & (elems - 1)) == 0;
  const int block_per_pair = packed ? 2 : (elems > 1024 ? 1 : 0);
  const int64_t blocks = packed ? (B + 8 * (32 / elems) - 1) / (8 * (32 / elems))
                                : (block_per_pair ? std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 8) : (B + 7) / 8);
  hs_inner_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(elems, B, (const cplx*)a, (const cplx*)b,
                                                                      (cplx*)out, block_per_pair);
  return qt_check_launch("hs_inner_kernel");
}
