// Batched Choi-matrix projections: CP, TP, TNI and physical (Dykstra).  See qt_choi.cuh for the math and
// the reference citations (operator_tools/project_superoperators.py:19-144).
#include "qt_choi.cuh"
#include "../../include/qtomo.h"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

// n = 4, 5 (matrices that do not fit shared memory): qt_eigh_large.cu
int qt_large_proj_cp(int n, int64_t B, const void* in, void* out, void* ws, int64_t ws_bytes, cudaStream_t st);
int qt_choi2kraus_lowrank64(int64_t B, const void* choi, double tol, double* evals_out, void* kraus_out,
                            int32_t* count_out, cudaStream_t st);
int qt_large_tp_correction(int n, int64_t B, const void* in, void* out, int make_tp, cudaStream_t st);
int64_t qt_large_physical_workspace_bytes(int n, int64_t B);
int qt_large_proj_physical(int n, int64_t B, const void* in, void* out, int make_tp, void* ws, int64_t ws_bytes,
                           int* calls, int* status, cudaStream_t st);

template <int N>
struct ProjCfg {
  static constexpr int NT = (N >= 3) ? QT_N3_THREADS : 32;           // threads per group
  static constexpr int GPB = (N >= 3) ? 1 : 4;             // groups per block
  using Sync = typename std::conditional<(N >= 3), SyncBlock, SyncWarp>::type;
  using G = ChoiGroup<N, NT, Sync>;
  static constexpr size_t group_smem = G::group_smem;      // X, V, T (padded) + small scratch
};

// ---- CP ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(ProjCfg<N>::NT * ProjCfg<N>::GPB)
    proj_cp_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  using C = ProjCfg<N>;
  using G = typename C::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw + C::group_smem * gib);
  cplx* V = X + G::MP;
  double* small = reinterpret_cast<double*>(V + 2 * G::MP);
  const int64_t b = (int64_t)blockIdx.x * C::GPB + gib;
  if (b >= B) return;
  const cplx* src = in + b * G::MM;
  auto herm = [&](int r, int c) {
    const cplx x = src[r * G::M + c], y = src[c * G::M + r];
    return cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
  };
  for (int e = tid; e < G::MM; e += C::NT) X[G::sidx(e)] = herm(e / G::M, e % G::M);
  C::Sync::sync();
  jacobi_eigh<G::M, C::NT, typename C::Sync, true, G::LD>(X, V, small, small + G::M, tid);
  G::recompose_psd(X, V, small, herm, tid);
  cplx* dst = out + b * G::MM;
  for (int e = tid; e < G::MM; e += C::NT) dst[e] = X[G::sidx(e)];
}

// ---- closest unitary ---------------------------------------------------------------------------------
// proj_choi_to_unitary (project_superoperators.py:147-175): eigh of the Hermitian part, the eigenvector of the
// largest eigenvalue un-vec'ed (column stacking) into the dominant Kraus operator K, the unitary polar factor of K
// (U @ Vh of the reference's SVD == K (K^dagger K)^{-1/2}: one d x d eigh by the first warp of the group), global
// phase fixed so that element (0, 0) is real and non-negative, and kraus2choi of the result.  The eigenvector's
// arbitrary phase cancels in the last two steps.  K must have full rank (any process with a dominant Kraus
// operator close to a unitary); a singular K has no unique polar factor.
template <int N>
__global__ void __launch_bounds__(ProjCfg<N>::NT * ProjCfg<N>::GPB)
    proj_unitary_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  using C = ProjCfg<N>;
  using G = typename C::G;
  constexpr int D = G::D, DD = D * D, M = G::M, LD = G::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw + C::group_smem * gib);
  cplx* V = X + G::MP;
  double* small = reinterpret_cast<double*>(V + 2 * G::MP);
  double* base = small + M + JacobiScratch<M>::doubles;
  cplx* K = reinterpret_cast<cplx*>(base);
  cplx* P = K + DD;
  cplx* W = P + DD;
  cplx* U = W + DD;
  double* pev = reinterpret_cast<double*>(U + DD);
  double* pscr = pev + D + (D & 1);
  const int64_t b = (int64_t)blockIdx.x * C::GPB + gib;
  if (b >= B) return;
  const cplx* src = in + b * G::MM;
  for (int e = tid; e < G::MM; e += C::NT) {
    const int r = e / M, c = e % M;
    const cplx x = src[r * M + c], y = src[c * M + r];
    X[G::sidx(e)] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
  }
  C::Sync::sync();
  jacobi_eigh<M, C::NT, typename C::Sync, true, LD>(X, V, small, small + M, tid);
  int kmax = 0;
  double best = small[0];
  for (int k = 1; k < M; ++k)
    if (small[k] > best) {
      best = small[k];
      kmax = k;
    }
  for (int e = tid; e < DD; e += C::NT) K[e] = V[((e % D) * D + e / D) * LD + kmax];  // K[i][j] = v[j d + i]
  C::Sync::sync();
  if (tid < 32) {
    for (int e = tid; e < DD; e += 32) {
      const int a = e / D, c = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int k = 0; k < D; ++k) cfma(acc, cconj(K[k * D + a]), K[k * D + c]);
      if (a == c) acc.y = 0.0;
      P[e] = acc;
    }
    __syncwarp();
    jacobi_eigh<D, 32, SyncWarp, true>(P, W, pev, pscr, tid);
    for (int e = tid; e < DD; e += 32) {  // (K^dagger K)^{-1/2}
      const int a = e / D, c = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(W[a * D + k], rsqrt(pev[k])), W[c * D + k]);
      P[e] = acc;
    }
    __syncwarp();
    for (int e = tid; e < DD; e += 32) {
      const int i = e / D, j = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int k = 0; k < D; ++k) cfma(acc, K[i * D + k], P[k * D + j]);
      U[e] = acc;
    }
    __syncwarp();
    const cplx u00 = U[0];
    const double mag = sqrt(cabs2(u00));
    const cplx ph = mag > 0.0 ? cmake(u00.x / mag, -u00.y / mag) : cmake(1.0, 0.0);
    __syncwarp();
    for (int e = tid; e < DD; e += 32) U[e] = cmul(U[e], ph);
  }
  C::Sync::sync();
  cplx* dst = out + b * G::MM;
  for (int e = tid; e < G::MM; e += C::NT) {
    const int r = e / M, c = e % M;  // vec(U)[j d + i] = U[i][j]
    dst[e] = cmul(U[(r % D) * D + r / D], cconj(U[(c % D) * D + c / D]));
  }
}

// ---- CP for n = 1: one 4x4 matrix per THREAD, Jacobi entirely in registers --------------------------------
// A warp per 4x4 matrix (the generic kernel above) leaves 3/4 of the lanes idle and pays shared-memory
// latency on every rotation: 1.8 % of the HBM roofline (profiles/r01_bench_streaming_v1.json).  Here the
// Hermitian matrix is kept as diagonal d[4] + upper triangle o[r][c] (r < c), V as 16 complex registers, and
// the three round-robin steps (0,1)(2,3) | (0,2)(1,3) | (0,3)(1,2) are unrolled with compile-time indices.
template <int R, int C>
__device__ __forceinline__ cplx hget(const cplx (&o)[4][4]) {
  if constexpr (R < C) return o[R][C];
  else return cconj(o[C][R]);
}
template <int R, int C>
__device__ __forceinline__ void hset(cplx (&o)[4][4], cplx v) {
  if constexpr (R < C) o[R][C] = v;
  else o[C][R] = cconj(v);
}
template <int P, int Q, int K>
__device__ __forceinline__ void rot4_k(cplx (&o)[4][4], double c, cplx s) {
  if constexpr (K != P && K != Q) {
    const cplx a = hget<K, P>(o), b = hget<K, Q>(o);
    hset<K, P>(o, csub(cscale(a, c), cmul(cconj(s), b)));
    hset<K, Q>(o, cadd(cmul(s, a), cscale(b, c)));
  }
}
template <int P, int Q>
__device__ __forceinline__ void rot4(double (&d)[4], cplx (&o)[4][4], cplx (&v)[4][4]) {
  double c, an, gn;
  cplx s;
  jacobi_rotation(d[P], d[Q], o[P][Q], c, s, an, gn);
  d[P] = an;
  d[Q] = gn;
  o[P][Q] = cmake(0.0, 0.0);
  rot4_k<P, Q, 0>(o, c, s);
  rot4_k<P, Q, 1>(o, c, s);
  rot4_k<P, Q, 2>(o, c, s);
  rot4_k<P, Q, 3>(o, c, s);
  const cplx cs = cconj(s);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const cplx v0 = v[r][P], v1 = v[r][Q];
    v[r][P] = csub(cscale(v0, c), cmul(cs, v1));
    v[r][Q] = cadd(cmul(s, v0), cscale(v1, c));
  }
}

__global__ void __launch_bounds__(128) proj_cp4_thread_kernel(int64_t B, const cplx* __restrict__ in,
                                                              cplx* __restrict__ out) {
  __shared__ cplx tile[16 * QT_TS];  // element-major: tile[e * QT_TS + item] -> conflict-free per-thread access, padded stride for the transposing side
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * 128;
  const int nb = (int)min((int64_t)128, B - b0);
  for (int e = tid; e < nb * 16; e += 128) tile[(e % 16) * QT_TS + e / 16] = in[b0 * 16 + e];
  __syncthreads();
  if (tid < nb) {
    double d[4];
    cplx o[4][4], v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      d[r] = tile[(r * 4 + r) * QT_TS + tid].x;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[r][c] = cmake(r == c ? 1.0 : 0.0, 0.0);
        if (r < c) {  // Hermitian part (C + C^dagger) / 2, project_superoperators.py:30
          const cplx x = tile[(r * 4 + c) * QT_TS + tid], y = tile[(c * 4 + r) * QT_TS + tid];
          o[r][c] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
        }
      }
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
      double off = 0.0, tot = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        tot = fma(d[r], d[r], tot);
#pragma unroll
        for (int c = r + 1; c < 4; ++c) off += cabs2(o[r][c]);
      }
      off *= 2.0;
      tot += off;
      if (off <= 1e-26 * tot || tot == 0.0) break;  // relative off-diagonal norm 1e-13 (quadratic convergence: usually far below)
      rot4<0, 1>(d, o, v);
      rot4<2, 3>(d, o, v);
      rot4<0, 2>(d, o, v);
      rot4<1, 3>(d, o, v);
      rot4<0, 3>(d, o, v);
      rot4<1, 2>(d, o, v);
    }
    // OUT = V max(lambda, 0) V^dagger (upper triangle computed, lower mirrored)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = r; c < 4; ++c) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 4; ++k) cfma_conj(acc, cscale(v[r][k], fmax(d[k], 0.0)), v[c][k]);
        if (r == c) acc.y = 0.0;
        tile[(r * 4 + c) * QT_TS + tid] = acc;
        if (r != c) tile[(c * 4 + r) * QT_TS + tid] = cconj(acc);
      }
  }
  __syncthreads();
  for (int e = tid; e < nb * 16; e += 128) out[b0 * 16 + e] = tile[(e % 16) * QT_TS + e / 16];
}

// ---- TP / TNI: streaming kernel, several items per block for small n --------------------------------
template <int N, bool make_tp>
__global__ void proj_tp_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, int items_per_block) {
  constexpr int D = 1 << N, M = D * D, MM = M * M;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* tile = reinterpret_cast<cplx*>(smem_raw);         // [items][MM]
  cplx* Eall = tile + (size_t)items_per_block * MM;       // [items][D*D]
  cplx* scratch = Eall + (size_t)items_per_block * D * D; // TNI only: per item pt, P, W (3 D^2) + pev + jacobi
  const int64_t b0 = (int64_t)blockIdx.x * items_per_block;
  const int nb = (int)min((int64_t)items_per_block, B - b0);
  for (int e = threadIdx.x; e < nb * MM; e += blockDim.x) tile[e] = in[b0 * MM + e];
  __syncthreads();
  if constexpr (make_tp) {
    for (int w = threadIdx.x; w < nb * D * D; w += blockDim.x) {
      const int bi = w / (D * D), a = (w / D) % D, c = w % D;
      const cplx* Cm = tile + (size_t)bi * MM;
      cplx s = cmake(0.0, 0.0);
      for (int bb = 0; bb < D; ++bb) s = cadd(s, Cm[(a * D + bb) * M + c * D + bb]);
      if (a == c) s.x -= 1.0;
      Eall[w] = cscale(s, 1.0 / D);
    }
  } else if constexpr (D <= 4) {
    // n <= 2: one THREAD per item -- partial trace, d x d eigendecomposition (Jacobi unrolled in registers) and
    // the correction; a warp per 2x2 / 4x4 problem left this streaming kernel at 0.17 of the HBM roof
    for (int bi = threadIdx.x; bi < nb; bi += blockDim.x) {
      const cplx* Cm = tile + (size_t)bi * MM;
      cplx pt[4][4];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          cplx sacc = cmake(0.0, 0.0);
#pragma unroll
          for (int bb = 0; bb < D; ++bb) sacc = cadd(sacc, Cm[(a * D + bb) * M + c * D + bb]);
          pt[a][c] = sacc;
        }
      double dg[4];
      cplx o[4][4], v[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        dg[a] = (a < D) ? pt[a][a].x : 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          v[a][c] = cmake(a == c ? 1.0 : 0.0, 0.0);
          o[a][c] = cmake(0.0, 0.0);
          if (a < c && c < D) o[a][c] = cmake(0.5 * (pt[a][c].x + pt[c][a].x), 0.5 * (pt[a][c].y - pt[c][a].y));
        }
      }
      if constexpr (D == 2) {
        rot4<0, 1>(dg, o, v);  // a 2x2 Hermitian matrix is diagonalised by one rotation
      } else {
        for (int sweep = 0; sweep < 30; ++sweep) {
          double off = 0.0, tot = 0.0;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            tot = fma(dg[a], dg[a], tot);
#pragma unroll
            for (int c = a + 1; c < 4; ++c) off += cabs2(o[a][c]);
          }
          off *= 2.0;
          tot += off;
          if (off <= 1e-26 * tot || tot == 0.0) break;  // relative off-diagonal norm 1e-13 (quadratic convergence: usually far below)
          rot4<0, 1>(dg, o, v);
          rot4<2, 3>(dg, o, v);
          rot4<0, 2>(dg, o, v);
          rot4<1, 3>(dg, o, v);
          rot4<0, 3>(dg, o, v);
          rot4<1, 2>(dg, o, v);
        }
      }
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          cplx acc = cmake(0.0, 0.0);
#pragma unroll
          for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(v[a][k], fmin(dg[k], 1.0)), v[c][k]);
          Eall[bi * D * D + a * D + c] = cscale(csub(pt[a][c], acc), 1.0 / D);
        }
    }
  } else {
    // one warp per item computes pt, its eigen-decomposition and the correction
    constexpr int PER = 3 * D * D * 2 + D + JacobiScratch<D>::doubles + (D % 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int bi = warp; bi < nb; bi += nw) {
      double* base = reinterpret_cast<double*>(scratch) + (size_t)bi * ((PER + 1) / 2 * 2);
      cplx* pt = reinterpret_cast<cplx*>(base);
      cplx* P = pt + D * D;
      cplx* W = P + D * D;
      double* pev = reinterpret_cast<double*>(W + D * D);
      double* pscr = pev + D;
      const cplx* Cm = tile + (size_t)bi * MM;
      for (int e = lane; e < D * D; e += 32) {
        const int a = e / D, c = e % D;
        cplx s = cmake(0.0, 0.0);
        for (int bb = 0; bb < D; ++bb) s = cadd(s, Cm[(a * D + bb) * M + c * D + bb]);
        pt[e] = s;
      }
      __syncwarp();
      for (int e = lane; e < D * D; e += 32) {
        const int a = e / D, c = e % D;
        const cplx x = pt[e], y = pt[c * D + a];
        P[e] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
      }
      __syncwarp();
      jacobi_eigh<D, 32, SyncWarp, true>(P, W, pev, pscr, lane);
      for (int e = lane; e < D * D; e += 32) {
        const int a = e / D, c = e % D;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(W[a * D + k], fmin(pev[k], 1.0)), W[c * D + k]);
        Eall[bi * D * D + e] = cscale(csub(pt[e], acc), 1.0 / D);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * MM; e += blockDim.x) {
    const int bi = e / MM, r = (e / M) % M, c = e % M;
    cplx v = tile[e];
    if ((r % D) == (c % D)) v = csub(v, Eall[bi * D * D + (r / D) * D + (c / D)]);
    out[b0 * MM + e] = v;
  }
}

// ---- TNI in two passes (out-of-place calls) -----------------------------------------------------------
// The fused kernel above holds a tile of items in shared memory while ONE thread (n = 2) or warp (n = 3) per item
// runs the d x d eigendecomposition: the other threads idle and the tile pins the occupancy (0.18 / 0.32 of the HBM
// roof).  Here pass 1 reads only the elements the partial trace needs ((a b),(c b): a quarter (n = 3) to a half
// (n = 2) of the sectors), one thread / warp per item over the whole batch, and leaves the d x d correction E in the
// first d^2 elements of out[b]; pass 2 is a pure streaming kernel out = in - kron(E, I) that first lifts the E of
// its items into shared memory (it is about to overwrite them).
template <int N>
__global__ void __launch_bounds__(N <= 2 ? 128 : 256)
    tni_correction_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  constexpr int D = 1 << N, M = D * D, MM = M * M;
  if constexpr (N <= 2) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const cplx* Cm = in + b * MM;
    cplx pt[4][4];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        cplx sacc = cmake(0.0, 0.0);
#pragma unroll
        for (int bb = 0; bb < D; ++bb) sacc = cadd(sacc, Cm[(a * D + bb) * M + c * D + bb]);
        pt[a][c] = sacc;
      }
    double dg[4];
    cplx o[4][4], v[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      dg[a] = (a < D) ? pt[a][a].x : 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[a][c] = cmake(a == c ? 1.0 : 0.0, 0.0);
        o[a][c] = cmake(0.0, 0.0);
        if (a < c && c < D) o[a][c] = cmake(0.5 * (pt[a][c].x + pt[c][a].x), 0.5 * (pt[a][c].y - pt[c][a].y));
      }
    }
    if constexpr (D == 2) {
      rot4<0, 1>(dg, o, v);  // a 2x2 Hermitian matrix is diagonalised by one rotation
    } else {
      for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, tot = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          tot = fma(dg[a], dg[a], tot);
#pragma unroll
          for (int c = a + 1; c < 4; ++c) off += cabs2(o[a][c]);
        }
        off *= 2.0;
        tot += off;
        if (off <= 1e-26 * tot || tot == 0.0) break;
        rot4<0, 1>(dg, o, v);
        rot4<2, 3>(dg, o, v);
        rot4<0, 2>(dg, o, v);
        rot4<1, 3>(dg, o, v);
        rot4<0, 3>(dg, o, v);
        rot4<1, 2>(dg, o, v);
      }
    }
    cplx* E = out + b * MM;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(v[a][k], fmin(dg[k], 1.0)), v[c][k]);
        E[a * D + c] = cscale(csub(pt[a][c], acc), 1.0 / D);
      }
  } else {
    // one warp per item
    constexpr int PER = 3 * D * D * 2 + D + JacobiScratch<D>::doubles + (D % 2);
    constexpr int PERA = (PER + 1) / 2 * 2;
    __shared__ __align__(16) double scratch[8 * PERA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + warp;
    if (b >= B) return;
    double* base = scratch + (size_t)warp * PERA;
    cplx* pt = reinterpret_cast<cplx*>(base);
    cplx* P = pt + D * D;
    cplx* W = P + D * D;
    double* pev = reinterpret_cast<double*>(W + D * D);
    double* pscr = pev + D;
    const cplx* Cm = in + b * MM;
    for (int e = lane; e < D * D; e += 32) {
      const int a = e / D, c = e % D;
      cplx sacc = cmake(0.0, 0.0);
#pragma unroll
      for (int bb = 0; bb < D; ++bb) sacc = cadd(sacc, Cm[(a * D + bb) * M + c * D + bb]);
      pt[e] = sacc;
    }
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
      const int a = e / D, c = e % D;
      const cplx x = pt[e], y = pt[c * D + a];
      P[e] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
    }
    __syncwarp();
    jacobi_eigh<D, 32, SyncWarp, true>(P, W, pev, pscr, lane);
    cplx* E = out + b * MM;
    for (int e = lane; e < D * D; e += 32) {
      const int a = e / D, c = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(W[a * D + k], fmin(pev[k], 1.0)), W[c * D + k]);
      E[e] = cscale(csub(pt[e], acc), 1.0 / D);
    }
  }
}

// n = 1 in ONE pass: a 4 x 4 Choi matrix is 16 registers of one thread (128 items per block through the padded
// transposition tile), its 2 x 2 partial trace is diagonalised by a single rotation, and the corrected matrix goes
// straight back out: 512 bytes of traffic per item instead of the ~900 of the two-pass route (the correction parked in
// out[b] cost a quarter of a read and of a write on top of the two full reads).
__global__ void __launch_bounds__(128) tni_fused1_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out) {
  constexpr int D = 2, M = 4;
  __shared__ cplx tile[16 * QT_TS];
  const int tid = threadIdx.x;
  for (int64_t b0 = (int64_t)blockIdx.x * 128; b0 < B; b0 += (int64_t)gridDim.x * 128) {
    const int nb = (int)min((int64_t)128, B - b0);
    for (int e = tid; e < nb * 16; e += 128) tile[(e % 16) * QT_TS + e / 16] = in[b0 * 16 + e];
    __syncthreads();
    if (tid < nb) {
      cplx pt[4][4];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c)
          pt[a][c] = cadd(tile[((a * D) * M + c * D) * QT_TS + tid], tile[((a * D + 1) * M + c * D + 1) * QT_TS + tid]);
      double dg[4];
      cplx o[4][4], v[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        dg[a] = (a < D) ? pt[a][a].x : 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          v[a][c] = cmake(a == c ? 1.0 : 0.0, 0.0);
          o[a][c] = cmake(0.0, 0.0);
          if (a < c && c < D) o[a][c] = cmake(0.5 * (pt[a][c].x + pt[c][a].x), 0.5 * (pt[a][c].y - pt[c][a].y));
        }
      }
      rot4<0, 1>(dg, o, v);  // a 2x2 Hermitian matrix is diagonalised by one rotation
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          cplx acc = cmake(0.0, 0.0);
#pragma unroll
          for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(v[a][k], fmin(dg[k], 1.0)), v[c][k]);
          const cplx E = cscale(csub(pt[a][c], acc), 1.0 / D);
#pragma unroll
          for (int bb = 0; bb < D; ++bb) {  // out = in - kron(E, I): elements ((a bb), (c bb))
            const int idx = ((a * D + bb) * M + c * D + bb) * QT_TS + tid;
            tile[idx] = csub(tile[idx], E);
          }
        }
    }
    __syncthreads();
    for (int e = tid; e < nb * 16; e += 128) out[b0 * 16 + e] = tile[(e % 16) * QT_TS + e / 16];
    __syncthreads();
  }
}

template <int N>
__global__ void __launch_bounds__(256)
    tni_apply_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, int items_per_block) {
  constexpr int D = 1 << N, M = D * D, MM = M * M;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Esm = reinterpret_cast<cplx*>(smem_raw);  // [items][D*D]
  const int64_t b0 = (int64_t)blockIdx.x * items_per_block;
  const int nb = (int)min((int64_t)items_per_block, B - b0);
  for (int w = threadIdx.x; w < nb * D * D; w += blockDim.x) Esm[w] = out[(b0 + w / (D * D)) * MM + w % (D * D)];
  __syncthreads();
  for (int e = threadIdx.x; e < nb * MM; e += blockDim.x) {
    const int bi = e / MM, r = (e / M) % M, c = e % M;
    cplx v = in[b0 * MM + e];
    if ((r % D) == (c % D)) v = csub(v, Esm[bi * D * D + (r / D) * D + (c / D)]);
    out[b0 * MM + e] = v;
  }
}

// ---- physical (Dykstra): persistent groups, S = out[b], Q / CPREV in the workspace -------------------
template <int N>
__global__ void __launch_bounds__(ProjCfg<N>::NT * ProjCfg<N>::GPB)
    proj_physical_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, int make_tp,
                         cplx* __restrict__ workspace, int* __restrict__ eigh_calls, int* __restrict__ status_out,
                         double rel2) {
  using C = ProjCfg<N>;
  using G = typename C::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw + C::group_smem * gib);
  cplx* V = X + G::MP;
  cplx* T = V + G::MP;
  double* small = reinterpret_cast<double*>(T + G::MP);
  double* red = small + G::SMALL_DOUBLES - 64;
  const int64_t group = (int64_t)blockIdx.x * C::GPB + gib;
  const int64_t n_groups = (int64_t)gridDim.x * C::GPB;
  cplx* Q = workspace + group * 2 * G::MM;
  cplx* CPREV = Q + G::MM;
  for (int64_t b = group; b < B; b += n_groups) {
    const cplx* src = in + b * G::MM;
    cplx* S = out + b * G::MM;
    // S = Hermitian part of the input (all the iterates ever see, project_superoperators.py:30); the
    // anti-Hermitian part only enters the stopping rule (see ChoiGroup::project_physical)
    double anti2 = 0.0;
    for (int e = tid; e < G::MM; e += C::NT) {
      const int r = e / G::M, c = e % G::M;
      const cplx x = src[e], y = src[c * G::M + r];
      S[e] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
      anti2 += 0.25 * ((x.x - y.x) * (x.x - y.x) + (x.y + y.y) * (x.y + y.y));
    }
    anti2 = group_sum<C::NT, typename C::Sync>(anti2, red, tid);
    C::Sync::sync();
    bool v_valid = false;  // items are unrelated: the first decomposition of each starts cold
    int st = 0;
    int calls;
    if constexpr (N >= 3) {
      calls = G::project_physical_v2(S, CPREV, X, V, T, small, make_tp != 0, tid, v_valid, nullptr, rel2,
                                     QT_DYKSTRA_MAX_ITER, anti2 > 0.0 ? src : nullptr, &st);
    } else {
      calls = G::project_physical(S, Q, CPREV, X, V, T, small, make_tp != 0, tid, v_valid, nullptr, rel2,
                                  QT_DYKSTRA_MAX_ITER, anti2 > 0.0 ? src : nullptr, &st);
    }
    if (tid == 0 && eigh_calls) eigh_calls[b] = calls;
    if (tid == 0 && status_out) status_out[b] = st;
    C::Sync::sync();
  }
}


// ---- choi2kraus: eigh + sqrt(lambda) unvec(v) for |lambda| > tol ---------------------------------------
// operator_tools/superoperator_transformations.py:325-336.  np.linalg.eigh reads the LOWER triangle and
// returns ascending eigenvalues; the Kraus operators keep that order, compacted to the front of
// kraus_out[b] (count_out[b] of them, the rest zero-filled).  Negative eigenvalues give i*sqrt(|lambda|)
// (np.lib.scimath.sqrt).  Eigenvector phases are a gauge: compare through kraus2choi(choi2kraus(C)) == C.
template <int N>
__global__ void __launch_bounds__(ProjCfg<N>::NT * ProjCfg<N>::GPB)
    choi2kraus_kernel(int64_t B, const cplx* __restrict__ in, double tol, double* __restrict__ evals_out,
                      cplx* __restrict__ kraus_out, int* __restrict__ count_out, int skip_done) {
  using C = ProjCfg<N>;
  using G = typename C::G;
  constexpr int M = G::M, D = G::D, LD = G::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw + C::group_smem * gib);
  cplx* V = X + G::MP;
  double* small = reinterpret_cast<double*>(V + 2 * G::MP);
  double* ev = small;
  int* pos = reinterpret_cast<int*>(small + M + JacobiScratch<M>::doubles);  // [M] slot of eigenpair k, or -1
  int* rank = pos + M;                                                       // [M]
  const int64_t b = (int64_t)blockIdx.x * C::GPB + gib;
  if (b >= B) return;
  if (skip_done && count_out[b] >= 0) return;  // finished by the certified low-rank path (uniform over the group)
  const cplx* src = in + b * G::MM;
  for (int e = tid; e < G::MM; e += C::NT) {
    const int r = e / M, c = e % M;
    cplx v = (r >= c) ? src[r * M + c] : cconj(src[c * M + r]);
    if (r == c) v.y = 0.0;
    X[r * LD + c] = v;
  }
  C::Sync::sync();
  jacobi_eigh<M, C::NT, typename C::Sync, true, LD>(X, V, ev, small + M, tid);
  for (int k = tid; k < M; k += C::NT) {
    int rk = 0;
    for (int j = 0; j < M; ++j) rk += (ev[j] < ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
    rank[k] = rk;
  }
  C::Sync::sync();
  for (int k = tid; k < M; k += C::NT) {
    int before = 0;
    for (int j = 0; j < M; ++j) before += (rank[j] < rank[k] && fabs(ev[j]) > tol) ? 1 : 0;
    pos[k] = (fabs(ev[k]) > tol) ? before : -1;
    evals_out[b * M + rank[k]] = ev[k];
  }
  C::Sync::sync();
  int kept = 0;
  for (int k = 0; k < M; ++k) kept += (pos[k] >= 0) ? 1 : 0;
  if (tid == 0) count_out[b] = kept;
  cplx* dst = kraus_out + b * (int64_t)M * M;
  for (int e = tid; e < (M - kept) * M; e += C::NT) dst[kept * M + e] = cmake(0.0, 0.0);
  for (int e = tid; e < G::MM; e += C::NT) {
    const int k = e / M, r = e % M;  // eigenpair k, vec index r = j*D + i  ->  K[i][j]
    if (pos[k] < 0) continue;
    const double lam = ev[k];
    const double sq = sqrt(fabs(lam));
    const cplx v = V[r * LD + k];
    const cplx val = (lam >= 0.0) ? cscale(v, sq) : cmake(-sq * v.y, sq * v.x);
    dst[pos[k] * M + (r % D) * D + (r / D)] = val;
  }
}

template <int N>
static int launch_choi2kraus(int64_t B, const void* in, double tol, double* evals, void* kraus, int* count,
                             cudaStream_t st) {
  using C = ProjCfg<N>;
  const size_t smem = C::group_smem * C::GPB;
  int skip_done = 0;
  if constexpr (N == 3) {  // channels with <= 8 Kraus operators never reach the 64 x 64 Jacobi solver
    const int rc = qt_choi2kraus_lowrank64(B, in, tol, evals, kraus, count, st);
    if (rc) return rc;
    skip_done = 1;
  }
  QT_CUDA(cudaFuncSetAttribute(choi2kraus_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  choi2kraus_kernel<N><<<(unsigned)((B + C::GPB - 1) / C::GPB), C::NT * C::GPB, smem, st>>>(
      B, (const cplx*)in, tol, evals, (cplx*)kraus, count, skip_done);
  return qt_check_launch("choi2kraus_kernel");
}

template <int N>
static int64_t physical_grid(int64_t B) {
  using C = ProjCfg<N>;
  const int per_sm = (N >= 3) ? 1 : 8;
  const int64_t max_blocks = (int64_t)QT_NUM_SMS * per_sm;
  return std::min<int64_t>((B + C::GPB - 1) / C::GPB, max_blocks);
}

template <int N>
static int launch_cp(int64_t B, const void* in, void* out, cudaStream_t st) {
  using C = ProjCfg<N>;
  const size_t smem = C::group_smem * C::GPB;
  QT_CUDA(cudaFuncSetAttribute(proj_cp_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proj_cp_kernel<N><<<(unsigned)((B + C::GPB - 1) / C::GPB), C::NT * C::GPB, smem, st>>>(B, (const cplx*)in,
                                                                                         (cplx*)out);
  return qt_check_launch("proj_cp_kernel");
}

template <int N>
static int launch_unitary(int64_t B, const void* in, void* out, cudaStream_t st) {
  using C = ProjCfg<N>;
  const size_t smem = C::group_smem * C::GPB;
  QT_CUDA(cudaFuncSetAttribute(proj_unitary_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proj_unitary_kernel<N><<<(unsigned)((B + C::GPB - 1) / C::GPB), C::NT * C::GPB, smem, st>>>(B, (const cplx*)in,
                                                                                              (cplx*)out);
  return qt_check_launch("proj_unitary_kernel");
}

template <int N>
static int launch_tp(int64_t B, const void* in, void* out, int make_tp, cudaStream_t st) {
  constexpr int D = 1 << N, MM = D * D * D * D;
  // tile of items per block: QT_TP_TILE elements (16 KB) for the streaming TP pass -- several blocks per SM overlap their
  // load / partial-trace / store phases; the fused TNI keeps the larger tile (its per-item eigensolver wants threads)
  static const int tp_tile = []() {
    const char* e = getenv("QT_TP_TILE");
    return e ? atoi(e) : 1024;
  }();
  const int ipb = make_tp ? std::max(1, tp_tile / MM) : std::max(1, std::min(64, 4096 / MM));
  constexpr int PER = 3 * D * D * 2 + D + JacobiScratch<D>::doubles + (D % 2);
  const size_t smem = sizeof(cplx) * ((size_t)ipb * MM + (size_t)ipb * D * D) +
                      ((make_tp || D <= 4) ? 0 : sizeof(double) * (size_t)ipb * ((PER + 1) / 2 * 2));
  const unsigned blocks = (unsigned)((B + ipb - 1) / ipb);
  {
    if (!make_tp && N == 1) {  // one pass, in place or not
      const unsigned fb = (unsigned)std::min<int64_t>((B + 127) / 128, (int64_t)QT_NUM_SMS * 16);
      tni_fused1_kernel<<<fb, 128, 0, st>>>(B, (const cplx*)in, (cplx*)out);
      return qt_check_launch("tni_fused1_kernel");
    }
    if (!make_tp && in != out) {  // two-pass TNI (the correction is parked in out[b], so not for in-place calls)
      if (N <= 2)
        tni_correction_kernel<N><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, (const cplx*)in, (cplx*)out);
      else
        tni_correction_kernel<N><<<(unsigned)((B + 7) / 8), 256, 0, st>>>(B, (const cplx*)in, (cplx*)out);
      int rc = qt_check_launch("tni_correction_kernel");
      if (rc) return rc;
      const int ia = std::max(1, (N == 1 ? 1024 : 8192) / MM);
      tni_apply_kernel<N><<<(unsigned)((B + ia - 1) / ia), 256, sizeof(cplx) * ia * D * D, st>>>(B, (const cplx*)in,
                                                                                                  (cplx*)out, ia);
      return qt_check_launch("tni_apply_kernel");
    }
  }
  if (make_tp) {
    QT_CUDA(cudaFuncSetAttribute(proj_tp_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj_tp_kernel<N, true><<<blocks, 256, smem, st>>>(B, (const cplx*)in, (cplx*)out, ipb);
  } else {
    QT_CUDA(cudaFuncSetAttribute(proj_tp_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // n <= 2: the per-item eigendecomposition runs on one thread with ~200 registers; small blocks keep several
    // resident per SM
    proj_tp_kernel<N, false><<<blocks, (D <= 4) ? 64 : 256, smem, st>>>(B, (const cplx*)in, (cplx*)out, ipb);
  }
  return qt_check_launch("proj_tp_kernel");
}

// n = 4, 5: correction matrix by one block per item (qt_eigh_large.cu), then the streaming out = in - kron(E, I)
template <int N>
static int launch_tp_large(int64_t B, const void* in, void* out, int make_tp, cudaStream_t st) {
  constexpr int D = 1 << N;
  QT_REQUIRE(in != out, "TP / TNI projection at n >= 4 is out-of-place");
  int rc = qt_large_tp_correction(N, B, in, out, make_tp, st);
  if (rc) return rc;
  tni_apply_kernel<N><<<(unsigned)B, 256, sizeof(cplx) * D * D, st>>>(B, (const cplx*)in, (cplx*)out, 1);
  return qt_check_launch("tni_apply_kernel");
}

template <int N>
static int launch_physical(int64_t B, const void* in, void* out, int make_tp, double rel2, void* ws,
                           int64_t ws_bytes, int* eigh_calls, int* status, cudaStream_t st) {
  using C = ProjCfg<N>;
  const int64_t blocks = physical_grid<N>(B);
  const int64_t need = blocks * C::GPB * 2 * C::G::MM * (int64_t)sizeof(cplx);
  if (ws_bytes < need) {
    qt_set_error("qt_proj_physical_batch: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
                 (long long)need);
    return QT_ERR_WORKSPACE;
  }
  const size_t smem = C::group_smem * C::GPB;
  QT_CUDA(cudaFuncSetAttribute(proj_physical_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proj_physical_kernel<N><<<(unsigned)blocks, C::NT * C::GPB, smem, st>>>(B, (const cplx*)in, (cplx*)out, make_tp,
                                                                          (cplx*)ws, eigh_calls, status, rel2);
  return qt_check_launch("proj_physical_kernel");
}

#define DISPATCH_N3(n, CALL)                                              \
  switch (n) {                                                            \
    case 1: return CALL(1);                                               \
    case 2: return CALL(2);                                               \
    case 3: return CALL(3);                                               \
    default:                                                              \
      qt_set_error("this Choi operation supports n = 1..3 qubits (got %d)", n); \
      return QT_ERR_UNSUPPORTED;                                          \
  }

extern "C" int qt_proj_cp_ws_batch(int n, int64_t B, const void* choi, void* out, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  if (n == 4 || n == 5) {
    if (B == 0) return QT_OK;
    QT_REQUIRE(choi && out, "qt_proj_cp_ws_batch: null argument");
    return qt_large_proj_cp(n, B, choi, out, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return qt_proj_cp_batch(n, B, choi, out, stream);
}

extern "C" int qt_proj_cp_batch(int n, int64_t B, const void* choi, void* out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && out, "qt_proj_cp_batch: null argument");
  if (n == 4 || n == 5) {
    qt_set_error("qt_proj_cp_batch: n = %d needs a workspace: call qt_proj_cp_ws_batch", n);
    return QT_ERR_WORKSPACE;
  }
  if (n == 1) {
    proj_cp4_thread_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, (const cplx*)choi,
                                                                                          (cplx*)out);
    return qt_check_launch("proj_cp4_thread_kernel");
  }
#define CALL(N) launch_cp<N>(B, choi, out, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}

extern "C" int qt_proj_unitary_batch(int n, int64_t B, const void* choi, void* out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && out, "qt_proj_unitary_batch: null argument");
#define CALL(N) launch_unitary<N>(B, choi, out, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}

extern "C" int qt_proj_tp_batch(int n, int64_t B, const void* choi, void* out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && out, "qt_proj_tp_batch: null argument");
  if (n == 4) return launch_tp_large<4>(B, choi, out, 1, (cudaStream_t)stream);
  if (n == 5) return launch_tp_large<5>(B, choi, out, 1, (cudaStream_t)stream);
#define CALL(N) launch_tp<N>(B, choi, out, 1, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}

extern "C" int qt_proj_tni_batch(int n, int64_t B, const void* choi, void* out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && out, "qt_proj_tni_batch: null argument");
  if (n == 4) return launch_tp_large<4>(B, choi, out, 0, (cudaStream_t)stream);
  if (n == 5) return launch_tp_large<5>(B, choi, out, 0, (cudaStream_t)stream);
#define CALL(N) launch_tp<N>(B, choi, out, 0, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}

extern "C" int64_t qt_proj_physical_workspace_bytes(int n, int64_t B) {
  switch (n) {
    case 1: return physical_grid<1>(B) * ProjCfg<1>::GPB * 2 * ProjCfg<1>::G::MM * (int64_t)sizeof(cplx);
    case 2: return physical_grid<2>(B) * ProjCfg<2>::GPB * 2 * ProjCfg<2>::G::MM * (int64_t)sizeof(cplx);
    case 3: return physical_grid<3>(B) * ProjCfg<3>::GPB * 2 * ProjCfg<3>::G::MM * (int64_t)sizeof(cplx);
    case 4:
    case 5: return qt_large_physical_workspace_bytes(n, B);
    default: return -1;
  }
}

extern "C" int qt_proj_physical_batch(int n, int64_t B, const void* choi, void* out, int make_trace_preserving,
                                      double eigh_rel_tol, void* workspace, int64_t workspace_bytes,
                                      int32_t* eigh_calls_out, int32_t* status_out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && out && workspace, "qt_proj_physical_batch: null argument");
  QT_REQUIRE(choi != out, "qt_proj_physical_batch: in-place call not supported (the input is re-read by the stopping rule)");
  double rel2;
  if (qt_eigh_rel2_from_tol(eigh_rel_tol, n, &rel2, "qt_proj_physical_batch") != QT_OK) return QT_ERR_ARG;
  if (n == 4 || n == 5)  // one-sided Jacobi out of global memory: runs to its own (tight) convergence test
    return qt_large_proj_physical(n, B, choi, out, make_trace_preserving, workspace, workspace_bytes, eigh_calls_out,
                                  status_out, (cudaStream_t)stream);
#define CALL(N)                                                                                                  \
  launch_physical<N>(B, choi, out, make_trace_preserving, rel2, workspace, workspace_bytes, eigh_calls_out, \
                     status_out, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}

extern "C" int qt_choi2kraus_batch(int n, int64_t B, const void* choi, double tol, double* evals_out,
                                   void* kraus_out, int32_t* count_out, void* stream) {
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && evals_out && kraus_out && count_out, "qt_choi2kraus_batch: null argument");
#define CALL(N) launch_choi2kraus<N>(B, choi, tol, evals_out, kraus_out, count_out, (cudaStream_t)stream)
  DISPATCH_N3(n, CALL)
#undef CALL
}
