// choi2kraus for n = 4, 5 (operator_tools/superoperator_transformations.py:325-336): the 256 x 256 / 1024 x 1024
// Hermitian eigenproblem does not fit shared memory (1 MB / 16 MB per matrix), so the shared-memory ring solver of
// qt_eigh.cuh cannot be used.  One 1024-thread block per matrix runs a ONE-SIDED (Hestenes) Jacobi out of an
// L2-resident workspace:
//   U = (A + c I),  c = ||A||_inf >= spectral radius, so U is PSD and its singular values lambda + c are distinct
//                   exactly when the eigenvalues are (a plain one-sided sweep on an indefinite matrix cannot separate
//                   +lambda from -lambda);
//   each step orthogonalises M/2 disjoint column pairs (round-robin order): a warp forms the 2 x 2 Gram matrix of its
//   pair with coalesced loads + shuffles, and applies the rotation that diagonalises it to the two columns of U.
//   Pairs of one step touch disjoint columns: one __syncthreads per step, nothing else.
// Converged when no pair needed a rotation (|u_p . u_q|^2 <= 1e-29 |u_p|^2 |u_q|^2).  The accumulated rotations V are
// never formed: at convergence U = (A + c I) V has orthogonal columns, so V diagonalises (A + c I)^2 and
// u_k = (lambda_k + c) v_k, i.e.  lambda_k = |u_k| - c  and  v_k = u_k / |u_k|  (c is taken 1/16 above ||A||_inf so
// that |u_k| >= c / 16 > 0).  The Kraus operators are then written exactly like choi2kraus_kernel does for n <= 3.
// Columns are stored as contiguous ROWS of the workspace array (Ut[k][r] = U[r][k]).
#include "qt_choi.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

template <int M>
struct LargeCfg {
  static constexpr int NT = (M == 256) ? 512 : 1024, NW = NT / 32, HP = M / 2, PER = M / 32;
  static constexpr int D = (M == 256) ? 16 : 32;
  // M = 256 (512 threads, 128 registers): the pair's two columns stay in registers between the Gram pass and the update;
  // M = 1024 re-reads them
  static constexpr bool HOLD = (M == 256);
};

template <int M>
struct LargeShared {
  double ev[M], sig[M];
  double red[LargeCfg<M>::NW];
  int rotated;
};

// Eigendecomposition of the M x M Hermitian matrix herm(r, c) by the whole block.  U: M x M workspace (global / L2),
// column k stored as the contiguous row U[k * M ..].  On exit sh.ev[k] are the eigenvalues (unordered), and
// v_k = U[k, :] / sh.sig[k] the eigenvectors.  Returns the number of sweeps.
template <int M, class HermFn>
__device__ int large_eigh(HermFn herm, cplx* __restrict__ U, LargeShared<M>& sh) {
  using C = LargeCfg<M>;
  constexpr int NT = C::NT, NW = C::NW, HP = C::HP, PER = C::PER;
  constexpr bool HOLD = C::HOLD;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double rowmax = 0.0;
  for (int r = wid; r < M; r += NW) {
    double s = 0.0;
    for (int c = lane; c < M; c += 32) s += sqrt(cabs2(herm(r, c)));
    s = warp_sum(s);
    rowmax = fmax(rowmax, s);
  }
  __syncthreads();
  if (lane == 0) sh.red[wid] = rowmax;
  __syncthreads();
  double shift = 0.0;
  for (int w = 0; w < NW; ++w) shift = fmax(shift, sh.red[w]);
  shift = 1.0625 * shift + 1e-300;
  __syncthreads();
  for (int e = tid; e < M * M; e += NT) {
    const int k = e / M, r = e % M;  // column k, row r:  U[r][k] = conj(herm(k, r))
    cplx v = cconj(herm(k, r));
    if (k == r) v.x += shift;
    U[e] = v;
  }
  __syncthreads();
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    if (tid == 0) sh.rotated = 0;
    __syncthreads();
    for (int step = 0; step < M - 1; ++step) {
      for (int i = wid; i < HP; i += NW) {
        int p, q;
        rr_pair(M, step, i, p, q);
        cplx* up = U + (size_t)p * M;
        cplx* uq = U + (size_t)q * M;
        cplx a[HOLD ? PER : 1], c2[HOLD ? PER : 1];
        double alpha = 0.0, gamma = 0.0;
        cplx beta = cmake(0.0, 0.0);
#pragma unroll
        for (int t = 0; t < PER; ++t) {
          const cplx x = up[lane + 32 * t], y = uq[lane + 32 * t];
          if (HOLD) {
            a[t] = x;
            c2[t] = y;
          }
          alpha += cabs2(x);
          gamma += cabs2(y);
          cfma(beta, cconj(x), y);  // beta = u_p^dagger u_q
        }
        alpha = warp_sum(alpha);
        gamma = warp_sum(gamma);
        beta.x = warp_sum(beta.x);
        beta.y = warp_sum(beta.y);
        if (cabs2(beta) <= 1e-29 * alpha * gamma) continue;  // warp-uniform
        double c, an, gn;
        cplx s;
        jacobi_rotation(alpha, gamma, beta, c, s, an, gn);
        const cplx cs = cconj(s);
        if (lane == 0) sh.rotated = 1;
#pragma unroll
        for (int t = 0; t < PER; ++t) {
          const cplx x = HOLD ? a[t] : up[lane + 32 * t], y = HOLD ? c2[t] : uq[lane + 32 * t];
          up[lane + 32 * t] = csub(cscale(x, c), cmul(cs, y));
          uq[lane + 32 * t] = cadd(cmul(s, x), cscale(y, c));
        }
      }
      __syncthreads();
    }
    if (!sh.rotated) break;
    __syncthreads();
  }
  // eigenvalues: column norms of the shifted matrix minus the shift
  for (int k = wid; k < M; k += NW) {
    double acc = 0.0;
    for (int r = lane; r < M; r += 32) acc += cabs2(U[(size_t)k * M + r]);
    acc = warp_sum(acc);
    if (lane == 0) {
      sh.sig[k] = sqrt(acc);
      sh.ev[k] = sh.sig[k] - shift;
    }
  }
  __syncthreads();
  return sweep;
}

template <int M>
__global__ void __launch_bounds__(M == 256 ? 512 : 1024)
    choi2kraus_large_kernel(int64_t B, const cplx* __restrict__ in, double tol, double* __restrict__ evals_out,
                            cplx* __restrict__ kraus_out, int* __restrict__ count_out, cplx* __restrict__ ws,
                            int* __restrict__ sweeps_out, int skip_done) {
  using C = LargeCfg<M>;
  constexpr int NT = C::NT, D = C::D;
  __shared__ LargeShared<M> sh;
  __shared__ int rank[M], pos[M];
  double* ev = sh.ev;
  double* sig = sh.sig;
  const int tid = threadIdx.x;
  cplx* U = ws + (size_t)blockIdx.x * M * M;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    if (skip_done && count_out[b] >= 0) continue;  // finished by choi2kraus_lowrank_kernel (block-uniform)
    const cplx* src = in + b * (int64_t)M * M;
    // Hermitian matrix np.linalg.eigh sees (lower triangle)
    auto herm = [&](int r, int c) {
      cplx v = (r >= c) ? src[(size_t)r * M + c] : cconj(src[(size_t)c * M + r]);
      if (r == c) v.y = 0.0;
      return v;
    };
    const int sweep = large_eigh<M>(herm, U, sh);
    if (tid == 0 && sweeps_out) sweeps_out[b] = sweep;
    for (int k = tid; k < M; k += NT) {
      int rk = 0;
      for (int j = 0; j < M; ++j) rk += (ev[j] < ev[k] || (ev[j] == ev[k] && j < k)) ? 1 : 0;
      rank[k] = rk;
    }
    __syncthreads();
    for (int k = tid; k < M; k += NT) {
      int before = 0;
      for (int j = 0; j < M; ++j) before += (rank[j] < rank[k] && fabs(ev[j]) > tol) ? 1 : 0;
      pos[k] = (fabs(ev[k]) > tol) ? before : -1;
      evals_out[b * M + rank[k]] = ev[k];
    }
    __syncthreads();
    int kept = 0;
    for (int k = 0; k < M; ++k) kept += (pos[k] >= 0) ? 1 : 0;
    if (tid == 0) count_out[b] = kept;
    cplx* dst = kraus_out + b * (int64_t)M * M;
    for (int e = tid; e < (M - kept) * M; e += NT) dst[(size_t)kept * M + e] = cmake(0.0, 0.0);
    for (int e = tid; e < M * M; e += NT) {
      const int k = e / M, r = e % M;  // eigenpair k, vec index r = j*D + i  ->  K[i][j]
      if (pos[k] < 0) continue;
      const double lam = ev[k];
      const double sq = sqrt(fabs(lam)) / sig[k];  // v_k = u_k / |u_k|
      const cplx v = U[e];
      const cplx val = (lam >= 0.0) ? cscale(v, sq) : cmake(-sq * v.y, sq * v.x);
      dst[(size_t)pos[k] * M + (r % D) * D + (r / D)] = val;
    }
    __syncthreads();
  }
}

// =============================================================================================
// choi2kraus, n = 4, 5: certified low-rank fast path.
// The Choi matrix of a channel with a handful of Kraus operators has that many non-zero eigenvalues, and choi2kraus only
// returns eigenpairs with |lambda| > tol (superoperator_transformations.py:334-336).  One block per matrix, one thread
// per row, three passes over the matrix (np.linalg.eigh's view of it: the lower triangle):
//   1. Y = A Omega           (Omega: M x 8 pseudo-random probe; rows of Y stay in registers, ||A||_F on the way)
//      Q = orth(Y)           (modified Gram-Schmidt, twice, block reductions)
//   2. Z = A Q,  B = Q^dagger Z  (8 x 8),  B = W diag(mu) W^dagger  (warp Jacobi)
//   3. residual = || A - Q B Q^dagger ||_F, element by element (no cancellation).
// If residual <= 0.01 tol, Weyl's inequality bounds every eigenvalue outside the captured ones by 0.01 tol -- they are
// below choi2kraus's threshold -- and the captured ones are exact to the same margin: the Kraus operators are
// sqrt(mu_k) unvec(Q w_k), written exactly like the general kernel does (ascending order, i sqrt(|mu|) for mu < 0; the
// eigenvalues reported for the discarded directions are 0).  Otherwise count_out[b] = -1 and the general one-sided
// Jacobi kernel below handles the matrix: full-rank inputs cost one wasted scan (~1 % of the solver's time).
// =============================================================================================
__device__ __forceinline__ double lr_rand(unsigned a, unsigned b) {  // deterministic hash -> (-1, 1)
  unsigned long long z = ((unsigned long long)a << 32 | b) + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;  // 53 bits -> [0, 2) - 1
}

template <int M>
__global__ void __launch_bounds__(M <= 256 ? M : 1024)
    choi2kraus_lowrank_kernel(int64_t B, const cplx* __restrict__ in, double tol, double* __restrict__ evals_out,
                              cplx* __restrict__ kraus_out, int* __restrict__ count_out) {
  constexpr int KC = 8, NT = M, NW = NT / 32, D = (M == 64) ? 8 : (M == 256) ? 16 : 32;
  static_assert(M == 64 || M == 256 || M == 1024, "one thread per row, whole warps");
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* Wm = reinterpret_cast<cplx*>(raw);            // [M][KC]: Omega, then Q
  cplx* Bs = Wm + (size_t)M * KC;                     // [KC][KC]
  cplx* Ws = Bs + KC * KC;                            // [KC][KC] eigenvectors of B
  double* mu = reinterpret_cast<double*>(Ws + KC * KC);  // [KC] + Jacobi scratch
  double* red = mu + KC + JacobiScratch<KC>::doubles;    // [NW][2]
  __shared__ int order[KC], posk[KC], kept_s, ok_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, r = tid;
  auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) {
      red[2 * wid] = a;
      red[2 * wid + 1] = b;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < NW; ++w) {
      sa += red[2 * w];
      sb += red[2 * w + 1];
    }
    oa = sa;
    ob = sb;
  };
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* A = in + b * (int64_t)M * M;
    auto a_of = [&](int c) {  // row r of the Hermitian matrix np.linalg.eigh sees
      cplx v = (c <= r) ? A[(size_t)r * M + c] : cconj(A[(size_t)c * M + r]);
      if (c == r) v.y = 0.0;
      return v;
    };
    auto apply = [&](cplx (&acc)[KC], double* norm2) {  // acc = (A W)[r, :]
#pragma unroll
      for (int j = 0; j < KC; ++j) acc[j] = cmake(0.0, 0.0);
      double n2 = 0.0;
      for (int c = 0; c < M; ++c) {
        const cplx a = a_of(c);
        n2 += cabs2(a);
#pragma unroll
        for (int j = 0; j < KC; ++j) cfma(acc[j], a, Wm[c * KC + j]);
      }
      if (norm2) *norm2 = n2;
    };
    __syncthreads();
#pragma unroll
    for (int j = 0; j < KC; ++j) Wm[r * KC + j] = cmake(lr_rand(r, 2 * j), lr_rand(r, 2 * j + 1));
    __syncthreads();
    cplx y[KC];
    double n2row;
    apply(y, &n2row);
    double normA2, dummy;
    block_sum2(n2row, 0.0, normA2, dummy);
    // Q = orth(Y): modified Gram-Schmidt, two rounds (rows in registers, columns reduced over the block)
    for (int round = 0; round < 2; ++round) {
#pragma unroll
      for (int j = 0; j < KC; ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) {
          double dr, di;  // <q_i, y_j>
          block_sum2(y[i].x * y[j].x + y[i].y * y[j].y, y[i].x * y[j].y - y[i].y * y[j].x, dr, di);
          y[j] = csub(y[j], cmul(cmake(dr, di), y[i]));
        }
        double nn, d2;
        block_sum2(cabs2(y[j]), 0.0, nn, d2);
        const double inv = nn > 0.0 ? rsqrt(nn) : 0.0;
        y[j] = cscale(y[j], inv);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < KC; ++j) Wm[r * KC + j] = y[j];  // Q
    __syncthreads();
    cplx z[KC];
    apply(z, nullptr);
    // B = Q^dagger Z (Hermitian part)
    for (int i = 0; i < KC; ++i)
      for (int j = i; j < KC; ++j) {
        double br, bi;
        block_sum2(y[i].x * z[j].x + y[i].y * z[j].y, y[i].x * z[j].y - y[i].y * z[j].x, br, bi);
        if (tid == 0) {
          Bs[i * KC + j] = cmake(br, i == j ? 0.0 : bi);
          Bs[j * KC + i] = cmake(br, i == j ? 0.0 : -bi);
        }
      }
    __syncthreads();
    // residual || A - Q B Q^dagger ||_F^2 :  (Q B)[r, :] in registers
    cplx qb[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) {
      cplx acc = cmake(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < KC; ++i) cfma(acc, y[i], Bs[i * KC + j]);
      qb[j] = acc;
    }
    double res2 = 0.0;
    for (int c = 0; c < M; ++c) {
      cplx pred = cmake(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < KC; ++j) cfma_conj(pred, qb[j], Wm[c * KC + j]);
      res2 += cabs2(csub(a_of(c), pred));
    }
    double resid2, d3;
    block_sum2(res2, 0.0, resid2, d3);
    const bool certified = resid2 <= (0.01 * tol) * (0.01 * tol);
    if (!certified) {
      if (tid == 0) count_out[b] = -1;  // the general kernel takes this matrix
      continue;
    }
    // eigen-decomposition of the captured 8 x 8 block (warp 0), ascending order
    if (tid < 32) {
      jacobi_eigh<KC, 32, SyncWarp, true>(Bs, Ws, mu, mu + KC, lane);
      if (lane < KC) {
        int rk = 0;
        for (int j = 0; j < KC; ++j) rk += (mu[j] < mu[lane] || (mu[j] == mu[lane] && j < lane)) ? 1 : 0;
        order[rk] = lane;
      }
      __syncwarp();
      if (lane == 0) {
        int kept = 0;
        for (int q = 0; q < KC; ++q) {
          const int k = order[q];
          posk[k] = (fabs(mu[k]) > tol) ? kept++ : -1;
        }
        kept_s = kept;
      }
    }
    __syncthreads();
    const int kept = kept_s;
    // all M eigenvalues, ascending: negative captured ones, the M - KC discarded directions (reported as 0), the rest
    {
      int nneg = 0;
      for (int q = 0; q < KC; ++q) nneg += (mu[order[q]] < 0.0) ? 1 : 0;
      double v = 0.0;
      if (r < nneg) v = mu[order[r]];
      else if (r >= M - (KC - nneg)) v = mu[order[r - (M - KC)]];
      evals_out[b * M + r] = v;
    }
    if (tid == 0) count_out[b] = kept;
    cplx* dst = kraus_out + b * (int64_t)M * M;
    for (size_t e = (size_t)kept * M + tid; e < (size_t)M * M; e += NT) dst[e] = cmake(0.0, 0.0);
    for (int k = 0; k < KC; ++k) {
      if (posk[k] < 0) continue;
      cplx v = cmake(0.0, 0.0);  // (Q w_k)[r]
#pragma unroll
      for (int j = 0; j < KC; ++j) cfma(v, y[j], Ws[j * KC + k]);
      const double lam = mu[k], sq = sqrt(fabs(lam));
      const cplx val = (lam >= 0.0) ? cscale(v, sq) : cmake(-sq * v.y, sq * v.x);
      dst[(size_t)posk[k] * M + (r % D) * D + (r / D)] = val;  // vec index r = j*D + i  ->  K[i][j]
    }
    __syncthreads();
  }
}

// =============================================================================================
// Choi-matrix projections for n = 4, 5 (operator_tools/project_superoperators.py:19-144).  The reference functions are
// size-agnostic; the shared-memory kernels of qt_project.cu stop at 64 x 64.  Same algorithms here with every matrix in
// global memory (L2-resident per block) and large_eigh as the eigensolver: correct, and bound by that solver
// (~30 ms per 256 x 256 decomposition per SM) -- the reference's LAPACK call takes ~15 ms on one core at n = 4.
// =============================================================================================
// OUT = V max(ev, 0) V^dagger from the solver's output (v_k = U[k, :] / sig[k]); or X - V min(ev, 0) V^dagger when fewer
// eigenvalues are negative.  x(r, c) returns the decomposed matrix.
template <int M, class XFn>
__device__ void large_recompose_psd(cplx* __restrict__ OUT, const cplx* __restrict__ U, const LargeShared<M>& sh, XFn x) {
  constexpr int NT = LargeCfg<M>::NT;
  const int tid = threadIdx.x;
  int npos = 0;
  for (int k = 0; k < M; ++k) npos += (sh.ev[k] > 0.0) ? 1 : 0;
  const bool use_pos = npos <= M / 2;
  // thread = (column c, group of 4 rows): U[k][c] is read once per k for 4 outputs; lanes walk c (coalesced)
  for (int w = tid; w < (M / 4) * M; w += NT) {
    const int c = w % M, r0 = (w / M) * 4;
    cplx acc[4] = {cmake(0.0, 0.0), cmake(0.0, 0.0), cmake(0.0, 0.0), cmake(0.0, 0.0)};
    for (int k = 0; k < M; ++k) {
      const double lam = sh.ev[k];
      const double wk = use_pos ? fmax(lam, 0.0) : fmax(-lam, 0.0);
      if (wk == 0.0) continue;
      const double wgt = wk / (sh.sig[k] * sh.sig[k]);
      const cplx uc = U[(size_t)k * M + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) cfma_conj(acc[i], cscale(U[(size_t)k * M + r0 + i], wgt), uc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) OUT[(size_t)(r0 + i) * M + c] = use_pos ? acc[i] : cadd(x(r0 + i, c), acc[i]);
  }
  __syncthreads();
}

template <int M>
__global__ void __launch_bounds__(M == 256 ? 512 : 1024)
    proj_cp_large_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, cplx* __restrict__ ws) {
  __shared__ LargeShared<M> sh;
  cplx* U = ws + (size_t)blockIdx.x * M * M;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* src = in + b * (int64_t)M * M;
    auto herm = [&](int r, int c) {  // (C + C^dagger) / 2, project_superoperators.py:30
      const cplx x = src[(size_t)r * M + c], y = src[(size_t)c * M + r];
      return cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
    };
    large_eigh<M>(herm, U, sh);
    large_recompose_psd<M>(out + b * (int64_t)M * M, U, sh, herm);
  }
}

// TP / TNI correction matrix of one item (block per item): E = (Tr_out C - I)/d  or  (Tr_out C - clamp_le1)/d, parked in
// the first d^2 elements of out[b] for tni_apply_kernel (qt_project.cu) -- the n <= 3 two-pass protocol.
template <int N, bool MAKE_TP>
__global__ void __launch_bounds__(256) tp_correction_large_kernel(int64_t B, const cplx* __restrict__ in,
                                                                  cplx* __restrict__ out) {
  using G = ChoiGroup<N, 256, SyncBlock>;
  constexpr int D = G::D, M = G::M;
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* pt = reinterpret_cast<cplx*>(raw);
  cplx* E = pt + D * D;
  cplx* P = E + D * D;
  cplx* W = P + D * D;
  double* pev = reinterpret_cast<double*>(W + D * D);
  double* pscr = pev + D;
  const int tid = threadIdx.x;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    G::partial_trace_out(in + b * (int64_t)M * M, M, pt, tid);
    G::tp_correction(pt, E, MAKE_TP, P, W, pev, pscr, tid);
    cplx* dst = out + b * (int64_t)M * M;
    for (int e = tid; e < D * D; e += 256) dst[e] = E[e];
    __syncthreads();
  }
}

// Dykstra (project_superoperators.py:87-144) with every matrix in global memory.  Per block: Q, CPREV, X (the CP
// projection), U (solver workspace); S = out[b].  Same bookkeeping as ChoiGroup::project_physical.
template <int N>
__global__ void __launch_bounds__(N == 4 ? 512 : 1024)
    proj_physical_large_kernel(int64_t B, const cplx* __restrict__ in, cplx* __restrict__ out, int make_tp,
                               cplx* __restrict__ ws, int* __restrict__ eigh_calls, int* __restrict__ status_out) {
  constexpr int M = 1 << (2 * N), D = 1 << N, NT = LargeCfg<M>::NT;
  constexpr size_t MM = (size_t)M * M;
  using G = ChoiGroup<N, NT, SyncBlock>;
  __shared__ LargeShared<M> sh;
  extern __shared__ __align__(16) unsigned char raw[];
  cplx* E = reinterpret_cast<cplx*>(raw);
  cplx* En = E + D * D;
  cplx* ptS = En + D * D;
  cplx* ptC = ptS + D * D;
  cplx* P = ptC + D * D;
  cplx* W = P + D * D;
  double* pev = reinterpret_cast<double*>(W + D * D);
  double* pscr = pev + D;
  double* red = pscr + JacobiScratch<D>::doubles;
  const int tid = threadIdx.x;
  cplx* Q = ws + (size_t)blockIdx.x * 4 * MM;
  cplx* CPREV = Q + MM;
  cplx* X = CPREV + MM;
  cplx* U = X + MM;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* src = in + b * MM;
    cplx* S = out + b * MM;
    double anti2 = 0.0;
    for (size_t e = tid; e < MM; e += NT) {
      const size_t r = e / M, c = e % M;
      const cplx x = src[e], y = src[c * M + r];
      S[e] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
      anti2 += 0.25 * ((x.x - y.x) * (x.x - y.x) + (x.y + y.y) * (x.y + y.y));
      Q[e] = cmake(0.0, 0.0);
      CPREV[e] = cmake(0.0, 0.0);
    }
    anti2 = group_sum<NT, SyncBlock>(anti2, red, tid);
    const bool raw_in = anti2 > 0.0;
    for (int e = tid; e < D * D; e += NT) E[e] = cmake(0.0, 0.0);
    __threadfence_block();
    __syncthreads();
    int n_eigh = 0, st = 0;
    while (true) {
      auto pre_cp = [&](int r, int c) { return csub(S[(size_t)r * M + c], Q[(size_t)r * M + c]); };  // Hermitian
      const int sw = large_eigh<M>(pre_cp, U, sh);
      if (sw >= 60) st |= 2;
      ++n_eigh;
      large_recompose_psd<M>(X, U, sh, pre_cp);
      __threadfence_block();
      __syncthreads();
      double n_dcp = 0.0;
      cplx ip_q = cmake(0.0, 0.0);
      for (size_t e = tid; e < MM; e += NT) {
        const cplx cp = X[e], s = S[e], q = Q[e], cprev = CPREV[e];
        const cplx d1 = csub(cp, s);
        n_dcp += cabs2(d1);
        cfma_conj(ip_q, csub(cp, cprev), q);
        if (raw_in) {
          const cplx x = src[e], y = src[(e % M) * M + e / M];
          const cplx a = cmake(0.5 * (x.x - y.x), 0.5 * (x.y + y.y));
          if (n_eigh == 1) n_dcp += cabs2(a);
          else cfma_conj(ip_q, csub(cprev, cp), a);
        }
        Q[e] = cadd(d1, q);
        CPREV[e] = cp;
      }
      G::partial_trace_out(S, M, ptS, tid);
      G::partial_trace_out(X, M, ptC, tid);
      for (int e = tid; e < D * D; e += NT) ptC[e] = cadd(ptC[e], cscale(E[e], (double)D));
      __syncthreads();
      G::tp_correction(ptC, En, make_tp != 0, P, W, pev, pscr, tid);
      for (size_t e = tid; e < MM; e += NT) {
        const int r = (int)(e / M), c = (int)(e % M);
        cplx v = X[e];
        if ((r % D) == (c % D)) v = cadd(v, csub(E[(r / D) * D + c / D], En[(r / D) * D + c / D]));
        S[e] = v;
      }
      double n_dtp = 0.0;
      cplx ip_t = cmake(0.0, 0.0);
      for (int e = tid; e < D * D; e += NT) {
        n_dtp += cabs2(csub(En[e], E[e]));
        const cplx dpt = csub(csub(ptC[e], cscale(En[e], (double)D)), ptS[e]);
        cfma_conj(ip_t, dpt, E[e]);
      }
      n_dcp = group_sum<NT, SyncBlock>(n_dcp, red, tid);
      n_dtp = group_sum<NT, SyncBlock>(n_dtp, red, tid);
      ip_q.x = group_sum<NT, SyncBlock>(ip_q.x, red, tid);
      ip_q.y = group_sum<NT, SyncBlock>(ip_q.y, red, tid);
      ip_t.x = group_sum<NT, SyncBlock>(ip_t.x, red, tid);
      ip_t.y = group_sum<NT, SyncBlock>(ip_t.y, red, tid);
      const double crit = n_dcp + D * n_dtp + 2.0 * sqrt(cabs2(ip_t)) + 2.0 * sqrt(cabs2(ip_q));
      __threadfence_block();
      __syncthreads();
      if (crit < 1e-4) break;
      if (n_eigh >= QT_DYKSTRA_MAX_ITER) {
        st |= 1;
        break;
      }
      for (int e = tid; e < D * D; e += NT) E[e] = En[e];
      __syncthreads();
    }
    if (tid == 0 && eigh_calls) eigh_calls[b] = n_eigh;
    if (tid == 0 && status_out) status_out[b] = st;
    __syncthreads();
  }
}

static int64_t large_grid(int64_t B) { return std::min<int64_t>(B, QT_NUM_SMS); }

extern "C" int64_t qt_choi2kraus_large_workspace_bytes(int n, int64_t B) {
  if (n != 4 && n != 5) return -1;
  const int64_t M = 1LL << (2 * n);
  return large_grid(B) * M * M * (int64_t)sizeof(cplx);
}

extern "C" int qt_choi2kraus_large_batch(int n, int64_t B, const void* choi, double tol, double* evals_out,
                                         void* kraus_out, int32_t* count_out, void* workspace, int64_t workspace_bytes,
                                         int32_t* sweeps_out, void* stream) {
  QT_REQUIRE(n == 4 || n == 5, "qt_choi2kraus_large_batch: n=%d (use qt_choi2kraus_batch for n <= 3)", n);
  if (B == 0) return QT_OK;
  QT_REQUIRE(choi && evals_out && kraus_out && count_out && workspace, "qt_choi2kraus_large_batch: null argument");
  QT_REQUIRE(workspace_bytes >= qt_choi2kraus_large_workspace_bytes(n, B),
             "qt_choi2kraus_large_batch: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
             (long long)qt_choi2kraus_large_workspace_bytes(n, B));
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)large_grid(B);
  if (sweeps_out) QT_CUDA(cudaMemsetAsync(sweeps_out, 0, sizeof(int32_t) * B, st));
  // certified low-rank fast path first (count_out[b] = -1 where it does not apply), then the general solver
  const int64_t M = 1LL << (2 * n);
  const size_t smem = sizeof(cplx) * (M * 8 + 2 * 64) + sizeof(double) * (8 + JacobiScratch<8>::doubles + 2 * 32 + 8);
  const unsigned fast_grid = (unsigned)std::min<int64_t>(B, (int64_t)QT_NUM_SMS * (n == 4 ? 4 : 1));
  if (n == 4) {
    QT_CUDA(cudaFuncSetAttribute(choi2kraus_lowrank_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    choi2kraus_lowrank_kernel<256><<<fast_grid, 256, smem, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                               count_out);
  } else {
    QT_CUDA(cudaFuncSetAttribute(choi2kraus_lowrank_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    choi2kraus_lowrank_kernel<1024><<<fast_grid, 1024, smem, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                                 count_out);
  }
  int rc = qt_check_launch("choi2kraus_lowrank_kernel");
  if (rc) return rc;
  if (n == 4)
    choi2kraus_large_kernel<256><<<grid, 512, 0, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                         count_out, (cplx*)workspace, sweeps_out, 1);
  else
    choi2kraus_large_kernel<1024><<<grid, 1024, 0, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                          count_out, (cplx*)workspace, sweeps_out, 1);
  return qt_check_launch("choi2kraus_large_kernel");
}

// n = 3 (64 x 64): the same certified low-rank path in front of the shared-memory Jacobi kernel of qt_project.cu, which
// then only visits the items left at count_out[b] = -1.  55 ms -> ~1 ms per 16384 two-Kraus-operator channels.
int qt_choi2kraus_lowrank64(int64_t B, const void* choi, double tol, double* evals_out, void* kraus_out,
                            int32_t* count_out, cudaStream_t st) {
  const size_t smem = sizeof(cplx) * (64 * 8 + 2 * 64) + sizeof(double) * (8 + JacobiScratch<8>::doubles + 2 * 32 + 8);
  const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 16);
  choi2kraus_lowrank_kernel<64><<<grid, 64, smem, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out, count_out);
  return qt_check_launch("choi2kraus_lowrank_kernel");
}

// ---- n = 4, 5 projections: host side -------------------------------------------------------------------------
extern "C" int64_t qt_proj_cp_workspace_bytes(int n, int64_t B) {
  if (n >= 1 && n <= 3) return 0;
  if (n != 4 && n != 5) return -1;
  const int64_t M = 1LL << (2 * n);
  return large_grid(B) * M * M * (int64_t)sizeof(cplx);
}

int qt_large_proj_cp(int n, int64_t B, const void* in, void* out, void* ws, int64_t ws_bytes, cudaStream_t st) {
  QT_REQUIRE(ws && ws_bytes >= qt_proj_cp_workspace_bytes(n, B),
             "qt_proj_cp_batch: n = %d needs a workspace of qt_proj_cp_workspace_bytes(n, B) = %lld bytes", n,
             (long long)qt_proj_cp_workspace_bytes(n, B));
  QT_REQUIRE(in != out, "qt_proj_cp_batch: n >= 4 is out-of-place");
  const unsigned grid = (unsigned)large_grid(B);
  if (n == 4) proj_cp_large_kernel<256><<<grid, 512, 0, st>>>(B, (const cplx*)in, (cplx*)out, (cplx*)ws);
  else proj_cp_large_kernel<1024><<<grid, 1024, 0, st>>>(B, (const cplx*)in, (cplx*)out, (cplx*)ws);
  return qt_check_launch("proj_cp_large_kernel");
}

template <int N, bool MAKE_TP>
static int launch_tp_corr_large(int64_t B, const void* in, void* out, cudaStream_t st) {
  constexpr int D = 1 << N;
  const size_t smem = sizeof(cplx) * 4 * D * D + sizeof(double) * (D + JacobiScratch<D>::doubles);
  QT_CUDA(cudaFuncSetAttribute(tp_correction_large_kernel<N, MAKE_TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tp_correction_large_kernel<N, MAKE_TP><<<(unsigned)std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 4), 256, smem, st>>>(
      B, (const cplx*)in, (cplx*)out);
  return qt_check_launch("tp_correction_large_kernel");
}

// first pass of the TP / TNI projection at n = 4, 5 (the second pass is tni_apply_kernel in qt_project.cu)
int qt_large_tp_correction(int n, int64_t B, const void* in, void* out, int make_tp, cudaStream_t st) {
  if (n == 4) return make_tp ? launch_tp_corr_large<4, true>(B, in, out, st) : launch_tp_corr_large<4, false>(B, in, out, st);
  return make_tp ? launch_tp_corr_large<5, true>(B, in, out, st) : launch_tp_corr_large<5, false>(B, in, out, st);
}

int64_t qt_large_physical_workspace_bytes(int n, int64_t B) {
  const int64_t M = 1LL << (2 * n);
  return large_grid(B) * 4 * M * M * (int64_t)sizeof(cplx);
}

template <int N>
static int launch_physical_large(int64_t B, const void* in, void* out, int make_tp, void* ws, int* calls, int* status,
                                 cudaStream_t st) {
  constexpr int D = 1 << N;
  const size_t smem = sizeof(cplx) * 6 * D * D + sizeof(double) * (D + JacobiScratch<D>::doubles + 64);
  QT_CUDA(cudaFuncSetAttribute(proj_physical_large_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  proj_physical_large_kernel<N><<<(unsigned)large_grid(B), N == 4 ? 512 : 1024, smem, st>>>(
      B, (const cplx*)in, (cplx*)out, make_tp, (cplx*)ws, calls, status);
  return qt_check_launch("proj_physical_large_kernel");
}

int qt_large_proj_physical(int n, int64_t B, const void* in, void* out, int make_tp, void* ws, int64_t ws_bytes,
                           int* calls, int* status, cudaStream_t st) {
  QT_REQUIRE(ws_bytes >= qt_large_physical_workspace_bytes(n, B), "qt_proj_physical_batch: workspace too small (%lld < %lld bytes)",
             (long long)ws_bytes, (long long)qt_large_physical_workspace_bytes(n, B));
  if (n == 4) return launch_physical_large<4>(B, in, out, make_tp, ws, calls, status, st);
  return launch_physical_large<5>(B, in, out, make_tp, ws, calls, status, st);
}
