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
#include "qt_eigh.cuh"
#include "../../include/qtomo.h"

#include <algorithm>

template <int M>
__global__ void __launch_bounds__(M == 256 ? 512 : 1024)
    choi2kraus_large_kernel(int64_t B, const cplx* __restrict__ in, double tol, double* __restrict__ evals_out,
                            cplx* __restrict__ kraus_out, int* __restrict__ count_out, cplx* __restrict__ ws,
                            int* __restrict__ sweeps_out) {
  constexpr int NT = (M == 256) ? 512 : 1024, NW = NT / 32, HP = M / 2, PER = M / 32;
  constexpr int D = (M == 256) ? 16 : 32;
  // M = 256 (512 threads, 128 registers): the pair's two columns stay in registers between the Gram pass and the update;
  // M = 1024 re-reads them
  constexpr bool HOLD = (M == 256);
  __shared__ double ev[M], sig[M];
  __shared__ int rank[M], pos[M];
  __shared__ double red[NW];
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  cplx* U = ws + (size_t)blockIdx.x * M * M;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* src = in + b * (int64_t)M * M;
    // Hermitian matrix np.linalg.eigh sees (lower triangle), infinity norm, and the shifted start
    auto herm = [&](int r, int c) {
      cplx v = (r >= c) ? src[(size_t)r * M + c] : cconj(src[(size_t)c * M + r]);
      if (r == c) v.y = 0.0;
      return v;
    };
    double rowmax = 0.0;
    for (int r = wid; r < M; r += NW) {
      double s = 0.0;
      for (int c = lane; c < M; c += 32) s += sqrt(cabs2(herm(r, c)));
      s = warp_sum(s);
      rowmax = fmax(rowmax, s);
    }
    if (lane == 0) red[wid] = rowmax;
    __syncthreads();
    double shift = 0.0;
    for (int w = 0; w < NW; ++w) shift = fmax(shift, red[w]);
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
      if (tid == 0) rotated = 0;
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
          if (lane == 0) rotated = 1;
#pragma unroll
          for (int t = 0; t < PER; ++t) {
            const cplx x = HOLD ? a[t] : up[lane + 32 * t], y = HOLD ? c2[t] : uq[lane + 32 * t];
            up[lane + 32 * t] = csub(cscale(x, c), cmul(cs, y));
            uq[lane + 32 * t] = cadd(cmul(s, x), cscale(y, c));
          }
        }
        __syncthreads();
      }
      if (!rotated) break;
      __syncthreads();
    }
    if (tid == 0 && sweeps_out) sweeps_out[b] = sweep;
    // eigenvalues: column norms of the shifted matrix minus the shift
    for (int k = wid; k < M; k += NW) {
      double acc = 0.0;
      for (int r = lane; r < M; r += 32) acc += cabs2(U[(size_t)k * M + r]);
      acc = warp_sum(acc);
      if (lane == 0) {
        sig[k] = sqrt(acc);
        ev[k] = sig[k] - shift;
      }
    }
    __syncthreads();
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
  if (n == 4)
    choi2kraus_large_kernel<256><<<grid, 512, 0, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                         count_out, (cplx*)workspace, sweeps_out);
  else
    choi2kraus_large_kernel<1024><<<grid, 1024, 0, st>>>(B, (const cplx*)choi, tol, evals_out, (cplx*)kraus_out,
                                                          count_out, (cplx*)workspace, sweeps_out);
  return qt_check_launch("choi2kraus_large_kernel");
}
