// Batched iterative (diluted) maximum-likelihood state tomography -- the R rho R fixed point.
//
// Replaces, per experiment, the loop of forest/benchmarking/tomography.py:168-270 and the operator
// `_R` of tomography.py:273-338.  One experiment = one thread (n <= 2, state in registers) or one
// warp (n = 3..5, state in shared memory); the batch of independent experiments is the parallel axis.
//
// Algebra used (checked against the imported reference to 1e-16):
//   t_j  = Tr(P_j rho)                                   (forward Pauli transform of rho, real)
//   a+-  = ((1 +- m_k)/2) / ((1 +- c_k t_j)/2 + tiny)    for every result k whose observable is c_k P_j
//   R    = (1/K) sum_k [ (a+ + a-)/2 * I + c_k (a+ - a-)/2 * P_j ]      (K = len(results))
//   M    = I + eps (R - I) (+ maxent / hedging terms);   rho <- M rho M / tr;  stop on ||drho||_F < tol
// Loop semantics replicated exactly: at most maxiter-1 updates (tomography.py:244), division by
// len(results) (:338), `tiny` added to the predicted probability (:321,336).
#include "qt_common.cuh"
#include "qt_eigh.cuh"
#include "../../include/qtomo.h"

#include <algorithm>
#include <vector>

struct qt_mle_plan {
  int n, K, S;            // qubits, len(results), 4^n
  int unit_coeff;         // every coefficient == 1 -> register fast path allowed
  int* d_slot_ptr;        // [S+1]  CSR over canonical Pauli slots
  int* d_member_col;      // [K]    column of `expect` for each member
  double* d_member_coeff; // [K]
  double* d_member_linw;  // [K]    linear-inversion weight c_k / (d * sum_{k' in slot} c_k'^2)
  int* d_mask2idx;        // [S]    (x*D + z) -> canonical Pauli index
};

// =============================================================================================
// Register kernel: one experiment per thread, n = 1 or 2, unit coefficients, vanilla MLE.
// rho is kept Hermitian-packed: h[r][c] (r<c) = Re, h[c][r] = Im of element (r,c); h[r][r] = diagonal.
// =============================================================================================
template <int D>
struct Herm {
  double h[D][D];
  __device__ __forceinline__ double re(int r, int c) const { return r <= c ? h[r][c] : h[c][r]; }
  __device__ __forceinline__ double im(int r, int c) const {
    return r == c ? 0.0 : (r < c ? h[c][r] : -h[r][c]);
  }
};

template <int N>
__global__ void __launch_bounds__(32) mle_reg_kernel(int64_t B, int K, const int* __restrict__ slot_ptr,
                                                     const int* __restrict__ member_col,
                                                     const double* __restrict__ expect, double eps, double tol,
                                                     int maxiter, cplx* __restrict__ rho_out,
                                                     int* __restrict__ iters_out) {
  constexpr int D = 1 << N, S = 1 << (2 * N);
  constexpr double TINY = 2.2250738585072014e-308;
  __shared__ double fp[S][32], fm[S][32];
  const int tid = threadIdx.x;
  const int64_t b = (int64_t)blockIdx.x * 32 + tid;
  if (b >= B) return;

  // aggregate (1 +- m_k)/2 per canonical Pauli slot
  for (int s = 0; s < S; ++s) {
    double ap = 0.0, am = 0.0;
    for (int m = slot_ptr[s]; m < slot_ptr[s + 1]; ++m) {
      double e = expect[b * K + member_col[m]];
      ap += 0.5 * (1.0 + e);
      am += 0.5 * (1.0 - e);
    }
    fp[s][tid] = ap;
    fm[s][tid] = am;
  }
  const double invK = 1.0 / (double)K;

  Herm<D> rho;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) rho.h[r][c] = (r == c) ? 1.0 / D : 0.0;

  int it = 1;
  while (true) {
    if (it >= maxiter) break;
    // ---- forward Pauli transform + likelihood ratios --------------------------------------
    double w[S];
    double w0 = 0.0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int x = pauli_xmask(j, N), z = pauli_zmask(j, N);
      const int ph = popc_c(x & z) & 3;
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int r = c ^ x;  // element rho[c, r]
        const double sgn = (popc_c(z & c) & 1) ? -1.0 : 1.0;
        // Re( i^ph * rho[c, r] )
        double v;
        if (ph == 0) v = rho.re(c, r);
        else if (ph == 1) v = -rho.im(c, r);
        else if (ph == 2) v = -rho.re(c, r);
        else v = rho.im(c, r);
        t += sgn * v;
      }
      // one MUFU-seeded reciprocal for both ratios (an IEEE DDIV costs ~133 issue cycles per warp)
      const double pp = 0.5 * (1.0 + t) + TINY, pm = 0.5 * (1.0 - t) + TINY;
      const double ipm = fast_rcp(pp * pm);
      const double ap = fp[j][tid] * pm * ipm, am = fm[j][tid] * pp * ipm;
      w0 += 0.5 * (ap + am);
      w[j] = 0.5 * (ap - am);
    }
    w[0] += w0;
    // ---- M = (1 - eps) I + eps R,  R = (1/K) sum_j w_j P_j  (Hermitian packed) --------------
    Herm<D> M;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) {
        const int x = r ^ c;
        double sre = 0.0, sim = 0.0;
#pragma unroll
        for (int z = 0; z < D; ++z) {
          const int j = pauli_from_masks(x, z, N);
          const int ph = popc_c(x & z) & 3;
          const double sgn = (popc_c(z & c) & 1) ? -1.0 : 1.0;
          // P_j[r, c] = i^ph * sgn
          if (ph == 0) sre += sgn * w[j];
          else if (ph == 1) sim += sgn * w[j];
          else if (ph == 2) sre -= sgn * w[j];
          else sim -= sgn * w[j];
        }
        if (r == c) {
          M.h[r][r] = (1.0 - eps) + eps * invK * sre;
        } else {
          M.h[r][c] = eps * invK * sre;
          M.h[c][r] = eps * invK * sim;
        }
      }
    // ---- T = M rho (full), rho' = T M (upper triangle) ---------------------------------------
    double Tr_[D][D], Ti_[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double mr = M.re(r, k), mi = M.im(r, k), pr = rho.re(k, c), pi = rho.im(k, c);
          ar = fma(mr, pr, ar);
          if (k != c) ai = fma(mr, pi, ai);
          if (r != k) {
            if (k != c) ar = fma(-mi, pi, ar);
            ai = fma(mi, pr, ai);
          }
        }
        Tr_[r][c] = ar;
        Ti_[r][c] = ai;
      }
    Herm<D> nw;
    double tr = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) {
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double mr = M.re(k, c), mi = M.im(k, c);
          ar = fma(Tr_[r][k], mr, ar);
          if (k != c) ar = fma(-Ti_[r][k], mi, ar);
          if (r != c) {
            ai = fma(Ti_[r][k], mr, ai);
            if (k != c) ai = fma(Tr_[r][k], mi, ai);
          }
        }
        if (r == c) {
          nw.h[r][r] = ar;
          tr += ar;
        } else {
          nw.h[r][c] = ar;
          nw.h[c][r] = ai;
        }
      }
    const double inv = fast_rcp(tr);
    double diff = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double v = nw.h[r][c] * inv;
        const double dlt = v - rho.h[r][c];
        diff = fma((r == c) ? 1.0 : 2.0, dlt * dlt, diff);
        rho.h[r][c] = v;
      }
    if (sqrt(diff) < tol) break;
    ++it;
  }
  cplx* out = rho_out + b * D * D;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) out[r * D + c] = cmake(rho.re(r, c), rho.im(r, c));
  iters_out[b] = it;
}

// =============================================================================================
// Quad kernel: one 2-qubit experiment per 4 lanes (8 experiments per warp), unit coefficients, vanilla MLE.
//
// With one experiment per thread a batch of 4096 fills only 128 warps (< 1 per SM sub-partition) and
// every warp is latency-bound on its own dependency chains (ncu: 5 cycles per issued instruction, FP64
// pipe 6 % busy -- profiles/r01_ncu_mle_reg_kernel_v0.md).  Splitting an experiment over the 4 lanes of a
// quad cuts the per-warp critical path ~3x and puts a warp on ~86 % of the 592 sub-partitions.
//
// Lane c (= lane & 3) owns column c of rho' = M rho M.  To keep every register index static it works in
// the frame permuted by the X-type Pauli Pi_c (row r -> r ^ c):  R_c = Pi_c rho Pi_c, M_c = Pi_c M Pi_c, so
// that "column c" is always column 0:  v = M_c (R_c M_c[:,0]) = rho'[. ^ c, c].  Conjugation by Pi_c only
// flips signs in the Pauli basis:  Pi_c P_j Pi_c = (-1)^{|z_j & c|} P_j,  hence
//     Tr(P_j rho) = (-1)^{|z_j & c|} Tr(P_j R_c),     M_c = (1-eps) I + (eps/K) sum_j (-1)^{|z_j & c|} w_j P_j.
// The likelihood ratios of the 16 Pauli slots are split 4 per lane (slot j = 4c + i) and all-gathered with
// 32 shuffles; the new state is all-gathered with 18 (Hermitian: R_c[a][b] = shfl_xor(v[a ^ b], b)).
// =============================================================================================
__device__ __forceinline__ double flip_sign(double x, int mask) {
  return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
}
__device__ __forceinline__ double shfl_xor_d(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }

__global__ void __launch_bounds__(32) mle_quad_kernel(int64_t B, int K, const int* __restrict__ slot_ptr,
                                                      const int* __restrict__ member_col,
                                                      const double* __restrict__ expect, double eps, double tol,
                                                      int maxiter, cplx* __restrict__ rho_out,
                                                      int* __restrict__ iters_out) {
  constexpr int N = 2, D = 4, S = 16;
  constexpr double TINY = 2.2250738585072014e-308;
  const int lane = threadIdx.x;
  const int c = lane & 3, qbase = lane & ~3;
  const int64_t b = (int64_t)blockIdx.x * 8 + (lane >> 2);
  const bool valid = b < B;

  // this lane's 4 Pauli slots j = 4c + i: aggregated (1 +- m_k)/2 and the frame sign (-1)^{|z_j & c|}
  double fp[4], fm[4];
  int sgm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = 4 * c + i;
    double ap = 0.0, am = 0.0;
    if (valid) {
      for (int m = slot_ptr[j]; m < slot_ptr[j + 1]; ++m) {
        const double e = expect[b * K + member_col[m]];
        ap += 0.5 * (1.0 + e);
        am += 0.5 * (1.0 - e);
      }
    }
    fp[i] = ap;
    fm[i] = am;
    sgm[i] = (__popc(pauli_zmask(j, N) & c) & 1) ? (int)0x80000000 : 0;
  }
  const int m_z1 = (c & 1) ? (int)0x80000000 : 0, m_z2 = (c & 2) ? (int)0x80000000 : 0;
  const double sc = 0.5 * eps / (double)K;
  const double tol2 = tol * tol;

  Herm<D> R;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int k = 0; k < D; ++k) R.h[r][k] = (r == k) ? 1.0 / D : 0.0;

  // Loop bookkeeping with ONE exit/store site (at the bottom): the reference tests `iteration >= maxiter` at the
  // top of the next trip (tomography.py:244), which is the same as testing it right after the increment.
  int it = 1;
  bool done = !valid;
  if (valid && it >= maxiter) {  // maxiter <= 1: the maximally mixed start is the answer
    cplx* out = rho_out + b * (D * D) + c * D;
#pragma unroll
    for (int k = 0; k < D; ++k) out[k] = cmake((k == c) ? 1.0 / D : 0.0, 0.0);
    if (c == 0) iters_out[b] = it;
    done = true;
  }
  // One R-rho-R update Rin -> Rout (+ the squared Frobenius distance between them).  The convergence test of a trip is
  // evaluated one trip LATER (`settle`), so its compare / shuffle / vote / branch chain never sits between two
  // updates; the two state register sets alternate, which keeps the state a late test has to store intact.
  auto update = [&](const Herm<D>& Rin, Herm<D>& Rout, double& diff_out) {
    // ---- th[j] = Tr(P_j R_c) / 2 for all 16 Paulis (static butterfly on the Hermitian-packed state) ----
    double th[S];
    {
      const double s01 = Rin.h[0][0] + Rin.h[1][1], m01 = Rin.h[0][0] - Rin.h[1][1];
      const double s23 = Rin.h[2][2] + Rin.h[3][3], m23 = Rin.h[2][2] - Rin.h[3][3];
      th[pauli_from_masks(0, 0, N)] = 0.5 * (s01 + s23);
      th[pauli_from_masks(0, 1, N)] = 0.5 * (m01 + m23);
      th[pauli_from_masks(0, 2, N)] = 0.5 * (s01 - s23);
      th[pauli_from_masks(0, 3, N)] = 0.5 * (m01 - m23);
    }
#pragma unroll
    for (int x = 1; x < D; ++x) {
#pragma unroll
      for (int z = 0; z < D; ++z) {
        const int ph = popc_c(x & z) & 3;
        const bool odd = ph & 1;
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int a = 0; a < D; ++a) {
          if (a < (a ^ x)) {
            // pair (a, a^x): even phase -> +-2 Re, odd phase -> -+2 Im of element (a, a^x)
            const bool neg = ((popc_c(z & a) & 1) != 0) != (odd ? (ph == 1) : (ph == 2));
            const double v = odd ? Rin.im(a, a ^ x) : Rin.re(a, a ^ x);
            if (first) acc = neg ? -v : v;
            else acc = neg ? acc - v : acc + v;
            first = false;
          }
        }
        th[pauli_from_masks(x, z, N)] = acc;
      }
    }
    // ---- this lane's 4 slots: likelihood ratios with ONE reciprocal per slot ----
    double w[4], w0 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double lo = (c & 1) ? th[4 + i] : th[i];
      const double hi = (c & 1) ? th[12 + i] : th[8 + i];
      const double tt = flip_sign((c & 2) ? hi : lo, sgm[i]);  // true-frame Tr(P_j rho)/2
      const double pp = (0.5 + tt) + TINY, pm = (0.5 - tt) + TINY;
      const double inv = fast_rcp(pp * pm);
      const double ap = fp[i] * pm * inv, am = fm[i] * pp * inv;
      w0 += ap + am;
      w[i] = sc * (ap - am);
    }
    w0 += shfl_xor_d(w0, 1);
    w0 += shfl_xor_d(w0, 2);
    // ---- all-gather the 16 coefficients, move them to this lane's frame ----
    double wf[S];
#pragma unroll
    for (int src = 0; src < 4; ++src)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * src + i;
        const int z = pauli_zmask(j, N);
        const int mask = ((z & 1) ? m_z1 : 0) ^ ((z & 2) ? m_z2 : 0);
        const double v = __shfl_sync(0xffffffffu, w[i], qbase + src);
        wf[j] = (z == 0) ? v : flip_sign(v, mask);
      }
    wf[0] = fma(sc, w0, wf[0]);
    // ---- M_c = (1 - eps) I + sum_j wf_j P_j  (Hermitian packed) ----
    Herm<D> M;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = r; k < D; ++k) {
        const int x = r ^ k;
        double sre = 0.0, sim = 0.0;
        bool fre = true, fim = true;
#pragma unroll
        for (int z = 0; z < D; ++z) {
          const double wj = wf[pauli_from_masks(x, z, N)];
          const int ph = popc_c(x & z) & 3;
          const bool neg = ((popc_c(z & k) & 1) != 0) != (ph >= 2);  // P_j[r, k] = i^ph * sgn
          if ((ph & 1) == 0) {
            sre = fre ? (neg ? -wj : wj) : (neg ? sre - wj : sre + wj);
            fre = false;
          } else {
            sim = fim ? (neg ? -wj : wj) : (neg ? sim - wj : sim + wj);
            fim = false;
          }
        }
        if (r == k) {
          M.h[r][r] = (1.0 - eps) + sre;
        } else {
          M.h[r][k] = sre;
          M.h[k][r] = sim;
        }
      }
    // ---- u = R_c M_c[:, 0],  v = M_c u  (column 0 of M_c R_c M_c) ----
    double ur[D], ui[D], vr[D], vi[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double ar = 0.0, ai = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double pr = Rin.re(r, k), pi = Rin.im(r, k), mr = M.re(k, 0), mi = M.im(k, 0);
        ar = fma(pr, mr, ar);
        if (k != 0) ai = fma(pr, mi, ai);
        if (r != k) {
          if (k != 0) ar = fma(-pi, mi, ar);
          ai = fma(pi, mr, ai);
        }
      }
      ur[r] = ar;
      ui[r] = ai;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double ar = 0.0, ai = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double mr = M.re(r, k), mi = M.im(r, k);
        ar = fma(mr, ur[k], ar);
        ai = fma(mr, ui[k], ai);
        if (r != k) {
          ar = fma(-mi, ui[k], ar);
          ai = fma(mi, ur[k], ai);
        }
      }
      vr[r] = ar;
      vi[r] = ai;
    }
    // ---- all-gather the new (unnormalised) state in this lane's frame ----
    Herm<D> nw;
    nw.h[0][0] = vr[0];
#pragma unroll
    for (int k = 1; k < D; ++k) {  // element (0, k) = conj(v[k])
      nw.h[0][k] = vr[k];
      nw.h[k][0] = -vi[k];
      nw.h[k][k] = shfl_xor_d(vr[0], k);
    }
#pragma unroll
    for (int a = 1; a < D; ++a)
#pragma unroll
      for (int k = a + 1; k < D; ++k) {  // element (a, k) = v[a ^ k] of lane c ^ k
        nw.h[a][k] = shfl_xor_d(vr[a ^ k], k);
        nw.h[k][a] = shfl_xor_d(vi[a ^ k], k);
      }
    // trace and ||rho' - rho||_F^2 summed pairwise: invariant under the lanes' index relabelling
    const double tr = (nw.h[0][0] + nw.h[1][1]) + (nw.h[2][2] + nw.h[3][3]);
    const double inv = fast_rcp(tr);
    double dd[D], dx[D];
#pragma unroll
    for (int x = 0; x < D; ++x) dx[x] = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double v = nw.h[r][k] * inv;
        const double dlt = v - Rin.h[r][k];
        if (r == k) dd[r] = dlt * dlt;
        else dx[r ^ k] = fma(dlt, dlt, dx[r ^ k]);
        Rout.h[r][k] = v;
      }
    diff_out = ((dd[0] + dd[1]) + (dd[2] + dd[3])) + 2.0 * ((dx[1] + dx[2]) + dx[3]);
  };
  // deferred bookkeeping of the PREVIOUS trip, whose result `Rnew` is the input of the trip that has just been issued
  auto settle = [&](const Herm<D>& Rnew, double diff_prev) {
    const int conv = __shfl_sync(0xffffffffu, (int)(diff_prev < tol2), qbase);
    if (!done && !conv) ++it;
    if (!done && (conv || it >= maxiter)) {
      cplx* out = rho_out + b * (D * D) + c * D;
      out[c] = cmake(Rnew.h[0][0], 0.0);
#pragma unroll
      for (int k = 1; k < D; ++k) out[k ^ c] = cmake(Rnew.re(0, k), Rnew.im(0, k));
      if (c == 0) iters_out[b] = it;
      done = true;
    }
  };
  // `it` starts one short and the first settle (nothing pending: distance = infinity) brings it to the
  // reference's initial 1 without storing anything
  Herm<D> R2;
  double diff_pend = 1e300, diff_new;
  if (!done) it = 0;
  while (true) {
    update(R, R2, diff_new);
    settle(R, diff_pend);
    diff_pend = diff_new;
    if (__all_sync(0xffffffffu, done)) break;
    update(R2, R, diff_new);
    settle(R2, diff_pend);
    diff_pend = diff_new;
    if (__all_sync(0xffffffffu, done)) break;
  }
}

// =============================================================================================
// Warp kernel: one experiment per warp, any n <= 5, arbitrary observable lists / coefficients,
// vanilla + maximum-entropy + hedged variants (tomography.py:252-260).  State in shared memory.
// =============================================================================================
template <int N>
struct MleWarpSmem {
  static constexpr int D = 1 << N, S = 1 << (2 * N);
  // rho, M, T (+ V for the eigen-decomposition used by the variants)
  __host__ __device__ static constexpr size_t bytes(bool variants) {
    return (sizeof(cplx) * D * D * (variants ? 4 : 3) + sizeof(double) * S * 4 +
            sizeof(double) * (D + JacobiScratch<D>::doubles) + 15) / 16 * 16;
  }
};

// Sign bookkeeping of the warp kernel without per-term popcounts, selects and phase switches.
// walsh_parity<N>(m): bit v (v < 2^N) = parity(m & v), built by doubling (the upper half of every 2^(b+1)-block is the
// lower half, complemented if bit b of m is set).  popc2_patterns<N>(m, p0, p1): bits 0 and 1 of popcount(m & v).
template <int N>
__device__ __forceinline__ unsigned walsh_parity(int m) {
  unsigned p = 0;
#pragma unroll
  for (int b = 0; b < N; ++b) {
    const unsigned half = (1u << (1 << b)) - 1u;
    const unsigned lo = p & half;
    p = lo | ((((m >> b) & 1) ? (~lo & half) : lo) << (1 << b));
  }
  return p;
}
template <int N>
__device__ __forceinline__ void popc2_patterns(int m, unsigned& p0, unsigned& p1) {
  p0 = 0;
  p1 = 0;
#pragma unroll
  for (int b = 0; b < N; ++b) {
    const unsigned half = (1u << (1 << b)) - 1u;
    const unsigned l0 = p0 & half, l1 = p1 & half;
    const bool add = (m >> b) & 1;  // count + 1 in the upper half: bit 0 flips, bit 1 takes the carry
    p0 = l0 | ((add ? (~l0 & half) : l0) << (1 << b));
    p1 = l1 | ((add ? ((l1 ^ l0) & half) : l1) << (1 << b));
  }
}

template <int N>
__global__ void mle_warp_kernel(int64_t B, int K, const int* __restrict__ slot_ptr,
                                const int* __restrict__ member_col, const double* __restrict__ member_coeff,
                                const int* __restrict__ mask2idx, const double* __restrict__ expect,
                                const double* __restrict__ counts, double eps, double entropy_penalty, double beta,
                                double tol, int maxiter, int unit_coeff, cplx* __restrict__ rho_out,
                                int* __restrict__ iters_out) {
  constexpr int D = 1 << N, S = 1 << (2 * N), DD = D * D;
  constexpr double TINY = 2.2250738585072014e-308;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const bool variants = (entropy_penalty > 0.0) || (beta > 0.0);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t per_warp = MleWarpSmem<N>::bytes(variants);
  unsigned char* base = smem_raw + per_warp * wib;
  cplx* rho = reinterpret_cast<cplx*>(base);
  cplx* M = rho + DD;
  cplx* T = M + DD;
  cplx* V = T + DD;                       // only when variants
  double* tv = reinterpret_cast<double*>(rho + (variants ? 4 : 3) * DD);
  double* wv = tv + S;
  double* fpv = wv + S;                   // unit coefficients: sum over the slot's members of (1 + m_k)/2 ...
  double* fmv = fpv + S;                  // ... and of (1 - m_k)/2: the loop then needs no global load and one reciprocal
  double* ev = fmv + S;                   // [D] eigenvalues, followed by the Jacobi scratch
  const int64_t b = (int64_t)blockIdx.x * wpb + wib;
  if (b >= B) return;
  const double* ex = expect + b * K;
  if (unit_coeff) {
    for (int j = lane; j < S; j += 32) {
      double ap = 0.0, am = 0.0;
      for (int m = slot_ptr[j]; m < slot_ptr[j + 1]; ++m) {
        const double e = ex[member_col[m]];
        ap += 0.5 * (1.0 + e);
        am += 0.5 * (1.0 - e);
      }
      fpv[j] = ap;
      fmv[j] = am;
    }
  }

  double num_meas = 0.0;
  if (beta > 0.0) {
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += counts[b * K + k];
    num_meas = warp_sum(s);
  }
  for (int e = lane; e < DD; e += 32) rho[e] = cmake((e / D == e % D) ? 1.0 / D : 0.0, 0.0);
  __syncwarp();

  int it = 1;
  while (true) {
    if (it >= maxiter) break;
    // t_j = Tr(P_j rho) and the ratio sums per slot
    double w0 = 0.0;
    for (int j = lane; j < S; j += 32) {
      const int x = pauli_xmask(j, N), z = pauli_zmask(j, N);
      const int ph = __popc(x & z) & 3;
      cplx acc = cmake(0.0, 0.0);
      if constexpr (N <= 4) {
        const unsigned zpar = walsh_parity<N>(z);  // bit c = parity(z & c): the sign of rho[c][c ^ x] in the trace
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const cplx v = rho[c * D + (c ^ x)];
          const int sw = (int)((zpar >> c) << 31);  // acc -= v  ==  acc += (-v), bit for bit
          acc.x += flip_sign(v.x, sw);
          acc.y += flip_sign(v.y, sw);
        }
      } else {  // d = 32: the unrolled form spills; keep the loop
        for (int c = 0; c < D; ++c) {
          cplx v = rho[c * D + (c ^ x)];
          if (__popc(z & c) & 1) acc = csub(acc, v); else acc = cadd(acc, v);
        }
      }
      const double t = cmul_ipow(acc, ph).x;
      double wj = 0.0;
      if (unit_coeff) {
        // sum_m f_m / (pr + tiny) = (sum_m f_m) / (pr + tiny): one MUFU-seeded reciprocal for both outcomes (an IEEE
        // division costs ~133 issue cycles per warp, profiles/r01_ubench_fp64.txt); same formulation as mle_reg_kernel
        const double pp = 0.5 * (1.0 + t) + TINY, pm = 0.5 * (1.0 - t) + TINY;
        const double ipm = fast_rcp(pp * pm);
        const double ap = fpv[j] * pm * ipm, am = fmv[j] * pp * ipm;
        w0 += 0.5 * (ap + am);
        wj = 0.5 * (ap - am);
      } else
      for (int m = slot_ptr[j]; m < slot_ptr[j + 1]; ++m) {
        const double e = ex[member_col[m]], cf = member_coeff[m];
        const double pred = cf * t;
        const double ap = (0.5 * (1.0 + e)) / (0.5 * (1.0 + pred) + TINY);
        const double am = (0.5 * (1.0 - e)) / (0.5 * (1.0 - pred) + TINY);
        w0 += 0.5 * (ap + am);
        wj += cf * 0.5 * (ap - am);
      }
      wv[j] = wj;
    }
    w0 = warp_sum(w0);
    __syncwarp();
    if (lane == 0) wv[0] += w0;
    __syncwarp();
    // Tk = R - I  (stored in M for now)
    const double invK = 1.0 / (double)K;
    for (int e = lane; e < DD; e += 32) {
      const int r = e / D, c = e % D, x = r ^ c;
      // term z is  w[(x, z)] i^{|x & z|} (-1)^{|z & c|}: real for an even phase, imaginary for an odd one, negative when
      // bit 1 of the phase and the parity of z & c differ.  The canonical index of (x, z) is linear in the bits of z:
      // digit b = x_b ? 1 + z_b : 3 z_b.  Same terms, same order as the phase switch + conditional add it replaces.
      cplx acc = cmake(0.0, 0.0);
      if constexpr (N <= 4) {
        unsigned p0, p1;
        popc2_patterns<N>(x, p0, p1);
        const unsigned neg = p1 ^ walsh_parity<N>(c);
        int idx[D];
        idx[0] = pauli_from_masks(x, 0, N);
#pragma unroll
        for (int z = 1; z < D; ++z) {
          const int b = 31 - __builtin_clz(z & -z);  // lowest set bit of z (compile time after unrolling)
          idx[z] = idx[z & (z - 1)] + ((((x >> b) & 1) ? 1 : 3) << (2 * b));
        }
#pragma unroll
        for (int z = 0; z < D; ++z) {
          const double sv = flip_sign(wv[idx[z]], (int)((neg >> z) << 31));
          if ((p0 >> z) & 1) acc.y += sv; else acc.x += sv;
        }
      } else {
        for (int z = 0; z < D; ++z) {
          const double wj = wv[pauli_from_masks(x, z, N)];  // == mask2idx[x * D + z], computed instead of loaded
          cplx term = cmul_ipow(cmake(wj, 0.0), __popc(x & z) & 3);
          if (__popc(z & c) & 1) acc = csub(acc, term); else acc = cadd(acc, term);
        }
      }
      acc = cscale(acc, invK);
      if (r == c) acc.x -= 1.0;
      M[e] = acc;
    }
    __syncwarp();
    if (variants) {
      // eigen-decomposition of the (Hermitian) current state: rho = V diag(ev) V^dagger
      for (int e = lane; e < DD; e += 32) T[e] = rho[e];
      __syncwarp();
      jacobi_eigh_warp<D>(T, V, ev, lane);
      if (entropy_penalty > 0.0) {
        // logm(rho) - I tr(rho logm rho)   (tomography.py:252-254)
        double s = 0.0;
        for (int k = lane; k < D; k += 32) s += ev[k] * log(ev[k]);
        s = warp_sum(s);
        for (int e = lane; e < DD; e += 32) {
          const int r = e / D, c = e % D;
          cplx acc = cmake(0.0, 0.0);
          for (int k = 0; k < D; ++k) {
            cplx vv = cscale(V[r * D + k], log(ev[k]));
            cfma_conj(acc, vv, V[c * D + k]);
          }
          if (r == c) acc.x -= s;
          M[e] = csub(M[e], cscale(acc, entropy_penalty));
        }
      }
      if (beta > 0.0) {
        // Tk *= num_meas/2;  Tk += beta (pinv(rho) - d I)/2   (tomography.py:257-260)
        // scipy.linalg.pinv cut-off: singular values <= max(M,N) * eps * sigma_max are dropped.
        double smax = 0.0;
        for (int k = 0; k < D; ++k) smax = fmax(smax, fabs(ev[k]));
        const double cut = D * 2.220446049250313e-16 * smax;
        for (int e = lane; e < DD; e += 32) {
          const int r = e / D, c = e % D;
          cplx acc = cmake(0.0, 0.0);
          for (int k = 0; k < D; ++k) {
            if (fabs(ev[k]) > cut) {
              cplx vv = cscale(V[r * D + k], 1.0 / ev[k]);
              cfma_conj(acc, vv, V[c * D + k]);
            }
          }
          if (r == c) acc.x -= (double)D;
          cplx tk = cscale(M[e], 0.5 * num_meas);
          M[e] = cadd(tk, cscale(acc, 0.5 * beta));
        }
      }
      __syncwarp();
    }
    // M = I + eps Tk
    for (int e = lane; e < DD; e += 32) {
      cplx v = cscale(M[e], eps);
      if (e / D == e % D) v.x += 1.0;
      M[e] = v;
    }
    __syncwarp();
    // T = M rho
    for (int e = lane; e < DD; e += 32) {
      const int r = e / D, c = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int k = 0; k < D; ++k) cfma(acc, M[r * D + k], rho[k * D + c]);
      T[e] = acc;
    }
    __syncwarp();
    // rho' = T M, trace, normalise, ||rho' - rho||_F
    double tr = 0.0;
    cplx nv[(DD + 31) / 32];
#pragma unroll
    for (int i = 0; i < (DD + 31) / 32; ++i) {
      const int e = lane + 32 * i;
      cplx acc = cmake(0.0, 0.0);
      if (e < DD) {
        const int r = e / D, c = e % D;
        for (int k = 0; k < D; ++k) cfma(acc, T[r * D + k], M[k * D + c]);
        if (r == c) tr += acc.x;
      }
      nv[i] = acc;
    }
    // the reference divides by the complex trace; its imaginary part is rounding noise (~1e-17)
    tr = warp_sum(tr);
    const double inv = fast_rcp(tr);
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < (DD + 31) / 32; ++i) {
      const int e = lane + 32 * i;
      if (e < DD) {
        cplx v = cscale(nv[i], inv);
        diff += cabs2(csub(v, rho[e]));
        nv[i] = v;
      }
    }
    diff = warp_sum(diff);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < (DD + 31) / 32; ++i) {
      const int e = lane + 32 * i;
      if (e < DD) rho[e] = nv[i];
    }
    __syncwarp();
    if (sqrt(diff) < tol) break;
    ++it;
  }
  cplx* out = rho_out + b * DD;
  for (int e = lane; e < DD; e += 32) out[e] = rho[e];
  if (lane == 0) iters_out[b] = it;
}

// =============================================================================================
// single-step streaming kernel: one R rho R update per experiment (n = 1, 2), rho read from and
// written to HBM.  This is the HBM-roofline view of the update: 632 B / item at n=2.
// =============================================================================================
template <int N>
__global__ void mle_step_kernel(int64_t B, int K, const double* __restrict__ expect_canon,
                                const cplx* __restrict__ rho_in, double eps, cplx* __restrict__ rho_out) {
  // expect_canon: [S-1, B] canonical-order expectations, item-minor (coalesced); K = S-1 results.
  constexpr int D = 1 << N, S = 1 << (2 * N), DD = D * D;
  constexpr double TINY = 2.2250738585072014e-308;
  __shared__ cplx tile[QT_TS * DD];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * 128;
  const int nb = (int)min((int64_t)128, B - b0);
  // coalesced load of nb matrices (16 B per element, consecutive threads -> consecutive elements)
  for (int e = tid; e < nb * DD; e += 128) tile[(e % DD) * QT_TS + (e / DD)] = rho_in[b0 * DD + e];
  __syncthreads();
  if (tid < nb) {
    cplx r[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) r[i][j] = tile[(i * D + j) * QT_TS + tid];
    double w[S];
    double w0 = 0.0;
    w[0] = 0.0;
#pragma unroll
    for (int j = 1; j < S; ++j) {
      const int x = pauli_xmask(j, N), z = pauli_zmask(j, N);
      cplx acc = cmake(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        if (popc_c(z & c) & 1) acc = csub(acc, r[c][c ^ x]); else acc = cadd(acc, r[c][c ^ x]);
      }
      const double t = cmul_ipow(acc, popc_c(x & z) & 3).x;
      const double e = expect_canon[(int64_t)(j - 1) * B + b0 + tid];
      // one reciprocal for both ratios (a correctly rounded DDIV costs ~133 issue cycles per warp and made
      // this streaming kernel FP64-bound instead of HBM-bound)
      const double pp = 0.5 * (1.0 + t) + TINY, pm = 0.5 * (1.0 - t) + TINY;
      const double ipm = fast_rcp(pp * pm);
      const double ap = (0.5 * (1.0 + e)) * pm * ipm;
      const double am = (0.5 * (1.0 - e)) * pp * ipm;
      w0 += 0.5 * (ap + am);
      w[j] = 0.5 * (ap - am);
    }
    w[0] = w0;
    const double invK = 1.0 / (double)K;
    cplx M[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int x = i ^ c;
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int z = 0; z < D; ++z) {
          cplx term = cmul_ipow(cmake(w[pauli_from_masks(x, z, N)], 0.0), popc_c(x & z) & 3);
          if (popc_c(z & c) & 1) acc = csub(acc, term); else acc = cadd(acc, term);
        }
        acc = cscale(acc, eps * invK);
        if (i == c) acc.x += 1.0 - eps;
        M[i][c] = acc;
      }
    cplx T[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) cfma(acc, M[i][k], r[k][c]);
        T[i][c] = acc;
      }
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) cfma(acc, T[i][k], M[k][c]);
        r[i][c] = acc;
        if (i == c) tr += acc.x;
      }
    const double inv = fast_rcp(tr);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) tile[(i * D + c) * QT_TS + tid] = cscale(r[i][c], inv);
  }
  __syncthreads();
  for (int e = tid; e < nb * DD; e += 128) rho_out[b0 * DD + e] = tile[(e % DD) * QT_TS + (e / DD)];
}

// n = 2 with the Hermitian-packed arithmetic of mle_reg_kernel (rho is a state: its upper triangle is read): ~750
// FP64 instructions per item instead of ~1160 with general complex products, and 16 state registers instead of 32.
__global__ void __launch_bounds__(128) mle_step_herm_kernel(int64_t B, int K, const double* __restrict__ expect_canon,
                                                            const cplx* __restrict__ rho_in, double eps,
                                                            cplx* __restrict__ rho_out) {
  constexpr int N = 2, D = 4, S = 16, DD = 16;
  constexpr double TINY = 2.2250738585072014e-308;
  __shared__ cplx tile[QT_TS * DD];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * 128;
  const int nb = (int)min((int64_t)128, B - b0);
  for (int e = tid; e < nb * DD; e += 128) tile[(e % DD) * QT_TS + (e / DD)] = rho_in[b0 * DD + e];
  __syncthreads();
  if (tid < nb) {
    Herm<D> rho;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) {
        const cplx e = tile[(r * D + c) * QT_TS + tid];
        if (r == c) {
          rho.h[r][r] = e.x;
        } else {
          rho.h[r][c] = e.x;
          rho.h[c][r] = e.y;
        }
      }
    double w[S];
    double w0 = 0.0;
    w[0] = 0.0;
#pragma unroll
    for (int j = 1; j < S; ++j) {
      const int x = pauli_xmask(j, N), z = pauli_zmask(j, N);
      const int ph = popc_c(x & z) & 3;
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int r = c ^ x;  // Re( i^ph * rho[c, r] )
        double v;
        if (ph == 0) v = rho.re(c, r);
        else if (ph == 1) v = -rho.im(c, r);
        else if (ph == 2) v = -rho.re(c, r);
        else v = rho.im(c, r);
        t += ((popc_c(z & c) & 1) ? -1.0 : 1.0) * v;
      }
      const double e = expect_canon[(int64_t)(j - 1) * B + b0 + tid];
      const double pp = 0.5 * (1.0 + t) + TINY, pm = 0.5 * (1.0 - t) + TINY;
      const double ipm = fast_rcp(pp * pm);
      const double ap = (0.5 * (1.0 + e)) * pm * ipm, am = (0.5 * (1.0 - e)) * pp * ipm;
      w0 += 0.5 * (ap + am);
      w[j] = 0.5 * (ap - am);
    }
    w[0] = w0;
    const double sc = eps / (double)K;
    Herm<D> M;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) {
        const int x = r ^ c;
        double sre = 0.0, sim = 0.0;
#pragma unroll
        for (int z = 0; z < D; ++z) {
          const int j = pauli_from_masks(x, z, N);
          const int ph = popc_c(x & z) & 3;
          const double sgn = (popc_c(z & c) & 1) ? -1.0 : 1.0;
          if (ph == 0) sre += sgn * w[j];
          else if (ph == 1) sim += sgn * w[j];
          else if (ph == 2) sre -= sgn * w[j];
          else sim -= sgn * w[j];
        }
        if (r == c) {
          M.h[r][r] = (1.0 - eps) + sc * sre;
        } else {
          M.h[r][c] = sc * sre;
          M.h[c][r] = sc * sim;
        }
      }
    double Tr_[D][D], Ti_[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double mr = M.re(r, k), mi = M.im(r, k), pr = rho.re(k, c), pi = rho.im(k, c);
          ar = fma(mr, pr, ar);
          if (k != c) ai = fma(mr, pi, ai);
          if (r != k) {
            if (k != c) ar = fma(-mi, pi, ar);
            ai = fma(mi, pr, ai);
          }
        }
        Tr_[r][c] = ar;
        Ti_[r][c] = ai;
      }
    Herm<D> nw;
    double tr = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = r; c < D; ++c) {
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double mr = M.re(k, c), mi = M.im(k, c);
          ar = fma(Tr_[r][k], mr, ar);
          if (k != c) ar = fma(-Ti_[r][k], mi, ar);
          if (r != c) {
            ai = fma(Ti_[r][k], mr, ai);
            if (k != c) ai = fma(Tr_[r][k], mi, ai);
          }
        }
        if (r == c) {
          nw.h[r][r] = ar;
          tr += ar;
        } else {
          nw.h[r][c] = ar;
          nw.h[c][r] = ai;
        }
      }
    const double inv = fast_rcp(tr);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) tile[(r * D + c) * QT_TS + tid] = cmake(nw.re(r, c) * inv, nw.im(r, c) * inv);
  }
  __syncthreads();
  for (int e = tid; e < nb * DD; e += 128) rho_out[b0 * DD + e] = tile[(e % DD) * QT_TS + (e / DD)];
}

// =============================================================================================
// Linear inversion (tomography.py:130-165): rho = unvec(pinv(M) e) + I/d with rows M_k = c_k vec(P_k)^dagger.
// Distinct Paulis are orthogonal, so M^dagger M is diagonal in the Pauli basis and the pseudo-inverse is a
// per-slot weighted average:  rho = I/d + sum_j [ sum_{k in j} c_k e_k / (d sum_{k in j} c_k^2) ] P_j
// for ANY list of Pauli observables (duplicates, missing terms, non-unit coefficients) -- one inverse Pauli
// transform per experiment, no SVD.  One block per experiment.
// =============================================================================================
template <int N>
__global__ void linear_inv_kernel(int64_t B, int K, const int* __restrict__ slot_ptr, const int* __restrict__ member_col,
                                  const double* __restrict__ member_linw, const int* __restrict__ mask2idx,
                                  const double* __restrict__ expect, cplx* __restrict__ rho_out) {
  constexpr int D = 1 << N, S = 1 << (2 * N), DD = D * D;
  __shared__ double w[S];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const double* ex = expect + b * K;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      double acc = 0.0;
      for (int m = slot_ptr[j]; m < slot_ptr[j + 1]; ++m) acc = fma(member_linw[m], ex[member_col[m]], acc);
      w[j] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < DD; e += blockDim.x) {
      const int r = e / D, c = e % D, x = r ^ c;
      cplx acc = cmake(r == c ? 1.0 / D : 0.0, 0.0);
      for (int z = 0; z < D; ++z) {
        const cplx term = cmul_ipow(cmake(w[mask2idx[x * D + z]], 0.0), __popc(x & z) & 3);
        if (__popc(z & c) & 1) acc = csub(acc, term); else acc = cadd(acc, term);
      }
      rho_out[b * DD + e] = acc;
    }
    __syncthreads();
  }
}

// =============================================================================================
// state_log_likelihood (tomography.py:341-375), one experiment per block:
//   ll = sum_k sum_{sign} n_k (1 + sign e_k)/2 * log10((1 + sign c_k tr(P_k rho))/2),  terms with pr <= 0 skipped.
// The Pauli expectations tr(P rho) of the state are formed once per block for all 4^n slots.
// =============================================================================================
template <int N>
__global__ void log_likelihood_kernel(int64_t B, int K, const int* __restrict__ slot_ptr,
                                      const int* __restrict__ member_col, const double* __restrict__ member_coeff,
                                      const int* __restrict__ mask2idx, const cplx* __restrict__ rho,
                                      const double* __restrict__ expect, const double* __restrict__ counts,
                                      double* __restrict__ ll_out) {
  constexpr int D = 1 << N, S = 1 << (2 * N), DD = D * D;
  __shared__ double t[S];
  __shared__ double red[32];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const cplx* r = rho + b * DD;
    for (int e = threadIdx.x; e < S; e += blockDim.x) {
      const int x = e / D, z = e % D;
      cplx acc = cmake(0.0, 0.0);
      for (int c = 0; c < D; ++c) {  // tr(P rho) = sum_c P[c^x][c] rho[c][c^x],  P[c^x][c] = i^{|x&z|} (-1)^{|z&c|}
        const cplx v = r[c * D + (c ^ x)];
        if (__popc(z & c) & 1) acc = csub(acc, v); else acc = cadd(acc, v);
      }
      t[mask2idx[e]] = cmul_ipow(acc, __popc(x & z) & 3).x;
    }
    __syncthreads();
    double ll = 0.0;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      for (int m = slot_ptr[j]; m < slot_ptr[j + 1]; ++m) {
        const int col = member_col[m];
        const double meas = expect[b * K + col], n = counts[b * K + col], pred = member_coeff[m] * t[j];
        const double pp = 0.5 * (1.0 + pred), pm = 0.5 * (1.0 - pred);
        if (pp > 0.0) ll += n * (1.0 + meas) * 0.5 * log10(pp);
        if (pm > 0.0) ll += n * (1.0 - meas) * 0.5 * log10(pm);
      }
    }
    ll = warp_sum(ll);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ll;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      ll_out[b] = tot;
    }
    __syncthreads();
  }
}

// n = 1, 2: 4^n lanes per experiment, 32 / 4^n experiments per warp, no shared memory.  Lane e = (x, z) forms
// tr(P rho) of ITS Pauli from the d elements rho[c][c ^ x] (the lanes of one x read the same 16-byte elements: a
// broadcast) and then walks the members of that Pauli's slot; the block-per-experiment kernel above spent its time on
// two block barriers and one 16-element pass per experiment (0.09 of the HBM roof at n = 2).
template <int N>
__global__ void __launch_bounds__(256)
    log_likelihood_packed_kernel(int64_t B, int K, const int* __restrict__ slot_ptr, const int* __restrict__ member_col,
                                 const double* __restrict__ member_coeff, const int* __restrict__ mask2idx,
                                 const cplx* __restrict__ rho, const double* __restrict__ expect,
                                 const double* __restrict__ counts, double* __restrict__ ll_out) {
  constexpr int D = 1 << N, S = 1 << (2 * N), DD = D * D, G = 32 / S;
  const int lane = threadIdx.x & 31, e = lane % S;
  const int x = e / D, z = e % D;
  const int j = mask2idx[e], m0 = slot_ptr[j], m1 = slot_ptr[j + 1];
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rounds = (B + nwarps * G - 1) / (nwarps * G);
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t b = (it * nwarps + warp) * G + lane / S;
    const bool live = b < B;
    double ll = 0.0;
    if (live) {
      const cplx* r = rho + b * DD;
      cplx acc = cmake(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const cplx v = r[c * D + (c ^ x)];
        if (__popc(z & c) & 1) acc = csub(acc, v); else acc = cadd(acc, v);
      }
      const double t = cmul_ipow(acc, __popc(x & z) & 3).x;
      for (int m = m0; m < m1; ++m) {
        const int col = member_col[m];
        const double meas = expect[b * K + col], n = counts[b * K + col], pred = member_coeff[m] * t;
        const double pp = 0.5 * (1.0 + pred), pm = 0.5 * (1.0 - pred);
        if (pp > 0.0) ll += n * (1.0 + meas) * 0.5 * log10(pp);
        if (pm > 0.0) ll += n * (1.0 - meas) * 0.5 * log10(pm);
      }
    }
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
    if (live && e == 0) ll_out[b] = ll;
  }
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int qt_mle_plan_create(int n, int K, const int32_t* pauli_idx, const double* coeff,
                                  qt_mle_plan** plan_out) {
  QT_REQUIRE(n >= 1 && n <= 5, "qt_mle_plan_create: n=%d out of range 1..5", n);
  QT_REQUIRE(K >= 1 && pauli_idx && coeff && plan_out, "qt_mle_plan_create: bad arguments");
  const int S = 1 << (2 * n), D = 1 << n;
  std::vector<int> slot_ptr(S + 1, 0), col(K), cursor(S, 0), m2i(S);
  std::vector<double> cf(K), linw(K), slot_c2(S, 0.0);
  int unit = 1;
  for (int k = 0; k < K; ++k) {
    QT_REQUIRE(pauli_idx[k] >= 0 && pauli_idx[k] < S, "qt_mle_plan_create: pauli_idx[%d]=%d out of range", k,
               pauli_idx[k]);
    slot_ptr[pauli_idx[k] + 1]++;
    if (coeff[k] != 1.0) unit = 0;
  }
  for (int s = 0; s < S; ++s) slot_ptr[s + 1] += slot_ptr[s];
  for (int k = 0; k < K; ++k) {  // stable: members keep the order of `results`
    int s = pauli_idx[k];
    int pos = slot_ptr[s] + cursor[s]++;
    col[pos] = k;
    cf[pos] = coeff[k];
    slot_c2[s] += coeff[k] * coeff[k];
  }
  for (int s = 0; s < S; ++s)
    for (int m = slot_ptr[s]; m < slot_ptr[s + 1]; ++m) linw[m] = slot_c2[s] > 0.0 ? cf[m] / (D * slot_c2[s]) : 0.0;
  for (int x = 0; x < D; ++x)
    for (int z = 0; z < D; ++z) m2i[x * D + z] = pauli_from_masks(x, z, n);
  qt_mle_plan* p = new qt_mle_plan();
  p->n = n; p->K = K; p->S = S; p->unit_coeff = unit;
  p->d_slot_ptr = nullptr; p->d_member_col = nullptr; p->d_member_coeff = nullptr; p->d_mask2idx = nullptr;
  p->d_member_linw = nullptr;
  QT_CUDA(cudaMalloc(&p->d_member_linw, sizeof(double) * K));
  QT_CUDA(cudaMemcpy(p->d_member_linw, linw.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMalloc(&p->d_slot_ptr, sizeof(int) * (S + 1)));
  QT_CUDA(cudaMalloc(&p->d_member_col, sizeof(int) * K));
  QT_CUDA(cudaMalloc(&p->d_member_coeff, sizeof(double) * K));
  QT_CUDA(cudaMalloc(&p->d_mask2idx, sizeof(int) * S));
  QT_CUDA(cudaMemcpy(p->d_slot_ptr, slot_ptr.data(), sizeof(int) * (S + 1), cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_member_col, col.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_member_coeff, cf.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_mask2idx, m2i.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
  *plan_out = p;
  return QT_OK;
}

extern "C" int qt_mle_plan_destroy(qt_mle_plan* p) {
  if (!p) return QT_OK;
  cudaFree(p->d_slot_ptr);
  cudaFree(p->d_member_col);
  cudaFree(p->d_member_coeff);
  cudaFree(p->d_member_linw);
  cudaFree(p->d_mask2idx);
  delete p;
  return QT_OK;
}

template <int N>
static int launch_warp(const qt_mle_plan* p, int64_t B, const double* expect, const double* counts, double eps,
                       double entropy_penalty, double beta, double tol, int maxiter, cplx* rho_out, int* iters_out,
                       cudaStream_t st) {
  const bool variants = entropy_penalty > 0.0 || beta > 0.0;
  const size_t per_warp = MleWarpSmem<N>::bytes(variants);
  int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (96 * 1024) / per_warp));
  // A batch that fits the machine in one wave (every block resident at once) runs as long as its fullest SM: with 8 warps
  // per block 2048 experiments land as 16 warps on 108 SMs and 8 on the other 40; smaller blocks spread them 14 / 13.
  while (wpb > 2 && (B + wpb - 1) / wpb < (int64_t)QT_NUM_SMS * 8) wpb /= 2;
  const size_t smem = per_warp * wpb;
  QT_CUDA(cudaFuncSetAttribute(mle_warp_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (B + wpb - 1) / wpb;
  mle_warp_kernel<N><<<(unsigned)blocks, 32 * wpb, smem, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col,
                                                               p->d_member_coeff, p->d_mask2idx, expect, counts, eps,
                                                               entropy_penalty, beta, tol, maxiter, p->unit_coeff,
                                                               rho_out, iters_out);
  return qt_check_launch("mle_warp_kernel");
}

extern "C" int qt_mle_state_batch(const qt_mle_plan* p, int64_t B, const double* expect, const double* counts,
                                  double epsilon, double entropy_penalty, double beta, double tol, int maxiter,
                                  int kernel_variant, void* rho_out, int32_t* iters_out, void* stream) {
  QT_REQUIRE(p, "qt_mle_state_batch: null plan");
  if (B == 0) return QT_OK;
  QT_REQUIRE(expect && rho_out && iters_out, "qt_mle_state_batch: null argument");
  QT_REQUIRE(!(entropy_penalty != 0.0 && beta != 0.0),
             "qt_mle_state_batch: entropy_penalty and beta cannot both be non-zero (tomography.py:225)");
  QT_REQUIRE(beta <= 0.0 || counts, "qt_mle_state_batch: hedged MLE needs counts");
  if (B == 0) return QT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cplx* out = (cplx*)rho_out;
  const bool variants = entropy_penalty > 0.0 || beta > 0.0;
  const bool reg_ok = p->n <= 2 && p->unit_coeff && !variants;
  QT_REQUIRE(kernel_variant != QT_MLE_KERNEL_REGISTER || reg_ok,
             "qt_mle_state_batch: register kernel needs n<=2, unit coefficients, vanilla MLE");
  QT_REQUIRE(kernel_variant != QT_MLE_KERNEL_QUAD || (reg_ok && p->n == 2),
             "qt_mle_state_batch: quad kernel needs n==2, unit coefficients, vanilla MLE");
  // AUTO at n = 2: 4 lanes per experiment while the batch cannot fill the FP64 pipes with one thread each
  // (the quad kernel executes ~1.8x the FP64 instructions per experiment but has 4x the parallelism)
  const bool auto_quad = kernel_variant == QT_MLE_KERNEL_AUTO && B <= 8192;  // measured crossover, profiles/r01_exp_mle_batch.txt
  if (reg_ok && p->n == 2 && (kernel_variant == QT_MLE_KERNEL_QUAD || auto_quad)) {
    const unsigned blocks = (unsigned)((B + 7) / 8);
    mle_quad_kernel<<<blocks, 32, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col, expect, epsilon, tol, maxiter, out,
                                           iters_out);
    return qt_check_launch("mle_quad_kernel");
  }
  if (reg_ok && kernel_variant != QT_MLE_KERNEL_WARP) {
    const unsigned blocks = (unsigned)((B + 31) / 32);
    if (p->n == 1)
      mle_reg_kernel<1><<<blocks, 32, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col, expect, epsilon, tol, maxiter,
                                               out, iters_out);
    else
      mle_reg_kernel<2><<<blocks, 32, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col, expect, epsilon, tol, maxiter,
                                               out, iters_out);
    return qt_check_launch("mle_reg_kernel");
  }
  switch (p->n) {
    case 1: return launch_warp<1>(p, B, expect, counts, epsilon, entropy_penalty, beta, tol, maxiter, out, iters_out, st);
    case 2: return launch_warp<2>(p, B, expect, counts, epsilon, entropy_penalty, beta, tol, maxiter, out, iters_out, st);
    case 3: return launch_warp<3>(p, B, expect, counts, epsilon, entropy_penalty, beta, tol, maxiter, out, iters_out, st);
    case 4: return launch_warp<4>(p, B, expect, counts, epsilon, entropy_penalty, beta, tol, maxiter, out, iters_out, st);
    default: return launch_warp<5>(p, B, expect, counts, epsilon, entropy_penalty, beta, tol, maxiter, out, iters_out, st);
  }
}

extern "C" int qt_mle_step_batch(int n, int64_t B, const double* expect_canon, const void* rho_in, double epsilon,
                                 void* rho_out, void* stream) {
  QT_REQUIRE(n == 1 || n == 2, "qt_mle_step_batch: n must be 1 or 2");
  if (B == 0) return QT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((B + 127) / 128);
  const int K = (1 << (2 * n)) - 1;
  if (n == 1)
    mle_step_kernel<1><<<blocks, 128, 0, st>>>(B, K, expect_canon, (const cplx*)rho_in, epsilon, (cplx*)rho_out);
  else
    mle_step_herm_kernel<<<blocks, 128, 0, st>>>(B, K, expect_canon, (const cplx*)rho_in, epsilon, (cplx*)rho_out);
  return qt_check_launch("mle_step_kernel");
}

extern "C" int qt_linear_inv_state_batch(const qt_mle_plan* p, int64_t B, const double* expect, void* rho_out,
                                         void* stream) {
  QT_REQUIRE(p, "qt_linear_inv_state_batch: null plan");
  if (B == 0) return QT_OK;
  QT_REQUIRE(expect && rho_out, "qt_linear_inv_state_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 16);
  const int threads = p->n <= 2 ? 32 : 256;
#define LAUNCH(N)                                                                                                  \
  linear_inv_kernel<N><<<blocks, threads, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col, p->d_member_linw, \
                                                   p->d_mask2idx, expect, (cplx*)rho_out)
  switch (p->n) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    default: LAUNCH(5); break;
  }
#undef LAUNCH
  return qt_check_launch("linear_inv_kernel");
}

extern "C" int qt_state_log_likelihood_batch(const qt_mle_plan* p, int64_t B, const void* rho, const double* expect,
                                             const double* counts, double* ll_out, void* stream) {
  QT_REQUIRE(p, "qt_state_log_likelihood_batch: null plan");
  if (B == 0) return QT_OK;
  QT_REQUIRE(rho && expect && counts && ll_out, "qt_state_log_likelihood_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)std::min<int64_t>(B, (int64_t)QT_NUM_SMS * 16);
  const int threads = 256;
#define LAUNCH(N)                                                                                                   \
  log_likelihood_kernel<N><<<blocks, threads, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col, p->d_member_coeff, \
                                                       p->d_mask2idx, (const cplx*)rho, expect, counts, ll_out)
#define LAUNCH_PACKED(N)                                                                                            \
  do {                                                                                                              \
    const int64_t per_block = 8 * (32 >> (2 * N));                                                                  \
    const unsigned pb = (unsigned)std::min<int64_t>((B + per_block - 1) / per_block, (int64_t)QT_NUM_SMS * 8);      \
    log_likelihood_packed_kernel<N><<<pb, 256, 0, st>>>(B, p->K, p->d_slot_ptr, p->d_member_col,                   \
                                                        p->d_member_coeff, p->d_mask2idx, (const cplx*)rho, expect, \
                                                        counts, ll_out);                                            \
  } while (0)
  switch (p->n) {
    case 1: LAUNCH_PACKED(1); break;
    case 2: LAUNCH_PACKED(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    default: LAUNCH(5); break;
  }
#undef LAUNCH
#undef LAUNCH_PACKED
  return qt_check_launch("log_likelihood_kernel");
}
