// Batched complex-Hermitian eigensolver: parallel cyclic two-sided Jacobi, matrix resident in shared
// memory, one matrix per warp (M <= 32) or per thread block (M = 64 ...).
//
// Stands in for the LAPACK calls the reference makes through scipy.linalg.eigh / np.linalg.eigh
// (operator_tools/project_superoperators.py:31, calculational.py:88, superoperator_transformations.py:334).
// Jacobi is used because every rotation of a round-robin step is independent (M/2 per step), the whole
// matrix lives on-chip, and its results (V max(L,0) V^dagger, sum sqrt(l)) are gauge-independent, so
// they agree with LAPACK to ~1e-14.
#pragma once
#include "qt_common.cuh"

// Round-robin ("circle") pairing: M players, step s in [0, M-1), pair i in [0, M/2).
__device__ __forceinline__ void rr_pair(int M, int s, int i, int& p, int& q) {
  const int m1 = M - 1;
  int a = s + i;
  if (a >= m1) a -= m1;
  int b2;
  if (i == 0) {
    b2 = m1;
  } else {
    b2 = s - i;
    if (b2 < 0) b2 += m1;
  }
  p = a < b2 ? a : b2;
  q = a < b2 ? b2 : a;
}

// Rotation J = [[c, s],[-conj(s), c]] (c real) that diagonalises [[alpha, beta],[conj(beta), gamma]].
__device__ __forceinline__ void jacobi_rotation(double alpha, double gamma, cplx beta, double& c, cplx& s,
                                                double& alpha_new, double& gamma_new) {
  const double ab2 = cabs2(beta);
  const double scale = fabs(alpha) + fabs(gamma);
  if (ab2 <= 1e-36 * scale * scale || ab2 == 0.0) {
    c = 1.0;
    s = cmake(0.0, 0.0);
    alpha_new = alpha;
    gamma_new = gamma;
    return;
  }
  const double ab = sqrt(ab2);
  const double tau = (gamma - alpha) / (2.0 * ab);
  const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  c = rsqrt(1.0 + t * t);
  const double sn = t * c;
  const double inv = sn / ab;
  s = cmake(beta.x * inv, beta.y * inv);  // s * e^{i phi}
  alpha_new = alpha - t * ab;
  gamma_new = gamma + t * ab;
}

struct SyncWarp {
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};
struct SyncBlock {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};

// Group-wide sum for NT threads (NT == 32: shuffle; else via shared scratch `red` of >= NT/32 doubles).
template <int NT, class Sync>
__device__ __forceinline__ double group_sum(double v, double* red, int tid) {
  v = warp_sum(v);
  if (NT == 32) return v;
  Sync::sync();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  Sync::sync();
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) tot += red[i];
  return tot;
}

// Scratch layout (doubles, 16-byte aligned base): rs[M/2] (complex) | rc[M/2] | red[32]
template <int M>
struct JacobiScratch {
  static constexpr int doubles = M / 2 + M + 32;
};

// A: M x M Hermitian in shared memory (row-major, leading dimension M), overwritten (diagonal = eigenvalues).
// V: M x M in shared memory; on exit column k is the eigenvector of ev[k].  If `init_v` the routine
//    starts from V = I; otherwise V is taken as given and A must already be expressed in that basis
//    (warm start: A = V0^dagger A0 V0).
// Returns the number of sweeps performed.
template <int M, int NT, class Sync, bool WANT_V>
__device__ int jacobi_eigh(cplx* A, cplx* V, double* ev, double* scratch, int tid, bool init_v = true,
                           int max_sweeps = 30) {
  constexpr int HP = M / 2;
  cplx* rs = reinterpret_cast<cplx*>(scratch);  // scratch must be 16-byte aligned
  double* rc = scratch + M;
  double* red = scratch + M + HP;

  if (WANT_V && init_v) {
    for (int e = tid; e < M * M; e += NT) V[e] = cmake((e / M == e % M) ? 1.0 : 0.0, 0.0);
  }
  Sync::sync();
  if (M == 1) {
    if (tid == 0) ev[0] = A[0].x;
    Sync::sync();
    return 0;
  }
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    double off = 0.0, tot = 0.0;
    for (int e = tid; e < M * M; e += NT) {
      const double a2 = cabs2(A[e]);
      tot += a2;
      if (e / M != e % M) off += a2;
    }
    off = group_sum<NT, Sync>(off, red, tid);
    tot = group_sum<NT, Sync>(tot, red, tid);
    if (off <= (1e-30 * M * M) * tot || tot == 0.0) break;
    for (int step = 0; step < M - 1; ++step) {
      // ---- rotation parameters, one pair per thread ----
      for (int i = tid; i < HP; i += NT) {
        int p, q;
        rr_pair(M, step, i, p, q);
        double c, an, gn;
        cplx s;
        jacobi_rotation(A[p * M + p].x, A[q * M + q].x, A[p * M + q], c, s, an, gn);
        rc[i] = c;
        rs[i] = s;
      }
      Sync::sync();
      // ---- A <- J^dagger A J, one 2x2 block (pair I rows, pair J cols) per work item ----
      for (int w = tid; w < HP * HP; w += NT) {
        const int I = w / HP, J = w % HP;
        int pi, qi, pj, qj;
        rr_pair(M, step, I, pi, qi);
        rr_pair(M, step, J, pj, qj);
        const double cI = rc[I], cJ = rc[J];
        const cplx sI = rs[I], sJ = rs[J];
        const cplx b00 = A[pi * M + pj], b01 = A[pi * M + qj], b10 = A[qi * M + pj], b11 = A[qi * M + qj];
        // X = B J_J
        const cplx csJ = cconj(sJ);
        cplx x00 = csub(cscale(b00, cJ), cmul(csJ, b01));
        cplx x01 = cadd(cmul(sJ, b00), cscale(b01, cJ));
        cplx x10 = csub(cscale(b10, cJ), cmul(csJ, b11));
        cplx x11 = cadd(cmul(sJ, b10), cscale(b11, cJ));
        // Y = J_I^dagger X
        const cplx csI = cconj(sI);
        cplx y00 = csub(cscale(x00, cI), cmul(sI, x10));
        cplx y01 = csub(cscale(x01, cI), cmul(sI, x11));
        cplx y10 = cadd(cmul(csI, x00), cscale(x10, cI));
        cplx y11 = cadd(cmul(csI, x01), cscale(x11, cI));
        if (I == J) {  // exact diagonal block: real diagonal, zero off-diagonal
          y00.y = 0.0;
          y11.y = 0.0;
          y01 = cmake(0.0, 0.0);
          y10 = cmake(0.0, 0.0);
        }
        A[pi * M + pj] = y00;
        A[pi * M + qj] = y01;
        A[qi * M + pj] = y10;
        A[qi * M + qj] = y11;
      }
      // ---- V <- V J ----
      if (WANT_V) {
        for (int w = tid; w < M * HP; w += NT) {
          const int r = w / HP, J = w % HP;
          int pj, qj;
          rr_pair(M, step, J, pj, qj);
          const double cJ = rc[J];
          const cplx sJ = rs[J];
          const cplx v0 = V[r * M + pj], v1 = V[r * M + qj];
          V[r * M + pj] = csub(cscale(v0, cJ), cmul(cconj(sJ), v1));
          V[r * M + qj] = cadd(cmul(sJ, v0), cscale(v1, cJ));
        }
      }
      Sync::sync();
    }
  }
  for (int k = tid; k < M; k += NT) ev[k] = A[k * M + k].x;
  Sync::sync();
  return sweep;
}

// warp convenience wrapper used by the MLE variants and the distance kernels (scratch must hold
// JacobiScratch<D>::doubles doubles, placed by the caller right after `ev`).
template <int D>
__device__ __forceinline__ int jacobi_eigh_warp(cplx* A, cplx* V, double* ev, int lane) {
  return jacobi_eigh<D, 32, SyncWarp, true>(A, V, ev, ev + D, lane);
}
