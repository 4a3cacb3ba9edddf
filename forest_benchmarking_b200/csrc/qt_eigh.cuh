// Batched complex-Hermitian eigensolver: parallel cyclic two-sided Jacobi, matrix resident in shared
// memory, one matrix per warp (M <= 32) or per thread block (M = 64).
//
// Stands in for the LAPACK calls the reference makes through scipy.linalg.eigh / np.linalg.eigh
// (operator_tools/project_superoperators.py:31, calculational.py:88, superoperator_transformations.py:334).
// Jacobi is used because every rotation of a round-robin step is independent (M/2 per step), the whole
// matrix lives on-chip, it can be WARM-STARTED from a previous eigenbasis (the Dykstra / PGD loops
// decompose a slowly moving matrix hundreds of times), and its results (V max(L,0) V^dagger, sum sqrt(l))
// are gauge-independent, so they agree with LAPACK to ~1e-14.
//
// Step pipeline (one round-robin step = M/2 disjoint rotations):
//     A <- J^dagger A J   (all threads)                                   | sync
//     rotation parameters of step+1 (warp 0)  ||  V <- V J (other warps)  | sync
// so the latency-bound parameter phase (rsqrt / reciprocal chains) hides behind the V update.
// Matrices may have a padded leading dimension LD (M + 1 for M >= 16): column accesses of the
// recomposition / basis-change products are then bank-conflict free.
#pragma once
#include <type_traits>
#include "qt_common.cuh"

// Round-robin ("circle") pairing: M players, step s in [0, M-1), pair i in [0, M/2):
//   first(i) = (s + i) mod (M-1),   second(0) = M-1,   second(i) = (s - i) mod (M-1).
// The pair is NOT ordered by index: consecutive pairs then touch consecutive (ascending / descending)
// rows and columns, which keeps the 16-byte shared-memory accesses of a warp on distinct banks.
__device__ __forceinline__ void rr_pair(int M, int s, int i, int& p, int& q) {
  const int m1 = M - 1;
  p = s + i;
  if (p >= m1) p -= m1;
  q = s - i;
  if (q < 0) q += m1;
  if (i == 0) q = m1;
}

// Rotation J = [[c, s],[-conj(s), c]] (c real) that diagonalises [[alpha, beta],[conj(beta), gamma]];
// the rotated diagonal is (alpha - t|beta|, gamma + t|beta|).
// Square roots and the division go through MUFU seeds + Newton steps (rsqrt / fast_rcp): a correctly
// rounded DDIV / DSQRT costs ~130 issue cycles per warp on B200 (profiles/r01_ubench_fp64.txt).
__device__ __forceinline__ void jacobi_rotation(double alpha, double gamma, cplx beta, double& c, cplx& s,
                                                double& alpha_new, double& gamma_new) {
  // branch-free (the caller's instruction stream stays one basic block, so independent work can be scheduled
  // into the latency of this dependent chain): a negligible off-diagonal element runs the arithmetic on a
  // dummy value and selects the identity at the end
  const double ab2 = cabs2(beta);
  const double scale = fabs(alpha) + fabs(gamma);
  const bool skip = (ab2 <= 1e-36 * scale * scale) || (ab2 == 0.0);
  const double ab2s = skip ? 1.0 : ab2;
  // half-angle form (two dependent rsqrt instead of rsqrt -> rsqrt -> rcp -> rsqrt):
  //   cos 2t = |D| / r,  r = sqrt(D^2 + 4|b|^2),  D = gamma - alpha   (|t| <= pi/4, sign t = sign D)
  //   c = sqrt((1 + cos 2t) / 2),   sin t = sign(D) |b| / (r c),   t|b| = sign(D) |b|^2 / (r c^2)
  const double dlt = gamma - alpha;
  const double ir = fast_rsqrt(fma(dlt, dlt, 4.0 * ab2s));  // 1 / r
  const double h = fma(0.5 * fabs(dlt), ir, 0.5);           // c^2
  const double ic = fast_rsqrt(h);                           // 1 / c
  const double sg = (dlt >= 0.0) ? ir * ic : -(ir * ic);     // sign(D) / (r c)
  const double tab = ab2s * sg * ic;                         // tan(t) |beta|
  c = skip ? 1.0 : h * ic;
  s = skip ? cmake(0.0, 0.0) : cmake(beta.x * sg, beta.y * sg);  // sin(t) e^{i phi}
  alpha_new = skip ? alpha : alpha - tab;
  gamma_new = skip ? gamma : gamma + tab;
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

// Scratch layout (doubles, 16-byte aligned base): rs[2][M/2] (complex) | rc[2][M/2] | red[32]
// (ring solvers: red[16] | .. | mbarrier at +32)
template <int M>
struct JacobiScratch {
  static constexpr int doubles = 3 * M + 32;
};

// ABL: ablation mask for scripts/ubench_jacobi.cu only (1: no V update, 2: no block update, 4: no rotation
// chain, 8: no barrier); 0 in every product instantiation.
template <int M, int NT, int LD, bool WANT_V, int ABL = 0>
__device__ int jacobi_eigh_ring(cplx* A, cplx* V, double* ev, double* scratch, int tid, bool init_v,
                                int max_sweeps, double rel2, bool clean_a = false);

// A: M x M Hermitian in shared memory (row-major, leading dimension LD), overwritten (diagonal = eigenvalues).
//    Both triangles are stored and kept exactly conjugate: only the blocks above the block diagonal are
//    computed, their mirror images are written as conjugates; the 2x2 diagonal blocks are set analytically.
// V: M x M in shared memory (leading dimension LD); on exit column k is the eigenvector of ev[k].  If
//    `init_v` the routine starts from V = I; otherwise V is taken as given and A must already be expressed
//    in that basis (warm start: A = V0^dagger A0 V0).
// Stops when  off(A)^2 <= rel2 * ||A||_F^2  (rel2 <= 0: the default 1e-30 * M^2, i.e. a relative off-diagonal
// Frobenius norm of 1e-15 * M).  Returns the number of sweeps performed.
// clean_a (ring solver only): on exit A is the full Hermitian matrix V^dagger A0 V the iteration arrived at -- both
// triangles exactly conjugate, eigenvalue estimates on the diagonal, the NOT yet annihilated remainder off it -- so
// that a caller who stops early (large rel2) can use the remainder (first-order correction, next warm start).
template <int M, int NT, class Sync, bool WANT_V, int LD = M>
__device__ int jacobi_eigh(cplx* A, cplx* V, double* ev, double* scratch, int tid, bool init_v = true,
                           int max_sweeps = 30, double rel2 = 0.0, bool clean_a = false) {
  if constexpr ((M == 64 && (NT == 512 || NT == 256) && WANT_V) || ((M == 16 || M == 8) && NT == 32)) {
    return jacobi_eigh_ring<M, NT, LD, WANT_V>(A, V, ev, scratch, tid, init_v, max_sweeps, rel2, clean_a);
  }
  constexpr int HP = (M / 2 > 0) ? M / 2 : 1;
  constexpr int NOFF = HP * (HP - 1) / 2;                        // 2x2 blocks above the block diagonal
  constexpr int NBA = (NOFF + NT - 1) / NT > 0 ? (NOFF + NT - 1) / NT : 1;
  constexpr bool SPLIT = (NT > 32) && WANT_V;                    // warp 0: parameters, other warps: V update
  constexpr int VNT = SPLIT ? NT - 32 : NT;
  static_assert(NT == 32 || HP <= 32, "parameter phase is one warp wide");
  static_assert(VNT % HP == 0, "a thread keeps one column pair across its V rows");
  constexpr int RSTRIDE = VNT / HP;                              // V rows covered per round
  constexpr int NBV = (M + RSTRIDE - 1) / RSTRIDE;
  cplx* rs = reinterpret_cast<cplx*>(scratch);  // scratch must be 16-byte aligned
  double* rc = scratch + 4 * HP;
  double* red = rc + 2 * HP;

  if (WANT_V && init_v) {
    for (int e = tid; e < M * M; e += NT) V[(e / M) * LD + e % M] = cmake((e / M == e % M) ? 1.0 : 0.0, 0.0);
  }
  Sync::sync();
  if (M == 1) {
    if (tid == 0) ev[0] = A[0].x;
    Sync::sync();
    return 0;
  }
  // static assignment of the off-diagonal blocks (I < J) to threads: w -> (I, J) in row-major triangular order
  int bI[NBA], bJ[NBA];
#pragma unroll
  for (int k = 0; k < NBA; ++k) {
    int w = tid + k * NT, I = -1, J = 0;
    if (w < NOFF) {
      I = 0;
      while (w >= HP - 1 - I) {
        w -= HP - 1 - I;
        ++I;
      }
      J = I + 1 + w;
    }
    bI[k] = I;
    bJ[k] = J;
  }
  const int vt = SPLIT ? tid - 32 : tid;
  const int vJ = vt % HP, vr0 = vt / HP;

  // rotation parameters of `step` into buffer `buf`; the pair's own 2x2 diagonal block is finished here
  auto params = [&](int step, int buf) {
    for (int i = tid; i < HP; i += NT) {
      int p, q;
      rr_pair(M, step, i, p, q);
      double c, an, gn;
      cplx s;
      jacobi_rotation(A[p * LD + p].x, A[q * LD + q].x, A[p * LD + q], c, s, an, gn);
      rc[buf * HP + i] = c;
      rs[buf * HP + i] = s;
      A[p * LD + p] = cmake(an, 0.0);
      A[q * LD + q] = cmake(gn, 0.0);
      A[p * LD + q] = cmake(0.0, 0.0);
      A[q * LD + p] = cmake(0.0, 0.0);
    }
  };
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    double off = 0.0, tot = 0.0;
    for (int e = tid; e < M * M; e += NT) {
      const int r = e / M, c = e % M;
      const double a2 = cabs2(A[r * LD + c]);
      tot += a2;
      if (r != c) off += a2;
    }
    off = group_sum<NT, Sync>(off, red, tid);
    tot = group_sum<NT, Sync>(tot, red, tid);
    if (off <= (rel2 > 0.0 ? rel2 : 1e-30 * M * M) * tot || tot == 0.0) break;
    Sync::sync();  // the norm pass above read all of A (a warp-wide group_sum is shuffle-only): order it before
    params(0, 0);  // the diagonal-block writes of the first parameter phase (racecheck: WAR hazard otherwise)
    Sync::sync();
    for (int step = 0; step < M - 1; ++step) {
      const int cur = step & 1;
      const double* rcc = rc + cur * HP;
      const cplx* rsc = rs + cur * HP;
      // ---- A <- J^dagger A J on the blocks above the block diagonal (+ conjugate mirror) ----
#pragma unroll
      for (int k = 0; k < NBA; ++k) {
        const int I = bI[k], J = bJ[k];
        if (I < 0) continue;
        int pi, qi, pj, qj;
        rr_pair(M, step, I, pi, qi);
        rr_pair(M, step, J, pj, qj);
        const cplx b00 = A[pi * LD + pj], b01 = A[pi * LD + qj], b10 = A[qi * LD + pj], b11 = A[qi * LD + qj];
        const double cI = rcc[I], cJ = rcc[J];
        const cplx sI = rsc[I], sJ = rsc[J];
        // X = B J_J
        const cplx csJ = cconj(sJ);
        const cplx x00 = csub(cscale(b00, cJ), cmul(csJ, b01));
        const cplx x01 = cadd(cmul(sJ, b00), cscale(b01, cJ));
        const cplx x10 = csub(cscale(b10, cJ), cmul(csJ, b11));
        const cplx x11 = cadd(cmul(sJ, b10), cscale(b11, cJ));
        // Y = J_I^dagger X
        const cplx csI = cconj(sI);
        const cplx y00 = csub(cscale(x00, cI), cmul(sI, x10));
        const cplx y01 = csub(cscale(x01, cI), cmul(sI, x11));
        const cplx y10 = cadd(cmul(csI, x00), cscale(x10, cI));
        const cplx y11 = cadd(cmul(csI, x01), cscale(x11, cI));
        A[pi * LD + pj] = y00;
        A[pi * LD + qj] = y01;
        A[qi * LD + pj] = y10;
        A[qi * LD + qj] = y11;
        A[pj * LD + pi] = cconj(y00);
        A[qj * LD + pi] = cconj(y01);
        A[pj * LD + qi] = cconj(y10);
        A[qj * LD + qi] = cconj(y11);
      }
      Sync::sync();
      // ---- parameters of the next step (first warp)  ||  V <- V J (the other warps) ----
      if (!SPLIT || tid < 32) {
        if (step + 1 < M - 1) params(step + 1, cur ^ 1);
      }
      if (WANT_V && (!SPLIT || tid >= 32) && vr0 < M) {
        int pj, qj;
        rr_pair(M, step, vJ, pj, qj);
        const double cJ = rcc[vJ];
        const cplx sJ = rsc[vJ], csJ = cconj(sJ);
#pragma unroll
        for (int k0 = 0; k0 < NBV; k0 += 4) {
          cplx v0[4], v1[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int r = vr0 + (k0 + kk) * RSTRIDE;
            if (k0 + kk < NBV && r < M) {
              v0[kk] = V[r * LD + pj];
              v1[kk] = V[r * LD + qj];
            }
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int r = vr0 + (k0 + kk) * RSTRIDE;
            if (k0 + kk < NBV && r < M) {
              V[r * LD + pj] = csub(cscale(v0[kk], cJ), cmul(csJ, v1[kk]));
              V[r * LD + qj] = cadd(cmul(sJ, v0[kk]), cscale(v1[kk], cJ));
            }
          }
        }
      }
      Sync::sync();
    }
  }
  for (int k = tid; k < M; k += NT) ev[k] = A[k * LD + k].x;
  Sync::sync();
  return sweep;
}

// Split-phase CTA barrier on an mbarrier in shared memory: arrive (release) right after the last shared-memory
// store of a step, wait (acquire) right before the first load of the next one; the register-only V update sits
// in between, so a warp that is ahead does useful work instead of idling at the barrier.
__device__ __forceinline__ void mbar_init(double* slot, int count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(double* slot) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(double* slot, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(a), "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Ring eigensolver: ONE barrier per round-robin step, rotations / diagonal / V in registers.
//   M = 64 with 256 or 512 threads (one matrix per block, split-phase mbarrier), and
//   M = 16 or 8 with 32 threads (one matrix per warp, __syncwarp).
//
// The generic routine above spends most of its time at its two barriers per step and in the serial
// parameter phase (profiles/r01_ncu_pgdb3_kernel_v5.md: barrier 23 % of samples, FP64 pipe 29 % busy).  Here,
// with HP = M/2 pairs per step and the NT threads split into NT/HP segments of HP consecutive threads:
//  * every segment computes the HP rotations of the step itself (thread pr of a segment = pair pr; redundant
//    across segments, ~60 DFMA-class instructions per thread) from the pair's off-diagonal element and its
//    diagonal entries -- no separate parameter phase, no second barrier; a thread fetches the rotations of its
//    block (I, J) by shuffle;
//  * diagonal entries and the eigenvector matrix V live in REGISTERS (segment g: rows RV g .. RV g + RV-1 of V,
//    thread pr: the two columns of pair pr) and travel one position along the round-robin ring after each step
//    (p_i <- p_{i+1}, q_i <- q_{i-1}, q_{HP-1} -> p_{HP-1}, p_0 -> q_1, index M-1 stays in pair 0);
//  * of A only the 2x2 blocks above the block diagonal are kept (HP (HP-1) / 2 blocks, one or two per thread):
//    element (x, y) is valid at the position written by the block that last produced it.  One step later that is
//    the direct position for every element a block needs except (p_I, q_J) with J = I+1, which is read
//    transposed, and the three elements whose two indices formed a pair in the previous step (annihilated: read
//    as zero).  A's diagonal is NOT maintained: the eigenvalues are returned in `ev` only.
// ---------------------------------------------------------------------------------------------
template <int M, int NT, int LD, bool WANT_V, int ABL>
__device__ int jacobi_eigh_ring(cplx* A, cplx* V, double* ev, double* scratch, int tid, bool init_v,
                                int max_sweeps, double rel2, bool clean_a) {
  constexpr int HP = M / 2, M1 = M - 1, NOFF = HP * (HP - 1) / 2;
  constexpr bool WARP = (NT == 32);               // one matrix per warp: __syncwarp instead of the mbarrier
  static_assert(HP <= 32 && (HP & (HP - 1)) == 0 && NT % HP == 0 && (WARP || HP == 32), "segment = shuffle group");
  constexpr int NSEG = NT / HP, RV = WANT_V ? M / NSEG : 1;  // V rows per segment
  static_assert(!WANT_V || (M % NSEG == 0 && M >= NSEG), "V rows divide over the segments");
  constexpr int NB = (NOFF + NT - 1) / NT;        // 2x2 blocks per thread
  using Sync = typename std::conditional<WARP, SyncWarp, SyncBlock>::type;
  double* red = scratch;  // >= 16 doubles (block-wide variant only)
  const int lane = tid & 31, pr = tid % HP, seg = tid / HP;
  if (WANT_V && init_v) {
    for (int e = tid; e < M * M; e += NT) V[(e / M) * LD + e % M] = cmake((e / M == e % M) ? 1.0 : 0.0, 0.0);
  }
  Sync::sync();
  // static block assignment: block w = tid + k NT < NOFF -> (I, J), I < J, row-major triangular order; spare
  // slots run the same instructions on element (0, 0) with their stores predicated off (no branches in the step)
  bool has_block[NB], near1[NB], near2[NB], last2[NB], first2[NB];
  int bI[NB], bJ[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    int w = tid + k * NT, I = 0, J = 1;
    has_block[k] = w < NOFF;
    if (has_block[k]) {
      while (w >= HP - 1 - I) {
        w -= HP - 1 - I;
        ++I;
      }
      J = I + 1 + w;
    }
    bI[k] = I;
    bJ[k] = J;
    near1[k] = (J == I + 1);
    near2[k] = (J == I + 2);
    last2[k] = (I == HP - 2);
    first2[k] = (I == 0 && J == 1);
  }
  // V in registers, columns in the arrangement of step 0: pair i = (i, M-1 - i), pair 0 = (0, M-1)
  const int col0q = (pr == 0) ? M1 : M1 - pr;
  cplx vp[RV], vq[RV];
  if constexpr (WANT_V) {
#pragma unroll
    for (int k = 0; k < RV; ++k) {
      vp[k] = V[(RV * seg + k) * LD + pr];
      vq[k] = V[(RV * seg + k) * LD + col0q];
    }
  }
  double dp = A[pr * LD + pr].x, dq = A[col0q * LD + col0q].x;
  bool fresh = true;  // both triangles of A valid, nothing annihilated yet
  // split-phase barrier state: `pending` = an arrive of this thread has not been matched by a wait yet
  double* mbar = scratch + 32;
  unsigned parity = 0;
  bool pending = false;
  if constexpr (!WARP) {
    if (tid == 0) mbar_init(mbar, NT);
    __syncthreads();
  }
  // QT_JACOBI_PLAIN_BARRIER (validation builds only, scripts/ubench_jacobi.cu): the split-phase barrier becomes a
  // plain __syncthreads() at the wait point, which compute-sanitizer's racecheck can follow (it does not model
  // mbarrier ordering and reports every step-to-step dependency of the default build as a hazard).
  auto wait_pending = [&]() {
    if (pending) {
      if constexpr (WARP) {
        __syncwarp();
      } else {
#ifdef QT_JACOBI_PLAIN_BARRIER
        __syncthreads();
#else
        if constexpr (!(ABL & 8)) mbar_wait(mbar, parity);
#endif
      }
      parity ^= 1u;
      pending = false;
    }
  };

  auto v_update = [&](double c_prev, cplx s_prev) {  // V <- V J, then every column moves one ring position
    const cplx cs = cconj(s_prev);
#pragma unroll
    for (int k = 0; k < RV; ++k) {
      const cplx np = csub(cscale(vp[k], c_prev), cmul(cs, vq[k]));
      const cplx nq = cadd(cmul(s_prev, vp[k]), cscale(vq[k], c_prev));
      const cplx up = (pr == 0) ? np : nq;  // pair 0 hands its first column to pair 1's second slot
      cplx rp, rq;
      rp.x = __shfl_down_sync(0xffffffffu, np.x, 1, HP);
      rp.y = __shfl_down_sync(0xffffffffu, np.y, 1, HP);
      rq.x = __shfl_up_sync(0xffffffffu, up.x, 1, HP);
      rq.y = __shfl_up_sync(0xffffffffu, up.y, 1, HP);
      vp[k] = (pr == HP - 1) ? nq : rp;
      vq[k] = (pr == 0) ? nq : rq;
    }
  };

  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    // ---- off-diagonal / total Frobenius mass from the valid elements (arrangement of step 0) ----
    wait_pending();
    int p = pr, q = col0q;                         // this thread's pair
    int pi[NB], qi[NB], pj[NB], qj[NB];            // this thread's blocks
    {
      double off = 0.0;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        rr_pair(M, 0, bI[k], pi[k], qi[k]);
        rr_pair(M, 0, bJ[k], pj[k], qj[k]);
        if (!has_block[k]) pi[k] = qi[k] = pj[k] = qj[k] = 0;
        cplx b00 = A[pi[k] * LD + pj[k]];
        cplx b01 = near1[k] ? A[qj[k] * LD + pi[k]] : A[pi[k] * LD + qj[k]];
        const cplx b10 = A[qi[k] * LD + pj[k]];
        cplx b11 = A[qi[k] * LD + qj[k]];
        if (!fresh && near2[k]) b01 = cmake(0.0, 0.0);
        if (!fresh && last2[k]) b00 = cmake(0.0, 0.0);
        if (!fresh && first2[k]) b11 = cmake(0.0, 0.0);
        if (has_block[k]) off += cabs2(b00) + cabs2(b01) + cabs2(b10) + cabs2(b11);
      }
      if (seg == 0) off += cabs2(A[q * LD + p]);  // the off-diagonal element of pair `pr` itself
      const double dg = (seg == 0) ? dp * dp + dq * dq : 0.0;
      off = 2.0 * group_sum<NT, Sync>(off, red, tid);
      const double tot = off + group_sum<NT, Sync>(dg, red, tid);
      if (off <= (rel2 > 0.0 ? rel2 : 1e-30 * M * M) * tot || tot == 0.0) break;
    }
#pragma unroll 3
    for (int step = 0; step < M1; ++step) {
      // ---- loads of this step: the pair's off-diagonal element and the thread's 2x2 blocks ----
      wait_pending();
      const cplx beta = cconj(A[q * LD + p]);
      cplx b00[NB], b01[NB], b10[NB], b11[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        b00[k] = A[pi[k] * LD + pj[k]];
        b01[k] = A[near1[k] ? qj[k] * LD + pi[k] : pi[k] * LD + qj[k]];  // one load, address selected
        b01[k].y = near1[k] ? -b01[k].y : b01[k].y;
        b10[k] = A[qi[k] * LD + pj[k]];
        b11[k] = A[qi[k] * LD + qj[k]];
        b01[k] = (!fresh && near2[k]) ? cmake(0.0, 0.0) : b01[k];
        b00[k] = (!fresh && last2[k]) ? cmake(0.0, 0.0) : b00[k];
        b11[k] = (!fresh && first2[k]) ? cmake(0.0, 0.0) : b11[k];
      }
      // ---- rotation of pair `pr` (every segment computes all HP) ----
      double c, an, gn;
      cplx s;
      if constexpr (ABL & 4) {
        c = 0.8;
        s = cmake(0.36 + 1e-9 * beta.x, 0.48);
        an = dp;
        gn = dq;
      } else {
        jacobi_rotation(dp, dq, beta, c, s, an, gn);
      }
      // ---- A <- J^dagger A J on this thread's blocks (rotations of pairs I and J fetched by shuffle from the
      //      first segment of the warp: lane index = pair index) ----
#pragma unroll
      for (int k = 0; k < ((ABL & 2) ? 0 : NB); ++k) {
        double cI, cJ;
        cplx sI, sJ;
        if constexpr (ABL & 16) {
          cI = c; cJ = c; sI = s; sJ = cconj(s);
        } else {
          cI = __shfl_sync(0xffffffffu, c, bI[k]);
          cJ = __shfl_sync(0xffffffffu, c, bJ[k]);
          sI.x = __shfl_sync(0xffffffffu, s.x, bI[k]);
          sI.y = __shfl_sync(0xffffffffu, s.y, bI[k]);
          sJ.x = __shfl_sync(0xffffffffu, s.x, bJ[k]);
          sJ.y = __shfl_sync(0xffffffffu, s.y, bJ[k]);
        }
        const cplx csJ = cconj(sJ), csI = cconj(sI);
        const cplx x00 = csub(cscale(b00[k], cJ), cmul(csJ, b01[k]));
        const cplx x01 = cadd(cmul(sJ, b00[k]), cscale(b01[k], cJ));
        const cplx x10 = csub(cscale(b10[k], cJ), cmul(csJ, b11[k]));
        const cplx x11 = cadd(cmul(sJ, b10[k]), cscale(b11[k], cJ));
        const cplx y00 = csub(cscale(x00, cI), cmul(sI, x10));
        const cplx y01 = csub(cscale(x01, cI), cmul(sI, x11));
        const cplx y10 = cadd(cmul(csI, x00), cscale(x10, cI));
        const cplx y11 = cadd(cmul(csI, x01), cscale(x11, cI));
        if (has_block[k] && (!(ABL & 32) || y00.x == 1.2345)) {
          A[pi[k] * LD + pj[k]] = y00;
          A[pi[k] * LD + qj[k]] = y01;
          A[qi[k] * LD + pj[k]] = y10;
          A[qi[k] * LD + qj[k]] = y11;
        }
      }
      // ---- all shared-memory stores of the step are issued: arrive, then do the register-only work ----
      if constexpr (!WARP) {
#ifndef QT_JACOBI_PLAIN_BARRIER
        if constexpr (!(ABL & 8)) mbar_arrive(mbar);
#endif
      }
      pending = true;
      if constexpr (WANT_V && !(ABL & 1)) v_update(c, s);
      // ---- diagonal entries move along the ring; next step's indices ----
      {
        const double upd = (pr == 0) ? an : gn;
        const double rp = __shfl_down_sync(0xffffffffu, an, 1, HP), rq = __shfl_up_sync(0xffffffffu, upd, 1, HP);
        dp = (pr == HP - 1) ? gn : rp;
        dq = (pr == 0) ? gn : rq;
      }
      fresh = false;
      const int nxt = (step + 1 == M1) ? 0 : step + 1;
      rr_pair(M, nxt, pr, p, q);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        rr_pair(M, nxt, bI[k], pi[k], qi[k]);
        rr_pair(M, nxt, bJ[k], pj[k], qj[k]);
        if (!has_block[k]) pi[k] = qi[k] = pj[k] = qj[k] = 0;  // spare slot: read A[0][0], which no step writes
      }
    }
  }
  wait_pending();
  // a whole number of sweeps returns every column to its step-0 slot
  if (clean_a) {
    // Every unordered index pair is valid at exactly one position, owned by one thread (its 2x2 blocks, or the
    // pair element of segment 0): re-read it as the norm pass does and write both triangles.  A thread only touches
    // the two positions of pairs it owns, so no barrier is needed before the writes.
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (!has_block[k]) continue;
      int pi, qi, pj, qj;
      rr_pair(M, 0, bI[k], pi, qi);
      rr_pair(M, 0, bJ[k], pj, qj);
      cplx b00 = A[pi * LD + pj];
      cplx b01 = near1[k] ? cconj(A[qj * LD + pi]) : A[pi * LD + qj];
      const cplx b10 = A[qi * LD + pj];
      cplx b11 = A[qi * LD + qj];
      if (!fresh && near2[k]) b01 = cmake(0.0, 0.0);
      if (!fresh && last2[k]) b00 = cmake(0.0, 0.0);
      if (!fresh && first2[k]) b11 = cmake(0.0, 0.0);
      A[pi * LD + pj] = b00; A[pj * LD + pi] = cconj(b00);
      A[pi * LD + qj] = b01; A[qj * LD + pi] = cconj(b01);
      A[qi * LD + pj] = b10; A[pj * LD + qi] = cconj(b10);
      A[qi * LD + qj] = b11; A[qj * LD + qi] = cconj(b11);
    }
    if (seg == 0) {
      const cplx v = A[col0q * LD + pr];
      A[pr * LD + col0q] = cconj(v);
      A[pr * LD + pr] = cmake(dp, 0.0);
      A[col0q * LD + col0q] = cmake(dq, 0.0);
    }
  }
  if constexpr (WANT_V) {
#pragma unroll
    for (int k = 0; k < RV; ++k) {
      V[(RV * seg + k) * LD + pr] = vp[k];
      V[(RV * seg + k) * LD + col0q] = vq[k];
    }
  }
  if (seg == 0) {
    ev[pr] = dp;
    ev[col0q] = dq;
  }
  Sync::sync();
  if constexpr (!WARP) {
    if (tid == 0) {
      const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
    }
  }
  (void)lane;
  return sweep;
}

// warp convenience wrapper used by the MLE variants and the distance kernels (scratch must hold
// JacobiScratch<D>::doubles doubles, placed by the caller right after `ev`).
template <int D>
__device__ __forceinline__ int jacobi_eigh_warp(cplx* A, cplx* V, double* ev, int lane) {
  return jacobi_eigh<D, 32, SyncWarp, true>(A, V, ev, ev + D, lane);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory M x M complex products with strided register tiles (TR x TC outputs per work item:
// rows tr + i*M/TR, columns tc + j*M/TC -- so the lanes of a warp read consecutive columns and at most two
// distinct rows: conflict-free with the padded leading dimension).
//   MODE 0: C = A * B            MODE 1: C = A^dagger * B            ACCUM: C += ...
// C must not alias A or B.
// ---------------------------------------------------------------------------------------------
template <int M, int NT, int LD, int MODE, bool ACCUM = false>
__device__ void smem_matmul(cplx* __restrict__ C, const cplx* __restrict__ A, const cplx* __restrict__ B, int tid) {
  // M = 64: 4x4 tiles on 256 work items -- 8 shared loads per 16 complex FMAs keeps the LDS pipe (4 cycles per
  // 16-byte warp load) level with the FP64 pipe; 2x4 tiles on 512 threads were LDS-bound
  constexpr int TR = (M >= 64) ? 4 : (M >= 16 ? 2 : 1), TC = (M >= 16) ? 4 : 1;
  constexpr int NR = M / TR, NC = M / TC;  // tile grid
  for (int t = tid; t < NR * NC; t += NT) {
    const int tr = t / NC, tc = t % NC;
    cplx acc[TR][TC];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) acc[i][j] = cmake(0.0, 0.0);
#pragma unroll 4
    for (int k = 0; k < M; ++k) {
      cplx a[TR], b[TC];
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int r = tr + i * NR;
        a[i] = (MODE == 0) ? A[r * LD + k] : cconj(A[k * LD + r]);
      }
#pragma unroll
      for (int j = 0; j < TC; ++j) b[j] = B[k * LD + tc + j * NC];
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) cfma(acc[i][j], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        cplx* dst = &C[(tr + i * NR) * LD + tc + j * NC];
        *dst = ACCUM ? cadd(*dst, acc[i][j]) : acc[i][j];
      }
  }
}
