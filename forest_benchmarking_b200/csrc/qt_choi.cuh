// Device-side Choi-matrix projections shared by the standalone projection kernels (qt_project.cu) and
// the process-tomography kernel (qt_pgdb.cu).
//
//   proj CP   operator_tools/project_superoperators.py:19-34   (C+C^dagger)/2 -> eigh -> clamp -> V L V^dagger
//   proj TP   :62-84 (+ calculational.py:5-35)                 C - kron((Tr_out C - I)/d, I)
//   proj TNI  :37-59                                           C - kron((Tr_out C - clamp_le1(Tr_out C))/d, I)
//   physical  :87-144   Dykstra alternating projections with the Birgin-Raydan stopping rule
//
// One Choi matrix (m = 4^n) is one GROUP's work: a warp for n <= 2, a 512-thread block for n = 3.
// The eigensolver operands X, V and the basis-change temporary T live in shared memory with a padded
// leading dimension LD; the Dykstra state (S = last_state, Q = old_CP_change, CPREV = last_CP_projection)
// lives behind plain dense pointers (global/L2).
// Structure used: both the TP and TNI corrections are -kron(E, I_d) with a d x d matrix E, so
// old_TP_change is carried as E (d^2 numbers) and
//   ||dTP||^2 = d ||E_new - E||_F^2,   <old_TP, S_new - S> = -sum conj(E) (Tr_out S_new - Tr_out S),
//   new_CP_change - old_CP_change = CP - S   (so ||dCP||^2 = ||CP - S||_F^2).
// Warm start: consecutive CP-projection inputs differ by a small kron-structured term
// (pre_CP_{k+1} = pre_CP_k + kron(E_{k-1} - E_k, I)), so every eigendecomposition after the first starts
// from the previous eigenbasis: A' = V^dagger A V (two 64^3 products) leaves ~1e-3 of the Frobenius mass
// off the diagonal and Jacobi needs 2-3 sweeps instead of 9-10.
#pragma once
#include "qt_common.cuh"
#include "qt_eigh.cuh"

// threads of the block that owns one 64 x 64 Choi matrix (n = 3): 256 or 512 (see jacobi_eigh_block64)
#ifndef QT_N3_THREADS
#define QT_N3_THREADS 512
#endif
// safety caps that the reference does not have (its loops are unbounded); hitting one is reported through the
// per-item status word of the C ABI (QT_STATUS_* in include/qtomo.h)
#define QT_DYKSTRA_MAX_ITER 10000
#define QT_JACOBI_MAX_SWEEPS 30
#define QT_PGDB_MAX_OUTER 100000

template <int N, int NT, class Sync>
struct ChoiGroup {
  static constexpr int D = 1 << N;
  static constexpr int M = D * D;
  static constexpr int MM = M * M;                      // dense element count (global buffers)
  static constexpr int LD = (M >= 16) ? M + 1 : M;      // padded leading dimension (shared buffers)
  static constexpr int MP = M * LD;                     // padded element count
  static constexpr int TR = (M >= 16) ? 2 : 1, TC = (M >= 16) ? 4 : 1;  // register tile of the recomposition
  // small shared scratch (doubles): ev[M] | jacobi scratch | E[2*D*D] | En[2*D*D] | ptS[2*D*D] | ptC[2*D*D] |
  //                                 small eigh: P[2*D*D] W[2*D*D] pev[D] + jacobi scratch<D> | red2[64]
  //                                 v2 extras: Eprev[2*D*D] | sel[M ints] | counts[2 ints (+pad)]
  static constexpr int SMALL_DOUBLES =
      M + JacobiScratch<M>::doubles + 12 * D * D + D + JacobiScratch<D>::doubles + 64 + 2 * D * D + M / 2 + 2;
  static constexpr size_t group_smem = (sizeof(cplx) * 3 * MP + sizeof(double) * SMALL_DOUBLES + 15) / 16 * 16;

  static __device__ __forceinline__ int sidx(int e) { return (e / M) * LD + e % M; }

  // OUT (shared, padded) = V max(ev,0) V^dagger, or PRE + V max(-ev,0) V^dagger when fewer negatives.
  // pre(r, c) is a callable returning the Hermitian matrix that was decomposed.
  template <class PreFn>
  static __device__ void recompose_psd(cplx* OUT, const cplx* V, const double* ev, PreFn pre, int tid) {
    int npos = 0;
    for (int k = 0; k < M; ++k) npos += (ev[k] > 0.0) ? 1 : 0;
    const bool use_pos = npos <= M / 2;
    constexpr int NR = M / TR, NC = M / TC;
    for (int t = tid; t < NR * NC; t += NT) {
      const int tr = t / NC, tc = t % NC;
      cplx acc[TR][TC];
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = cmake(0.0, 0.0);
      for (int k = 0; k < M; ++k) {
        const double lam = ev[k];
        const double w = use_pos ? fmax(lam, 0.0) : fmax(-lam, 0.0);
        if (w == 0.0) continue;
        cplx vr[TR], vc[TC];
#pragma unroll
        for (int i = 0; i < TR; ++i) vr[i] = cscale(V[(tr + i * NR) * LD + k], w);
#pragma unroll
        for (int j = 0; j < TC; ++j) vc[j] = V[(tc + j * NC) * LD + k];
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int j = 0; j < TC; ++j) cfma_conj(acc[i][j], vr[i], vc[j]);
      }
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) {
          const int r = tr + i * NR, c = tc + j * NC;
          OUT[r * LD + c] = use_pos ? acc[i][j] : cadd(pre(r, c), acc[i][j]);
        }
    }
    Sync::sync();
  }

  // partial trace over the output factor: pt[a*D + c] = sum_b C[(a*D+b), (c*D+b)]; C has leading dimension ldc
  static __device__ void partial_trace_out(const cplx* C, int ldc, cplx* pt, int tid) {
    for (int e = tid; e < D * D; e += NT) {
      const int a = e / D, c = e % D;
      cplx s = cmake(0.0, 0.0);
      for (int b = 0; b < D; ++b) s = cadd(s, C[(a * D + b) * ldc + c * D + b]);
      pt[e] = s;
    }
    Sync::sync();
  }

  // E = correction matrix of the TP / TNI projection of a matrix whose partial trace is `pt`:
  //   TP : (pt - I)/d             TNI: (pt - V min(l,1) V^dagger)/d, V l V^dagger = eigh((pt+pt^dagger)/2)
  // P, W, pev, pscr: d x d scratch for the TNI eigen-decomposition (first warp of the group only).
  static __device__ void tp_correction(const cplx* pt, cplx* E, bool make_tp, cplx* P, cplx* W, double* pev,
                                       double* pscr, int tid) {
    if (make_tp) {
      for (int e = tid; e < D * D; e += NT) {
        cplx v = pt[e];
        if (e / D == e % D) v.x -= 1.0;
        E[e] = cscale(v, 1.0 / D);
      }
      Sync::sync();
      return;
    }
    if (tid < 32) {
      for (int e = tid; e < D * D; e += 32) {
        const int a = e / D, c = e % D;
        const cplx x = pt[e], y = pt[c * D + a];
        P[e] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
      }
      __syncwarp();
      jacobi_eigh<D, 32, SyncWarp, true>(P, W, pev, pscr, tid);
      for (int e = tid; e < D * D; e += 32) {
        const int a = e / D, c = e % D;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < D; ++k) cfma_conj(acc, cscale(W[a * D + k], fmin(pev[k], 1.0)), W[c * D + k]);
        E[e] = cscale(csub(pt[e], acc), 1.0 / D);
      }
    }
    Sync::sync();
  }

  // Eigendecomposition of the Hermitian matrix held in X (padded, shared).  With a valid previous eigenbasis
  // in V the matrix is first rotated into it (X <- V^dagger X V through T) and Jacobi continues from V.
  static __device__ int eigh_warm(cplx* X, cplx* V, cplx* T, double* ev, double* jscr, bool& v_valid, int tid,
                                  double rel2) {
    if (v_valid) {
      smem_matmul<M, NT, LD, 0>(T, X, V, tid);
      Sync::sync();
      smem_matmul<M, NT, LD, 1>(X, V, T, tid);
      Sync::sync();
      // restore exact Hermiticity (the two products round differently above and below the diagonal)
      for (int e = tid; e < MM; e += NT) {
        const int r = e / M, c = e % M;
        if (r < c) {
          const cplx a = X[r * LD + c], b = X[c * LD + r];
          const cplx h = cmake(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
          X[r * LD + c] = h;
          X[c * LD + r] = cconj(h);
        } else if (r == c) {
          X[r * LD + r].y = 0.0;
        }
      }
      Sync::sync();
      return jacobi_eigh<M, NT, Sync, true, LD>(X, V, ev, jscr, tid, /*init_v=*/false, QT_JACOBI_MAX_SWEEPS, rel2);
    }
    v_valid = true;
    return jacobi_eigh<M, NT, Sync, true, LD>(X, V, ev, jscr, tid, /*init_v=*/true, QT_JACOBI_MAX_SWEEPS, rel2);
  }

  // Dykstra.  On entry S holds the (Hermitian) matrix to project; on exit S holds the projection.
  // X, V, T: shared padded work matrices; S, Q, CPREV: dense M x M state buffers (any address space); small:
  // shared scratch of SMALL_DOUBLES doubles (16-byte aligned).  v_valid: V holds an eigenbasis to warm-start
  // from (kept up to date here).  Returns the number of CP projections (eigh calls); *sweeps_acc accumulates
  // the Jacobi sweeps they took.
  // RAW (optional): the caller's original, NON-Hermitian input (dense M x M, global).  The reference feeds it to
  // the loop as is (project_superoperators.py:108-111); its CP step Hermitises (:30), so the anti-Hermitian part
  // A = (RAW - RAW^dagger)/2 never reaches the iterates, but it lives on in old_CP_change (= Q_H - A) and therefore
  // enters the stopping rule: ||A||_F^2 in the first trip's ||dCP||^2, and -<A, CP_k - CP_{k-1}> (purely imaginary)
  // inside the abs() of the old_CP_change inner product on every later trip.
  // status (optional, one int per group): bit 0 set when the Dykstra loop hit max_iter, bit 1 when a Jacobi call
  // used all its sweeps -- both caps are additions over the reference, so the caller must be able to see them.
  static __device__ int project_physical(cplx* S, cplx* Q, cplx* CPREV, cplx* X, cplx* V, cplx* T, double* small,
                                         bool make_tp, int tid, bool& v_valid, int* sweeps_acc = nullptr,
                                         double rel2 = 0.0, int max_iter = QT_DYKSTRA_MAX_ITER,
                                         const cplx* RAW = nullptr, int* status = nullptr) {
    double* ev = small;
    double* jscr = ev + M;
    cplx* E = reinterpret_cast<cplx*>(jscr + JacobiScratch<M>::doubles);
    cplx* En = E + D * D;
    cplx* ptS = En + D * D;
    cplx* ptC = ptS + D * D;
    cplx* P = ptC + D * D;
    cplx* W = P + D * D;
    double* pev = reinterpret_cast<double*>(W + D * D);
    double* pscr = pev + D;
    double* red = pscr + JacobiScratch<D>::doubles;

    for (int e = tid; e < MM; e += NT) {
      Q[e] = cmake(0.0, 0.0);
      CPREV[e] = cmake(0.0, 0.0);
    }
    for (int e = tid; e < D * D; e += NT) E[e] = cmake(0.0, 0.0);
    Sync::sync();
    int n_eigh = 0;
    while (true) {
      // X = pre_CP = S - Q
      for (int e = tid; e < MM; e += NT) X[sidx(e)] = csub(S[e], Q[e]);
      Sync::sync();
      const int sw = eigh_warm(X, V, T, ev, jscr, v_valid, tid, rel2);
      if (sweeps_acc) *sweeps_acc += sw;
      if (status && sw >= QT_JACOBI_MAX_SWEEPS) *status |= 2;
      ++n_eigh;
      recompose_psd(X, V, ev, [&](int r, int c) { return csub(S[r * M + c], Q[r * M + c]); }, tid);  // X = CP
      // criterion pieces + state update of Q, CPREV
      double n_dcp = 0.0;
      cplx ip_q = cmake(0.0, 0.0);
      for (int e = tid; e < MM; e += NT) {
        const cplx cp = X[sidx(e)], s = S[e], q = Q[e], cprev = CPREV[e];
        const cplx d1 = csub(cp, s);
        n_dcp += cabs2(d1);
        cfma_conj(ip_q, csub(cp, cprev), q);  // conj(q) * (cp - cprev)
        if (RAW) {
          const cplx x = RAW[e], y = RAW[(e % M) * M + e / M];
          const cplx a = cmake(0.5 * (x.x - y.x), 0.5 * (x.y + y.y));  // anti-Hermitian part of the input
          if (n_eigh == 1) n_dcp += cabs2(a);
          else cfma_conj(ip_q, csub(cprev, cp), a);  // - conj(a) * (cp - cprev)
        }
        Q[e] = cadd(d1, q);                   // new_CP_change = CP - pre_CP = CP - S + Q
        CPREV[e] = cp;
      }
      partial_trace_out(S, M, ptS, tid);
      partial_trace_out(X, LD, ptC, tid);
      // pre_TP = CP - old_TP_change = CP + kron(E, I):  Tr_out(pre_TP) = ptC + d E
      for (int e = tid; e < D * D; e += NT) ptC[e] = cadd(ptC[e], cscale(E[e], (double)D));
      Sync::sync();
      tp_correction(ptC, En, make_tp, P, W, pev, pscr, tid);
      // new_state = pre_TP - kron(En, I) = CP + kron(E - En, I)
      for (int e = tid; e < MM; e += NT) {
        const int r = e / M, c = e % M;
        cplx v = X[r * LD + c];
        if ((r % D) == (c % D)) {
          const int a = r / D, cc = c / D;
          v = cadd(v, csub(E[a * D + cc], En[a * D + cc]));
        }
        S[e] = v;
      }
      // ||dTP||^2 = d ||En - E||^2 ;  <old_TP, dS> = -sum conj(E) (Tr_out Snew - Tr_out S), Tr_out Snew = ptC - d En
      double n_dtp = 0.0;
      cplx ip_t = cmake(0.0, 0.0);
      for (int e = tid; e < D * D; e += NT) {
        n_dtp += cabs2(csub(En[e], E[e]));
        const cplx dpt = csub(csub(ptC[e], cscale(En[e], (double)D)), ptS[e]);
        cfma_conj(ip_t, dpt, E[e]);
      }
      n_dcp = group_sum<NT, Sync>(n_dcp, red, tid);
      n_dtp = group_sum<NT, Sync>(n_dtp, red, tid);
      ip_q.x = group_sum<NT, Sync>(ip_q.x, red, tid);
      ip_q.y = group_sum<NT, Sync>(ip_q.y, red, tid);
      ip_t.x = group_sum<NT, Sync>(ip_t.x, red, tid);
      ip_t.y = group_sum<NT, Sync>(ip_t.y, red, tid);
      const double crit = n_dcp + D * n_dtp + 2.0 * sqrt(cabs2(ip_t)) + 2.0 * sqrt(cabs2(ip_q));
      Sync::sync();
      if (crit < 1e-4) break;
      if (n_eigh >= max_iter) {
        if (status) *status |= 1;
        break;
      }
      for (int e = tid; e < D * D; e += NT) E[e] = En[e];
      Sync::sync();
    }
    return n_eigh;
  }

  // ------------------------------------------------------------------------------------------------------------
  // Dykstra, second generation (used for M = 64, one matrix per thread block).  Same iteration, same stopping rule,
  // same trip counts as project_physical above; what changes is the work per trip:
  //
  //  * State.  With E_k the d x d TP/TNI correction carried into trip k (E_0 = 0) the iterates telescope:
  //        pre_CP_k = X_0 - kron(E_k, I),      last_state_k = CP_{k-1} + kron(E_{k-1} - E_k, I)   (k >= 1),
  //        old_CP_change_k = CP_{k-1} - pre_CP_{k-1},
  //    so the only dense state besides the input X_0 (kept in S) is the previous CP projection (CPREV).
  //  * Warm start without two full products.  The eigensolver's work matrix A = V^dagger pre_CP V is kept from trip to
  //    trip (Jacobi leaves it as D + R, see clean_a in qt_eigh.cuh); the next trip only adds the kron-structured change
  //        A += V^dagger kron(E_{k-1} - E_k, I) V      (a 64 x 8-term product and ONE full 64^3 product).
  //  * Early stop + first-order correction.  Jacobi stops at a relative off-diagonal norm tau (rel2 = tau^2, default
  //    1e-5 instead of 1e-8: ~1 sweep less per call).  With A = D + R the PSD projection is expanded around D,
  //        P_+(D + R) = D_+ + L o R + O(|R|^2 / gap),   L_ij = (d_i^+ - d_j^+) / (d_i - d_j)  in [0, 1],
  //    the Daleckii-Krein (Loewner matrix) derivative of x -> max(x, 0).  L vanishes on (negative, negative) pairs, so
  //    with `sel` = the smaller of the positive / non-positive index sets (K members, ~10 for unitary-like truths)
  //        CP = G Vs^dagger + Vs G^dagger,   G = V W (64 x K),   W = core[:, sel] with the (sel, sel) block halved,
  //    a rank-2K product, fused with the stopping-rule sums (no CP buffer, no extra passes).
  //    scripts/proto/early_stop_cp.py (CPU) shows the estimate moves by ~1e-9 and every trip count stays put.
  //
  // S (global, dense): in = the Hermitian matrix to project, out = the projection.  CPREV: global scratch, dense.
  // X, V, T: shared padded work matrices (X is the eigensolver's A; V the eigenbasis, warm across calls when v_valid).
  // ------------------------------------------------------------------------------------------------------------
  static __device__ __forceinline__ cplx kron_term(const cplx* E, int r, int c) {  // kron(E, I_D)[r, c]
    return ((r % D) == (c % D)) ? E[(r / D) * D + c / D] : cmake(0.0, 0.0);
  }

  static __device__ int project_physical_v2(cplx* S, cplx* CPREV, cplx* X, cplx* V, cplx* T, double* small,
                                            bool make_tp, int tid, bool& v_valid, int* sweeps_acc, double rel2,
                                            int max_iter, const cplx* RAW, int* status) {
    static_assert(M == 64 && NT == 512, "v2 tile mapping assumes a 64 x 64 matrix on 512 threads");
    double* ev = small;
    double* jscr = ev + M;
    cplx* E = reinterpret_cast<cplx*>(jscr + JacobiScratch<M>::doubles);
    cplx* En = E + D * D;
    cplx* ptS = En + D * D;
    cplx* ptC = ptS + D * D;
    cplx* P = ptC + D * D;
    cplx* W = P + D * D;
    double* pev = reinterpret_cast<double*>(W + D * D);
    double* pscr = pev + D;
    double* red = pscr + JacobiScratch<D>::doubles;
    cplx* Ep = reinterpret_cast<cplx*>(red + 64);
    int* sel = reinterpret_cast<int*>(Ep + D * D);
    int* cnts = sel + M;
    constexpr int LDG = 33;  // leading dimension of G (K <= 32 columns) inside T

    // output tile of this thread in the fused recomposition: rows tr, tr + 32; columns tc + 16 j
    const int tr = tid / 16, tc = tid % 16;

    for (int e = tid; e < D * D; e += NT) {
      E[e] = cmake(0.0, 0.0);
      Ep[e] = cmake(0.0, 0.0);
      ptS[e] = cmake(0.0, 0.0);
    }
    Sync::sync();
    int n_eigh = 0;
    while (true) {
      // ---- A = V^dagger pre_CP_k V ----
      if (n_eigh == 0) {
        for (int e = tid; e < MM; e += NT) X[sidx(e)] = S[e];
        Sync::sync();
        if (v_valid) {
          smem_matmul<M, NT, LD, 0>(T, X, V, tid);
          Sync::sync();
          smem_matmul<M, NT, LD, 1>(X, V, T, tid);
          Sync::sync();
        }
      } else {
        // T = kron(Ep - E, I) V :  T[(a,b), j] = sum_c dE[a,c] V[(c,b), j];  thread = (a, b, 8 columns j = jg + 8 jj)
        {
          const int a = tid / (D * 8), b = (tid / 8) % D, jg = tid % 8;
          cplx de[D];
#pragma unroll
          for (int c = 0; c < D; ++c) de[c] = csub(Ep[a * D + c], E[a * D + c]);
#pragma unroll
          for (int jj = 0; jj < M / 8; ++jj) {
            const int j = jg + 8 * jj;
            cplx acc = cmake(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < D; ++c) cfma(acc, de[c], V[(c * D + b) * LD + j]);
            T[(a * D + b) * LD + j] = acc;
          }
        }
        Sync::sync();
        smem_matmul<M, NT, LD, 1, true>(X, V, T, tid);  // A += V^dagger T
        Sync::sync();
      }
      const bool warm = v_valid;
      v_valid = true;
      const int sw = jacobi_eigh<M, NT, Sync, true, LD>(X, V, ev, jscr, tid, /*init_v=*/!warm, QT_JACOBI_MAX_SWEEPS,
                                                       rel2, /*clean_a=*/true);
      if (sweeps_acc) *sweeps_acc += sw;
      if (status && sw >= QT_JACOBI_MAX_SWEEPS) *status |= 2;
      ++n_eigh;
      const bool first = (n_eigh == 1);

      // ---- selection: the smaller of {ev > 0} / {ev <= 0} ----
      if (tid < 32) {
        const bool p0 = ev[tid] > 0.0, p1 = ev[tid + 32] > 0.0;
        const unsigned m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
        const int npos = __popc(m0) + __popc(m1);
        const bool use_pos = npos <= M / 2;
        const unsigned s0 = use_pos ? m0 : ~m0, s1 = use_pos ? m1 : ~m1;
        const unsigned below = (1u << tid) - 1u;
        if ((s0 >> tid) & 1u) sel[__popc(s0 & below)] = tid;
        if ((s1 >> tid) & 1u) sel[__popc(s0) + __popc(s1 & below)] = tid + 32;
        if (tid == 0) {
          cnts[0] = use_pos ? npos : M - npos;
          cnts[1] = use_pos ? 1 : 0;
        }
      }
      for (int e = tid; e < D * D; e += NT) ptC[e] = cmake(0.0, 0.0);
      Sync::sync();
      const int K = cnts[0];
      const bool use_pos = cnts[1] != 0;

      // ---- G = V W  (64 x K into T): W[i, p] = A[i][sel_p] * (1/2 if i in sel, else |d_p| / (|d_p| + |d_i|)) ----
      for (int w = tid; w < M * K; w += NT) {
        const int r = w % M, pp = w / M;  // a warp shares pp: A and ev reads are broadcasts
        const int sp = sel[pp];
        const double dp_abs = fabs(ev[sp]);
        cplx acc = cmake(0.0, 0.0);
#pragma unroll 4
        for (int i = 0; i < M; ++i) {
          const double di = ev[i];
          const bool in_sel = use_pos ? (di > 0.0) : !(di > 0.0);
          const double wgt = in_sel ? 0.5 : dp_abs * fast_rcp(dp_abs + fabs(di) + 1e-300);
          cfma(acc, V[r * LD + i], cscale(X[i * LD + sp], wgt));
        }
        T[r * LDG + pp] = acc;
      }
      Sync::sync();
      // ---- this thread's tile of  Psel = G Vs^dagger + Vs G^dagger ----
      cplx acc[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cmake(0.0, 0.0);
      for (int pp = 0; pp < K; ++pp) {
        const int sp = sel[pp];
        cplx gr[2], vr[2], gc[4], vc[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          gr[i] = T[(tr + 32 * i) * LDG + pp];
          vr[i] = V[(tr + 32 * i) * LD + sp];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gc[j] = T[(tc + 16 * j) * LDG + pp];
          vc[j] = V[(tc + 16 * j) * LD + sp];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cfma_conj(acc[i][j], gr[i], vc[j]);
            cfma_conj(acc[i][j], vr[i], gc[j]);
          }
      }
      // ---- CP_k on the tile, stopping-rule sums, CP_{k-1} <- CP_k, partial traces ----
      // (X_0 and CP_{k-1} come from global / L2 here rather than earlier: 48 more live doubles across the products
      // above would not fit the 128-register budget of a 512-thread block, and the exposed latency is ~0.3 %)
      cplx x0[2][4], cprev[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = (tr + 32 * i) * M + tc + 16 * j;
          x0[i][j] = S[e];
          cprev[i][j] = first ? cmake(0.0, 0.0) : CPREV[e];
        }
      double n_dcp = 0.0;
      cplx ip_q = cmake(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tr + 32 * i, c = tc + 16 * j, e = r * M + c;
          const cplx ke = kron_term(E, r, c), kp = kron_term(Ep, r, c);
          const cplx xk = csub(x0[i][j], ke);                        // pre_CP_k
          const cplx cp = use_pos ? acc[i][j] : csub(xk, acc[i][j]);  // P_+ = X - P_-
          const cplx s = first ? x0[i][j] : cadd(cprev[i][j], csub(kp, ke));  // last_state_k
          const cplx q = first ? cmake(0.0, 0.0) : cadd(csub(cprev[i][j], x0[i][j]), kp);  // old_CP_change
          n_dcp += cabs2(csub(cp, s));
          cfma_conj(ip_q, csub(cp, cprev[i][j]), q);
          if (RAW) {
            const cplx x = RAW[e], y = RAW[c * M + r];
            const cplx a = cmake(0.5 * (x.x - y.x), 0.5 * (x.y + y.y));
            if (first) n_dcp += cabs2(a);
            else cfma_conj(ip_q, csub(cprev[i][j], cp), a);
          }
          CPREV[e] = cp;
          acc[i][j] = cp;
          if ((r % D) == (c % D)) {
            cplx* dst = &ptC[(r / D) * D + c / D];
            atomicAdd(&dst->x, cp.x);
            atomicAdd(&dst->y, cp.y);
            if (first) {
              cplx* ds = &ptS[(r / D) * D + c / D];
              atomicAdd(&ds->x, x0[i][j].x);
              atomicAdd(&ds->y, x0[i][j].y);
            }
          }
        }
      Sync::sync();
      // pre_TP = CP + kron(E, I):  Tr_out(pre_TP) = Tr_out(CP) + d E
      for (int e = tid; e < D * D; e += NT) ptC[e] = cadd(ptC[e], cscale(E[e], (double)D));
      Sync::sync();
      tp_correction(ptC, En, make_tp, P, W, pev, pscr, tid);
      double n_dtp = 0.0;
      cplx ip_t = cmake(0.0, 0.0);
      for (int e = tid; e < D * D; e += NT) {
        n_dtp += cabs2(csub(En[e], E[e]));
        const cplx dpt = csub(csub(ptC[e], cscale(En[e], (double)D)), ptS[e]);  // Tr_out(new_state - last_state)
        cfma_conj(ip_t, dpt, E[e]);
      }
      n_dcp = group_sum<NT, Sync>(n_dcp, red, tid);
      n_dtp = group_sum<NT, Sync>(n_dtp, red, tid);
      ip_q.x = group_sum<NT, Sync>(ip_q.x, red, tid);
      ip_q.y = group_sum<NT, Sync>(ip_q.y, red, tid);
      ip_t.x = group_sum<NT, Sync>(ip_t.x, red, tid);
      ip_t.y = group_sum<NT, Sync>(ip_t.y, red, tid);
      const double crit = n_dcp + D * n_dtp + 2.0 * sqrt(cabs2(ip_t)) + 2.0 * sqrt(cabs2(ip_q));
      Sync::sync();
      const bool capped = n_eigh >= max_iter;
      if (crit < 1e-4 || capped) {
        if (capped && !(crit < 1e-4) && status) *status |= 1;
        // new_state = CP + kron(E - En, I)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = tr + 32 * i, c = tc + 16 * j;
            S[r * M + c] = cadd(acc[i][j], csub(kron_term(E, r, c), kron_term(En, r, c)));
          }
        __threadfence_block();
        Sync::sync();
        break;
      }
      // carry: Tr_out(last_state_{k+1}) = Tr_out(CP_k) + d (E_k - E_{k+1});  E_{k-1} <- E_k <- En
      for (int e = tid; e < D * D; e += NT) {
        ptS[e] = csub(ptC[e], cscale(En[e], (double)D));
        Ep[e] = E[e];
        E[e] = En[e];
      }
      Sync::sync();  // (CPREV is written and re-read by the same thread: no fence needed)
    }
    return n_eigh;
  }
};
