// Batched process tomography by projected gradient descent with backtracking (PGDB).
//
// Replaces, per experiment, forest/benchmarking/tomography.py:542-594 (pgdb_process_estimate) with
// _extract_from_results (:494-539), _cost (:597-614) and _grad_cost (:617-633), and the Dykstra projection
// proj_choi_to_physical (operator_tools/project_superoperators.py:87-144) it calls on every outer step.
// One experiment = one warp (n <= 2) or one 256-thread block (n = 3), persistent over the batch.
//
// The reference's dense design matrix A (27216 x 4096 complex = 1.78 GB at n = 3) is never built.  With
// R = PTM of the current estimate E (real for Hermitian E) and r_i[j] = Tr(P_j rho_i) the Pauli vector of
// input state i (checked against the oracle to 1e-16):
//     t_i[k]   = Tr(P_k E(rho_i)) = sum_j R[k,j] r_i[j]
//     p_s(+-)  = (t_i[0] +- c_s t_i[k_s]) / (2 d^2)                       (== A vec(E), tomography.py:611)
//     cost     = - sum_s n_s+ log clip(p_s+) + n_s- log clip(p_s-)         (:613-614, clip at 1e-6)
//     gradient = -(1/d^2) pl2choi(Gp),  Gp[k,j] = sum_i w_i[k] r_i[j],
//                w_i[0] = sum_{s in i} (eta+ + eta-)/2,  w_i[k_s] += c_s (eta+ - eta-)/2,  eta = n / clip(p)
// Everything is linear in E, so the line search only needs T_est + alpha T_upd (one T build per outer step).
#include "qt_choi.cuh"
#include "qt_pauli.cuh"
#include "../../include/qtomo.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

struct qt_pgdb_plan {
  int n, S, n_in, canonical;
  int nb = 0;            // > 0: the input states are the full product set of nb single-qubit states (see PgdbView)
  double bvec[6][4];     // Pauli vectors (1, x, y, z) of those nb single-qubit states
  int* d_state_id;    // [S]
  int* d_pidx;        // [S]
  double* d_coeff;    // [S]
  double* d_svec;     // [n_in, 4^n]
  // linear inversion (built on first use by qt_linear_inv_process_batch): settings grouped by observable
  std::vector<int> h_sid, h_pidx;
  std::vector<double> h_coeff, h_svec;
  int* d_li_slot_ptr = nullptr;    // [4^n + 1]  CSR over the observable's canonical Pauli index
  int* d_li_member_col = nullptr;  // [S]        column of `expect` of each member
  double* d_li_member_w = nullptr; // [S, 4^n]   row of the slot's pseudo-inverse that multiplies that expectation
  std::mutex li_mutex;
};

struct PgdbView {
  int S, n_in, canonical;
  const int* state_id;
  const int* pidx;
  const double* coeff;
  const double* svec;
  // Kronecker structure of the input states (generate_process_tomography_experiment, tomography.py:71-97: itertools.product
  // over one single-qubit set, first qubit most significant): state i = (s_0 .. s_{n-1}) in base nb and
  // svec[i][j] = prod_q bvec[s_q][digit_q(j)].  nb == 0: no such structure, svec is used as a dense table.
  int nb;
  double bvec[6][4];
};

template <int N>
struct PgdbCfg {
  static constexpr int NT = (N >= 3) ? QT_N3_THREADS : 32;
  static constexpr int GPB = (N >= 3) ? 1 : 4;
  using Sync = typename std::conditional<(N >= 3), SyncBlock, SyncWarp>::type;
  using G = ChoiGroup<N, NT, Sync>;
  static constexpr size_t group_smem = G::group_smem;  // X, V, T (padded) + small scratch
  // per-group global workspace, in doubles
  static __host__ __device__ int64_t ws_doubles(int n_in) {
    return 5LL * 2 * G::MM + 3LL * n_in * G::M;
  }
};

template <int N>
struct Pgdb {
  using C = PgdbCfg<N>;
  using G = typename C::G;
  using Sync = typename C::Sync;
  static constexpr int NT = C::NT, D = G::D, M = G::M, MM = G::MM, LD = G::LD;

  // in-place choi <-> superop reshuffle of a shared M x M matrix (swap index digits 0 and 3)
  static __device__ void reshuffle_inplace(cplx* X, int tid) {
    for (int e = tid; e < MM; e += NT) {
      const int r = e / M, c = e % M;
      const int i0 = r / D, i1 = r % D, i2 = c / D, i3 = c % D;
      if (i0 < i3) {
        const int e1 = r * LD + c, e2 = (i3 * D + i1) * LD + i2 * D + i0;
        const cplx a = X[e1], b = X[e2];
        X[e1] = b;
        X[e2] = a;
      }
    }
    Sync::sync();
  }

  // X (shared, Choi) -> Pauli-Liouville coefficients left at butterfly positions: R[k][j] = X[pos(k)*M + pos(j)] / d
  static __device__ void choi_to_pl_positions(cplx* X, int tid) {
    reshuffle_inplace(X, tid);
    pauli_butterfly_smem<true, true, Sync>(X, N, M, LD, 1, tid, NT);
    pauli_butterfly_smem<true, false, Sync>(X, N, M, 1, LD, tid, NT);
  }
  // inverse: X holds PL coefficients at butterfly positions -> Choi (times d; caller scales)
  static __device__ void pl_positions_to_choi(cplx* X, int tid) {
    pauli_butterfly_smem<false, true, Sync>(X, N, M, LD, 1, tid, NT);
    pauli_butterfly_smem<false, false, Sync>(X, N, M, 1, LD, tid, NT);
    reshuffle_inplace(X, tid);
  }

  // T[i][k] = sum_j R[k][j] svec[i][j], R read from butterfly positions of X (real part) with scale 1/d.
  // The real PTM is first compacted (transposed, Pauli order) into `rt` (M*M doubles of shared scratch) so that
  // the dot products read consecutive words: lanes walk k, svec[i][j] is a broadcast.
  static __device__ void build_T(const cplx* X, const PgdbView& pv, double* T, double* rt, int tid) {
    const double scale = 1.0 / D;
    for (int e = tid; e < MM; e += NT) {
      const int j = e / M, k = e % M;
      rt[e] = X[pauli_to_pos(k, N) * LD + pauli_to_pos(j, N)].x * scale;  // rt[j][k] = R[k][j] / d
    }
    Sync::sync();
    if (pv.nb > 0) {
      // svec = bvec (x) ... (x) bvec: contract the leading n-1 digits of j against the state prefix, then expand the
      // last digit -- 4^n (n + 1) + 4 nb flops per (k, prefix) instead of 4^n nb per (k, prefix); no dense table read.
      const int nb = pv.nb;
      int npre = 1;
      for (int q = 0; q < N - 1; ++q) npre *= nb;
      for (int w = tid; w < npre * M; w += NT) {
        const int k = w % M, pre = w / M;  // consecutive lanes = consecutive k: conflict-free rt reads, coalesced T stores
        int sq[N > 1 ? N - 1 : 1];
        {
          int r = pre;
          for (int q = N - 2; q >= 0; --q) {
            sq[q] = r % nb;
            r /= nb;
          }
        }
        double u[4] = {0.0, 0.0, 0.0, 0.0};
        for (int jp = 0; jp < M / 4; ++jp) {  // jp = the leading n-1 base-4 digits of j
          double wgt = 1.0;
#pragma unroll
          for (int q = 0; q < N - 1; ++q) wgt *= pv.bvec[sq[q]][(jp >> (2 * (N - 2 - q))) & 3];
#pragma unroll
          for (int jl = 0; jl < 4; ++jl) u[jl] = fma(rt[(jp * 4 + jl) * M + k], wgt, u[jl]);
        }
        for (int sl = 0; sl < nb; ++sl) {
          const double* b = pv.bvec[sl];
          T[(pre * nb + sl) * M + k] = fma(u[0], b[0], fma(u[1], b[1], fma(u[2], b[2], u[3] * b[3])));
        }
      }
    } else {
      for (int e = tid; e < pv.n_in * M; e += NT) {
        const int i = e / M, k = e % M;
        const double* sv = pv.svec + (int64_t)i * M;
        double acc = 0.0;
#pragma unroll 8
        for (int j = 0; j < M; ++j) acc = fma(rt[j * M + k], sv[j], acc);
        T[e] = acc;
      }
    }
    Sync::sync();
  }

  // Gp[k][j] = sum_i W[i][k] svec[i][j] into X at butterfly positions (the adjoint of build_T)
  static __device__ void build_gp(const PgdbView& pv, const double* W, cplx* X, int tid) {
    if (pv.nb > 0) {
      const int nb = pv.nb;
      int npre = 1;
      for (int q = 0; q < N - 1; ++q) npre *= nb;
      for (int w = tid; w < (M / 4) * M; w += NT) {
        const int k = w % M, jp = w / M;  // jp = leading n-1 digits of j; this thread fills j = 4 jp + (0..3)
        double y[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // y[s_last] = sum over state prefixes
        for (int pre = 0; pre < npre; ++pre) {
          double wgt = 1.0;
          int r = pre;
#pragma unroll
          for (int q = N - 2; q >= 0; --q) {
            wgt *= pv.bvec[r % nb][(jp >> (2 * (N - 2 - q))) & 3];
            r /= nb;
          }
          for (int sl = 0; sl < nb; ++sl) y[sl] = fma(W[(pre * nb + sl) * M + k], wgt, y[sl]);
        }
#pragma unroll
        for (int jl = 0; jl < 4; ++jl) {
          double acc = 0.0;
          for (int sl = 0; sl < nb; ++sl) acc = fma(y[sl], pv.bvec[sl][jl], acc);
          X[pauli_to_pos(k, N) * LD + pauli_to_pos(jp * 4 + jl, N)] = cmake(acc, 0.0);
        }
      }
    } else {
      for (int e = tid; e < MM; e += NT) {
        const int k = e / M, j = e % M;
        double acc = 0.0;
        for (int i = 0; i < pv.n_in; ++i) acc = fma(W[i * M + k], pv.svec[(int64_t)i * M + j], acc);
        X[pauli_to_pos(k, N) * LD + pauli_to_pos(j, N)] = cmake(acc, 0.0);
      }
    }
    Sync::sync();
  }

  struct Data {
    const double* ex;
    const double* cnt;
    double inv_total;
  };

  static __device__ __forceinline__ void setting_of(const PgdbView& pv, int s, int& i, int& k, double& c) {
    if (pv.canonical) {
      i = s / (M - 1);
      k = s % (M - 1) + 1;
      c = 1.0;
    } else {
      i = pv.state_id[s];
      k = pv.pidx[s];
      c = pv.coeff[s];
    }
  }

  // cost of est + alpha * upd from the T arrays
  static __device__ double cost(const PgdbView& pv, const Data& dt, const double* Te, const double* Tu, double alpha,
                                double* red, int tid) {
    const double h = 1.0 / (2.0 * D * D);
    double acc = 0.0;
    for (int s = tid; s < pv.S; s += NT) {
      int i, k;
      double c;
      setting_of(pv, s, i, k, c);
      const double t0 = fma(alpha, Tu[i * M], Te[i * M]);
      const double tk = c * fma(alpha, Tu[i * M + k], Te[i * M + k]);
      const double pp = fmax((t0 + tk) * h, 1e-6), pm = fmax((t0 - tk) * h, 1e-6);
      const double plus = 0.5 * (1.0 + dt.ex[s]);
      const double np_ = dt.cnt[s] * plus * dt.inv_total, nm = dt.cnt[s] * (1.0 - plus) * dt.inv_total;
      acc -= np_ * log(pp) + nm * log(pm);
    }
    return group_sum<NT, Sync>(acc, red, tid);
  }

  // w_i[k] (global scratch W [n_in][M]) from the current T_est
  static __device__ void build_w(const PgdbView& pv, const Data& dt, const double* Te, double* W, int tid) {
    const double h = 1.0 / (2.0 * D * D);
    if (pv.canonical) {
      for (int i = tid; i < pv.n_in; i += NT) {
        const double t0 = Te[i * M];
        double w0 = 0.0;
        for (int k = 1; k < M; ++k) {
          const int s = i * (M - 1) + k - 1;
          const double tk = Te[i * M + k];
          const double pp = fmax((t0 + tk) * h, 1e-6), pm = fmax((t0 - tk) * h, 1e-6);
          const double plus = 0.5 * (1.0 + dt.ex[s]);
          const double ep = dt.cnt[s] * plus * dt.inv_total / pp, em = dt.cnt[s] * (1.0 - plus) * dt.inv_total / pm;
          w0 += 0.5 * (ep + em);
          W[i * M + k] = 0.5 * (ep - em);
        }
        W[i * M] = w0;
      }
    } else {
      for (int e = tid; e < pv.n_in * M; e += NT) W[e] = 0.0;
      Sync::sync();
      for (int s = tid; s < pv.S; s += NT) {
        int i, k;
        double c;
        setting_of(pv, s, i, k, c);
        const double t0 = Te[i * M], tk = c * Te[i * M + k];
        const double pp = fmax((t0 + tk) * h, 1e-6), pm = fmax((t0 - tk) * h, 1e-6);
        const double plus = 0.5 * (1.0 + dt.ex[s]);
        const double ep = dt.cnt[s] * plus * dt.inv_total / pp, em = dt.cnt[s] * (1.0 - plus) * dt.inv_total / pm;
        atomicAdd(&W[i * M], 0.5 * (ep + em));
        atomicAdd(&W[i * M + k], c * 0.5 * (ep - em));
      }
    }
    __threadfence_block();
    Sync::sync();
  }

  // One experiment.  EST = choi_out[b] (global).  ws: per-group workspace.  X, V, small: shared.
  static __device__ void run(const PgdbView& pv, const Data& dt, bool make_tp, cplx* EST, double* ws, cplx* X,
                             cplx* V, cplx* T, double* small, int* counters, int* status_out, int tid, double rel2) {
    cplx* Gr = reinterpret_cast<cplx*>(ws);
    cplx* U = Gr + MM;
    cplx* S = U + MM;
    cplx* Q = S + MM;
    cplx* CPREV = Q + MM;
    double* Te = reinterpret_cast<double*>(CPREV + MM);
    double* Tu = Te + (int64_t)pv.n_in * M;
    double* W = Tu + (int64_t)pv.n_in * M;
    double* red = small + G::SMALL_DOUBLES - 64;
    const double mu = 3.0 / (2.0 * D * D), gamma = 0.3;

    // est = I / d ; T_est from it
    for (int e = tid; e < MM; e += NT) {
      const cplx v = cmake((e / M == e % M) ? 1.0 / D : 0.0, 0.0);
      EST[e] = v;
      X[G::sidx(e)] = v;
    }
    for (int e = tid; e < pv.n_in * M; e += NT) Tu[e] = 0.0;
    Sync::sync();
    choi_to_pl_positions(X, tid);
    build_T(X, pv, Te, reinterpret_cast<double*>(T), tid);
    double old_cost = cost(pv, dt, Te, Tu, 0.0, red, tid);
    int outer = 0, cost_evals = 1, eighs = 0, sweeps = 0, status = 0;
    bool v_valid = false;  // V keeps the last eigenbasis across Dykstra AND outer iterations (warm start)
    while (true) {
      ++outer;
      // ---- gradient ----
      build_w(pv, dt, Te, W, tid);
      build_gp(pv, W, X, tid);
      pl_positions_to_choi(X, tid);
      // gradient = -(1/d^2) * (1/d) * X ; S = est - gradient / mu (Hermitian by construction up to rounding)
      const double gs = -1.0 / ((double)D * D * D);
      for (int e = tid; e < MM; e += NT) {
        const int r = e / M, c = e % M;
        const cplx x = X[r * LD + c], y = X[c * LD + r];
        const cplx g = cmake(0.5 * gs * (x.x + y.x), 0.5 * gs * (x.y - y.y));
        Gr[e] = g;
        S[e] = csub(EST[e], cscale(g, 1.0 / mu));
      }
      Sync::sync();
      // ---- projection ----
      if constexpr (N >= 3) {
        eighs += G::project_physical_v2(S, CPREV, X, V, T, small, make_tp, tid, v_valid, &sweeps, rel2,
                                        QT_DYKSTRA_MAX_ITER, nullptr, &status);
      } else {
        eighs += G::project_physical(S, Q, CPREV, X, V, T, small, make_tp, tid, v_valid, &sweeps, rel2,
                                     QT_DYKSTRA_MAX_ITER, nullptr, &status);
      }
      // ---- update direction, its PTM image, <update, gradient> ----
      double ip = 0.0;
      for (int e = tid; e < MM; e += NT) {
        const cplx u = csub(S[e], EST[e]);
        U[e] = u;
        X[G::sidx(e)] = u;
        const cplx g = Gr[e];
        ip += u.x * g.x + u.y * g.y;
      }
      ip = group_sum<NT, Sync>(ip, red, tid);
      Sync::sync();
      choi_to_pl_positions(X, tid);
      build_T(X, pv, Tu, reinterpret_cast<double*>(T), tid);
      // ---- backtracking line search (tomography.py:574-585) ----
      double alpha = 1.0;
      double new_cost = cost(pv, dt, Te, Tu, alpha, red, tid);
      ++cost_evals;
      double change = gamma * alpha * ip;
      while (new_cost > old_cost + change) {
        alpha *= 0.5;
        change *= 0.5;
        new_cost = cost(pv, dt, Te, Tu, alpha, red, tid);
        ++cost_evals;
        if (alpha < 1e-15) break;
      }
      // ---- est += alpha * update ----
      for (int e = tid; e < MM; e += NT) {
        const cplx u = U[e];
        cplx v = EST[e];
        v.x = fma(alpha, u.x, v.x);
        v.y = fma(alpha, u.y, v.y);
        EST[e] = v;
      }
      for (int e = tid; e < pv.n_in * M; e += NT) Te[e] = fma(alpha, Tu[e], Te[e]);
      __threadfence_block();
      Sync::sync();
      if (old_cost - new_cost < 1e-10) break;
      if (outer >= QT_PGDB_MAX_OUTER) {
        status |= QT_STATUS_PGDB_CAP;
        break;
      }
      old_cost = new_cost;
    }
    if (tid == 0 && counters) {
      counters[0] = outer;
      counters[1] = cost_evals;
      counters[2] = eighs;
      counters[3] = sweeps;
    }
    if (tid == 0 && status_out) *status_out = status;
  }
};

template <int N>
__global__ void __launch_bounds__(PgdbCfg<N>::NT * PgdbCfg<N>::GPB) pgdb_kernel(PgdbView pv, int64_t B, const double* __restrict__ expect,
                            const double* __restrict__ counts, int make_tp, cplx* __restrict__ choi_out,
                            int* __restrict__ counters, int* __restrict__ status_out, double* __restrict__ workspace,
                            double rel2) {
  using C = PgdbCfg<N>;
  using G = typename C::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw + C::group_smem * gib);
  cplx* V = X + G::MP;
  cplx* T = V + G::MP;
  double* small = reinterpret_cast<double*>(T + G::MP);
  const int64_t group = (int64_t)blockIdx.x * C::GPB + gib;
  // workspace = [ 256-byte header: work-queue counter | per-group scratch ... ]
  unsigned long long* queue = reinterpret_cast<unsigned long long*>(workspace);
  double* ws = workspace + 32 + group * C::ws_doubles(pv.n_in);
  double* red = small + G::SMALL_DOUBLES - 64;
  long long* next = reinterpret_cast<long long*>(red + 48);
  // Experiments take 0.6x .. 1.5x the mean time (data-dependent trip counts), so groups pull the next
  // experiment from a global counter instead of owning a fixed stride of the batch.
  while (true) {
    if (tid == 0) *next = (long long)atomicAdd(queue, 1ULL);
    C::Sync::sync();
    const int64_t b = *next;
    C::Sync::sync();
    if (b >= B) break;
    typename Pgdb<N>::Data dt;
    dt.ex = expect + b * pv.S;
    dt.cnt = counts + b * pv.S;
    double tot = 0.0;
    for (int s = tid; s < pv.S; s += C::NT) tot += dt.cnt[s];
    tot = group_sum<C::NT, typename C::Sync>(tot, red, tid);
    dt.inv_total = 1.0 / tot;
    C::Sync::sync();
    Pgdb<N>::run(pv, dt, make_tp != 0, choi_out + b * G::MM, ws, X, V, T, small,
                 counters ? counters + 4 * b : nullptr, status_out ? status_out + b : nullptr, tid, rel2);
    C::Sync::sync();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static void bloch_of_state(int code, double v[4]) {
  static const double s2 = 1.4142135623730951, s6 = 2.449489742783178;
  static const double tab[10][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1},
                                    {0, 0, 1}, {2 * s2 / 3, 0, -1.0 / 3}, {-s2 / 3, -s6 / 3, -1.0 / 3},
                                    {-s2 / 3, s6 / 3, -1.0 / 3}};
  v[0] = 1.0;
  v[1] = tab[code][0];
  v[2] = tab[code][1];
  v[3] = tab[code][2];
}

extern "C" int qt_pgdb_plan_create(int n, int S, const int32_t* state_codes, const int32_t* pauli_idx,
                                   const double* coeff, qt_pgdb_plan** plan_out) {
  QT_REQUIRE(n >= 1 && n <= 3, "qt_pgdb_plan_create: n=%d out of range 1..3", n);
  QT_REQUIRE(S >= 1 && state_codes && pauli_idx && coeff && plan_out, "qt_pgdb_plan_create: bad arguments");
  const int M = 1 << (2 * n);
  std::map<std::vector<int>, int> ids;
  std::vector<std::vector<int>> states;
  std::vector<int> sid(S);
  for (int s = 0; s < S; ++s) {
    std::vector<int> key(state_codes + (size_t)s * n, state_codes + (size_t)(s + 1) * n);
    for (int q = 0; q < n; ++q)
      QT_REQUIRE(key[q] >= 0 && key[q] <= 9, "qt_pgdb_plan_create: state code %d out of range 0..9", key[q]);
    QT_REQUIRE(pauli_idx[s] >= 0 && pauli_idx[s] < M, "qt_pgdb_plan_create: pauli_idx[%d]=%d out of range", s,
               pauli_idx[s]);
    auto it = ids.find(key);
    if (it == ids.end()) {
      it = ids.emplace(key, (int)states.size()).first;
      states.push_back(key);
    }
    sid[s] = it->second;
  }
  const int n_in = (int)states.size();
  std::vector<double> svec((size_t)n_in * M);
  for (int i = 0; i < n_in; ++i)
    for (int j = 0; j < M; ++j) {
      double v = 1.0;
      for (int q = 0; q < n; ++q) {
        double b4[4];
        bloch_of_state(states[i][q], b4);
        v *= b4[(j >> (2 * (n - 1 - q))) & 3];
      }
      svec[(size_t)i * M + j] = v;
    }
  // Kronecker structure: state i = base-nb digits over one single-qubit set (last qubit fastest)
  int nb = 0;
  {
    int cand = 1;
    while (cand <= 6) {
      long long pw = 1;
      for (int q = 0; q < n; ++q) pw *= cand;
      if (pw == n_in) break;
      ++cand;
    }
    if (cand <= 6 && n_in >= cand) {
      bool ok = true;
      for (int i = 0; i < n_in && ok; ++i) {
        int r = i;
        for (int q = n - 1; q >= 0 && ok; --q) {
          ok = states[i][q] == states[r % cand][n - 1];  // digit r % cand of the base set, read off the first nb states
          r /= cand;
        }
      }
      for (int a = 0; a < cand && ok; ++a)  // the base set itself must have distinct members with all leading codes equal
        for (int q = 0; q + 1 < n && ok; ++q) ok = states[a][q] == states[0][n - 1];
      if (ok) nb = cand;
    }
  }
  int canonical = (S == n_in * (M - 1));
  for (int s = 0; s < S && canonical; ++s)
    canonical = (sid[s] == s / (M - 1)) && (pauli_idx[s] == s % (M - 1) + 1) && (coeff[s] == 1.0);
  qt_pgdb_plan* p = new qt_pgdb_plan();
  p->n = n; p->S = S; p->n_in = n_in; p->canonical = canonical;
  p->nb = nb;
  for (int a = 0; a < 6; ++a) {
    double b4[4] = {0.0, 0.0, 0.0, 0.0};
    if (a < nb) bloch_of_state(states[a][n - 1], b4);
    for (int c = 0; c < 4; ++c) p->bvec[a][c] = b4[c];
  }
  p->d_state_id = nullptr; p->d_pidx = nullptr; p->d_coeff = nullptr; p->d_svec = nullptr;
  QT_CUDA(cudaMalloc(&p->d_state_id, sizeof(int) * S));
  QT_CUDA(cudaMalloc(&p->d_pidx, sizeof(int) * S));
  QT_CUDA(cudaMalloc(&p->d_coeff, sizeof(double) * S));
  QT_CUDA(cudaMalloc(&p->d_svec, sizeof(double) * svec.size()));
  QT_CUDA(cudaMemcpy(p->d_state_id, sid.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_pidx, pauli_idx, sizeof(int) * S, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_coeff, coeff, sizeof(double) * S, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_svec, svec.data(), sizeof(double) * svec.size(), cudaMemcpyHostToDevice));
  p->h_sid = sid;
  p->h_pidx.assign(pauli_idx, pauli_idx + S);
  p->h_coeff.assign(coeff, coeff + S);
  p->h_svec = svec;
  *plan_out = p;
  return QT_OK;
}

extern "C" int qt_pgdb_plan_destroy(qt_pgdb_plan* p) {
  if (!p) return QT_OK;
  cudaFree(p->d_state_id);
  cudaFree(p->d_pidx);
  cudaFree(p->d_coeff);
  cudaFree(p->d_svec);
  cudaFree(p->d_li_slot_ptr);
  cudaFree(p->d_li_member_col);
  cudaFree(p->d_li_member_w);
  delete p;
  return QT_OK;
}

extern "C" int qt_pgdb_plan_info(const qt_pgdb_plan* p, int32_t* n_in_out, int32_t* canonical_out) {
  QT_REQUIRE(p, "qt_pgdb_plan_info: null plan");
  if (n_in_out) *n_in_out = p->n_in;
  if (canonical_out) *canonical_out = p->canonical;
  return QT_OK;
}

template <int N>
static int64_t pgdb_grid(int64_t B) {
  using C = PgdbCfg<N>;
  const int per_sm = (N >= 3) ? 1 : 8;
  return std::min<int64_t>((B + C::GPB - 1) / C::GPB, (int64_t)QT_NUM_SMS * per_sm);
}

extern "C" int64_t qt_pgdb_workspace_bytes(const qt_pgdb_plan* p, int64_t B) {
  if (!p) return -1;
  switch (p->n) {
    case 1: return 256 + pgdb_grid<1>(B) * PgdbCfg<1>::GPB * PgdbCfg<1>::ws_doubles(p->n_in) * 8;
    case 2: return 256 + pgdb_grid<2>(B) * PgdbCfg<2>::GPB * PgdbCfg<2>::ws_doubles(p->n_in) * 8;
    default: return 256 + pgdb_grid<3>(B) * PgdbCfg<3>::GPB * PgdbCfg<3>::ws_doubles(p->n_in) * 8;
  }
}

template <int N>
static int launch_pgdb(const qt_pgdb_plan* p, int64_t B, const double* expect, const double* counts, int make_tp,
                       double rel2, void* choi_out, int* counters, int* status, void* ws, cudaStream_t st) {
  using C = PgdbCfg<N>;
  PgdbView pv{p->S, p->n_in, p->canonical, p->d_state_id, p->d_pidx, p->d_coeff, p->d_svec, p->nb, {}};
  for (int a = 0; a < 6; ++a)
    for (int c = 0; c < 4; ++c) pv.bvec[a][c] = p->bvec[a][c];
  const size_t smem = C::group_smem * C::GPB;
  QT_CUDA(cudaMemsetAsync(ws, 0, 256, st));  // work-queue counter
  QT_CUDA(cudaFuncSetAttribute(pgdb_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pgdb_kernel<N><<<(unsigned)pgdb_grid<N>(B), C::NT * C::GPB, smem, st>>>(pv, B, expect, counts, make_tp,
                                                                          (cplx*)choi_out, counters, status, (double*)ws, rel2);
  return qt_check_launch("pgdb_kernel");
}

extern "C" int qt_pgdb_process_batch(const qt_pgdb_plan* p, int64_t B, const double* expect, const double* counts,
                                     int trace_preserving, double eigh_rel_tol, void* choi_out,
                                     int32_t* counters_out, int32_t* status_out, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
  QT_REQUIRE(p, "qt_pgdb_process_batch: null plan");
  if (B == 0) return QT_OK;
  QT_REQUIRE(expect && counts && choi_out && workspace, "qt_pgdb_process_batch: null argument");
  if (workspace_bytes < qt_pgdb_workspace_bytes(p, B)) {
    qt_set_error("qt_pgdb_process_batch: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
                 (long long)qt_pgdb_workspace_bytes(p, B));
    return QT_ERR_WORKSPACE;
  }
  double rel2;
  if (qt_eigh_rel2_from_tol(eigh_rel_tol, p->n, &rel2, "qt_pgdb_process_batch") != QT_OK) return QT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define QT_PGDB_ARGS p, B, expect, counts, trace_preserving, rel2, choi_out, counters_out, status_out, workspace, st
  switch (p->n) {
    case 1: return launch_pgdb<1>(QT_PGDB_ARGS);
    case 2: return launch_pgdb<2>(QT_PGDB_ARGS);
    default: return launch_pgdb<3>(QT_PGDB_ARGS);
  }
#undef QT_PGDB_ARGS
}

// =============================================================================================
// linear_inv_process_estimate (tomography.py:459-491): choi = unvec(pinv(M) e) + I/d with one row
// vec(conj(rho_in) (x) c P_k)^dagger of M per setting.
//
// In Pauli-Liouville coordinates (an isometry of vec(choi) up to a factor, so the minimum-norm least-squares
// solution is the same) a setting with observable P_k only involves row k of the PTM R:
//     e_s = c_s sum_j R[k_s, j] r_{i_s}[j],   r_i = Pauli vector of input state i,
// so M is block diagonal over k and  R[k, :] = pinv(X_k) e_k  with X_k = [c_s r_{i_s}] over the settings of slot k.
// The dense 540 x 256 (n = 2) / 13608 x 4096 (n = 3) complex pseudo-inverse of the reference is never formed: the plan
// keeps, per setting, the 4^n-vector  w_s = (X_k^T X_k)^+ c_s r_{i_s}  and the kernel accumulates
// R[k, :] = sum_s w_s e_s, adds the identity term (R[0, 0] += 1, the eye(d^2)/d of the reference) and goes
// PTM -> Choi in shared memory.  Checked against the oracle's dense pinv to 1e-14.
// =============================================================================================
// FROM_X: the PTM (at butterfly positions, real parts) is already in choi_out, left there by linproc_accum_kernel
template <int N, bool FROM_X>
__global__ void __launch_bounds__(PgdbCfg<N>::NT * PgdbCfg<N>::GPB)
    linproc_kernel(int64_t B, int S, const int* __restrict__ slot_ptr, const int* __restrict__ member_col,
                   const double* __restrict__ member_w, const double* __restrict__ expect, cplx* __restrict__ choi_out) {
  using C = PgdbCfg<N>;
  using G = typename C::G;
  constexpr int M = G::M, MM = G::MM, LD = G::LD, D = G::D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gib = threadIdx.x / C::NT, tid = threadIdx.x % C::NT;
  cplx* X = reinterpret_cast<cplx*>(smem_raw) + (size_t)G::MP * gib;
  for (int64_t b = (int64_t)blockIdx.x * C::GPB + gib; b < B; b += (int64_t)gridDim.x * C::GPB) {
    if constexpr (FROM_X) {
      const cplx* src = choi_out + b * MM;
      for (int e = tid; e < MM; e += C::NT) X[(e / M) * LD + e % M] = src[e];
    } else {
      const double* ex = expect + b * S;
      for (int e = tid; e < MM; e += C::NT) {
        const int k = e / M, j = e % M;
        double acc = (e == 0) ? 1.0 : 0.0;
        for (int m = slot_ptr[k]; m < slot_ptr[k + 1]; ++m) acc = fma(member_w[(int64_t)m * M + j], ex[member_col[m]], acc);
        X[pauli_to_pos(k, N) * LD + pauli_to_pos(j, N)] = cmake(acc, 0.0);
      }
    }
    C::Sync::sync();
    Pgdb<N>::pl_positions_to_choi(X, tid);
    cplx* dst = choi_out + b * MM;
    for (int e = tid; e < MM; e += C::NT) dst[e] = cscale(X[(e / M) * LD + e % M], 1.0 / D);
    C::Sync::sync();
  }
}

// n = 2, 3: the accumulation R[k, :] = sum_s w_s e_s re-reads the whole weight table (n = 2: 69 KB, n = 3: 7 MB) for
// every experiment when one block works on one experiment (0.07 / 0.02 of the HBM roof).  Here a block takes EB
// experiments at once: thread (kg, j) walks the members of slot k = k0 + kg, loads w_s[j] ONCE and applies it to the EB
// expectation values of that setting (staged through shared memory in chunks of CH members, read back as broadcasts).
// The PTM goes to choi_out (butterfly positions, real); linproc_kernel<N, true> turns it into the Choi matrix in place.
template <int N, int EB, int CH>
__global__ void __launch_bounds__(256)
    linproc_accum_kernel(int64_t B, int S, const int* __restrict__ slot_ptr, const int* __restrict__ member_col,
                         const double* __restrict__ member_w, const double* __restrict__ expect, cplx* __restrict__ ptm_out) {
  constexpr int M = 1 << (2 * N), MM = M * M, KG = 256 / M;  // KG slots per pass
  static_assert(M <= 256 && EB % 2 == 0, "one thread per (slot of the pass, column)");
  extern __shared__ __align__(16) double et[];  // [KG][CH][EB]
  const int tid = threadIdx.x, j = tid % M, kg = tid / M;
  const int64_t b0 = (int64_t)blockIdx.x * EB;
  const int nb = (int)min((int64_t)EB, B - b0);
  {
    const int k0 = blockIdx.y * KG;  // one pass of KG slots per block: (B / EB) x (M / KG) blocks fill the machine at small B
    const int k = k0 + kg, m_lo = slot_ptr[k], len = slot_ptr[k + 1] - m_lo;
    int lenmax = 0;
    for (int kk = 0; kk < KG; ++kk) lenmax = max(lenmax, slot_ptr[k0 + kk + 1] - slot_ptr[k0 + kk]);
    double acc[EB];
#pragma unroll
    for (int bb = 0; bb < EB; ++bb) acc[bb] = 0.0;
    for (int c0 = 0; c0 < lenmax; c0 += CH) {
      __syncthreads();
      const int clen = min(CH, lenmax - c0);  // members actually present in this chunk (the longest slot of the pass)
      for (int idx = tid; idx < KG * clen * EB; idx += 256) {
        const int bb = idx % EB, mm = (idx / EB) % clen, kk = idx / (EB * clen);
        const int lo = slot_ptr[k0 + kk], ln = slot_ptr[k0 + kk + 1] - lo;
        double v = 0.0;
        if (c0 + mm < ln && bb < nb) v = expect[(b0 + bb) * S + member_col[lo + c0 + mm]];
        et[(kk * CH + mm) * EB + bb] = v;
      }
      __syncthreads();
      const int cnt = min(CH, len - c0);
      const double* erow = et + (size_t)kg * CH * EB;
#pragma unroll 8
      for (int mm = 0; mm < cnt; ++mm) {  // unrolled: eight independent weight loads in flight (L2 latency)
        const double wv = member_w[(int64_t)(m_lo + c0 + mm) * M + j];
#pragma unroll
        for (int bb = 0; bb < EB; bb += 2) {
          const double2 e2 = *reinterpret_cast<const double2*>(erow + mm * EB + bb);
          acc[bb] = fma(wv, e2.x, acc[bb]);
          acc[bb + 1] = fma(wv, e2.y, acc[bb + 1]);
        }
      }
    }
    const int pos = pauli_to_pos(k, N) * M + pauli_to_pos(j, N);
    const double idt = (k == 0 && j == 0) ? 1.0 : 0.0;  // the eye(d^2) / d term of the reference
#pragma unroll
    for (int bb = 0; bb < EB; ++bb)
      if (bb < nb) ptm_out[(b0 + bb) * MM + pos] = cmake(acc[bb] + idt, 0.0);
  }
}

// cyclic Jacobi for a small real symmetric matrix on the host (plan setup only): a -> eigenvalues on the diagonal,
// v -> eigenvectors in columns
static void host_sym_jacobi(int m, std::vector<double>& a, std::vector<double>& v) {
  v.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) v[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, tot = 0.0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        tot += a[(size_t)i * m + j] * a[(size_t)i * m + j];
        if (i != j) off += a[(size_t)i * m + j] * a[(size_t)i * m + j];
      }
    if (off <= 1e-30 * tot) break;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = a[(size_t)p * m + q];
        if (apq == 0.0) continue;
        const double theta = (a[(size_t)q * m + q] - a[(size_t)p * m + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < m; ++k) {  // columns p, q
          const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
          a[(size_t)k * m + p] = c * akp - sn * akq;
          a[(size_t)k * m + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {  // rows p, q
          const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
          a[(size_t)p * m + k] = c * apk - sn * aqk;
          a[(size_t)q * m + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < m; ++k) {
          const double vkp = v[(size_t)k * m + p], vkq = v[(size_t)k * m + q];
          v[(size_t)k * m + p] = c * vkp - sn * vkq;
          v[(size_t)k * m + q] = sn * vkp + c * vkq;
        }
      }
  }
}

static int build_linear_inversion(qt_pgdb_plan* p) {
  const int M = 1 << (2 * p->n), S = p->S;
  std::vector<int> slot_ptr(M + 1, 0), col(S), cursor(M, 0);
  for (int s = 0; s < S; ++s) slot_ptr[p->h_pidx[s] + 1]++;
  for (int k = 0; k < M; ++k) slot_ptr[k + 1] += slot_ptr[k];
  for (int s = 0; s < S; ++s) {  // stable: members keep the order of `results`
    const int k = p->h_pidx[s];
    col[slot_ptr[k] + cursor[k]++] = s;
  }
  std::vector<double> w((size_t)S * M);
  // slots that share the same (input state, coefficient) list share the pseudo-inverse (every slot of a complete design)
  std::map<std::vector<std::pair<int, double>>, int> seen;
  std::vector<std::vector<double>> ginv;
  for (int k = 0; k < M; ++k) {
    const int m0 = slot_ptr[k], m1 = slot_ptr[k + 1];
    if (m0 == m1) continue;
    std::vector<std::pair<int, double>> key;
    for (int m = m0; m < m1; ++m) key.emplace_back(p->h_sid[col[m]], p->h_coeff[col[m]]);
    auto it = seen.find(key);
    if (it == seen.end()) {
      std::vector<double> g((size_t)M * M, 0.0), v;
      for (const auto& kv : key) {
        const double* r = &p->h_svec[(size_t)kv.first * M];
        const double c2 = kv.second * kv.second;
        for (int a = 0; a < M; ++a)
          for (int b = 0; b < M; ++b) g[(size_t)a * M + b] += c2 * r[a] * r[b];
      }
      host_sym_jacobi(M, g, v);
      double lmax = 0.0;
      for (int a = 0; a < M; ++a) lmax = std::max(lmax, g[(size_t)a * M + a]);
      std::vector<double> gi((size_t)M * M, 0.0);
      for (int e = 0; e < M; ++e) {
        const double lam = g[(size_t)e * M + e];
        if (!(lam > 1e-11 * lmax)) continue;  // null space of the design: minimum-norm solution, like pinv
        for (int a = 0; a < M; ++a)
          for (int b = 0; b < M; ++b) gi[(size_t)a * M + b] += v[(size_t)a * M + e] * v[(size_t)b * M + e] / lam;
      }
      it = seen.emplace(key, (int)ginv.size()).first;
      ginv.push_back(std::move(gi));
    }
    const std::vector<double>& gi = ginv[it->second];
    for (int m = m0; m < m1; ++m) {
      const int s = col[m];
      const double* r = &p->h_svec[(size_t)p->h_sid[s] * M];
      for (int a = 0; a < M; ++a) {
        double acc = 0.0;
        for (int b = 0; b < M; ++b) acc += gi[(size_t)a * M + b] * r[b];
        w[(size_t)m * M + a] = acc * p->h_coeff[s];
      }
    }
  }
  QT_CUDA(cudaMalloc(&p->d_li_slot_ptr, sizeof(int) * (M + 1)));
  QT_CUDA(cudaMalloc(&p->d_li_member_col, sizeof(int) * S));
  QT_CUDA(cudaMalloc(&p->d_li_member_w, sizeof(double) * w.size()));
  QT_CUDA(cudaMemcpy(p->d_li_slot_ptr, slot_ptr.data(), sizeof(int) * (M + 1), cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_li_member_col, col.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
  QT_CUDA(cudaMemcpy(p->d_li_member_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice));
  return QT_OK;
}

template <int N>
static int launch_linproc(const qt_pgdb_plan* p, int64_t B, const double* expect, void* choi_out, cudaStream_t st) {
  using C = PgdbCfg<N>;
  const size_t smem = sizeof(cplx) * C::G::MP * C::GPB;
  const int per_sm = (N >= 3) ? 3 : 8;
  const int64_t grid = std::min<int64_t>((B + C::GPB - 1) / C::GPB, (int64_t)QT_NUM_SMS * per_sm);
  if constexpr (N >= 2) {
    constexpr int EB = 8, CH = 64, KG = 256 >> (2 * N);
    const size_t smem_a = sizeof(double) * KG * CH * EB;
    QT_CUDA(cudaFuncSetAttribute(linproc_accum_kernel<N, EB, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    linproc_accum_kernel<N, EB, CH><<<dim3((unsigned)((B + EB - 1) / EB), (1u << (2 * N)) / KG), 256, smem_a, st>>>(
        B, p->S, p->d_li_slot_ptr, p->d_li_member_col, p->d_li_member_w, expect, (cplx*)choi_out);
    int rc = qt_check_launch("linproc_accum_kernel");
    if (rc) return rc;
    QT_CUDA(cudaFuncSetAttribute(linproc_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linproc_kernel<N, true><<<(unsigned)grid, C::NT * C::GPB, smem, st>>>(B, p->S, p->d_li_slot_ptr, p->d_li_member_col,
                                                                          p->d_li_member_w, expect, (cplx*)choi_out);
    return qt_check_launch("linproc_kernel");
  } else {
    QT_CUDA(cudaFuncSetAttribute(linproc_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linproc_kernel<N, false><<<(unsigned)grid, C::NT * C::GPB, smem, st>>>(B, p->S, p->d_li_slot_ptr, p->d_li_member_col,
                                                                           p->d_li_member_w, expect, (cplx*)choi_out);
    return qt_check_launch("linproc_kernel");
  }
}

extern "C" int qt_linear_inv_process_batch(qt_pgdb_plan* p, int64_t B, const double* expect, void* choi_out,
                                           void* stream) {
  QT_REQUIRE(p, "qt_linear_inv_process_batch: null plan");
  {
    std::lock_guard<std::mutex> guard(p->li_mutex);  // two host threads may share one plan
    if (!p->d_li_member_w) {
      const int rc = build_linear_inversion(p);
      if (rc != QT_OK) return rc;
    }
  }
  if (B == 0) return QT_OK;
  QT_REQUIRE(expect && choi_out, "qt_linear_inv_process_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->n) {
    case 1: return launch_linproc<1>(p, B, expect, choi_out, st);
    case 2: return launch_linproc<2>(p, B, expect, choi_out, st);
    default: return launch_linproc<3>(p, B, expect, choi_out, st);
  }
}
