/* libqtomo -- C ABI of the B200 (sm_100a) batched quantum-tomography / superoperator-algebra engine.
 *
 * The reference (rigetti/forest-benchmarking) is pure Python and has no FFI: its boundary for this
 * path is a set of Python call signatures taking numpy arrays / ExperimentResult lists.
 * Each entry point below is the batched device-side replacement of one of those functions; the Python
 * package forest_benchmarking_b200 binds them with ctypes and re-exposes the reference's signatures.
 * File:line citations are relative to /root/reference/forest/benchmarking/.
 *
 * Conventions
 *   - every function returns 0 on success, a negative QT_ERR_* code otherwise; qt_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - all array arguments are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer; nothing is allocated behind the caller's back except inside *_plan_create;
 *   - complex matrices are interleaved complex128, row-major [B, rows, cols] (numpy C order);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous;
 *   - Pauli index: base-4 digits I=0 X=1 Y=2 Z=3, first qubit most significant (utils.py:146-156);
 *   - input-state code per qubit: 0..5 = +X,-X,+Y,-Y,+Z,-Z (tomography.py:89), 6..9 = SIC0..3 (:71).
 */
#ifndef QTOMO_H
#define QTOMO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QT_OK 0
#define QT_ERR_ARG (-1)
#define QT_ERR_CUDA (-2)
#define QT_ERR_UNSUPPORTED (-3)
#define QT_ERR_WORKSPACE (-4)

int qt_version(void);
/* copies the last error message of this thread into buf (NUL-terminated, truncated to len) */
int qt_last_error(char* buf, int len);

/* `eigh_rel_tol` argument of qt_proj_physical_batch / qt_pgdb_process_batch: relative off-diagonal Frobenius norm at
 * which the Jacobi eigensolver behind proj_choi_to_completely_positive declares convergence.  Negative = the default
 * below; 0 = tight (1e-15 * 4^n); otherwise < 1e-3.  A per-call argument: the library has no mutable global state.
 * n <= 2: the remainder is dropped (error ~ tol).  n = 3: the remainder enters a first-order correction of the PSD
 * projection (error ~ tol^2 / spectral gap), which is why its default can be 1e-5. */
#define QT_EIGH_REL_TOL_DEFAULT 1e-8
#define QT_EIGH_REL_TOL_DEFAULT_CORRECTED 1e-5
/* per-item status bits (status_out arrays): the reference's loops are unbounded, ours carry safety caps */
#define QT_STATUS_DYKSTRA_CAP 1 /* proj_choi_to_physical stopped at 10000 CP projections without meeting its rule */
#define QT_STATUS_JACOBI_CAP 2  /* an eigendecomposition used all 30 sweeps */
#define QT_STATUS_PGDB_CAP 4    /* pgdb_process_estimate stopped at 100000 outer iterations */

/* FP64 FMA throughput probe (bench utility): blocks*threads*8*iters FMAs; scratch = 1 double on device */
int qt_fp64_probe(int blocks, int threads, int iters, double* scratch, void* stream);

/* ---- state tomography: iterative_mle_state_estimate (tomography.py:168-270) + _R (:273-338) ---- */
typedef struct qt_mle_plan qt_mle_plan;
/* pauli_idx_host[K], coeff_host[K]: observable of results[k] = coeff * Pauli(pauli_idx); K = len(results) */
int qt_mle_plan_create(int n, int K, const int32_t* pauli_idx_host, const double* coeff_host, qt_mle_plan** plan);
int qt_mle_plan_destroy(qt_mle_plan* plan);
#define QT_MLE_KERNEL_AUTO 0
#define QT_MLE_KERNEL_REGISTER 1 /* one experiment per thread, n<=2, unit coefficients, vanilla MLE */
#define QT_MLE_KERNEL_WARP 2     /* one experiment per warp, any n<=5, all variants */
#define QT_MLE_KERNEL_QUAD 3     /* one experiment per 4 lanes, n==2, unit coefficients, vanilla MLE (AUTO picks it) */
/* expect[B,K], counts[B,K] (may be NULL unless beta>0) -> rho_out[B,d,d] complex, iters_out[B]
 * iters_out = the reference's loop counter at exit (== maxiter when the cap was hit, tomography.py:244) */
int qt_mle_state_batch(const qt_mle_plan* plan, int64_t B, const double* expect, const double* counts,
                       double epsilon, double entropy_penalty, double beta, double tol, int maxiter,
                       int kernel_variant, void* rho_out, int32_t* iters_out, void* stream);
/* linear_inv_state_estimate (tomography.py:130-165): expect[B,K] -> rho_out[B,d,d].  Uses the plan's observable
 * list; valid for any list of Pauli observables (the pseudo-inverse is diagonal in the Pauli basis). */
int qt_linear_inv_state_batch(const qt_mle_plan* plan, int64_t B, const double* expect, void* rho_out, void* stream);
/* state_log_likelihood (tomography.py:341-375): rho[B,d,d], expect[B,K], counts[B,K] -> ll_out[B] (log10 likelihood
 * of the measured +-1 frequencies under rho; outcomes of non-positive predicted probability are skipped). */
int qt_state_log_likelihood_batch(const qt_mle_plan* plan, int64_t B, const void* rho, const double* expect,
                                  const double* counts, double* ll_out, void* stream);
/* ONE R rho R update, rho streamed HBM -> HBM (n = 1, 2; complete canonical Pauli set, K = 4^n - 1).
 * expect_canon[K, B] (item-minor); rho_in must be Hermitian (a state: at n = 2 only its upper triangle is read).
 * The HBM-roofline view of the update. */
int qt_mle_step_batch(int n, int64_t B, const double* expect_canon, const void* rho_in, double epsilon,
                      void* rho_out, void* stream);

/* ---- raw shots -> expectation / variance (observable_estimation.py) ------------------------------ */
/* shots_to_obs_moments (:804-853) for B settings at once.  bits[B, n_shots, n_qubits] bytes of 0/1; col_mask[b] has bit q
 * set when column q belongs to setting b's observable (0 = identity term: mean = coeff, var = 0); coeff[B] the
 * observable's real coefficient.  mean_out[B], var_out[B] (variance of the mean; Beta-posterior moments when
 * use_beta_prior != 0). */
int qt_shots_to_obs_moments_batch(int64_t B, int64_t n_shots, int n_qubits, const uint8_t* bits,
                                  const uint32_t* col_mask, const double* coeff, int use_beta_prior,
                                  double* mean_out, double* var_out, void* stream);
/* the arithmetic of calibrate_observable_estimates (:1033-1049) with ratio_variance (:1052-1090):
 * mean_out = mean / cal_mean, var_out = var / cal_mean^2 + mean^2 cal_var / cal_mean^4 */
int qt_calibrate_estimates_batch(int64_t B, const double* mean, const double* var, const double* cal_mean,
                                 const double* cal_var, double* mean_out, double* var_out, void* stream);

/* ---- superoperator conversions (operator_tools/superoperator_transformations.py) ------------- */
/* kraus[B, n_kraus, d, d] -> choi[B, d^2, d^2] = sum_k vec(K) vec(K)^dagger            (:159-182) */
int qt_kraus2choi_batch(int d, int n_kraus, int64_t B, const void* kraus, void* choi_out, void* stream);
/* kraus[B, n_kraus, d, d] -> superop[B, d^2, d^2] = sum_k conj(K) (x) K                 (:100-145) */
int qt_kraus2superop_batch(int d, int n_kraus, int64_t B, const void* kraus, void* superop_out, void* stream);
/* choi2superop == superop2choi: reshape [d]*4, swapaxes(0,3); out-of-place    (:267-277, :351-361) */
int qt_choi_superop_reshuffle_batch(int d, int64_t B, const void* in, void* out, void* stream);
/* superop2pauli_liouville (:253-264) / pauli_liouville2superop (:301-312), n qubits, [B,4^n,4^n].
 * workspace: NULL for n<=3; B*16^n*16 bytes of device memory for n = 4, 5 (two-pass path). */
int qt_superop2pl_batch(int n, int64_t B, const void* superop, void* pl_out, void* workspace, void* stream);
int qt_pl2superop_batch(int n, int64_t B, const void* pl, void* superop_out, void* workspace, void* stream);
/* The same two conversions with the algorithm selectable (forward != 0: superop -> PL).  BUTTERFLY = the calls above
 * (Kronecker-factored add-only stages, HBM-bound, the default everywhere).  DENSE_DMMA = the reference's formulation,
 * two dense complex products with the 4^n x 4^n basis matrix (:253-264, :301-312) on the FP64 tensor path
 * (mma.sync m8n8k4.f64); n = 2, 3 only; kept for the measured comparison (profiles/r02_ptm_dense_vs_butterfly.md). */
#define QT_PL_VARIANT_BUTTERFLY 0
#define QT_PL_VARIANT_DENSE_DMMA 1
int qt_superop_pl_batch_variant(int n, int64_t B, const void* in, void* out, void* workspace, int forward, int variant,
                                void* stream);
/* choi2kraus (:325-336), n = 1..3: eigh of the lower triangle; evals_out[B,4^n] ascending (np.linalg.eigh order);
 * kraus_out[B,4^n,d,d]: sqrt(lambda_k) * unvec(v_k) for |lambda_k| > tol in ascending-eigenvalue order, compacted to
 * the front (count_out[b] operators, the rest zero); negative lambda -> i*sqrt(|lambda|) like np.lib.scimath.sqrt.
 * Eigenvector phases are a gauge: compare via kraus2choi(choi2kraus(C)) == C (the reference's own test). */
int qt_choi2kraus_batch(int n, int64_t B, const void* choi, double tol, double* evals_out, void* kraus_out,
                        int32_t* count_out, void* stream);
/* choi2kraus for n = 4, 5 (the 4^n x 4^n eigenproblem does not fit shared memory): one-sided Jacobi out of an L2-resident
 * workspace of qt_choi2kraus_large_workspace_bytes(n, B) bytes; same outputs as qt_choi2kraus_batch;
 * sweeps_out[B] (may be NULL) = Jacobi sweeps taken */
int64_t qt_choi2kraus_large_workspace_bytes(int n, int64_t B);
int qt_choi2kraus_large_batch(int n, int64_t B, const void* choi, double tol, double* evals_out, void* kraus_out,
                              int32_t* count_out, void* workspace, int64_t workspace_bytes, int32_t* sweeps_out,
                              void* stream);

/* ---- distance measures (distance_measures.py) -------------------------------------------------- */
/* rho, sigma: [B, 2^n, 2^n]; out[B] doubles */
int qt_fidelity_batch(int n, int64_t B, const void* rho, const void* sigma, double* out, void* stream);         /* :64-84 */
/* reference semantics: 0.5 * induced 1-norm (max column abs-sum), :100-114 */
int qt_trace_distance_batch(int n, int64_t B, const void* rho, const void* sigma, double* out, void* stream);
/* textbook 0.5 * nuclear norm (extra; not the reference's behaviour) */
int qt_trace_distance_nuclear_batch(int n, int64_t B, const void* rho, const void* sigma, double* out, void* stream);
int qt_purity_batch(int n, int64_t B, const void* rho, double* out, void* stream);                             /* :14-37 */
/* hilbert_schmidt_ip tr(A^dagger B) per pair (:198-216), a, b: [B, rows, cols], out[B] complex; entanglement_fidelity /
 * process_fidelity (:271-375) are this number on Pauli-Liouville matrices, rescaled on the host */
int qt_hs_inner_batch(int64_t rows, int64_t cols, int64_t B, const void* a, const void* b, void* out, void* stream);
/* project_state_matrix_to_physical (operator_tools/project_state_matrix.py:6-52): closest trace-one PSD matrix */
int qt_project_state_batch(int n, int64_t B, const void* rho, void* out, void* stream);

/* ---- Choi-matrix projections (operator_tools/project_superoperators.py), [B,4^n,4^n] ----
 * n = 1..3: shared-memory kernels.  n = 4, 5 (1 MB / 16.8 MB per matrix): the same algorithms out of global memory with
 * the one-sided Jacobi solver of qt_choi2kraus_large_batch; CP needs a workspace (qt_proj_cp_ws_batch), TP / TNI / physical
 * are out-of-place; proj_choi_to_unitary stays n <= 3. */
int qt_proj_cp_batch(int n, int64_t B, const void* choi, void* out, void* stream); /* n = 1..3 */
/* proj_choi_to_completely_positive for n = 1..5: workspace of qt_proj_cp_workspace_bytes(n, B) bytes (0 for n <= 3) */
int64_t qt_proj_cp_workspace_bytes(int n, int64_t B);
int qt_proj_cp_ws_batch(int n, int64_t B, const void* choi, void* out, void* workspace, int64_t workspace_bytes,
                        void* stream);
/* proj_choi_to_unitary (project_superoperators.py:147-175): Choi matrix of the unitary closest to the process
 * (dominant Kraus operator -> polar factor -> phase convention -> kraus2choi), n = 1..3 */
int qt_proj_unitary_batch(int n, int64_t B, const void* choi, void* out, void* stream);
  /* :19-34 */
int qt_proj_tp_batch(int n, int64_t B, const void* choi, void* out, void* stream);   /* :62-84 */
int qt_proj_tni_batch(int n, int64_t B, const void* choi, void* out, void* stream);  /* :37-59 */
/* Dykstra CP+TP (or CP+TNI) projection, :87-144.  A non-Hermitian input is handled like the reference does: the
 * iterates only see its Hermitian part, the anti-Hermitian part enters the stopping rule (so the trip count matches).
 * out must not alias choi.  workspace: device buffer of qt_proj_physical_workspace_bytes(n, B) bytes;
 * eigh_calls_out[B] (may be NULL) = number of CP projections each item needed; status_out[B] (may be NULL) =
 * QT_STATUS_* bits. */
int64_t qt_proj_physical_workspace_bytes(int n, int64_t B);
int qt_proj_physical_batch(int n, int64_t B, const void* choi, void* out, int make_trace_preserving,
                           double eigh_rel_tol, void* workspace, int64_t workspace_bytes, int32_t* eigh_calls_out,
                           int32_t* status_out, void* stream);

/* ---- process tomography: pgdb_process_estimate (tomography.py:542-594) ------------------------- */
typedef struct qt_pgdb_plan qt_pgdb_plan;
/* One entry per result, in the order of `results`: state_codes_host[S*n] (per-qubit input-state codes,
 * qubits[0] first), pauli_idx_host[S], coeff_host[S].  n = 1..3. */
int qt_pgdb_plan_create(int n, int S, const int32_t* state_codes_host, const int32_t* pauli_idx_host,
                        const double* coeff_host, qt_pgdb_plan** plan);
int qt_pgdb_plan_destroy(qt_pgdb_plan* plan);
/* number of distinct input states; canonical = settings are product(states) x all traceless Paulis */
int qt_pgdb_plan_info(const qt_pgdb_plan* plan, int32_t* n_in_out, int32_t* canonical_out);
int64_t qt_pgdb_workspace_bytes(const qt_pgdb_plan* plan, int64_t B);
/* expect[B,S], counts[B,S] -> choi_out[B,4^n,4^n]; counters_out[B,4] (may be NULL) = outer iterations,
 * cost evaluations, eigh calls (the trip counts of tomography.py:570, :576/:582 and project_superoperators.py:115)
 * and the total number of Jacobi sweeps those eigh calls took (a cost figure; no reference counterpart);
 * status_out[B] (may be NULL) = QT_STATUS_* bits; eigh_rel_tol: see QT_EIGH_REL_TOL_DEFAULT */
int qt_pgdb_process_batch(const qt_pgdb_plan* plan, int64_t B, const double* expect, const double* counts,
                          int trace_preserving, double eigh_rel_tol, void* choi_out, int32_t* counters_out,
                          int32_t* status_out, void* workspace, int64_t workspace_bytes, void* stream);
/* linear_inv_process_estimate (tomography.py:459-491): expect[B,S] -> choi_out[B,m,m], minimum-norm least squares
 * over the plan's settings list plus the identity term.  The per-observable pseudo-inverse weights are built from
 * the plan on first use (the reference's dense pinv of the S x 16^n measurement matrix is never formed). */
int qt_linear_inv_process_batch(qt_pgdb_plan* plan, int64_t B, const double* expect, void* choi_out, void* stream);

/* ---- multi-GPU: the one collective of the path for hosts without torch.distributed ----------------
 * Single process, one communicator per device (ncclCommInitAll; NCCL is bound at run time with dlopen).
 * qt_allgather_bytes: device r contributes sendbufs[r][0 .. nbytes_per_rank) and receives all slices in rank order in
 * recvbufs[r][0 .. ndev * nbytes_per_rank); streams[r] (or NULL: default streams) orders the collective on device r. */
typedef struct qt_comm qt_comm;
int qt_comm_init_all(int ndev, const int32_t* devices, qt_comm** comm_out);
int qt_allgather_bytes(qt_comm* comm, const void* const* sendbufs, void* const* recvbufs, int64_t nbytes_per_rank,
                       void* const* streams);
int qt_comm_destroy(qt_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* QTOMO_H */
