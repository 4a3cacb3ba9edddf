"""CPU oracle: a NumPy/SciPy restatement of the forest-benchmarking tomography hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``forest_benchmarking_b200/`` imports this module; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the checker / CPU baseline, never as the product path.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs every function here against the
unmodified reference imported from /root/reference (through ``oracle/pyquil_shim``) when that tree
is present, and ``tests/golden/*.npz`` (written by ``oracle/make_golden.py`` from the reference
itself) pin it on machines where the reference tree is absent (the GPU box).

All file:line citations are relative to /root/reference/forest/benchmarking/.

Conventions (same as the reference):
  * ``qubits[0]`` is the LEFT-most tensor factor (tomography.py:229-233).
  * Pauli index = base-4 number with digits I=0, X=1, Y=2, Z=3, first qubit most significant --
    the order of ``itertools.product('IXYZ', repeat=n)`` (utils.py:146-156, utils.py:398-409).
  * One-qubit input-state code: 0..5 = +X,-X,+Y,-Y,+Z,-Z (tomography.py:89), 6..9 = SIC0..SIC3
    (tomography.py:71); a product state is a tuple of codes, first qubit first.
  * ``vec`` stacks columns (operator_tools/superoperator_transformations.py:33-51).
"""
import itertools

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------------------------
# Pauli / state bookkeeping  (utils.py:146-156, 398-409; pyquil.simulation.matrices, restated)
# --------------------------------------------------------------------------------------------
_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
PAULIS_1Q = (_I2, _X, _Y, _Z)

STATE_LABELS = ("X+", "X-", "Y+", "Y-", "Z+", "Z-", "SIC0", "SIC1", "SIC2", "SIC3")


def one_qubit_state_vector(code):
    """pyquil.simulation.matrices.STATES restated; code as in the module docstring."""
    s2, s3 = np.sqrt(2.0), np.sqrt(3.0)
    table = [
        np.array([1, 1]) / s2, np.array([1, -1]) / s2,
        np.array([1, 1j]) / s2, np.array([1, -1j]) / s2,
        np.array([1, 0]), np.array([0, 1]),
        np.array([1, 0]),
        np.array([1, s2]) / s3,
        np.array([1, np.exp(-2j * np.pi / 3) * s2]) / s3,
        np.array([1, np.exp(2j * np.pi / 3) * s2]) / s3,
    ]
    return np.asarray(table[code], dtype=complex)


def pauli_digits(idx, n):
    """Base-4 digits of a Pauli index, first qubit first."""
    return [(idx >> (2 * (n - 1 - q))) & 3 for q in range(n)]


def pauli_labels(n):
    """Canonical label order, utils.py:153 (itertools.product('IXYZ', repeat=n))."""
    return ["".join(t) for t in itertools.product("IXYZ", repeat=n)]


def pauli_matrix(idx, n):
    """n-qubit Pauli matrix for a canonical index; equals ``n_qubit_pauli_basis(n).ops[idx]``
    (utils.py:398-409) and ``lifted_pauli(term, qubits[::-1])`` (tomography.py:233,327)."""
    m = np.eye(1, dtype=complex)
    for dgt in pauli_digits(idx, n):
        m = np.kron(m, PAULIS_1Q[dgt])
    return m


def product_state_matrix(codes):
    """Density matrix of a product input state, first qubit left-most
    (``lifted_state_operator(in_state, qubits[::-1])``, tomography.py:513)."""
    m = np.eye(1, dtype=complex)
    for c in codes:
        v = one_qubit_state_vector(c).reshape(2, 1)
        m = np.kron(m, v @ v.conj().T)
    return m


def process_tomography_settings(n, in_basis="pauli"):
    """(state codes, pauli idx) for every setting in generator order (tomography.py:71-97)."""
    codes_1q = range(0, 6) if in_basis.lower() == "pauli" else range(6, 10)
    out = []
    for st in itertools.product(codes_1q, repeat=n):
        for k in range(1, 4 ** n):
            out.append((tuple(st), k))
    return out


# --------------------------------------------------------------------------------------------
# vec / unvec and conversions  (operator_tools/superoperator_transformations.py)
# --------------------------------------------------------------------------------------------
def vec(m):
    """Column stacking, :33-51."""
    return np.asarray(m).T.reshape(-1, 1)


def unvec(v, shape=None):
    """:54-79."""
    v = np.asarray(v)
    if shape is None:
        d = int(round(np.sqrt(v.size)))
        shape = (d, d)
    return v.reshape(*shape).T


def _as_kraus_list(kraus_ops):
    if isinstance(kraus_ops, np.ndarray) and kraus_ops.ndim == 2:
        return [kraus_ops]
    return [np.asarray(k) for k in kraus_ops]


def kraus2choi(kraus_ops):
    """sum_k vec(K) vec(K)^dagger, :159-182."""
    ks = _as_kraus_list(kraus_ops)
    tot = 0
    for k in ks:
        v = vec(k)
        tot = tot + v @ v.conj().T
    return tot


def kraus2superop(kraus_ops):
    """sum_k conj(K) (x) K, :100-145 (non-square Kraus allowed)."""
    ks = _as_kraus_list(kraus_ops)
    r, c = ks[0].shape
    s = np.zeros((r * r, c * c), dtype=complex)
    for k in ks:
        s += np.kron(k.conj(), k)
    return s


def reshuffle(m):
    """choi2superop == superop2choi: reshape [d]*4, swap axes 0 and 3, :267-277, :351-361."""
    m = np.asarray(m)
    d = int(round(np.sqrt(m.shape[0])))
    return m.reshape(d, d, d, d).swapaxes(0, 3).reshape(d * d, d * d)


choi2superop = reshuffle
superop2choi = reshuffle


def pauli2computational_basis_matrix(dim):
    """Columns are vec(P_i), :374-408."""
    n = int(round(np.log2(dim)))
    m = np.zeros((dim * dim, dim * dim), dtype=complex)
    for i in range(4 ** n):
        m[:, i] = vec(pauli_matrix(i, n))[:, 0]
    return m


def computational2pauli_basis_matrix(dim):
    """:411-438."""
    return pauli2computational_basis_matrix(dim).conj().T / dim


def superop2pauli_liouville(s):
    """c2p @ S @ c2p^dagger * d, :253-264."""
    s = np.asarray(s)
    d = int(round(np.sqrt(s.shape[0])))
    c2p = computational2pauli_basis_matrix(d)
    return c2p @ s @ c2p.conj().T * d


def pauli_liouville2superop(pl):
    """p2c @ R @ p2c^dagger / d, :301-312."""
    pl = np.asarray(pl)
    d = int(round(np.sqrt(pl.shape[0])))
    p2c = pauli2computational_basis_matrix(d)
    return p2c @ pl @ p2c.conj().T / d


def choi2pauli_liouville(c):
    """:364-371."""
    return superop2pauli_liouville(choi2superop(c))


def pauli_liouville2choi(pl):
    """:315-322."""
    return superop2choi(pauli_liouville2superop(pl))


def kraus2pauli_liouville(kraus_ops):
    """:148-156."""
    return superop2pauli_liouville(kraus2superop(kraus_ops))


def choi2kraus(choi, tol=1e-9):
    """eigh; sqrt(lambda) * unvec(v) for |lambda| > tol (complex sqrt of negative lambda), :325-336."""
    w, v = np.linalg.eigh(choi)
    out = []
    for lam, col in zip(w, v.T):
        if abs(lam) > tol:
            out.append(np.emath.sqrt(lam) * unvec(col.reshape(-1, 1)))
    return out


def kraus2chi(kraus_ops):
    """chi (process) matrix: sum_k c_k c_k^dagger with c_k = c2p vec(K_k), :82-97."""
    ks = _as_kraus_list(kraus_ops)
    c2p = computational2pauli_basis_matrix(ks[0].shape[0])
    return sum((c2p @ vec(k)) @ (c2p @ vec(k)).conj().T for k in ks)


def chi2choi(chi):
    """p2c chi p2c^dagger, :217-226."""
    p2c = pauli2computational_basis_matrix(int(np.sqrt(np.asarray(chi).shape[0])))
    return p2c @ np.asarray(chi) @ p2c.conj().T


def chi2pauli_liouville(chi):
    """:185-192."""
    return choi2pauli_liouville(chi2choi(chi))


def chi2superop(chi):
    """:207-214."""
    return pauli_liouville2superop(chi2pauli_liouville(chi))


def choi2chi(choi):
    """kraus2chi(choi2kraus(choi)), :339-348 (drops |eigenvalue| <= 1e-9, takes |.| of negative ones)."""
    return kraus2chi(choi2kraus(choi))


def superop2chi(s):
    """:241-250."""
    return choi2chi(reshuffle(s))


def pauli_liouville2chi(pl):
    """:291-298."""
    return choi2chi(pauli_liouville2choi(pl))


# --------------------------------------------------------------------------------------------
# Projections  (operator_tools/project_superoperators.py, calculational.py)
# --------------------------------------------------------------------------------------------
def partial_trace_out(choi):
    """Tr over the second (output) factor of a d^2 x d^2 matrix: calculational.py:5-35 with
    keep=[0], dims=[d, d]."""
    d = int(round(np.sqrt(choi.shape[0])))
    return np.einsum("ijkj->ik", np.asarray(choi).reshape(d, d, d, d))


def proj_choi_to_completely_positive(choi):
    """Hermitise, eigh, clamp negative eigenvalues, recompose; project_superoperators.py:19-34."""
    h = (choi + choi.conj().T) / 2
    w, v = sla.eigh(h)
    w = np.where(w < 0, 0.0, w)
    return (v * w) @ v.conj().T


def proj_choi_to_trace_preserving(choi):
    """C - kron((Tr_out C - I)/d, I), project_superoperators.py:62-84."""
    d = int(round(np.sqrt(choi.shape[0])))
    pt = partial_trace_out(choi)
    return choi - np.kron((pt - np.eye(d)) / d, np.eye(d))


def proj_choi_to_trace_non_increasing(choi):
    """project_superoperators.py:37-59."""
    d = int(round(np.sqrt(choi.shape[0])))
    pt = partial_trace_out(choi)
    w, v = sla.eigh((pt + pt.conj().T) / 2)
    w = np.where(w > 1, 1.0, w)
    proj = (v * w) @ v.conj().T
    return choi - np.kron((pt - proj) / d, np.eye(d))


def proj_choi_to_physical(choi, make_trace_preserving=True, return_count=False):
    """Dykstra alternating projections with the Birgin-Raydan stop rule,
    project_superoperators.py:87-144."""
    q_cp = np.zeros_like(choi)      # old_CP_change
    q_tp = np.zeros_like(choi)      # old_TP_change
    cp_prev = np.zeros_like(choi)   # last_CP_projection
    state = choi
    n_eigh = 0
    while True:
        pre_cp = state - q_cp
        cp = proj_choi_to_completely_positive(pre_cp)
        n_eigh += 1
        d_cp = cp - pre_cp
        pre_tp = cp - q_tp
        if make_trace_preserving:
            new_state = proj_choi_to_trace_preserving(pre_tp)
        else:
            new_state = proj_choi_to_trace_non_increasing(pre_tp)
        d_tp = new_state - pre_tp
        crit = (np.linalg.norm(d_cp - q_cp) ** 2 + np.linalg.norm(d_tp - q_tp) ** 2
                + 2 * abs(np.vdot(q_tp, new_state - state))
                + 2 * abs(np.vdot(q_cp, cp - cp_prev)))
        if crit < 1e-4:
            break
        q_cp, q_tp, cp_prev, state = d_cp, d_tp, cp, new_state
    if return_count:
        return new_state, n_eigh
    return new_state


# --------------------------------------------------------------------------------------------
# Distance measures  (distance_measures.py:64-114, calculational.py:77-91)
# --------------------------------------------------------------------------------------------
def proj_choi_to_unitary(choi):
    """project_superoperators.py:147-175."""
    choi = np.asarray(choi, dtype=complex)
    d = int(round(np.sqrt(choi.shape[0])))
    vals, vs = sla.eigh((choi + choi.conj().T) / 2)
    kraus = unvec(vs[:, np.argmax(vals)].reshape(d * d, 1))
    u, _, vh = sla.svd(kraus)
    unitary = u @ vh
    return kraus2choi(np.exp(-1j * np.angle(unitary[0, 0])) * unitary)


def sqrtm_psd(m):
    """calculational.py:77-91."""
    w, v = sla.eigh(m)
    w = np.sqrt(np.maximum(w, 0))
    return (v * w) @ v.conj().T


def fidelity(rho, sigma):
    """(tr sqrt(sqrt(rho) sigma sqrt(rho)))^2, distance_measures.py:64-84 (real part returned)."""
    s = sqrtm_psd(rho)
    f = np.trace(sqrtm_psd(s @ sigma @ s)) ** 2
    return float(np.real(f))


def trace_distance(rho, sigma):
    """0.5 * induced 1-norm (max column abs-sum) -- sic, distance_measures.py:114; pinned by the
    reference's own test_distance_measures.py:73-82."""
    return 0.5 * float(np.max(np.sum(np.abs(np.asarray(rho) - np.asarray(sigma)), axis=0)))


def hilbert_schmidt_ip(a, b):
    """tr(A^dagger B), distance_measures.py:198-216."""
    return np.trace(np.asarray(a).conj().T @ np.asarray(b))


def entanglement_fidelity(pl0, pl1):
    """tr(E^dagger F) / dim^2, distance_measures.py:271-303."""
    return hilbert_schmidt_ip(pl0, pl1) / np.asarray(pl0).shape[0]


def process_fidelity(pl0, pl1):
    """(dim F_e + 1) / (dim + 1), distance_measures.py:306-361."""
    dim = int(np.sqrt(np.asarray(pl0).shape[0]))
    return (dim * entanglement_fidelity(pl0, pl1) + 1) / (dim + 1)


def purity(rho, dim_renorm=False):
    """distance_measures.py:14-37."""
    p = np.real(np.trace(rho @ rho))
    if dim_renorm:
        d = rho.shape[0]
        p = (d / (d - 1.0)) * (p - 1.0 / d)
    return float(p)


def impurity(rho, dim_renorm=False):
    """distance_measures.py:40-61."""
    imp = 1 - np.real(np.trace(rho @ rho))
    if dim_renorm:
        d = rho.shape[0]
        imp = (d / (d - 1.0)) * imp
    return float(imp)


def project_state_matrix_to_physical(rho):
    """Smolin "wizard" projection onto trace-one PSD matrices, operator_tools/project_state_matrix.py:6-52:
    normalise the trace, eigh, and if any eigenvalue is negative zero the smallest ones while spreading their
    (negative) mass evenly over the eigenvalues that stay (water filling from the bottom)."""
    rho = np.asarray(rho, dtype=complex)
    rho = rho / np.trace(rho)
    w, v = sla.eigh(rho)
    if w.min() >= 0:
        return rho
    d = len(w)
    desc = list(w[::-1])
    new = [0.0] * d
    i, acc = d, 0.0
    while desc[i - 1] + acc / float(i) < 0:
        acc += desc[i - 1]
        i -= 1
    for j in range(i):
        new[j] = desc[j] + acc / float(i)
    new.reverse()
    return (v * np.array(new)) @ v.conj().T


def shots_to_obs_moments(bitarray, idxs, coeff=1.0, use_beta_dist_unbiased_prior=False):
    """observable_estimation.py:804-853 with the observable given as (column indices, real coefficient)."""
    bitarray = np.asarray(bitarray)
    if len(idxs) == 0:
        return coeff, 0
    obs_vals = np.prod(1 - 2 * bitarray[:, list(idxs)].astype(np.int64), axis=1)
    if use_beta_dist_unbiased_prior:
        n_plus = int(np.sum(obs_vals == 1))
        n_minus = len(obs_vals) - n_plus
        a, b = n_plus + 1, n_minus + 1                      # scipy.stats.beta.mean / .var of beta(a, b)
        bm, bv = a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1))
        return (2 * bm - 1) * coeff, 4 * bv * coeff ** 2   # utils.py:446-458
    obs_vals = coeff * obs_vals
    return np.mean(obs_vals).item(), np.var(obs_vals).item() / len(bitarray)


def ratio_variance(a, var_a, b, var_b):
    """observable_estimation.py:1052-1090."""
    return var_a / b ** 2 + (a ** 2 * var_b) / b ** 4


def resample_expectations_with_beta(expectations, counts, prior_counts=1):
    """tomography.py:378-409 on arrays: one np.random.beta draw per result, in order (global NumPy RNG)."""
    out = np.empty(len(expectations))
    for k, (e, n) in enumerate(zip(expectations, counts)):
        num_plus = ((e + 1) / 2) * n
        num_minus = n - num_plus
        out[k] = 2 * np.random.beta(num_plus + prior_counts, num_minus + prior_counts) - 1
    return out


# --------------------------------------------------------------------------------------------
# State tomography  (tomography.py:130-338)
# --------------------------------------------------------------------------------------------
def linear_inv_state_estimate(pauli_idx, coeffs, expectations, n):
    """unvec(pinv(M) e) + I/d with rows vec(P_k)^dagger, tomography.py:130-165."""
    d = 2 ** n
    rows = [vec(c * pauli_matrix(k, n)).T.conj() for k, c in zip(pauli_idx, coeffs)]
    m = np.vstack(rows)
    r = sla.pinv(m) @ np.asarray(expectations, dtype=float)
    return unvec(r) + np.eye(d) / d


def state_log_likelihood(rho, pauli_idx, coeffs, expectations, counts, n):
    """tomography.py:341-375 (log10 likelihood; outcomes of non-positive predicted probability are skipped)."""
    ll = 0
    for k, c, meas, cnt in zip(pauli_idx, coeffs, expectations, counts):
        pred = np.real(np.trace(c * pauli_matrix(k, n) @ rho))
        for sign in (1, -1):
            f = cnt * (1 + sign * meas) / 2
            pr = (1 + sign * pred) / 2
            if pr <= 0:
                continue
            ll += f * np.log10(pr)
    return ll


def r_operator(rho, op_mats, expectations):
    """Eq. 4 of the diluted-MLE paper as the reference codes it, tomography.py:273-338."""
    tiny = np.finfo(float).tiny
    d = rho.shape[0]
    eye = np.eye(d)
    upd = np.zeros((d, d), dtype=complex)
    for op, m in zip(op_mats, expectations):
        pred = np.trace(op @ rho)
        for sgn in (1, -1):
            upd += ((1 + sgn * m) / 2) / ((1 + sgn * pred) / 2 + tiny) * (eye + sgn * op) / 2
    return upd / len(op_mats)


def mle_state_estimate(pauli_idx, coeffs, expectations, counts, n, epsilon=.1, entropy_penalty=0.0,
                       beta=0.0, tol=1e-9, maxiter=10_000, rebuild_paulis=False):
    """Scalar (one experiment) restatement of iterative_mle_state_estimate, tomography.py:168-270.

    Returns (rho, iteration) where ``iteration`` is the reference's loop counter at exit:
    the number of updates applied when converged, ``maxiter`` when the cap was hit (in which case
    maxiter-1 updates were applied, tomography.py:244-246).
    ``rebuild_paulis=True`` re-krons every Pauli matrix on every iteration like the reference does
    through ``lifted_pauli`` (tomography.py:327) -- used only for honest CPU-baseline timing.
    """
    if entropy_penalty != 0.0 and beta != 0.0:
        raise ValueError("entropy_penalty and beta cannot both be non-zero")
    d = 2 ** n
    eye = np.eye(d)
    num_meas = float(np.sum(counts))
    ops = [c * pauli_matrix(k, n) for k, c in zip(pauli_idx, coeffs)]
    rho = eye / d
    it = 1
    while True:
        prev = rho
        if it >= maxiter:
            break
        if rebuild_paulis:
            ops = [c * pauli_matrix(k, n) for k, c in zip(pauli_idx, coeffs)]
        tk = r_operator(rho, ops, expectations) - eye
        if entropy_penalty > 0.0:
            lg = sla.logm(rho)
            tk = tk - entropy_penalty * (lg - eye * np.trace(rho @ lg))
        if beta > 0.0:
            tk = tk * (num_meas / 2)
            tk = tk + beta * (sla.pinv(rho) - d * eye) / 2
        m = eye + epsilon * tk
        rho = m @ rho @ m
        rho = rho / np.trace(rho)
        if np.linalg.norm(rho - prev) < tol:
            break
        it += 1
    return rho, it


def mle_state_estimate_batch(pauli_idx, coeffs, expectations, n, epsilon=.1, tol=1e-9,
                             maxiter=10_000):
    """Vectorised-over-experiments form of the vanilla path above (entropy_penalty = beta = 0).
    Same arithmetic per item; finished items are frozen.  expectations: [B, K]."""
    e = np.asarray(expectations, dtype=float)
    b, k = e.shape
    d = 2 ** n
    tiny = np.finfo(float).tiny
    ops = np.stack([c * pauli_matrix(i, n) for i, c in zip(pauli_idx, coeffs)])  # [K,d,d]
    eye = np.eye(d)
    rho = np.broadcast_to(eye / d, (b, d, d)).astype(complex).copy()
    iters = np.ones(b, dtype=np.int32)
    active = np.ones(b, dtype=bool)
    while True:
        active &= iters < maxiter
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        r = rho[idx]
        pred = np.einsum("kij,bji->bk", ops, r)
        m = e[idx]
        ap = ((1 + m) / 2) / ((1 + pred) / 2 + tiny)
        am = ((1 - m) / 2) / ((1 - pred) / 2 + tiny)
        big_r = (np.einsum("bk,kij->bij", (ap - am) / 2, ops)
                 + ((ap + am) / 2).sum(axis=1)[:, None, None] * eye) / k
        mm = eye + epsilon * (big_r - eye)
        new = mm @ r @ mm
        new = new / np.trace(new, axis1=1, axis2=2)[:, None, None]
        diff = np.linalg.norm(new - r, axis=(1, 2))
        rho[idx] = new
        conv = diff < tol
        active[idx[conv]] = False
        iters[idx[~conv]] += 1
    return rho, iters


# --------------------------------------------------------------------------------------------
# Process tomography  (tomography.py:459-633)
# --------------------------------------------------------------------------------------------
def extract_design(settings, coeffs, expectations, counts, n):
    """Dense A (2K x d^4, / d^2) and n (2K x 1, / grand total), tomography.py:494-539.
    ``settings`` = list of (state codes, pauli idx)."""
    d = 2 ** n
    eye = np.eye(d)
    rows, cnt = [], []
    cache_s, cache_p = {}, {}
    for (codes, k), c, e, tot in zip(settings, coeffs, expectations, counts):
        if codes not in cache_s:
            cache_s[codes] = product_state_matrix(codes)
        if (k, c) not in cache_p:
            cache_p[(k, c)] = c * pauli_matrix(k, n)
        rin, op = cache_s[codes], cache_p[(k, c)]
        for sgn in (1, -1):
            proj = (eye + sgn * op) / 2
            rows.append(vec(np.kron(rin, proj.T)).T[0])
        plus = (1 + e) / 2
        cnt += [tot * plus, tot * (1 - plus)]
    a = np.asarray(rows) / d ** 2
    nn = np.asarray(cnt, dtype=float)[:, None] / float(np.sum(counts))
    return a, nn


def pgdb_cost(a, nn, est, eps=1e-6):
    """tomography.py:597-614.  Like the reference this keeps the complex dtype: np.clip / np.log act
    on complex numbers (lexicographic clip == clip of the real part; the imaginary parts are ~1e-17
    rounding noise) and the returned cost is a complex scalar whose comparisons are lexicographic."""
    p = a @ vec(est)
    p = np.clip(p, eps, None)
    return (-nn.T @ np.log(p))[0, 0]


def pgdb_grad(a, nn, est, eps=1e-6):
    """tomography.py:617-633."""
    p = a @ vec(est)
    p = np.clip(p, eps, None)
    return unvec(-a.conj().T @ (nn / p))


def pgdb_process_estimate(settings, coeffs, expectations, counts, n, trace_preserving=True,
                          return_counters=False):
    """Projected gradient descent with backtracking, tomography.py:542-594."""
    a, nn = extract_design(settings, coeffs, expectations, counts, n)
    d = 2 ** n
    est = np.eye(d * d, dtype=complex) / d
    old_cost = pgdb_cost(a, nn, est)
    mu = 3 / (2 * d ** 2)
    gamma = .3
    outer = cost_evals = eighs = 0
    cost_evals += 1
    while True:
        outer += 1
        g = pgdb_grad(a, nn, est)
        proj, ne = proj_choi_to_physical(est - g / mu, trace_preserving, return_count=True)
        eighs += ne
        upd = proj - est
        alpha = 1.0
        new_cost = pgdb_cost(a, nn, est + alpha * upd)
        cost_evals += 1
        change = gamma * alpha * np.vdot(upd, g)
        while new_cost > old_cost + change:
            alpha *= .5
            change *= .5
            new_cost = pgdb_cost(a, nn, est + alpha * upd)
            cost_evals += 1
            if alpha < 1e-15:
                break
        est = est + alpha * upd
        if old_cost - new_cost < 1e-10:
            break
        old_cost = new_cost
    if return_counters:
        return est, dict(outer=outer, cost_evals=cost_evals, eighs=eighs)
    return est


def linear_inv_process_estimate(settings, coeffs, expectations, n):
    """tomography.py:459-491."""
    d = 2 ** n
    rows = []
    for (codes, k), c in zip(settings, coeffs):
        rin = product_state_matrix(codes)
        rows.append(vec(np.kron(rin.conj(), c * pauli_matrix(k, n))).conj().T)
    m = np.vstack(rows)
    r = sla.pinv(m) @ np.asarray(expectations, dtype=float)
    return unvec(r) + np.eye(d * d) / d


# --------------------------------------------------------------------------------------------
# Synthetic data (SURVEY.md 8d generator recipe; random_operators.py:21-107 semantics)
# --------------------------------------------------------------------------------------------
def ginibre_state(rng, d, rank=None):
    """G G^dagger / tr with the real block drawn first (random_operators.py:90-107)."""
    r = d if rank is None else rank
    g = rng.standard_normal((d, r)) + 1j * rng.standard_normal((d, r))
    rho = g @ g.conj().T
    return rho / np.trace(rho)


def haar_unitary(rng, d):
    """QR of a Ginibre matrix with the diag(R)/|diag(R)| phase fix (random_operators.py:49-72)."""
    g = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / np.sqrt(2)
    q, r = np.linalg.qr(g)
    ph = np.diag(r) / np.abs(np.diag(r))
    return q * ph


def synth_state_tomography(seed, batch, n, shots=1000, rank=None):
    """Config-2 style data: returns (rho_true [B,d,d], pauli_idx [K], expectations [B,K],
    counts [B,K]) with K = 4^n - 1 canonical Paulis and binomial shot noise."""
    rng = np.random.default_rng(seed)
    d = 2 ** n
    k = 4 ** n - 1
    ops = np.stack([pauli_matrix(i, n) for i in range(1, 4 ** n)])
    truth = np.empty((batch, d, d), dtype=complex)
    ex = np.empty((batch, k))
    for b in range(batch):
        rho = ginibre_state(rng, d, rank)
        truth[b] = rho
        e = np.clip(np.real(np.einsum("kij,ji->k", ops, rho)), -1, 1)
        plus = rng.binomial(shots, (1 + e) / 2)
        ex[b] = (2 * plus - shots) / shots
    return truth, np.arange(1, 4 ** n, dtype=np.int32), ex, np.full((batch, k), shots, dtype=float)


def apply_choi(choi, rho):
    """Lambda(rho) = Tr_in[(rho^T (x) I) C] (apply_superoperator.py:60-90 semantics)."""
    d = rho.shape[0]
    return np.einsum("ac,abcd->bd", rho, np.asarray(choi).reshape(d, d, d, d))


def synth_process_tomography(seed, batch, n, shots=1000, in_basis="pauli", unitary=True):
    """Config-3 style data: Haar-random unitary channel per item (or a 2-Kraus mixture when
    ``unitary=False``); returns (choi_true [B,d^2,d^2], settings, expectations [B,S], counts [B,S])."""
    rng = np.random.default_rng(seed)
    d = 2 ** n
    settings = process_tomography_settings(n, in_basis)
    states = sorted({s for s, _ in settings}, key=lambda s: [c for c in s])
    ops = np.stack([pauli_matrix(i, n) for i in range(1, 4 ** n)])
    nset = len(settings)
    truth = np.empty((batch, d * d, d * d), dtype=complex)
    ex = np.empty((batch, nset))
    state_mats = {s: product_state_matrix(s) for s in states}
    for b in range(batch):
        if unitary:
            choi = kraus2choi([haar_unitary(rng, d)])
        else:
            choi = kraus2choi([np.sqrt(.7) * haar_unitary(rng, d), np.sqrt(.3) * haar_unitary(rng, d)])
        truth[b] = choi
        pos = 0
        for st in itertools.product(*[sorted({s[q] for s in states}) for q in range(n)]):
            out = apply_choi(choi, state_mats[tuple(st)])
            e = np.clip(np.real(np.einsum("kij,ji->k", ops, out)), -1, 1)
            plus = rng.binomial(shots, (1 + e) / 2)
            ex[b, pos:pos + len(e)] = (2 * plus - shots) / shots
            pos += len(e)
    return truth, settings, ex, np.full((batch, nset), shots, dtype=float)
