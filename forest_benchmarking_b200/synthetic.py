"""Vectorised synthetic tomography data (host side, NumPy) for benchmarks and examples.

Distributions follow SURVEY.md 8(d): Ginibre-random true states / Haar-random unitary processes
(reference operator_tools/random_operators.py:49-107 semantics), exact Pauli expectations, binomial
shot noise with N shots per setting.  This is input generation, not part of the estimators.
"""
import itertools

import numpy as np

_P1 = np.array([[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=complex)


def pauli_stack(n):
    """[4^n, d, d] Pauli matrices in canonical order (first qubit most significant)."""
    ops = _P1
    for _ in range(n - 1):
        ops = np.einsum("aij,bkl->abikjl", ops, _P1).reshape(ops.shape[0] * 4, ops.shape[1] * 2, ops.shape[2] * 2)
    return ops


def ginibre_states(rng, batch, d):
    g = rng.standard_normal((batch, d, d)) + 1j * rng.standard_normal((batch, d, d))
    rho = g @ g.conj().transpose(0, 2, 1)
    return rho / np.trace(rho, axis1=1, axis2=2)[:, None, None]


def haar_unitaries(rng, batch, d):
    g = (rng.standard_normal((batch, d, d)) + 1j * rng.standard_normal((batch, d, d))) / np.sqrt(2)
    q, r = np.linalg.qr(g)
    ph = np.diagonal(r, axis1=1, axis2=2)
    return q * (ph / np.abs(ph))[:, None, :]


def state_tomography_batch(seed, batch, n, shots=1000):
    """-> pauli_idx [K] int32, expectations [B, K], counts [B, K], rho_true [B, d, d]; K = 4^n - 1."""
    rng = np.random.default_rng(seed)
    d = 2 ** n
    rho = ginibre_states(rng, batch, d)
    ops = pauli_stack(n)[1:]
    e = np.clip(np.real(np.einsum("kij,bji->bk", ops, rho)), -1, 1)
    plus = rng.binomial(shots, (1 + e) / 2)
    ex = (2.0 * plus - shots) / shots
    return np.arange(1, 4 ** n, dtype=np.int32), ex, np.full(ex.shape, float(shots)), rho


_BLOCH = {0: (1, 0, 0), 1: (-1, 0, 0), 2: (0, 1, 0), 3: (0, -1, 0), 4: (0, 0, 1), 5: (0, 0, -1),
          6: (0, 0, 1),
          7: (2 * np.sqrt(2) / 3, 0, -1 / 3),
          8: (-np.sqrt(2) / 3, -np.sqrt(6) / 3, -1 / 3),
          9: (-np.sqrt(2) / 3, np.sqrt(6) / 3, -1 / 3)}


def input_state_pauli_vectors(n, in_basis="pauli"):
    """[n_in, 4^n] Pauli-expansion coefficients r_i[j] = Tr(P_j rho_i) of every product input state, in
    generator order (reference tomography.py:71-97)."""
    codes = range(0, 6) if in_basis.lower() == "pauli" else range(6, 10)
    one = {c: np.array((1.0,) + tuple(_BLOCH[c])) for c in codes}
    rows = []
    for st in itertools.product(codes, repeat=n):
        v = np.ones(1)
        for c in st:
            v = np.kron(v, one[c])
        rows.append(v)
    return np.array(rows)


def process_tomography_batch(seed, batch, n, shots=1000, in_basis="pauli"):
    """Haar-random unitary channels.  -> state_codes [S, n] int32, pauli_idx [S] int32,
    expectations [B, S], counts [B, S], ptm_true [B, 4^n, 4^n];  S = n_in * (4^n - 1)."""
    rng = np.random.default_rng(seed)
    d = 2 ** n
    u = haar_unitaries(rng, batch, d)
    ops = pauli_stack(n)
    # PTM R[k, j] = Tr(P_k U P_j U^dagger) / d
    upu = np.einsum("bij,ajk,blk->bail", u, ops, u.conj())          # [B, j, d, d]
    ptm = np.real(np.einsum("kil,bjli->bkj", ops, upu)) / d
    svec = input_state_pauli_vectors(n, in_basis)                    # [n_in, 4^n]
    t = np.einsum("bkj,ij->bik", ptm, svec)[:, :, 1:]                # Tr(P_k E(rho_i)), k >= 1
    e = np.clip(t.reshape(batch, -1), -1, 1)
    plus = rng.binomial(shots, (1 + e) / 2)
    ex = (2.0 * plus - shots) / shots
    codes = range(0, 6) if in_basis.lower() == "pauli" else range(6, 10)
    k = 4 ** n - 1
    state_codes = np.repeat(np.array(list(itertools.product(codes, repeat=n)), dtype=np.int32), k, axis=0)
    pauli_idx = np.tile(np.arange(1, 4 ** n, dtype=np.int32), len(svec))
    return state_codes, pauli_idx, ex, np.full(ex.shape, float(shots)), ptm
