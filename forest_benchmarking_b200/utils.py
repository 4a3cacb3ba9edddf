"""Integer Pauli / input-state bookkeeping (must be bit-exact with the reference).

Canonical Pauli index: base-4 number with digits I=0, X=1, Y=2, Z=3 and the FIRST qubit most
significant -- the order of ``itertools.product('IXYZ', repeat=n)`` used by
``all_traceless_pauli_terms`` (reference utils.py:146-156) and ``n_qubit_pauli_basis`` (utils.py:398-409).
One-qubit input-state code: 0..5 = +X,-X,+Y,-Y,+Z,-Z (tomography.py:89), 6..9 = SIC0..SIC3 (:71).
"""
import itertools
from typing import List, Sequence

import numpy as np

from .paulis import PauliTerm

_DIGIT = {"I": 0, "X": 1, "Y": 2, "Z": 3}
_STATE_CODE = {("X", 0): 0, ("X", 1): 1, ("Y", 0): 2, ("Y", 1): 3, ("Z", 0): 4, ("Z", 1): 5,
               ("SIC", 0): 6, ("SIC", 1): 7, ("SIC", 2): 8, ("SIC", 3): 9}


def str_to_pauli_term(pauli_str: str, qubit_labels=None) -> PauliTerm:
    """reference utils.py:127-143."""
    if qubit_labels is None:
        qubit_labels = list(range(len(pauli_str)))
    return PauliTerm.from_list(list(zip(pauli_str, qubit_labels)))


def all_traceless_pauli_terms(qubits: Sequence[int]) -> List[PauliTerm]:
    """reference utils.py:146-156."""
    strs = ["".join(x) for x in itertools.product("IXYZ", repeat=len(qubits))][1:]
    return [str_to_pauli_term(s, qubits) for s in strs]


def pauli_labels(n: int) -> List[str]:
    return ["".join(x) for x in itertools.product("IXYZ", repeat=n)]


def pauli_term_to_index(term, qubits: Sequence[int]) -> int:
    """Canonical index of a Pauli term on ``qubits`` (qubits[0] = most significant digit).
    Raises if the term acts on a qubit outside ``qubits`` (the reference would silently drop it)."""
    idx = 0
    for q in qubits:
        idx = idx * 4 + _DIGIT[term[q]]
    if hasattr(term, "get_qubits"):
        extra = set(term.get_qubits()) - set(qubits)
        if extra:
            raise ValueError(f"observable acts on qubits {sorted(extra)} outside {list(qubits)}")
    return idx


def pauli_index_to_term(idx: int, qubits: Sequence[int], coefficient=1.0) -> PauliTerm:
    n = len(qubits)
    ops = ["IXYZ"[(idx >> (2 * (n - 1 - k))) & 3] for k in range(n)]
    return PauliTerm.from_list(list(zip(ops, qubits)), coefficient)


def real_coefficient(term) -> float:
    c = complex(term.coefficient)
    if abs(c.imag) > 1e-12 * max(1.0, abs(c.real)):
        raise ValueError(f"observable coefficient {c} is not real")
    return float(c.real)


def in_state_codes(in_state, qubits: Sequence[int]):
    """Per-qubit state codes of a TensorProductState, qubits[0] first."""
    out = []
    for q in qubits:
        s = in_state[q]
        out.append(_STATE_CODE[(s.label, int(s.index))])
    return tuple(out)


def flatten_state_results(results, qubits):
    """results -> (pauli_idx[K] int32, coeff[K] f64, expectation[K] f64, counts[K] f64)."""
    k = len(results)
    idx = np.empty(k, dtype=np.int32)
    cf = np.empty(k)
    ex = np.empty(k)
    cnt = np.empty(k)
    for i, r in enumerate(results):
        idx[i] = pauli_term_to_index(r.setting.observable, qubits)
        cf[i] = real_coefficient(r.setting.observable)
        e = complex(r.expectation)
        ex[i] = e.real
        cnt[i] = r.total_counts
    return idx, cf, ex, cnt


def flatten_process_results(results, qubits):
    """results -> (state_codes[K,n] int32, pauli_idx[K] int32, coeff[K], expectation[K], counts[K])."""
    k, n = len(results), len(qubits)
    codes = np.empty((k, n), dtype=np.int32)
    idx = np.empty(k, dtype=np.int32)
    cf = np.empty(k)
    ex = np.empty(k)
    cnt = np.empty(k)
    for i, r in enumerate(results):
        codes[i] = in_state_codes(r.setting.in_state, qubits)
        idx[i] = pauli_term_to_index(r.setting.observable, qubits)
        cf[i] = real_coefficient(r.setting.observable)
        ex[i] = complex(r.expectation).real
        cnt[i] = r.total_counts
    return codes, idx, cf, ex, cnt
