"""Restated semantics of pyquil.simulation.tools.lifted_pauli / lifted_state_operator:
left-kron build-up over `qubits`, so the LAST qubit in the list is the left-most tensor factor."""
import numpy as np
from pyquil.simulation.matrices import I, X, Y, Z, STATES

_P = {"I": I, "X": X, "Y": Y, "Z": Z}


def lifted_pauli(pauli_sum, qubits):
    terms = getattr(pauli_sum, "terms", [pauli_sum])
    dim = 2 ** len(qubits)
    out = np.zeros((dim, dim), dtype=np.complex128)
    for term in terms:
        m = np.eye(1)
        for q in qubits:
            m = np.kron(_P[term[q]], m)
        out = out + term.coefficient * m
    return out


def lifted_state_operator(state, qubits):
    mat = np.eye(1)
    for q in qubits:
        oneq = state[q]
        v = np.asarray(STATES[oneq.label][oneq.index]).reshape(2, 1)
        mat = np.kron(v @ v.conj().T, mat)
    return mat
