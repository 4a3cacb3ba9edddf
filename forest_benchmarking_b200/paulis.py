"""Minimal Pauli-term record so results lists can be built without pyquil.

The estimators only read ``term[q]`` -> 'I'/'X'/'Y'/'Z' and ``term.coefficient`` (reference call
sites tomography.py:327,515 via pyquil's ``lifted_pauli``), so pyquil's own ``PauliTerm`` objects are
accepted unchanged (duck typing); this class exists for users and tests that do not have pyquil.
"""
from typing import Iterable, Tuple

_VALID = "IXYZ"


class PauliTerm:
    __slots__ = ("_ops", "coefficient")

    def __init__(self, op: str = "I", index: int = 0, coefficient=1.0):
        if op not in _VALID:
            raise ValueError(f"{op!r} is not one of I, X, Y, Z")
        self._ops = {} if op == "I" else {index: op}
        self.coefficient = complex(coefficient)

    @classmethod
    def from_list(cls, terms_list: Iterable[Tuple[str, int]], coefficient=1.0) -> "PauliTerm":
        term = cls("I", 0, coefficient)
        for op, q in terms_list:
            if op not in _VALID:
                raise ValueError(f"{op!r} is not one of I, X, Y, Z")
            if q in term._ops:
                raise ValueError(f"qubit {q} appears twice")
            if op != "I":
                term._ops[q] = op
        return term

    def __getitem__(self, qubit: int) -> str:
        return self._ops.get(qubit, "I")

    def __iter__(self):
        return iter(self._ops.items())

    def __len__(self):
        return len(self._ops)

    def get_qubits(self):
        return list(self._ops)

    def operations_as_set(self):
        return frozenset(self._ops.items())

    def __eq__(self, other):
        return (isinstance(other, PauliTerm) and self._ops == other._ops
                and abs(self.coefficient - other.coefficient) < 1e-12)

    def __hash__(self):
        return hash((self.operations_as_set(), round(self.coefficient.real, 12), round(self.coefficient.imag, 12)))

    def __rmul__(self, scalar):
        out = PauliTerm("I", 0, self.coefficient * scalar)
        out._ops = dict(self._ops)
        return out

    def __repr__(self):
        body = "*".join(f"{op}{q}" for q, op in self._ops.items()) or "I"
        return f"{self.coefficient}*{body}"


def sI(q=None):
    return PauliTerm("I", 0)


def sX(q):
    return PauliTerm("X", q)


def sY(q):
    return PauliTerm("Y", q)


def sZ(q):
    return PauliTerm("Z", q)
