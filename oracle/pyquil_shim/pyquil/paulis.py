"""Minimal PauliTerm with the semantics the reference reads on the tomography path
(term[q] -> 'I'/'X'/'Y'/'Z', .coefficient, iteration over (qubit, op), from_list)."""
from collections import OrderedDict

_MUL = {  # (a, b) -> (phase, op)
    ("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"),
    ("Y", "Z"): (1j, "X"), ("Z", "Y"): (-1j, "X"),
    ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y"),
}


class PauliTerm:
    def __init__(self, op, index, coefficient=1.0):
        if op not in "IXYZ":
            raise ValueError(f"{op} is not a Pauli operator")
        self._ops = OrderedDict()
        if op != "I":
            self._ops[index] = op
        self.coefficient = complex(coefficient)

    @classmethod
    def from_list(cls, terms_list, coefficient=1.0):
        t = cls("I", 0, coefficient)
        seen = set()
        for op, q in terms_list:
            if q in seen:
                raise ValueError("from_list: duplicate qubit")
            seen.add(q)
            if op not in "IXYZ":
                raise ValueError(f"{op} is not a Pauli operator")
            if op != "I":
                t._ops[q] = op
        return t

    def copy(self):
        t = PauliTerm("I", 0, self.coefficient)
        t._ops = OrderedDict(self._ops)
        return t

    def __getitem__(self, q):
        return self._ops.get(q, "I")

    def __iter__(self):
        return iter(self._ops.items())

    def __len__(self):
        return len(self._ops)

    def get_qubits(self):
        return list(self._ops.keys())

    def operations_as_set(self):
        return frozenset(self._ops.items())

    def id(self, sort_ops=True):
        items = sorted(self._ops.items()) if sort_ops else self._ops.items()
        return "".join(f"{op}{q}" for q, op in items)

    def __hash__(self):
        return hash((round(self.coefficient.real, 12), round(self.coefficient.imag, 12),
                     self.operations_as_set()))

    def __eq__(self, other):
        if not isinstance(other, PauliTerm):
            return NotImplemented
        return (self.operations_as_set() == other.operations_as_set()
                and abs(self.coefficient - other.coefficient) < 1e-12)

    def __mul__(self, other):
        if isinstance(other, (int, float, complex)):
            t = self.copy()
            t.coefficient *= other
            return t
        t = self.copy()
        t.coefficient *= other.coefficient
        for q, op in other:
            a = t._ops.get(q, "I")
            if a == "I":
                t._ops[q] = op
            elif a == op:
                del t._ops[q]
            else:
                ph, c = _MUL[(a, op)]
                t.coefficient *= ph
                t._ops[q] = c
        return t

    __rmul__ = __mul__

    def compact_str(self):
        return f"{self.coefficient}*{self.id(sort_ops=False)}"

    def __str__(self):
        body = "*".join(f"{op}{q}" for q, op in self._ops.items()) or "I"
        return f"{self.coefficient}*{body}"

    __repr__ = __str__


def sI(q=None):
    return PauliTerm("I", 0)


def sX(q):
    return PauliTerm("X", q)


def sY(q):
    return PauliTerm("Y", q)


def sZ(q):
    return PauliTerm("Z", q)


def is_identity(term):
    return len(term) == 0
