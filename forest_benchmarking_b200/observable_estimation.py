"""Input record types of the tomography estimators (the argument types of the drop-in boundary).

Mirrors the fields the reference estimators read from forest/benchmarking/observable_estimation.py:
``_OneQState`` (:36-74), ``TensorProductState`` (:77-128), the state factories (:131-172),
``ExperimentSetting`` (:175-213) and ``ExperimentResult`` (:694-733).  Data acquisition, grouping and
calibration (the rest of that module) drive a QVM/QPU and are out of scope.  The reference's own
objects are accepted wherever these are (duck typing).
"""
from dataclasses import dataclass
from typing import Iterable, Tuple, Union


@dataclass(frozen=True)
class _OneQState:
    label: str   # 'X', 'Y', 'Z' or 'SIC'
    index: int   # 0 = plus eigenstate, 1 = minus (X/Y/Z); 0..3 for SIC
    qubit: int

    def __str__(self):
        if self.label in ("X", "Y", "Z"):
            return f"{self.label}{'+-'[self.index]}_{self.qubit}"
        return f"{self.label}{self.index}_{self.qubit}"


class TensorProductState:
    def __init__(self, states: Iterable[_OneQState] = ()):
        self.states = tuple(states)

    def __mul__(self, other):
        return TensorProductState(self.states + other.states)

    def __getitem__(self, qubit):
        for s in self.states:
            if s.qubit == qubit:
                return s
        raise IndexError(qubit)

    def __iter__(self):
        return iter(self.states)

    def __len__(self):
        return len(self.states)

    def __eq__(self, other):
        return isinstance(other, TensorProductState) and frozenset(self.states) == frozenset(other.states)

    def __hash__(self):
        return hash(frozenset(self.states))

    def __repr__(self):
        return "TensorProductState[" + " * ".join(str(s) for s in self.states) + "]"


def _factory(label, index):
    def make(q):
        return TensorProductState((_OneQState(label, index, q),))
    return make


plusX, minusX = _factory("X", 0), _factory("X", 1)
plusY, minusY = _factory("Y", 0), _factory("Y", 1)
plusZ, minusZ = _factory("Z", 0), _factory("Z", 1)
SIC0, SIC1, SIC2, SIC3 = (_factory("SIC", i) for i in range(4))


def zeros_state(qubits: Iterable[int]):
    return TensorProductState(_OneQState("Z", 0, q) for q in qubits)


@dataclass(frozen=True)
class ExperimentSetting:
    in_state: TensorProductState
    observable: object  # PauliTerm (ours or pyquil's)


@dataclass(frozen=True)
class ExperimentResult:
    setting: ExperimentSetting
    expectation: Union[float, complex]
    total_counts: int
    std_err: Union[float, complex] = None
    raw_expectation: Union[float, complex] = None
    raw_std_err: float = None
    calibration_expectation: Union[float, complex] = None
    calibration_std_err: Union[float, complex] = None
    calibration_counts: int = None
