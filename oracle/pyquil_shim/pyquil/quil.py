from pyquil import Program, _Placeholder


class DefGate(_Placeholder):
    pass


class Pragma(_Placeholder):
    pass


def merge_programs(*a, **k):
    raise NotImplementedError("pyquil shim")


def address_qubits(*a, **k):
    raise NotImplementedError("pyquil shim")
