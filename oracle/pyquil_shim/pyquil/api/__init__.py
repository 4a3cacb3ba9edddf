from pyquil import _Placeholder


class QuantumComputer(_Placeholder):
    pass


class QVM(_Placeholder):
    pass


class BenchmarkConnection(_Placeholder):
    pass


class QPUCompiler(_Placeholder):
    pass


class WavefunctionSimulator(_Placeholder):
    pass


def get_benchmarker(*a, **k):
    raise NotImplementedError("pyquil shim")
