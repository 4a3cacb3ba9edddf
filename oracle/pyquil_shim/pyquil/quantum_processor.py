from pyquil import _Placeholder


class NxQuantumProcessor(_Placeholder):
    pass
