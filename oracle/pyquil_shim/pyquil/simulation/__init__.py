from pyquil import _Placeholder


class NumpyWavefunctionSimulator(_Placeholder):
    pass
