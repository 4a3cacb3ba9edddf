from pyquil import _Placeholder


class Delay(_Placeholder):
    pass


class Gate(_Placeholder):
    pass


class Pragma(_Placeholder):
    pass


class Measurement(_Placeholder):
    pass
