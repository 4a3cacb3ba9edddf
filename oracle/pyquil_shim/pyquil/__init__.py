"""Inert stand-in for the `pyquil` package -- TEST INFRASTRUCTURE ONLY.

pyquil (pinned pyquil==4.5.0 in the reference's requirements-ci.txt:91) is not installable in
this image (no network).  The reference modules on the tomography hot path import it at module
top (tomography.py:10-12, utils.py:11-15, observable_estimation.py:19-23), so this stub makes
`import forest.benchmarking.tomography` work *unmodified* from /root/reference.  Only four pieces
do arithmetic and are restated from pyquil's published semantics:
  pyquil.paulis.PauliTerm, pyquil.simulation.matrices, pyquil.simulation.tools.lifted_pauli,
  pyquil.simulation.tools.lifted_state_operator.
Everything else is a placeholder that raises when used.  Nothing in the product package imports
this; it is used by oracle/make_golden.py and tests/test_oracle_vs_reference.py.
"""


class _Placeholder:
    def __init__(self, *a, **k):
        raise NotImplementedError("pyquil shim: this object is a placeholder")


class Program(_Placeholder):
    pass


def get_qc(*a, **k):
    raise NotImplementedError("pyquil shim: no QVM available")
