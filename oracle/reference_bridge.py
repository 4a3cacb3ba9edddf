"""Import the UNMODIFIED reference (forest.benchmarking from /root/reference or $FOREST_REF_PATH)
through oracle/pyquil_shim, and build its ExperimentResult lists from plain arrays.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and tests/test_oracle_vs_reference.py).
The reference tree does not exist on the GPU box; `available()` is False there.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STATE_FACTORIES = ("plusX", "minusX", "plusY", "minusY", "plusZ", "minusZ",
                    "SIC0", "SIC1", "SIC2", "SIC3")


def ref_path():
    return os.environ.get("FOREST_REF_PATH", "/root/reference")


def available():
    return os.path.isdir(os.path.join(ref_path(), "forest", "benchmarking"))


def load():
    """Returns a namespace with the reference modules."""
    if not available():
        raise RuntimeError("reference tree not present")
    shim = os.path.join(_HERE, "pyquil_shim")
    for p in (ref_path(), shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    import forest.benchmarking.tomography as tomo
    import forest.benchmarking.distance_measures as dm
    import forest.benchmarking.operator_tools as ot
    import forest.benchmarking.observable_estimation as oe
    import forest.benchmarking.utils as ut
    from forest.benchmarking.operator_tools.project_state_matrix import project_state_matrix_to_physical
    from forest.benchmarking.operator_tools.calculational import partial_trace, sqrtm_psd
    return types.SimpleNamespace(tomo=tomo, dm=dm, ot=ot, oe=oe, ut=ut,
                                 project_state_matrix_to_physical=project_state_matrix_to_physical,
                                 partial_trace=partial_trace, sqrtm_psd=sqrtm_psd)


def pauli_term(ref, idx, qubits, coeff=1.0):
    """Canonical index -> reference PauliTerm on `qubits` (qubits[0] = most significant digit)."""
    from pyquil.paulis import PauliTerm
    n = len(qubits)
    ops = ["IXYZ"[(idx >> (2 * (n - 1 - q))) & 3] for q in range(n)]
    return PauliTerm.from_list(list(zip(ops, qubits)), coefficient=coeff)


def state_results(ref, pauli_idx, coeffs, expectations, counts, qubits):
    oe = ref.oe
    out = []
    for k, c, e, t in zip(pauli_idx, coeffs, expectations, counts):
        setting = oe.ExperimentSetting(oe.zeros_state(qubits), pauli_term(ref, int(k), qubits, c))
        out.append(oe.ExperimentResult(setting=setting, expectation=float(e), total_counts=int(t)))
    return out


def process_results(ref, settings, coeffs, expectations, counts, qubits):
    oe = ref.oe
    out = []
    for (codes, k), c, e, t in zip(settings, coeffs, expectations, counts):
        st = oe.TensorProductState()
        for code, q in zip(codes, qubits):
            st = st * getattr(oe, _STATE_FACTORIES[code])(q)
        setting = oe.ExperimentSetting(st, pauli_term(ref, int(k), qubits, c))
        out.append(oe.ExperimentResult(setting=setting, expectation=float(e), total_counts=int(t)))
    return out
