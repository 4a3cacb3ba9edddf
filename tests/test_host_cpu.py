"""CPU: host-side logic -- integer bookkeeping, library exports, error behaviour without a GPU."""
import ctypes
import itertools

import numpy as np
import pytest

from forest_benchmarking_b200 import _lib, utils
from forest_benchmarking_b200.observable_estimation import (ExperimentResult, ExperimentSetting, zeros_state,
                                                           plusX, minusZ, SIC2, TensorProductState)
from forest_benchmarking_b200.paulis import PauliTerm, sX, sZ
from oracle import ref_numpy as orc


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()  # builds with nvcc if the .so is missing
    names = _lib.declared_symbols()
    assert "qt_mle_state_batch" in names and len(names) >= 10
    for name in names:
        assert hasattr(lib, name), name
    assert lib.qt_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    plan = ctypes.c_void_p()
    idx = (ctypes.c_int32 * 1)(99)
    cf = (ctypes.c_double * 1)(1.0)
    rc = lib.qt_mle_plan_create(1, 1, idx, cf, ctypes.byref(plan))
    assert rc == -1 and "out of range" in _lib.last_error()
    assert lib.qt_mle_plan_create(9, 1, idx, cf, ctypes.byref(plan)) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from forest_benchmarking_b200 import distance_measures as dm
    with pytest.raises(_lib.QtomoError):
        dm.fidelity(np.eye(2) / 2, np.eye(2) / 2)


@pytest.mark.parametrize("n", [1, 2, 3])
def test_pauli_index_is_itertools_product_order(n):
    qubits = [5, 2, 9][:n]
    labels = ["".join(t) for t in itertools.product("IXYZ", repeat=n)]
    assert utils.pauli_labels(n) == labels == orc.pauli_labels(n)
    terms = utils.all_traceless_pauli_terms(qubits)
    assert len(terms) == 4 ** n - 1
    for k, t in enumerate(terms, start=1):
        assert utils.pauli_term_to_index(t, qubits) == k
        assert "".join(t[q] for q in qubits) == labels[k]
        assert utils.pauli_index_to_term(k, qubits) == t


def test_flatten_results():
    qubits = [1, 0]
    res = [ExperimentResult(ExperimentSetting(zeros_state(qubits), sX(1)), 0.25, 100),
           ExperimentResult(ExperimentSetting(zeros_state(qubits), PauliTerm.from_list([("Z", 0), ("Y", 1)], -0.5)),
                            -0.5, 200)]
    idx, cf, ex, cnt = utils.flatten_state_results(res, qubits)
    assert idx.tolist() == [4, 2 * 4 + 3] and cf.tolist() == [1.0, -0.5]
    assert ex.tolist() == [0.25, -0.5] and cnt.tolist() == [100, 200]
    st = plusX(0) * minusZ(1)
    assert utils.in_state_codes(st, [0, 1]) == (0, 5) and utils.in_state_codes(st, [1, 0]) == (5, 0)
    assert utils.in_state_codes(SIC2(3), [3]) == (8,)
    with pytest.raises(ValueError):
        utils.pauli_term_to_index(sZ(7), [0, 1])


def test_bootstrap_resampling_consumes_the_reference_rng_stream():
    """estimate_variance's vectorised Beta resampling == the reference's per-result scalar draws (same seed)."""
    from forest_benchmarking_b200.tomography import _resample_expectations_with_beta
    from oracle import ref_numpy as orc
    _, pidx, ex, cnt = orc.synth_state_tomography(77, 1, 2)
    np.random.seed(5)
    got = _resample_expectations_with_beta(ex[0], cnt[0], 3)
    np.random.seed(5)
    want = np.stack([orc.resample_expectations_with_beta(ex[0], cnt[0]) for _ in range(3)])
    assert np.array_equal(got, want)


def test_basis_change_matrices_match_the_oracle():
    """pauli2computational_basis_matrix / computational2pauli_basis_matrix (superoperator_transformations.py:374-438):
    host-side constants, identical to the oracle's (which is pinned against the reference), mutually inverse."""
    from oracle import ref_numpy as orc
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    for d in (2, 4, 8):
        p2c = st.pauli2computational_basis_matrix(d)
        assert p2c.dtype == np.complex128 and np.array_equal(p2c, orc.pauli2computational_basis_matrix(d))
        c2p = st.computational2pauli_basis_matrix(d)
        assert np.array_equal(c2p, orc.computational2pauli_basis_matrix(d))
        assert np.allclose(c2p @ p2c, np.eye(d * d), atol=1e-15)
    with pytest.raises(ValueError):
        st.pauli2computational_basis_matrix(3)
