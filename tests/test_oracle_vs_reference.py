"""Pins oracle/ref_numpy.py against the UNMODIFIED reference imported from /root/reference through
oracle/pyquil_shim.  Skips when the reference tree is absent (GPU box); there tests/golden pins it."""
import numpy as np
import pytest

from oracle import ref_numpy as orc
from oracle import reference_bridge as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return rb.load()


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize("n", [1, 2, 3])
def test_pauli_order_bit_exact(ref, n):
    qubits = list(range(n))
    basis = ref.ut.n_qubit_pauli_basis(n)
    assert basis.labels == orc.pauli_labels(n)
    terms = ref.ut.all_traceless_pauli_terms(qubits)
    from pyquil.simulation.tools import lifted_pauli
    for k in range(4 ** n):
        assert np.array_equal(basis.ops[k], orc.pauli_matrix(k, n))
        if k:
            assert np.array_equal(lifted_pauli(terms[k - 1], qubits[::-1]), orc.pauli_matrix(k, n))
            assert terms[k - 1] == rb.pauli_term(ref, k, qubits)


def test_input_state_order(ref):
    import itertools
    from pyquil.simulation.tools import lifted_state_operator
    qubits = [0, 1]
    for basis, fn in (("pauli", ref.tomo._pauli_process_tomo_settings), ("sic", ref.tomo._sic_process_tomo_settings)):
        ours = orc.process_tomography_settings(2, basis)
        theirs = list(fn(qubits))
        assert len(ours) == len(theirs)
        for (codes, k), s in zip(ours, theirs):
            assert s.observable == rb.pauli_term(ref, k, qubits)
            assert np.allclose(lifted_state_operator(s.in_state, qubits[::-1]),
                               orc.product_state_matrix(codes), atol=1e-15)


@pytest.mark.parametrize("n", [1, 2, 3])
def test_conversions(ref, n):
    rng = np.random.default_rng(10 + n)
    d = 2 ** n
    ks = [np.sqrt(.7) * orc.haar_unitary(rng, d), np.sqrt(.3) * orc.haar_unitary(rng, d)]
    ot = ref.ot
    assert relerr(orc.kraus2choi(ks), ot.kraus2choi(ks)) < 1e-14
    assert relerr(orc.kraus2superop(ks), ot.kraus2superop(ks)) < 1e-14
    c = ot.kraus2choi(ks)
    s = ot.choi2superop(c)
    assert np.array_equal(orc.choi2superop(c), s)
    assert np.array_equal(orc.superop2choi(s), ot.superop2choi(s))
    assert np.array_equal(orc.pauli2computational_basis_matrix(d), ot.pauli2computational_basis_matrix(d))
    pl = ot.superop2pauli_liouville(s)
    assert relerr(orc.superop2pauli_liouville(s), pl) < 1e-14
    assert relerr(orc.pauli_liouville2superop(pl), ot.pauli_liouville2superop(pl)) < 1e-14
    assert relerr(orc.choi2pauli_liouville(c), ot.choi2pauli_liouville(c)) < 1e-14
    assert relerr(orc.pauli_liouville2choi(pl), ot.pauli_liouville2choi(pl)) < 1e-14
    assert relerr(orc.kraus2pauli_liouville(ks), ot.kraus2pauli_liouville(ks)) < 1e-14
    k1, k2 = orc.choi2kraus(c), ot.choi2kraus(c)
    assert len(k1) == len(k2) == 2
    assert relerr(orc.kraus2choi(k1), ot.kraus2choi(k2)) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3])
def test_projections(ref, n):
    rng = np.random.default_rng(20 + n)
    d2 = 4 ** n
    x = rng.standard_normal((d2, d2)) + 1j * rng.standard_normal((d2, d2))
    x = (x + x.conj().T) / 2 / d2
    x += orc.kraus2choi([orc.haar_unitary(rng, 2 ** n)])
    ot = ref.ot
    assert relerr(orc.proj_choi_to_completely_positive(x), ot.proj_choi_to_completely_positive(x)) < 1e-13
    assert relerr(orc.proj_choi_to_trace_preserving(x), ot.proj_choi_to_trace_preserving(x)) < 1e-14
    assert relerr(orc.proj_choi_to_trace_non_increasing(x), ot.proj_choi_to_trace_non_increasing(x)) < 1e-13
    assert relerr(orc.proj_choi_to_physical(x), ot.proj_choi_to_physical(x)) < 1e-12
    assert relerr(orc.proj_choi_to_physical(x, False), ot.proj_choi_to_physical(x, False)) < 1e-12
    assert relerr(orc.partial_trace_out(x), ref.partial_trace(x, [0], [2 ** n, 2 ** n])) < 1e-15


@pytest.mark.parametrize("n", [1, 2, 4])
def test_distances(ref, n):
    rng = np.random.default_rng(30 + n)
    d = 2 ** n
    for _ in range(5):
        r, s = orc.ginibre_state(rng, d), orc.ginibre_state(rng, d)
        assert abs(orc.fidelity(r, s) - ref.dm.fidelity(r, s)) < 1e-13
        assert abs(orc.trace_distance(r, s) - ref.dm.trace_distance(r, s)) < 1e-15
        assert abs(orc.purity(r) - ref.dm.purity(r, dim_renorm=False)) < 1e-14
    z0 = np.diag([1.0, 0]); z1 = np.diag([0, 1.0])
    assert orc.trace_distance(z0, z1) == ref.dm.trace_distance(z0, z1) == 0.5


@pytest.mark.parametrize("n,kw", [
    (1, {}), (2, dict(tol=1e-5)), (2, dict(maxiter=50)),
    (2, dict(entropy_penalty=.001, tol=1e-4)), (2, dict(epsilon=1e-4, beta=.5, tol=1e-3)),
])
def test_mle(ref, n, kw):
    _, pidx, ex, cnt = orc.synth_state_tomography(77 + n, 2, n)
    qubits = list(range(n))
    coeffs = np.ones(len(pidx))
    for b in range(2):
        res = rb.state_results(ref, pidx, coeffs, ex[b], cnt[b], qubits)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = ref.tomo.iterative_mle_state_estimate(res, qubits, **kw)
        got, _ = orc.mle_state_estimate(pidx, coeffs, ex[b], cnt[b], n, **kw)
        assert relerr(got, want) < 1e-11
        if not kw.get("entropy_penalty") and not kw.get("beta"):
            kb = {k: v for k, v in kw.items() if k in ("epsilon", "tol", "maxiter")}
            gotb, _ = orc.mle_state_estimate_batch(pidx, coeffs, ex[b:b + 1], n, **kb)
            assert relerr(gotb[0], want) < 1e-11
        assert relerr(orc.linear_inv_state_estimate(pidx, coeffs, ex[b], n),
                      ref.tomo.linear_inv_state_estimate(res, qubits)) < 1e-12
        assert abs(orc.state_log_likelihood(want, pidx, coeffs, ex[b], cnt[b], n)
                   - ref.tomo.state_log_likelihood(want, res, qubits)) < 1e-9 * abs(
            ref.tomo.state_log_likelihood(want, res, qubits))


def test_r_operator_with_coefficients_and_identity(ref):
    rng = np.random.default_rng(5)
    rho = orc.ginibre_state(rng, 4)
    qubits = [3, 7]
    pidx, coeffs, ex = [0, 7, 7, 12], [1.0, -1.0, 0.5, 1.0], [1.0, -.2, .1, .4]
    res = rb.state_results(ref, pidx, coeffs, ex, [10] * 4, qubits)
    want = ref.tomo._R(rho, res, qubits[::-1])
    got = orc.r_operator(rho, [c * orc.pauli_matrix(k, 2) for k, c in zip(pidx, coeffs)], ex)
    assert relerr(got, want) < 1e-14


@pytest.mark.parametrize("n,basis,tp", [(1, "pauli", True), (1, "sic", True), (1, "pauli", False), (2, "sic", True)])
def test_pgdb(ref, n, basis, tp):
    _, settings, ex, cnt = orc.synth_process_tomography(91 + n, 1, n, in_basis=basis)
    qubits = list(range(n))
    coeffs = np.ones(len(settings))
    res = rb.process_results(ref, settings, coeffs, ex[0], cnt[0], qubits)
    a_ref, n_ref = ref.tomo._extract_from_results(res, qubits[::-1])
    a, nn = orc.extract_design(settings, coeffs, ex[0], cnt[0], n)
    assert relerr(a, a_ref) < 1e-15 and relerr(nn, n_ref) < 1e-15
    want = ref.tomo.pgdb_process_estimate(res, qubits, trace_preserving=tp)
    got = orc.pgdb_process_estimate(settings, coeffs, ex[0], cnt[0], n, trace_preserving=tp)
    assert relerr(got, want) < 1e-10
    assert relerr(orc.linear_inv_process_estimate(settings, coeffs, ex[0], n),
                  ref.tomo.linear_inv_process_estimate(res, qubits)) < 1e-11


def test_project_state_matrix_to_physical(ref):
    from forest.benchmarking.operator_tools.project_state_matrix import project_state_matrix_to_physical as ref_proj
    rng = np.random.default_rng(11)
    # the reference's own known answer (tests/test_project_state_matrix.py:12-14)
    eigs = np.diag(np.array(list(reversed([3.0 / 5, 1.0 / 2, 7.0 / 20, 1.0 / 10, -11.0 / 20]))))
    phys = orc.project_state_matrix_to_physical(eigs)
    assert np.allclose(phys, np.diag([0, 0, 1.0 / 5, 7.0 / 20, 9.0 / 20]))
    for d in (2, 4, 8, 16):
        for _ in range(6):
            g = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
            h = (g + g.conj().T) / 2
            h = h / np.trace(h).real + 0.05 * np.eye(d) * rng.standard_normal()
            assert relerr(orc.project_state_matrix_to_physical(h), ref_proj(h)) < 1e-12
        rho = orc.ginibre_state(rng, d)
        assert relerr(orc.project_state_matrix_to_physical(3.0 * rho), ref_proj(3.0 * rho)) < 1e-13


def test_resample_matches_reference_stream(ref):
    from forest.benchmarking.tomography import _resample_expectations_with_beta
    _, pidx, ex, cnt = orc.synth_state_tomography(77, 1, 2)
    results = rb.state_results(ref, pidx, np.ones(15), ex[0], cnt[0], [0, 1])
    np.random.seed(5)
    want = [r.expectation for r in _resample_expectations_with_beta(results)]
    np.random.seed(5)
    got = orc.resample_expectations_with_beta(ex[0], cnt[0])
    assert np.array_equal(np.array(want), got)


@pytest.mark.parametrize("n", [1, 2])
def test_chi_family(ref, n):
    ot = ref.ot
    rng = np.random.default_rng(40 + n)
    d = 2 ** n
    ks = [np.sqrt(.6) * orc.haar_unitary(rng, d), np.sqrt(.4) * orc.haar_unitary(rng, d)]
    chi = orc.kraus2chi(ks)
    assert relerr(chi, ot.kraus2chi(ks)) < 1e-13
    assert relerr(orc.chi2choi(chi), ot.chi2choi(chi)) < 1e-13
    assert relerr(orc.chi2pauli_liouville(chi), ot.chi2pauli_liouville(chi)) < 1e-13
    assert relerr(orc.chi2superop(chi), ot.chi2superop(chi)) < 1e-13
    choi = orc.kraus2choi(ks)
    assert relerr(orc.choi2chi(choi), ot.choi2chi(choi)) < 1e-12
    assert relerr(orc.superop2chi(orc.reshuffle(choi)), ot.superop2chi(orc.reshuffle(choi))) < 1e-12
    pl = orc.choi2pauli_liouville(choi)
    assert relerr(orc.pauli_liouville2chi(pl), ot.pauli_liouville2chi(pl)) < 1e-12


def test_process_fidelities(ref):
    rng = np.random.default_rng(8)
    for n in (1, 2):
        d = 2 ** n
        u, v = orc.haar_unitary(rng, d), orc.haar_unitary(rng, d)
        p0, p1 = orc.kraus2pauli_liouville([u]), orc.kraus2pauli_liouville([np.sqrt(.9) * u, np.sqrt(.1) * v])
        assert abs(orc.hilbert_schmidt_ip(p0, p1) - ref.dm.hilbert_schmidt_ip(p0, p1)) < 1e-12
        assert abs(orc.entanglement_fidelity(p0, p1) - ref.dm.entanglement_fidelity(p0, p1)) < 1e-13
        assert abs(orc.process_fidelity(p0, p1) - ref.dm.process_fidelity(p0, p1)) < 1e-13


@pytest.mark.parametrize("n", [1, 2, 3])
def test_proj_choi_to_unitary(ref, n):
    from forest.benchmarking.operator_tools.project_superoperators import proj_choi_to_unitary
    rng = np.random.default_rng(40 + n)
    d = 2 ** n
    u = orc.haar_unitary(rng, d)
    g = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
    choi = .9 * orc.kraus2choi(u) + .1 * (g @ g.conj().T) / d ** 2 + .01 * g  # noisy, not Hermitian
    assert relerr(orc.proj_choi_to_unitary(choi), proj_choi_to_unitary(choi)) < 1e-12


def test_shots_to_obs_moments(ref):
    from forest.benchmarking.observable_estimation import shots_to_obs_moments, ratio_variance
    from pyquil.paulis import PauliTerm
    rng = np.random.default_rng(9)
    qubits = [4, 7, 9]
    bits = (rng.random((257, 3)) < [.2, .5, .8]).astype(np.uint8)
    for ops, idxs, c in [([("Z", 4)], [0], 1.0), ([("X", 7), ("Z", 9)], [1, 2], -0.5), ([("Y", 4), ("Y", 7), ("X", 9)], [0, 1, 2], 2.0),
                         ([], [], 0.7)]:
        term = PauliTerm.from_list(ops, coefficient=c) if ops else PauliTerm("I", 0, c)
        for prior in (False, True):
            want = shots_to_obs_moments(bits.astype(np.int64), qubits, term, prior)  # qc.run returns int64 bits
            got = orc.shots_to_obs_moments(bits, idxs, c, prior)
            assert np.allclose(got, want, rtol=1e-13, atol=1e-16)
    assert np.isclose(orc.ratio_variance(.3, .01, .9, .002), ratio_variance(.3, .01, .9, .002), rtol=1e-15)
