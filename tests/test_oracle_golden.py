"""CPU: the oracle (oracle/ref_numpy.py) against the golden vectors produced by the reference itself
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest

from oracle import ref_numpy as orc
from util import golden, relerr, max_relerr


@pytest.mark.parametrize("n", [1, 2, 3])
def test_algebra(n):
    g = golden(f"algebra_n{n}")
    for b in range(g["kraus"].shape[0]):
        ks = list(g["kraus"][b])
        assert relerr(orc.kraus2choi(ks), g["choi"][b]) < 1e-14
        assert relerr(orc.kraus2superop(ks), g["kraus2superop"][b]) < 1e-14
        assert np.array_equal(orc.choi2superop(g["choi"][b]), g["superop"][b])
        assert relerr(orc.superop2pauli_liouville(g["superop"][b]), g["pauli_liouville"][b]) < 1e-14
        assert relerr(orc.pauli_liouville2superop(g["pauli_liouville"][b]), g["pl2superop"][b]) < 1e-14
        x = g["noisy"][b]
        assert relerr(orc.proj_choi_to_completely_positive(x), g["proj_cp"][b]) < 1e-13
        assert relerr(orc.proj_choi_to_trace_preserving(x), g["proj_tp"][b]) < 1e-14
        assert relerr(orc.proj_choi_to_trace_non_increasing(x), g["proj_tni"][b]) < 1e-13
        assert relerr(orc.proj_choi_to_physical(x), g["proj_physical"][b]) < 1e-12
        assert relerr(orc.proj_choi_to_physical(x, False), g["proj_physical_tni"][b]) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 4])
def test_distances(n):
    g = golden(f"distances_n{n}")
    for b in range(g["rho"].shape[0]):
        assert abs(orc.fidelity(g["rho"][b], g["sigma"][b]) - g["fidelity"][b]) < 1e-13
        assert abs(orc.trace_distance(g["rho"][b], g["sigma"][b]) - g["trace_distance"][b]) < 1e-15
        assert abs(orc.purity(g["rho"][b]) - g["purity"][b]) < 1e-14


def test_known_answers():
    # reference tests/test_distance_measures.py:49-82 (values re-derived): |0> vs |1>
    z0, z1 = np.diag([1.0, 0]).astype(complex), np.diag([0, 1.0]).astype(complex)
    assert orc.fidelity(z0, z1) == 0.0 and orc.fidelity(z0, z0) == pytest.approx(1.0)
    assert orc.trace_distance(z0, z1) == 0.5      # sic: induced 1-norm
    # amplitude damping (tests/test_superoperator_transformations.py:12-40, re-derived), p = 0.1
    p = 0.1
    k0 = np.array([[1, 0], [0, np.sqrt(1 - p)]]); k1 = np.array([[0, np.sqrt(p)], [0, 0]])
    pl = np.array([[1, 0, 0, 0], [0, np.sqrt(1 - p), 0, 0], [0, 0, np.sqrt(1 - p), 0], [p, 0, 0, 1 - p]])
    assert np.allclose(orc.kraus2pauli_liouville([k0, k1]), pl, atol=1e-15)
    choi = np.array([[1, 0, 0, np.sqrt(1 - p)], [0, 0, 0, 0], [0, 0, p, 0], [np.sqrt(1 - p), 0, 0, 1 - p]])
    assert np.allclose(orc.kraus2choi([k0, k1]), choi, atol=1e-15)
    # CP projection of -Z, X (tests/test_project_superoperators.py:15-31 re-derived)
    had = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    neg = orc.kraus2choi([had]) * -1
    assert np.allclose(orc.proj_choi_to_completely_positive(neg), 0, atol=1e-14)


@pytest.mark.parametrize("name,tol", [("mle_1q", 1e-11), ("mle_2q_tol1e-4", 1e-11), ("mle_2q_maxiter200", 1e-12),
                                      ("mle_2q_hedged", 1e-10), ("mle_3q_tol1e-5", 1e-10)])
def test_mle(name, tol):
    g = golden(name)
    kw = eval(str(g["kwargs"]))
    n = int(g["n"])
    coeffs = np.ones(len(g["pauli_idx"]))
    if not kw.get("entropy_penalty") and not kw.get("beta"):
        rho, iters = orc.mle_state_estimate_batch(g["pauli_idx"], coeffs, g["expectations"], n, **kw)
        assert max_relerr(rho, g["rho_ref"]) < tol
        assert np.array_equal(iters, g["iters_ref"])
    b = 0
    rho, it = orc.mle_state_estimate(g["pauli_idx"], coeffs, g["expectations"][b], g["counts"][b], n, **kw)
    assert relerr(rho, g["rho_ref"][b]) < tol
    assert it == g["iters_ref"][b]


@pytest.mark.slow
def test_mle_2q_default_and_maxent():
    g = golden("mle_2q")
    rho, iters = orc.mle_state_estimate_batch(g["pauli_idx"], np.ones(15), g["expectations"], 2)
    assert max_relerr(rho, g["rho_ref"]) < 1e-9
    assert np.array_equal(iters, g["iters_ref"])


@pytest.mark.parametrize("name", ["pgdb_1q_pauli", "pgdb_1q_sic", "pgdb_1q_pauli_tni", "pgdb_1q_pauli_mixed",
                                  "pgdb_2q_sic"])
def test_pgdb(name):
    g = golden(name)
    n = int(g["n"])
    settings = [(tuple(int(c) for c in s), int(k)) for s, k in zip(g["state_codes"], g["pauli_idx"])]
    nb = 2 if n == 2 else g["expectations"].shape[0]
    for b in range(nb):
        est, cnt = orc.pgdb_process_estimate(settings, np.ones(len(settings)), g["expectations"][b], g["counts"][b],
                                             n, trace_preserving=bool(g["trace_preserving"]), return_counters=True)
        assert relerr(est, g["choi_ref"][b]) < 1e-9
        assert (cnt["eighs"], cnt["cost_evals"]) == tuple(g["counters_ref"][b])


def test_next_rows():
    """SURVEY 8(f) rows against outputs of the reference's own functions (oracle/make_golden.py --only next)."""
    g = golden("next_rows")
    for tag, n in (("lip_1q_pauli", 1), ("lip_1q_sic", 1), ("lip_2q_sic", 2)):
        settings = [(tuple(int(c) for c in s), int(k)) for s, k in zip(g[tag + "_codes"], g[tag + "_pidx"])]
        for b in range(3):
            got = orc.linear_inv_process_estimate(settings, np.ones(len(settings)), g[tag + "_ex"][b], n)
            assert relerr(got, g[tag + "_choi"][b]) < 1e-11
    for n in (1, 2, 3):
        for x, want in zip(g[f"unitary_n{n}_in"], g[f"unitary_n{n}_out"]):
            assert relerr(orc.proj_choi_to_unitary(x), want) < 1e-11
    for b in range(4):
        got = orc.state_log_likelihood(g["ll_rho"][b], g["ll_pidx"], np.ones(len(g["ll_pidx"])), g["ll_ex"][b], g["ll_cnt"][b], 2)
        assert abs(got - g["ll_value"][b]) < 1e-10 * abs(g["ll_value"][b])
    for i in range(6):
        idxs = [c for c in range(3) if (g["mom_masks"][i] >> c) & 1]
        for prior in (0, 1):
            got = orc.shots_to_obs_moments(g["mom_bits"][i], idxs, g["mom_coeffs"][i], bool(prior))
            assert np.allclose(got, g["mom_out"][prior, i], rtol=1e-13, atol=1e-17)
    a, va, b, vb = g["rv_in"]
    assert np.allclose(orc.ratio_variance(a, va, b, vb), g["rv_out"], rtol=1e-15)


def test_more_known_answers():
    """Hand-derivable values the reference's own tests pin, re-derived here."""
    # column-stacking vec / unvec (superoperator_transformations.py:33-79)
    a = np.array([[1, 2], [3, 4]])
    assert np.array_equal(orc.vec(a).ravel(), [1, 3, 2, 4]) and np.array_equal(orc.unvec(orc.vec(a)), a)
    # superoperator of the diagonal unitary I (x) Z is diag(conj(u) (x) u)
    iz = np.diag([1.0, -1, 1, -1])
    assert np.array_equal(np.diag(orc.kraus2superop([iz])).real, np.kron(np.diag(iz), np.diag(iz)))
    # Smolin-Gambetta-Smith example: eigenvalues (3/5, 1/2, 7/20, 1/10, -11/20) -> (9/20, 7/20, 1/5, 0, 0)
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5)))
    rho = (q * np.array([3 / 5, 1 / 2, 7 / 20, 1 / 10, -11 / 20])) @ q.conj().T
    out = orc.project_state_matrix_to_physical(rho)
    assert np.allclose(np.sort(np.linalg.eigvalsh(out))[::-1], [9 / 20, 7 / 20, 1 / 5, 0, 0], atol=1e-14)
    assert np.allclose(out, (q * np.array([9 / 20, 7 / 20, 1 / 5, 0, 0])) @ q.conj().T, atol=1e-14)
    # a CPTP map is a fixed point of the physical projection; the completely depolarising channel has purity 1/d
    u = orc.haar_unitary(rng, 4)
    choi = orc.kraus2choi([u])
    assert np.allclose(orc.proj_choi_to_physical(choi), choi, atol=1e-12)
    assert np.isclose(orc.purity(np.eye(4) / 4), 0.25)
    # shots -> moments: 3 of 4 shots give +1 -> mean 1/2, variance of the mean (1 - 1/4) / 4
    bits = np.array([[0, 0], [1, 1], [0, 0], [1, 0]])
    assert orc.shots_to_obs_moments(bits, [0, 1]) == (0.5, 0.1875)


def test_nonhermitian_physical_projection_and_distance_defaults():
    """proj_choi_to_physical on NON-Hermitian inputs: result and number of CP projections as the reference produced
    them (the anti-Hermitian part changes the trip count: see the *_calls_hermitised arrays); purity / impurity
    called without keywords (dim_renorm defaults to False, distance_measures.py:14,40)."""
    g = golden("proj_physical_nonherm")
    changed = 0
    for n in (1, 2, 3):
        xs = g[f"n{n}_in"]
        for b in range(len(xs)):
            out, calls = orc.proj_choi_to_physical(xs[b], return_count=True)
            assert relerr(out, g[f"n{n}_out"][b]) < 1e-12 and calls == g[f"n{n}_calls"][b]
            out, calls = orc.proj_choi_to_physical(xs[b], False, return_count=True)
            assert relerr(out, g[f"n{n}_out_tni"][b]) < 1e-12 and calls == g[f"n{n}_calls_tni"][b]
        changed += int(np.sum(g[f"n{n}_calls"] != g[f"n{n}_calls_hermitised"]))
    assert changed >= 4  # the fixture really exercises the difference
    for b, rho in enumerate(g["rho"]):
        assert abs(orc.purity(rho) - g["purity_default"][b]) < 1e-15
        assert abs(orc.purity(rho, dim_renorm=True) - g["purity_renorm"][b]) < 1e-15
        assert abs(orc.impurity(rho) - g["impurity_default"][b]) < 1e-15
        assert abs(orc.impurity(rho, dim_renorm=True) - g["impurity_renorm"][b]) < 1e-15
        assert abs(orc.fidelity(g["rho"][0], rho) - g["fidelity_tol1e6"][b]) < 1e-13


def test_pgdb_2q_trace_non_increasing_golden():
    g = golden("pgdb_2q_pauli_tni")
    settings = [(tuple(int(c) for c in s), int(k)) for s, k in zip(g["state_codes"], g["pauli_idx"])]
    assert not bool(g["trace_preserving"])
    est, c = orc.pgdb_process_estimate(settings, np.ones(len(settings)), g["expectations"][0], g["counts"][0], 2,
                                       trace_preserving=False, return_counters=True)
    assert relerr(est, g["choi_ref"][0]) < 1e-10
    assert (c["eighs"], c["cost_evals"]) == tuple(int(v) for v in g["counters_ref"][0])
