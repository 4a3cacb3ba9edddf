"""GPU parity: batched MLE state tomography and distance measures vs the oracle / reference goldens.
Tolerance (BASELINE.json north_star): <= 1e-6 relative Frobenius error; integer bookkeeping bit-exact."""
import numpy as np
import pytest

from oracle import ref_numpy as orc
from util import golden, relerr, max_relerr

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _run_mle(torch, n, pidx, ex, cnt=None, coeffs=None, kernel=0, **kw):
    from forest_benchmarking_b200 import tomography as tm
    plan = tm.MlePlan(n, pidx, coeffs)
    e = torch.from_numpy(np.ascontiguousarray(ex)).cuda()
    c = None if cnt is None else torch.from_numpy(np.ascontiguousarray(cnt)).cuda()
    rho, iters = tm.iterative_mle_state_estimate_batch(plan, e, c, kernel=kernel, **kw)
    torch.cuda.synchronize()
    return rho.cpu().numpy(), iters.cpu().numpy()


@pytest.mark.parametrize("name", ["mle_1q", "mle_2q", "mle_2q_tol1e-4", "mle_2q_maxiter200", "mle_3q_tol1e-5",
                                  "mle_2q_maxent", "mle_2q_hedged"])
@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_mle_golden(torch, name, kernel):
    g = golden(name)
    kw = eval(str(g["kwargs"]))
    n = int(g["n"])
    variants = bool(kw.get("entropy_penalty")) or bool(kw.get("beta"))
    if kernel == 1 and (n > 2 or variants):
        pytest.skip("register kernel: n<=2 vanilla only")
    if kernel == 3 and (n != 2 or variants):
        pytest.skip("quad kernel: n==2 vanilla only")
    rho, iters = _run_mle(torch, n, g["pauli_idx"], g["expectations"], g["counts"], kernel=kernel, **kw)
    assert max_relerr(rho, g["rho_ref"]) < TOL
    # iteration counters are part of the contract (identical stopping rule): exact on the goldens
    print(name, "kernel", kernel, "iters", iters.tolist(), "ref", g["iters_ref"].tolist())
    assert np.array_equal(iters, g["iters_ref"]), (iters, g["iters_ref"])


@pytest.mark.parametrize("n,kernel", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2)])
def test_mle_vs_oracle_batch(torch, n, kernel):
    _, pidx, ex, cnt = orc.synth_state_tomography(500 + n, 40, n)
    kw = dict(tol=1e-6, maxiter=3000)
    rho, iters = _run_mle(torch, n, pidx, ex, cnt, kernel=kernel, **kw)
    want, witers = orc.mle_state_estimate_batch(pidx, np.ones(len(pidx)), ex, n, **kw)
    assert max_relerr(rho, want) < TOL
    print("mle_vs_oracle", n, kernel, "mismatches", np.nonzero(iters != witers)[0].tolist())
    assert np.all(np.abs(iters - witers) <= 1) and np.mean(iters != witers) <= 0.05
    # size-independent properties: Hermitian, unit trace, positive
    assert np.allclose(rho, rho.conj().transpose(0, 2, 1), atol=1e-13)
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1, atol=1e-12)
    assert np.linalg.eigvalsh(rho).min() > -1e-12


def test_mle_quad_kernel_ragged_batch_and_full_size(torch):
    """quad kernel (4 lanes per experiment): batch sizes that do not fill a warp; full BASELINE batch vs the
    register kernel (independent code path) and the size-independent properties."""
    for batch in (1, 5, 13):
        _, pidx, ex, cnt = orc.synth_state_tomography(700 + batch, batch, 2)
        kw = dict(tol=1e-7, maxiter=1500)
        rho, iters = _run_mle(torch, 2, pidx, ex, kernel=3, **kw)
        want, witers = orc.mle_state_estimate_batch(pidx, np.ones(15), ex, 2, **kw)
        assert max_relerr(rho, want) < TOL and np.all(np.abs(iters - witers) <= 2)
    from forest_benchmarking_b200 import synthetic as sy
    pidx, ex, _, _ = sy.state_tomography_batch(2002, 4096, 2)
    rho3, it3 = _run_mle(torch, 2, pidx, ex, kernel=3)
    rho1, it1 = _run_mle(torch, 2, pidx, ex, kernel=1)
    assert max_relerr(rho3, rho1) < TOL
    assert np.mean(it3 != it1) < 0.02 and np.all(np.abs(it3 - it1) <= 3)
    assert np.allclose(rho3, rho3.conj().transpose(0, 2, 1), atol=1e-13)
    assert np.allclose(np.trace(rho3, axis1=1, axis2=2), 1, atol=1e-12)
    assert np.linalg.eigvalsh(rho3).min() > -1e-12


def test_mle_4q_and_5q_warp_kernel(torch):
    for n, maxiter in ((4, 60), (5, 12)):
        _, pidx, ex, cnt = orc.synth_state_tomography(600 + n, 3, n)
        rho, iters = _run_mle(torch, n, pidx, ex, cnt, kernel=2, maxiter=maxiter)
        want, witers = orc.mle_state_estimate_batch(pidx, np.ones(len(pidx)), ex, n, maxiter=maxiter)
        assert max_relerr(rho, want) < TOL
        assert np.array_equal(iters, witers)


def test_mle_general_observable_lists(torch):
    """identity observable, duplicates, non-unit coefficients, incomplete sets (test_state_tomography.py:60-115)."""
    rng = np.random.default_rng(3)
    pidx = np.array([0, 7, 7, 12, 3, 1, 9], dtype=np.int32)
    coeffs = np.array([1.0, -1.0, 0.5, 1.0, 1.0, 2.0, 1.0])
    ex = rng.uniform(-.6, .6, size=(5, len(pidx)))
    ex[:, 0] = 1.0
    kw = dict(tol=1e-7, maxiter=500)
    rho, iters = _run_mle(torch, 2, pidx, ex, np.full_like(ex, 50.), coeffs=coeffs, **kw)
    for b in range(5):
        want, it = orc.mle_state_estimate(pidx, coeffs, ex[b], np.full(len(pidx), 50.), 2, **kw)
        assert relerr(rho[b], want) < TOL and abs(int(iters[b]) - it) <= 1
    # unit coefficients with duplicates + identity -> register kernel path
    coeffs1 = np.ones(len(pidx))
    rho1, _ = _run_mle(torch, 2, pidx, ex, None, coeffs=coeffs1, kernel=1, **kw)
    rho2, _ = _run_mle(torch, 2, pidx, ex, None, coeffs=coeffs1, kernel=2, **kw)
    rho3, _ = _run_mle(torch, 2, pidx, ex, None, coeffs=coeffs1, kernel=3, **kw)
    for b in range(5):
        want, _ = orc.mle_state_estimate(pidx, coeffs1, ex[b], np.full(len(pidx), 50.), 2, **kw)
        assert relerr(rho1[b], want) < TOL and relerr(rho2[b], want) < TOL and relerr(rho3[b], want) < TOL


def test_mle_dropin_signature_and_errors(torch):
    from forest_benchmarking_b200 import tomography as tm
    from forest_benchmarking_b200.observable_estimation import ExperimentResult, ExperimentSetting, zeros_state
    from forest_benchmarking_b200.utils import all_traceless_pauli_terms
    qubits = [4, 2]
    _, pidx, ex, cnt = orc.synth_state_tomography(9, 1, 2)
    res = [ExperimentResult(ExperimentSetting(zeros_state(qubits), t), e, int(c))
           for t, e, c in zip(all_traceless_pauli_terms(qubits), ex[0], cnt[0])]
    rho = tm.iterative_mle_state_estimate(res, qubits, tol=1e-6)
    want, _ = orc.mle_state_estimate(pidx, np.ones(15), ex[0], cnt[0], 2, tol=1e-6)
    assert relerr(rho, want) < TOL
    with pytest.raises(ValueError):
        tm.iterative_mle_state_estimate(res, qubits, entropy_penalty=.1, beta=.1)
    with pytest.warns(UserWarning):
        tm.iterative_mle_state_estimate(res, qubits, maxiter=5)
    assert tm.iterative_mle_state_estimate_batch(tm.MlePlan(2, pidx), torch.empty((0, 15), dtype=torch.float64,
                                                                                 device="cuda"))[0].shape[0] == 0


def test_mle_single_step_kernel(torch):
    from forest_benchmarking_b200 import tomography as tm
    for n in (1, 2):
        truth, pidx, ex, _ = orc.synth_state_tomography(40 + n, 300, n)
        rho0 = np.stack([orc.ginibre_state(np.random.default_rng(b), 2 ** n) for b in range(300)])
        out = tm.mle_step_batch(n, torch.from_numpy(np.ascontiguousarray(ex.T)).cuda(),
                                torch.from_numpy(rho0).cuda(), epsilon=.1).cpu().numpy()
        ops = [orc.pauli_matrix(int(k), n) for k in pidx]
        for b in range(0, 300, 37):
            m = np.eye(2 ** n) + .1 * (orc.r_operator(rho0[b], ops, ex[b]) - np.eye(2 ** n))
            want = m @ rho0[b] @ m
            assert relerr(out[b], want / np.trace(want)) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 4])
def test_distances_golden(torch, n):
    from forest_benchmarking_b200 import distance_measures as dm
    g = golden(f"distances_n{n}")
    rho, sigma = torch.from_numpy(g["rho"]).cuda(), torch.from_numpy(g["sigma"]).cuda()
    fid = dm.fidelity_batch(rho, sigma).cpu().numpy()
    td = dm.trace_distance_batch(rho, sigma).cpu().numpy()
    pur = dm.purity_batch(rho).cpu().numpy()
    assert np.max(np.abs(fid - g["fidelity"]) / np.maximum(np.abs(g["fidelity"]), 1e-12)) < TOL
    assert np.max(np.abs(td - g["trace_distance"])) < 1e-14
    assert np.max(np.abs(pur - g["purity"])) < 1e-13
    nuc = dm.trace_distance_nuclear_batch(rho, sigma).cpu().numpy()
    want = [0.5 * np.abs(np.linalg.eigvalsh(r - s)).sum() for r, s in zip(g["rho"], g["sigma"])]
    assert np.max(np.abs(nuc - want)) < 1e-12


@pytest.mark.parametrize("n", [3, 5])
def test_distances_vs_oracle(torch, n):
    from forest_benchmarking_b200 import distance_measures as dm
    rng = np.random.default_rng(n)
    d = 2 ** n
    rho = np.stack([orc.ginibre_state(rng, d) for _ in range(50)])
    sigma = np.stack([orc.ginibre_state(rng, d, rank=(1 if b % 5 == 0 else None)) for b in range(50)])
    fid = dm.fidelity_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sigma).cuda()).cpu().numpy()
    td = dm.trace_distance_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sigma).cuda()).cpu().numpy()
    for b in range(50):
        # rank-deficient sigma: d-1 eigenvalues of sqrt(rho) sigma sqrt(rho) are zero up to rounding and
        # sqrt() amplifies that noise to ~1e-8 in LAPACK and Jacobi alike -- north-star tolerance applies
        tol = 1e-6 * max(fid[b], 1e-3) if b % 5 == 0 else 1e-10
        assert abs(fid[b] - orc.fidelity(rho[b], sigma[b])) < tol
        assert abs(td[b] - orc.trace_distance(rho[b], sigma[b])) < 1e-14


def test_distance_known_answers_and_dropin(torch):
    from forest_benchmarking_b200 import distance_measures as dm
    z0, z1 = np.diag([1.0, 0]), np.diag([0, 1.0])
    assert dm.trace_distance(z0, z1) == 0.5          # reference tests/test_distance_measures.py:73-82 (sic)
    assert abs(dm.fidelity(z0, z1)) < 1e-15 and abs(dm.fidelity(z0, z0) - 1) < 1e-14
    assert abs(dm.purity(np.eye(2) / 2, dim_renorm=False) - 0.5) < 1e-15
    assert abs(dm.infidelity(z0, z0)) < 1e-14


def test_distance_defaults_match_reference(torch):
    """purity / impurity called WITHOUT keywords (reference default dim_renorm=False, distance_measures.py:14,40) and
    the fidelity return idiom (a Python float through np.real_if_close(.., tol), :82-84); goldens from the reference
    (oracle/make_golden.py --only nonherm)."""
    from forest_benchmarking_b200 import distance_measures as dm
    g = golden("proj_physical_nonherm")
    assert abs(dm.purity(np.eye(2) / 2) - 0.5) < 1e-15 and abs(dm.impurity(np.eye(2) / 2) - 0.5) < 1e-15
    assert abs(dm.purity(np.eye(2) / 2, dim_renorm=True)) < 1e-15
    for b, rho in enumerate(g["rho"]):
        assert abs(dm.purity(rho) - g["purity_default"][b]) < 1e-14
        assert abs(dm.purity(rho, dim_renorm=True) - g["purity_renorm"][b]) < 1e-14
        assert abs(dm.impurity(rho) - g["impurity_default"][b]) < 1e-14
        assert abs(dm.impurity(rho, True) - g["impurity_renorm"][b]) < 1e-14
        f = dm.fidelity(g["rho"][0], rho, tol=1e6)
        assert type(f) is float and abs(f - g["fidelity_tol1e6"][b]) < 1e-12
    assert type(dm.fidelity(g["rho"][0], g["rho"][1])) is float
    assert type(dm.purity(g["rho"][0])) is float
    # a non-Hermitian argument: tr(rho rho) is complex and the reference returns it as such
    x = g["rho"][0] + 0.1j * np.triu(np.ones((4, 4)), 1)
    want = np.trace(x @ x)
    got = dm.purity(x)
    assert isinstance(got, complex) and abs(got - want) < 1e-14


def test_wrappers_reject_bad_tensors(torch):
    """Every raw pointer handed to a kernel is checked first: dtype, device, exact shape, contiguity."""
    from forest_benchmarking_b200 import tomography as tm
    _, pidx, ex, cnt = orc.synth_state_tomography(77, 6, 2)
    plan = tm.MlePlan(2, pidx)
    e = torch.from_numpy(ex).cuda()
    rho = torch.eye(4, dtype=torch.complex128, device="cuda").repeat(6, 1, 1) / 4
    ec = torch.from_numpy(np.ascontiguousarray(ex.T)).cuda()
    tm.mle_step_batch(2, ec, rho)
    for bad in (lambda: tm.mle_step_batch(2, e, rho),                      # [B, K] instead of [K, B]
                lambda: tm.mle_step_batch(2, ec.float(), rho),              # dtype
                lambda: tm.mle_step_batch(2, ec, rho.to(torch.complex64)),
                lambda: tm.mle_step_batch(1, ec, rho),                      # n_qubits vs rho
                lambda: tm.mle_step_batch(2, ec, rho, out=rho[:3]),         # short out buffer
                lambda: tm.mle_step_batch(2, ec.cpu(), rho),                # host tensor
                lambda: tm.iterative_mle_state_estimate_batch(plan, e, out=torch.empty((5, 4, 4), dtype=torch.complex128, device="cuda")),
                lambda: tm.iterative_mle_state_estimate_batch(plan, e, iters_out=torch.empty((6,), dtype=torch.int64, device="cuda")),
                lambda: tm.iterative_mle_state_estimate_batch(plan, e, torch.from_numpy(cnt[:3]).cuda()),
                lambda: tm.iterative_mle_state_estimate_batch(plan, e, None, beta=0.5),
                lambda: tm.iterative_mle_state_estimate_batch(plan, e, out=torch.empty((6, 4, 8), dtype=torch.complex128, device="cuda")[:, :, ::2])):
        with pytest.raises(ValueError):
            bad()


def test_tensor_on_another_device_than_current(torch):
    """plans, streams and outputs must follow the operands' device, not whatever device is current."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from forest_benchmarking_b200 import tomography as tm, distance_measures as dm
    _, pidx, ex, cnt = orc.synth_state_tomography(78, 16, 2)
    with torch.cuda.device(1):
        plan1 = tm.MlePlan(2, pidx)
    e1 = torch.from_numpy(ex).to("cuda:1")
    torch.cuda.set_device(0)
    rho1, it1 = tm.iterative_mle_state_estimate_batch(plan1, e1, tol=1e-6, maxiter=2000)  # current device is 0
    assert rho1.device == e1.device
    plan0 = tm.MlePlan(2, pidx)
    rho0, it0 = tm.iterative_mle_state_estimate_batch(plan0, e1.to("cuda:0"), tol=1e-6, maxiter=2000)
    assert torch.equal(rho1.cpu(), rho0.cpu()) and torch.equal(it1.cpu(), it0.cpu())
    with pytest.raises(ValueError):
        tm.iterative_mle_state_estimate_batch(plan0, e1)  # plan on cuda:0, data on cuda:1
    f = dm.fidelity_batch(rho1, rho1)
    assert f.device == rho1.device and float((f - 1).abs().max()) < 1e-9


def test_linear_inv_state_estimate(torch):
    """a4 / BASELINE configs[0]: linear inversion vs the oracle's pinv, complete and general observable lists."""
    from forest_benchmarking_b200 import tomography as tm
    from forest_benchmarking_b200.observable_estimation import ExperimentResult, ExperimentSetting, zeros_state
    from forest_benchmarking_b200.utils import all_traceless_pauli_terms
    for n in (1, 2, 3):
        _, pidx, ex, cnt = orc.synth_state_tomography(1001 + n, 33, n)
        plan = tm.MlePlan(n, pidx)
        rho = tm.linear_inv_state_estimate_batch(plan, torch.from_numpy(ex).cuda()).cpu().numpy()
        want = np.stack([orc.linear_inv_state_estimate(pidx, np.ones(len(pidx)), ex[b], n) for b in range(33)])
        assert max_relerr(rho, want) < 1e-12
    # duplicates, identity, missing terms, non-unit coefficients: the pseudo-inverse is still diagonal in the Pauli basis
    rng = np.random.default_rng(5)
    pidx = np.array([0, 7, 7, 12, 3, 1, 9, 9, 9], dtype=np.int32)
    coeffs = np.array([1.0, -1.0, 0.5, 1.0, 1.0, 2.0, 1.0, 1.0, -3.0])
    ex = rng.uniform(-.6, .6, size=(4, len(pidx)))
    rho = tm.linear_inv_state_estimate_batch(tm.MlePlan(2, pidx, coeffs), torch.from_numpy(ex).cuda()).cpu().numpy()
    for b in range(4):
        assert relerr(rho[b], orc.linear_inv_state_estimate(pidx, coeffs, ex[b], 2)) < 1e-12
    # drop-in signature (config 1: 1 qubit, X Y Z)
    qubits = [3]
    _, pidx, ex, cnt = orc.synth_state_tomography(1001, 1, 1)
    res = [ExperimentResult(ExperimentSetting(zeros_state(qubits), t), e, int(c))
           for t, e, c in zip(all_traceless_pauli_terms(qubits), ex[0], cnt[0])]
    assert relerr(tm.linear_inv_state_estimate(res, qubits), orc.linear_inv_state_estimate(pidx, np.ones(3), ex[0], 1)) < 1e-12


def test_state_log_likelihood(torch):
    """"next" row 2: state_log_likelihood vs the oracle, incl. pure states (zero-probability outcomes
    are skipped), duplicated / weighted observables, and the drop-in signature."""
    from forest_benchmarking_b200 import tomography as tm
    from forest_benchmarking_b200.observable_estimation import ExperimentResult, ExperimentSetting, zeros_state
    from forest_benchmarking_b200.utils import all_traceless_pauli_terms
    for n in (1, 2, 3, 4):
        rho, pidx, ex, cnt = orc.synth_state_tomography(2001 + n, 9, n)
        rho = np.array(rho)
        rho[0] = 0.0
        rho[0][0, 0] = 1.0  # |0..0><0..0|: P(-1) = 0 for every Z-type observable
        plan = tm.MlePlan(n, pidx)
        got = tm.state_log_likelihood_batch(plan, torch.from_numpy(rho).cuda(), torch.from_numpy(ex).cuda(),
                                            torch.from_numpy(cnt.astype(np.float64)).cuda()).cpu().numpy()
        want = np.array([orc.state_log_likelihood(rho[b], pidx, np.ones(len(pidx)), ex[b], cnt[b], n) for b in range(9)])
        assert np.all(np.isfinite(got))
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-11
    rng = np.random.default_rng(6)
    pidx = np.array([0, 7, 7, 12, 3, 1, 9, 9, 9], dtype=np.int32)
    coeffs = np.array([1.0, -1.0, 0.5, 1.0, 1.0, 0.25, 1.0, 1.0, -0.3])
    ex = rng.uniform(-.6, .6, size=(3, len(pidx)))
    cnt = rng.integers(100, 1000, size=ex.shape).astype(np.float64)
    rho = np.stack([orc.ginibre_state(rng, 4) for _ in range(3)])
    got = tm.state_log_likelihood_batch(tm.MlePlan(2, pidx, coeffs), torch.from_numpy(rho).cuda(),
                                        torch.from_numpy(ex).cuda(), torch.from_numpy(cnt).cuda()).cpu().numpy()
    for b in range(3):
        want = orc.state_log_likelihood(rho[b], pidx, coeffs, ex[b], cnt[b], 2)
        assert abs(got[b] - want) < 1e-11 * abs(want)
    qubits = [3, 5]
    rho, pidx, ex, cnt = orc.synth_state_tomography(2002, 1, 2)
    res = [ExperimentResult(ExperimentSetting(zeros_state(qubits), t), e, int(c))
           for t, e, c in zip(all_traceless_pauli_terms(qubits), ex[0], cnt[0])]
    want = orc.state_log_likelihood(rho[0], pidx, np.ones(len(pidx)), ex[0], cnt[0], 2)
    assert abs(tm.state_log_likelihood(rho[0], res, qubits) - want) < 1e-11 * abs(want)


def test_project_state_matrix_and_estimate_variance(torch):
    """"next" row 1: wizard projection vs the oracle, and the bootstrap estimate_variance as ONE batch
    vs a replica-by-replica NumPy restatement with the same seeded RNG stream (tomography.py:412-453)."""
    from forest_benchmarking_b200 import distance_measures as dm, tomography as tm
    from forest_benchmarking_b200.operator_tools import project_state_matrix as psm
    from forest_benchmarking_b200.observable_estimation import ExperimentResult, ExperimentSetting, zeros_state
    from forest_benchmarking_b200.utils import all_traceless_pauli_terms
    rng = np.random.default_rng(21)
    for d in (2, 4, 8, 16, 32):
        mats = []
        for k in range(40):
            g = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
            h = (g + g.conj().T) / 2
            h = h / np.trace(h).real + 0.03 * rng.standard_normal() * np.eye(d)
            mats.append(h if k % 3 else 2.5 * orc.ginibre_state(rng, d))  # every third one is already physical
        mats = np.stack(mats)
        got = psm.project_state_matrix_to_physical_batch(torch.from_numpy(mats).cuda()).cpu().numpy()
        want = np.stack([orc.project_state_matrix_to_physical(m) for m in mats])
        assert max_relerr(got, want) < 1e-10
        assert np.allclose(np.trace(got, axis1=1, axis2=2), 1, atol=1e-12)
        assert np.linalg.eigvalsh(got).min() > -1e-12
    eigs = np.diag(np.array(list(reversed([3.0 / 5, 1.0 / 2, 7.0 / 20, 1.0 / 10, -11.0 / 20]))))
    # the reference's own known answer needs d = 5 (not a qubit dimension): embed it in d = 8
    emb = np.zeros((8, 8)); emb[:5, :5] = eigs
    assert np.allclose(psm.project_state_matrix_to_physical(emb)[:5, :5], np.diag([0, 0, 1.0 / 5, 7.0 / 20, 9.0 / 20]))

    qubits = [0, 1]
    truth, pidx, ex, cnt = orc.synth_state_tomography(31, 1, 2, shots=200)
    res = [ExperimentResult(ExperimentSetting(zeros_state(qubits), t), e, int(c))
           for t, e, c in zip(all_traceless_pauli_terms(qubits), ex[0], cnt[0])]
    for est, oest, functional, tgt, proj in (
            (tm.linear_inv_state_estimate, "lin", dm.purity, None, True),
            (tm.linear_inv_state_estimate, "lin", dm.fidelity, truth[0], True),
            (tm.iterative_mle_state_estimate, "mle", dm.trace_distance, truth[0], False)):
        np.random.seed(9)
        mean, var = tm.estimate_variance(res, qubits, est, functional, target_state=tgt, n_resamples=12,
                                         project_to_physical=proj)
        np.random.seed(9)
        vals = []
        for _ in range(12):
            e = orc.resample_expectations_with_beta(ex[0], cnt[0])
            if oest == "lin":
                rho = orc.linear_inv_state_estimate(pidx, np.ones(15), e, 2)
            else:
                rho, _ = orc.mle_state_estimate(pidx, np.ones(15), e, cnt[0], 2)
            if proj:
                rho = orc.project_state_matrix_to_physical(rho)
            if functional == dm.purity:
                vals.append(orc.purity(rho, dim_renorm=False))
            elif functional == dm.fidelity:
                vals.append(orc.fidelity(tgt, rho))
            else:
                vals.append(orc.trace_distance(tgt, rho))
        assert abs(mean - np.mean(vals)) < 1e-8 and abs(var - np.var(vals)) < 1e-8
    with pytest.raises(ValueError):
        tm.estimate_variance(res, qubits, tm.linear_inv_state_estimate, dm.fidelity)


def test_hilbert_schmidt_and_process_fidelity(torch):
    """tr(A^dagger B) streaming kernel and the process-fidelity family built on it (distance_measures.py:198-375)."""
    from forest_benchmarking_b200 import distance_measures as dm
    rng = np.random.default_rng(17)
    for m, batch in ((2, 13), (4, 101), (16, 37), (64, 9), (256, 3)):
        a = rng.standard_normal((batch, m, m)) + 1j * rng.standard_normal((batch, m, m))
        b = rng.standard_normal((batch, m, m)) + 1j * rng.standard_normal((batch, m, m))
        got = dm.hilbert_schmidt_ip_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
        want = np.array([orc.hilbert_schmidt_ip(x, y) for x, y in zip(a, b)])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-10)
    for n in (1, 2, 3):
        d = 2 ** n
        u, v = orc.haar_unitary(rng, d), orc.haar_unitary(rng, d)
        p0, p1 = orc.kraus2pauli_liouville([u]), orc.kraus2pauli_liouville([np.sqrt(.9) * u, np.sqrt(.1) * v])
        assert abs(dm.entanglement_fidelity(p0, p1) - orc.entanglement_fidelity(p0, p1).real) < 1e-12
        assert abs(dm.process_fidelity(p0, p1) - orc.process_fidelity(p0, p1).real) < 1e-12
        assert abs(dm.process_fidelity(p0, p0) - 1.0) < 1e-12
        assert abs(dm.process_infidelity(p0, p1) - (1 - orc.process_fidelity(p0, p1).real)) < 1e-12
        rho, sig = orc.ginibre_state(rng, d), orc.ginibre_state(rng, d)
        assert abs(dm.hilbert_schmidt_ip(rho, sig) - orc.hilbert_schmidt_ip(rho, sig).real) < 1e-13


def test_fidelity_cholesky_fast_path_and_fallback(torch):
    """fidelity: positive-definite rho goes through the Cholesky fast path, rank-deficient / non-PSD rho through the
    reference's eigh + sqrtm_psd sequence; both must match the oracle (distance_measures.py:64-84)."""
    from forest_benchmarking_b200 import distance_measures as dm
    rng = np.random.default_rng(77)
    for n in (2, 3, 4, 5):
        d = 2 ** n
        rho, sig = [], []
        for b in range(24):
            if b % 4 == 0:      # pure state: Cholesky pivots vanish -> fallback
                r = orc.ginibre_state(rng, d, rank=1)
            elif b % 4 == 1:    # slightly non-PSD Hermitian, unit trace (e.g. a linear-inversion estimate)
                r = orc.ginibre_state(rng, d)
                h = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
                r = r + 0.3 / d * (h + h.conj().T)
                r = r / np.trace(r).real
            elif b % 4 == 2:    # ill-conditioned but positive definite
                r = 0.999 * orc.ginibre_state(rng, d, rank=1) + 0.001 * np.eye(d) / d
            else:
                r = orc.ginibre_state(rng, d)
            rho.append(r)
            sig.append(orc.ginibre_state(rng, d))
        rho, sig = np.stack(rho), np.stack(sig)
        fid = dm.fidelity_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sig).cuda()).cpu().numpy()
        for b in range(24):
            want = np.real(orc.fidelity(rho[b], sig[b]))
            assert abs(fid[b] - want) < 1e-6 * max(abs(want), 1e-3), (n, b, fid[b], want)
            if b % 4 == 3:
                assert abs(fid[b] - want) < 1e-10


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_fidelity_tridiagonal_kernel_tails_and_structured_inputs(torch, n):
    """The fidelity kernels for d = 4 .. 32 (Cholesky -> L^dagger sigma L -> Householder tridiagonalisation by d lanes per
    pair, then one lane per pair runs the square-root-free QL; d = 32: one pair per warp, matrix in shared memory): batch sizes that leave partial warps / partial rounds, and
    inputs whose tridiagonal form has exact zeros (diagonal, identical, commuting, basis-state sigma), against the
    oracle (distance_measures.py:64-84)."""
    from forest_benchmarking_b200 import distance_measures as dm
    rng = np.random.default_rng(100 + n)
    d = 2 ** n
    for B in ((1, 31, 33, 391) if n < 5 else (1, 17, 33, 100)):
        rho, sig, kind = [], [], []
        for b in range(B):
            k = b % 8
            r = orc.ginibre_state(rng, d)
            s = orc.ginibre_state(rng, d)
            if k == 1:      # both diagonal: Y is diagonal, every e^2 is exactly zero
                r, s = np.diag(np.diag(r).real).astype(complex), np.diag(np.diag(s).real).astype(complex)
            elif k == 2:    # identical states: F = 1
                s = r.copy()
            elif k == 3:    # sigma a computational basis state (rank 1, exact zeros)
                s = np.zeros((d, d), complex)
                s[b % d, b % d] = 1.0
            elif k == 4:    # maximally mixed rho
                r = np.eye(d, dtype=complex) / d
            elif k == 5:    # commuting pair (same eigenbasis)
                w, v = np.linalg.eigh(r)
                p = rng.dirichlet(np.ones(d))
                s = (v * p) @ v.conj().T
            elif k == 6:    # block-diagonal pair: the tridiagonal matrix splits
                r[: d // 2, d // 2:] = r[d // 2:, : d // 2] = 0
                s[: d // 2, d // 2:] = s[d // 2:, : d // 2] = 0
                r, s = r / np.trace(r).real, s / np.trace(s).real
            rho.append(r), sig.append(s), kind.append(k)
        rho, sig = np.stack(rho), np.stack(sig)
        fid = dm.fidelity_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sig).cuda()).cpu().numpy()
        assert fid.shape == (B,) and np.all(np.isfinite(fid))
        for b in range(0, B, 1 if B < 64 else 3):
            want = np.real(orc.fidelity(rho[b], sig[b]))
            tol = 1e-6 * max(abs(want), 1e-3) if kind[b] == 3 else 1e-10
            assert abs(fid[b] - want) < tol, (n, B, b, kind[b], fid[b], want)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_purity_all_sizes(torch, n):
    """tr(rho rho) for Hermitian and for general complex matrices (the reference computes np.trace(rho @ rho))."""
    from forest_benchmarking_b200 import distance_measures as dm
    rng = np.random.default_rng(300 + n)
    d = 2 ** n
    x = rng.standard_normal((37, d, d)) + 1j * rng.standard_normal((37, d, d))
    x[::2] = np.stack([orc.ginibre_state(rng, d) for _ in range(19)])
    got = dm.purity_batch(torch.from_numpy(x).cuda()).cpu().numpy()
    want = np.array([np.real(np.trace(m @ m)) for m in x])
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_c_abi_single_process_allgather(torch):
    """qt_comm_init_all / qt_allgather_bytes: the path's one collective for hosts without torch.distributed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from forest_benchmarking_b200 import tomography as tm
    from forest_benchmarking_b200.sharding import SingleProcessComm, shard_range
    _, pidx, ex, cnt = orc.synth_state_tomography(79, 32, 2)
    comm = SingleProcessComm([0, 1])
    parts = []
    for r in range(2):
        lo, hi = shard_range(32, 2, r)
        with torch.cuda.device(r):
            plan = tm.MlePlan(2, pidx)
            rho, _ = tm.iterative_mle_state_estimate_batch(plan, torch.from_numpy(ex[lo:hi]).to(f"cuda:{r}"), tol=1e-6,
                                                           maxiter=2000)
        parts.append(rho)
    full = comm.all_gather(parts)
    for r in range(2):
        torch.cuda.synchronize(r)
    want = torch.cat([p.cpu() for p in parts])
    assert torch.equal(full[0].cpu(), want) and torch.equal(full[1].cpu(), want)
    comm.close()
