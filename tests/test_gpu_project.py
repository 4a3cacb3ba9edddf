"""GPU parity: Choi projections (CP / TP / TNI / Dykstra physical) vs reference goldens and the oracle."""
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


@pytest.mark.parametrize("n", [1, 2, 3])
def test_golden(torch, n):
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    g = golden(f"algebra_n{n}")
    x = torch.from_numpy(g["noisy"]).cuda()
    assert max_relerr(ps.proj_choi_to_completely_positive_batch(x).cpu().numpy(), g["proj_cp"]) < 1e-12
    assert max_relerr(ps.proj_choi_to_trace_preserving_batch(x).cpu().numpy(), g["proj_tp"]) < 1e-14
    assert max_relerr(ps.proj_choi_to_trace_non_increasing_batch(x).cpu().numpy(), g["proj_tni"]) < 1e-12
    assert max_relerr(ps.proj_choi_to_physical_batch(x).cpu().numpy(), g["proj_physical"]) < TOL
    assert max_relerr(ps.proj_choi_to_physical_batch(x, False).cpu().numpy(), g["proj_physical_tni"]) < TOL


@pytest.mark.parametrize("n,batch", [(1, 301), (2, 67), (3, 5)])
def test_vs_oracle_and_properties(torch, n, batch):
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    rng = np.random.default_rng(70 + n)
    d, m = 2 ** n, 4 ** n
    xs = []
    for b in range(batch):
        base = orc.kraus2choi([orc.haar_unitary(rng, d)]) if b % 2 else np.eye(m) / d
        noise = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        xs.append(base + (noise + noise.conj().T) * (0.3 / m))
    xs = np.stack(xs)
    xd = torch.from_numpy(xs).cuda()
    cp = ps.proj_choi_to_completely_positive_batch(xd).cpu().numpy()
    tp = ps.proj_choi_to_trace_preserving_batch(xd).cpu().numpy()
    phys, counts = ps.proj_choi_to_physical_batch(xd, return_counts=True)
    phys, counts = phys.cpu().numpy(), counts.cpu().numpy()
    picks = range(batch) if n < 3 else range(3)
    for b in picks:
        assert relerr(cp[b], orc.proj_choi_to_completely_positive(xs[b])) < 1e-12
        assert relerr(tp[b], orc.proj_choi_to_trace_preserving(xs[b])) < 1e-14
        want, ne = orc.proj_choi_to_physical(xs[b], return_count=True)
        assert relerr(phys[b], want) < TOL
        assert abs(int(counts[b]) - ne) <= 1
    # properties at every item: CP output PSD + idempotent; physical output CP and TP
    assert np.linalg.eigvalsh(cp).min() > -1e-12
    cp2 = ps.proj_choi_to_completely_positive_batch(torch.from_numpy(cp).cuda()).cpu().numpy()
    assert max_relerr(cp2, cp) < 1e-12
    pt = np.einsum("zijkj->zik", phys.reshape(batch, d, d, d, d))
    assert np.abs(pt - np.eye(d)).max() < 1e-10
    assert np.linalg.eigvalsh(phys).min() > -5e-3   # Dykstra stops at 1e-4 on squared norms


def test_known_answers_and_dropin(torch):
    """reference tests/test_project_superoperators.py:15-31, 79-90 (values re-derived)."""
    from forest_benchmarking_b200 import operator_tools as ot
    z = np.diag([1.0, -1.0]); x = np.array([[0, 1.0], [1, 0]]); y = np.array([[0, -1j], [1j, 0]])
    for op in (z, x, y):
        neg = -orc.kraus2choi([op])
        assert np.allclose(ot.proj_choi_to_completely_positive(neg), 0, atol=1e-14)
        pos = orc.kraus2choi([op])
        assert np.allclose(ot.proj_choi_to_completely_positive(pos), pos, atol=1e-14)
        assert np.allclose(ot.proj_choi_to_physical(pos), pos, atol=1e-12)
        assert np.allclose(ot.proj_choi_to_trace_preserving(pos), pos, atol=1e-15)
    bad = orc.kraus2choi([z]) * 1.3
    assert np.allclose(ot.proj_choi_to_trace_non_increasing(bad), orc.proj_choi_to_trace_non_increasing(bad), atol=1e-13)
    with pytest.raises(ValueError):
        ot.proj_choi_to_completely_positive(np.full((4, 4), np.nan))


@pytest.mark.parametrize("n,batch", [(1, 33), (2, 21), (3, 5)])
def test_proj_choi_to_unitary(torch, n, batch):
    """"next" row 3: closest unitary vs the oracle (eigh + SVD), noisy non-Hermitian inputs; a unitary
    channel is a fixed point up to the phase convention."""
    from forest_benchmarking_b200.operator_tools import project_superoperators as pj
    rng = np.random.default_rng(60 + n)
    d = 2 ** n
    chois, us = [], []
    for _ in range(batch):
        u = orc.haar_unitary(rng, d)
        g = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
        chois.append(.9 * orc.kraus2choi(u) + .1 * (g @ g.conj().T) / d ** 2 + .01 * g)
        us.append(u)
    chois = np.stack(chois)
    got = pj.proj_choi_to_unitary_batch(torch.from_numpy(chois).cuda()).cpu().numpy()
    for b in range(batch):
        assert relerr(got[b], orc.proj_choi_to_unitary(chois[b])) < 1e-9
        # the closest unitary is near the unitary the process was built around
        assert relerr(got[b], orc.kraus2choi(us[b] * np.exp(-1j * np.angle(us[b][0, 0])))) < 0.6
    exact = np.stack([orc.kraus2choi(u) for u in us])
    fixed = pj.proj_choi_to_unitary_batch(torch.from_numpy(exact).cuda()).cpu().numpy()
    assert max_relerr(fixed, exact) < 1e-10
    assert relerr(pj.proj_choi_to_unitary(chois[0]), orc.proj_choi_to_unitary(chois[0])) < 1e-9


@pytest.mark.parametrize("n", [1, 2, 3])
def test_tni_in_place_and_out_of_place_agree(torch, n):
    """The trace-non-increasing projection takes the two-pass path out of place (n >= 2) and the fused kernel in
    place: both must give the oracle's answer, also for inputs whose partial trace has eigenvalues above AND below 1."""
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    rng = np.random.default_rng(90 + n)
    d, m = 2 ** n, 4 ** n
    xs = []
    for b in range(37):
        g = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        xs.append((g @ g.conj().T) * (rng.uniform(.3, 2.0) / m) + 0.05 * g)
    xs = np.stack(xs)
    want = np.stack([orc.proj_choi_to_trace_non_increasing(x) for x in xs])
    xd = torch.from_numpy(xs).cuda()
    oop = ps.proj_choi_to_trace_non_increasing_batch(xd).cpu().numpy()
    inp = xd.clone()
    ps.proj_choi_to_trace_non_increasing_batch(inp, out=inp)
    assert max_relerr(oop, want) < 1e-12
    assert max_relerr(inp.cpu().numpy(), want) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3])
def test_nonhermitian_input_matches_reference_trip_counts(torch, n):
    """The reference does not Hermitise the input of proj_choi_to_physical: the anti-Hermitian part stays in
    old_CP_change and changes the Birgin-Raydan stopping rule (project_superoperators.py:112-136).  Goldens from the
    reference itself (oracle/make_golden.py --only nonherm): same projection AND same number of CP projections."""
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    g = golden("proj_physical_nonherm")
    x = torch.from_numpy(g[f"n{n}_in"]).cuda()
    for tp, key in ((True, ""), (False, "_tni")):
        out, calls, status = ps.proj_choi_to_physical_batch(x, tp, return_counts=True, return_status=True)
        assert max_relerr(out.cpu().numpy(), g[f"n{n}_out{key}"]) < TOL
        assert np.array_equal(calls.cpu().numpy(), g[f"n{n}_calls{key}"]), (calls.cpu().numpy(), g[f"n{n}_calls{key}"])
        assert not status.cpu().numpy().any()
    # Hermitian inputs of the same family: the counts the reference gives for the Hermitised matrices
    xh = (x + x.conj().transpose(1, 2)) / 2
    _, calls = ps.proj_choi_to_physical_batch(xh, True, return_counts=True)
    assert np.array_equal(calls.cpu().numpy(), g[f"n{n}_calls_hermitised"])
    # drop-in, single matrix
    assert relerr(ps.proj_choi_to_physical(g[f"n{n}_in"][0]), g[f"n{n}_out"][0]) < TOL
    with pytest.raises(ValueError):
        ps.proj_choi_to_physical_batch(x, out=x)


def test_eigh_tolerance_is_a_per_call_argument(torch):
    """No process-global tuning state: two calls with different tolerances, then the default again -> the default
    results are bit-identical, the tight result agrees to 1e-9."""
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    g = golden("algebra_n2")
    x = torch.from_numpy(g["noisy"]).cuda()
    a = ps.proj_choi_to_physical_batch(x)
    tight = ps.proj_choi_to_physical_batch(x, eigh_rel_tol=0.0)
    loose = ps.proj_choi_to_physical_batch(x, eigh_rel_tol=1e-5)
    b = ps.proj_choi_to_physical_batch(x)
    assert torch.equal(a, b)
    assert max_relerr(a.cpu().numpy(), tight.cpu().numpy()) < 1e-9
    assert max_relerr(loose.cpu().numpy(), g["proj_physical"]) < 1e-3
    with pytest.raises(Exception):
        ps.proj_choi_to_physical_batch(x, eigh_rel_tol=0.5)


@pytest.mark.parametrize("n,batch", [(4, 3), (5, 1)])
def test_large_choi_projections(torch, n, batch):
    """n = 4, 5 (256 x 256 / 1024 x 1024 Choi matrices: the reference functions are size-agnostic,
    project_superoperators.py:19-144): CP / TP / TNI vs the oracle; Dykstra physical projection with trip counts at
    n = 4 (Hermitian and non-Hermitian input)."""
    from forest_benchmarking_b200.operator_tools import project_superoperators as ps
    rng = np.random.default_rng(400 + n)
    d, m = 2 ** n, 4 ** n
    xs = []
    for b in range(batch):
        ks = [np.sqrt(.6) * orc.haar_unitary(rng, d), np.sqrt(.4) * orc.haar_unitary(rng, d)]
        noise = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        xs.append(orc.kraus2choi(ks) + (noise + noise.conj().T) * (0.2 / m) + (0.01 / m if b == 1 else 0.0) * noise)
    xs = np.stack(xs)
    xd = torch.from_numpy(xs).cuda()
    cp = ps.proj_choi_to_completely_positive_batch(xd).cpu().numpy()
    tp = ps.proj_choi_to_trace_preserving_batch(xd).cpu().numpy()
    tni = ps.proj_choi_to_trace_non_increasing_batch(xd).cpu().numpy()
    for b in range(batch):
        assert relerr(cp[b], orc.proj_choi_to_completely_positive(xs[b])) < 1e-11
        assert relerr(tp[b], orc.proj_choi_to_trace_preserving(xs[b])) < 1e-13
        assert relerr(tni[b], orc.proj_choi_to_trace_non_increasing(xs[b])) < 1e-11
    assert np.linalg.eigvalsh((cp[0] + cp[0].conj().T) / 2).min() > -1e-11
    with pytest.raises(ValueError):
        ps.proj_choi_to_trace_preserving_batch(xd, out=xd)
    if n == 4:
        phys, calls, status = ps.proj_choi_to_physical_batch(xd, return_counts=True, return_status=True)
        phys, calls = phys.cpu().numpy(), calls.cpu().numpy()
        assert not status.cpu().numpy().any()
        for b in range(batch):
            want, ne = orc.proj_choi_to_physical(xs[b], return_count=True)
            assert relerr(phys[b], want) < TOL and int(calls[b]) == ne, (calls[b], ne)
        pt = np.einsum("zijkj->zik", phys.reshape(batch, d, d, d, d))
        assert np.abs(pt - np.eye(d)).max() < 1e-9
        t, ct = ps.proj_choi_to_physical_batch(xd[:1], False, return_counts=True)
        want, ne = orc.proj_choi_to_physical(xs[0], False, return_count=True)
        assert relerr(t[0].cpu().numpy(), want) < TOL and int(ct[0]) == ne
