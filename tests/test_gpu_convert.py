"""GPU parity: superoperator conversions vs the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from oracle import ref_numpy as orc
from util import golden, relerr, max_relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("n", [1, 2, 3])
def test_golden_chain(torch, n):
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    g = golden(f"algebra_n{n}")
    kraus = torch.from_numpy(g["kraus"]).cuda()
    choi = st.kraus2choi_batch(kraus)
    assert max_relerr(choi.cpu().numpy(), g["choi"]) < 1e-14
    assert max_relerr(st.kraus2superop_batch(kraus).cpu().numpy(), g["kraus2superop"]) < 1e-14
    sup = st.reshuffle_batch(choi)
    assert np.array_equal(st.reshuffle_batch(torch.from_numpy(g["choi"]).cuda()).cpu().numpy(), g["superop"])  # bit-exact
    pl = st.superop2pauli_liouville_batch(torch.from_numpy(g["superop"]).cuda())
    assert max_relerr(pl.cpu().numpy(), g["pauli_liouville"]) < 1e-14
    back = st.pauli_liouville2superop_batch(torch.from_numpy(g["pauli_liouville"]).cuda())
    assert max_relerr(back.cpu().numpy(), g["pl2superop"]) < 1e-14
    assert max_relerr(st.pauli_liouville2choi_batch(pl).cpu().numpy(), g["choi"]) < 1e-13
    assert max_relerr(st.choi2pauli_liouville_batch(choi).cpu().numpy(), g["pauli_liouville"]) < 1e-13
    assert sup.shape == choi.shape


@pytest.mark.parametrize("n,batch", [(1, 1000), (2, 333), (3, 37), (4, 5), (5, 2)])
def test_sweep_vs_oracle_and_roundtrip(torch, n, batch):
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rng = np.random.default_rng(4004 + n)
    d = 2 ** n
    kraus = np.stack([np.stack([np.sqrt(.7) * orc.haar_unitary(rng, d), np.sqrt(.3) * orc.haar_unitary(rng, d)])
                      for _ in range(batch)])
    kd = torch.from_numpy(kraus).cuda()
    choi = st.kraus2choi_batch(kd)
    sup = st.reshuffle_batch(choi)
    pl = st.superop2pauli_liouville_batch(sup)
    sup2 = st.pauli_liouville2superop_batch(pl)
    choi2 = st.reshuffle_batch(sup2)
    torch.cuda.synchronize()
    # round trip (size-independent property) and reshuffle involution (bit-exact)
    assert max_relerr(choi2.cpu().numpy(), choi.cpu().numpy()) < 1e-13
    assert torch.equal(st.reshuffle_batch(sup), choi)
    # PTM of a CPTP map: real, first row (1,0,...,0)
    plh = pl.cpu().numpy()
    assert np.abs(plh.imag).max() < 1e-13 and np.allclose(plh[:, 0, 0], 1) and np.abs(plh[:, 0, 1:]).max() < 1e-13
    # spot checks against the oracle (dense c2p @ S @ c2p^dagger with a 1024 x 1024 basis matrix at n = 5: one item)
    picks = range(batch) if n <= 3 else [0]
    ch = choi.cpu().numpy(); sh = sup.cpu().numpy(); ks = st.kraus2superop_batch(kd).cpu().numpy()
    for b in picks:
        assert relerr(ch[b], orc.kraus2choi(list(kraus[b]))) < 1e-14
        assert np.array_equal(sh[b], orc.choi2superop(ch[b]))
        assert relerr(ks[b], orc.kraus2superop(list(kraus[b]))) < 1e-14
        want_pl = orc.superop2pauli_liouville(sh[b])
        assert relerr(plh[b], want_pl) < 1e-13
        # and the inverse direction from the ORACLE's PTM (not from our own forward result)
        back = st.pauli_liouville2superop_batch(torch.from_numpy(np.ascontiguousarray(want_pl[None])).cuda())
        assert relerr(back[0].cpu().numpy(), sh[b]) < 1e-13


def test_pauli_basis_index_pin(torch):
    """reference tests/test_superoperator_transformations.py:192-201: index 7 of the 2-qubit basis is X(x)Z."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    xz = np.kron(np.array([[0, 1], [1, 0]]), np.diag([1.0, -1.0]))
    sup = orc.kraus2superop([xz])  # conjugation by XZ: diagonal +-1 PTM
    pl = st.superop2pauli_liouville(sup)
    want = np.array([1 if np.allclose(xz @ orc.pauli_matrix(k, 2) @ xz.conj().T, orc.pauli_matrix(k, 2)) else -1
                     for k in range(16)], dtype=float)
    assert np.allclose(pl, np.diag(want), atol=1e-15)
    e7 = np.zeros((16, 16), dtype=complex); e7[7, 0] = 1.0
    s = st.pauli_liouville2superop(e7)  # column 0 of p2c-basis: vec(P_7)/d x vec(I)^dagger
    assert relerr(s, orc.pauli_liouville2superop(e7)) < 1e-15


def test_dropin_functions(torch):
    from forest_benchmarking_b200 import operator_tools as ot
    had = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    assert np.allclose(ot.kraus2pauli_liouville(had), np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, -1, 0], [0, 1, 0, 0]]),
                       atol=1e-15)  # Hadamard PTM, tests/test_superoperator_transformations.py:43-73 re-derived
    p = .1
    ad = [np.array([[1, 0], [0, np.sqrt(1 - p)]]), np.array([[0, np.sqrt(p)], [0, 0]])]
    c = ot.kraus2choi(ad)
    assert np.allclose(c, orc.kraus2choi(ad), atol=1e-16)
    assert np.allclose(ot.choi2superop(c), orc.choi2superop(c)) and np.allclose(ot.superop2choi(ot.choi2superop(c)), c)
    assert np.allclose(ot.choi2pauli_liouville(c), orc.choi2pauli_liouville(c), atol=1e-15)
    assert np.allclose(ot.pauli_liouville2choi(ot.choi2pauli_liouville(c)), c, atol=1e-15)
    assert np.allclose(ot.kraus2superop(ad), orc.kraus2superop(ad), atol=1e-16)
    assert np.array_equal(ot.vec(np.arange(4).reshape(2, 2)), np.array([[0], [2], [1], [3]]))
    assert np.array_equal(ot.unvec(ot.vec(np.arange(4).reshape(2, 2))), np.arange(4).reshape(2, 2))


@pytest.mark.parametrize("n,batch", [(1, 257), (2, 61), (3, 9)])
def test_choi2kraus(torch, n, batch):
    """a17 choi2kraus: not element-wise comparable (eigenvector gauge) -- check what the reference's own test
    checks (test_superoperator_transformations.py:263-271): kraus2choi(choi2kraus(C)) == C, plus sorted
    eigenvalues vs np.linalg.eigh, the |lambda| > tol filter and the complex sqrt of negative eigenvalues."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rng = np.random.default_rng(70 + n)
    d, m = 2 ** n, 4 ** n
    chois = []
    for b in range(batch):
        nk = 1 + b % 3
        ks = [np.sqrt(1.0 / nk) * orc.haar_unitary(rng, d) for _ in range(nk)]
        c = orc.kraus2choi(ks)
        if b % 4 == 3:  # indefinite Hermitian input: negative eigenvalues -> imaginary sqrt
            h = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
            c = c + 0.05 * (h + h.conj().T)
        chois.append(c)
    chois = np.stack(chois)
    kraus, counts, evals = st.choi2kraus_batch(torch.from_numpy(chois).cuda())
    kraus, counts, evals = kraus.cpu().numpy(), counts.cpu().numpy(), evals.cpu().numpy()
    for b in range(batch):
        w = np.linalg.eigvalsh(chois[b])
        assert np.allclose(evals[b], w, atol=1e-12)
        want = orc.choi2kraus(chois[b])
        assert counts[b] == len(want) == int(np.sum(np.abs(w) > 1e-9))
        assert np.all(kraus[b, counts[b]:] == 0)
        # sum_k vec(K) vec(K)^T-with-scimath-sqrt reproduces C:  sqrt(l)^2 = l also for negative l
        rec = np.zeros((m, m), dtype=complex)
        for k, lam in zip(kraus[b, :counts[b]], w[np.abs(w) > 1e-9]):
            v = orc.vec(k) / np.emath.sqrt(lam)
            rec += lam * (v @ v.conj().T)
        assert relerr(rec, chois[b]) < 1e-9
    # drop-in signature: list of operators, round trip through kraus2choi
    ks = st.choi2kraus(chois[0])
    assert relerr(st.kraus2choi(ks), chois[0]) < 1e-9


def test_chi_matrix_family(torch):
    """chi-matrix conversions as compositions of the conversion kernels, vs the oracle and the
    reference's known answers (test_superoperator_transformations.py:139-145, 184-189)."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    had = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    had_chi = 0.5 * np.array([[0, 0, 0, 0], [0, 1, 0, 1], [0, 0, 0, 0], [0, 1, 0, 1]])
    assert np.allclose(st.kraus2chi(had), had_chi)
    p = 0.1
    ad = [np.array([[1, 0], [0, np.sqrt(1 - p)]]), np.array([[0, np.sqrt(p)], [0, 0]])]
    assert np.allclose(st.kraus2chi(ad), orc.kraus2chi(ad))
    assert np.allclose(st.chi2pauli_liouville(orc.kraus2chi(ad)), orc.kraus2pauli_liouville(ad))
    rng = np.random.default_rng(90)
    for n, batch in ((1, 50), (2, 20), (3, 4)):
        d = 2 ** n
        kraus = np.stack([[np.sqrt(.7) * orc.haar_unitary(rng, d), np.sqrt(.3) * orc.haar_unitary(rng, d)]
                          for _ in range(batch)])
        chi = st.kraus2chi_batch(torch.from_numpy(kraus).cuda())
        want = np.stack([orc.kraus2chi(list(k)) for k in kraus])
        assert max_relerr(chi.cpu().numpy(), want) < 1e-12
        choi = st.chi2choi_batch(chi).cpu().numpy()
        assert max_relerr(choi, np.stack([orc.kraus2choi(list(k)) for k in kraus])) < 1e-12
        c0 = orc.kraus2choi(list(kraus[0]))
        assert relerr(st.chi2superop(want[0]), orc.chi2superop(want[0])) < 1e-12
        assert relerr(st.choi2chi(c0), orc.choi2chi(c0)) < 1e-9
        assert relerr(st.superop2chi(orc.reshuffle(c0)), orc.superop2chi(orc.reshuffle(c0))) < 1e-9
        assert relerr(st.pauli_liouville2chi(orc.choi2pauli_liouville(c0)), want[0]) < 1e-9
        ks = st.chi2kraus(want[0])
        assert relerr(st.kraus2choi(ks), c0) < 1e-9


@pytest.mark.parametrize("n,batch", [(4, 4), (5, 2)])
def test_choi2kraus_large(torch, n, batch):
    """a17 at n = 4, 5: eigenvalues vs numpy's eigh of the lower triangle, the reference's own round trip
    kraus2choi(choi2kraus(C)) == C, and the operator count.  Item 0 is full rank and item 3 has rank 10 (general one-sided
    Jacobi solver out of a global workspace); items 1, 2 have rank 3 (certified low-rank fast path, 8 probe columns)."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rng = np.random.default_rng(70 + n)
    d, m = 2 ** n, 4 ** n
    chois = []
    for i in range(batch):
        ks = [rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)) for _ in range(10 if i == 3 else 3)]
        c = sum(orc.kraus2choi(k) for k in ks) / (3 * d)
        if i == 0:
            g = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
            c = c + 1e-3 * (g + g.conj().T)  # full rank, indefinite
        chois.append(c)
    chois = np.stack(chois)
    kraus, counts, evals = st.choi2kraus_batch(torch.from_numpy(chois).cuda())
    kraus, counts, evals = kraus.cpu().numpy(), counts.cpu().numpy(), evals.cpu().numpy()
    for b in range(batch):
        want = np.linalg.eigvalsh(chois[b])
        assert np.abs(evals[b] - want).max() < 1e-11 * np.abs(want).max()
        assert counts[b] == int((np.abs(want) > 1e-9).sum())
        back = sum(orc.kraus2choi(k) for k in kraus[b, :counts[b]])
        # negative eigenvalues come back as i sqrt|l| v, whose outer product is +|l| v v^dagger: compare the PSD part
        vals, vecs = np.linalg.eigh(chois[b])
        ref = (vecs * np.abs(np.where(np.abs(vals) > 1e-9, vals, 0.0))) @ vecs.conj().T
        assert relerr(back, ref) < 1e-9
        assert np.abs(kraus[b, counts[b]:]).max(initial=0.0) == 0.0
    if batch > 1:
        assert counts[1] == 3
    if batch > 3:
        assert counts[3] == 10


@pytest.mark.parametrize("n,batch", [(2, 37), (3, 19)])
def test_dense_fp64_mma_ptm_variant(torch, n, batch):
    """The reference's dense formulation (two complex products with the 4^n x 4^n basis matrix,
    superoperator_transformations.py:253-264, 301-312) on the FP64 tensor path vs the oracle and vs the default
    butterfly kernels, both directions, ragged batch sizes."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rng = np.random.default_rng(900 + n)
    m = 4 ** n
    x = rng.standard_normal((batch, m, m)) + 1j * rng.standard_normal((batch, m, m))  # generic (non-physical) matrices
    xd = torch.from_numpy(x).cuda()
    fwd = st.superop2pauli_liouville_batch(xd, variant="dense_mma").cpu().numpy()
    inv = st.pauli_liouville2superop_batch(xd, variant="dense_mma").cpu().numpy()
    for b in range(0, batch, 6):
        assert relerr(fwd[b], orc.superop2pauli_liouville(x[b])) < 1e-13
        assert relerr(inv[b], orc.pauli_liouville2superop(x[b])) < 1e-13
    assert max_relerr(fwd, st.superop2pauli_liouville_batch(xd).cpu().numpy()) < 1e-13
    assert max_relerr(inv, st.pauli_liouville2superop_batch(xd).cpu().numpy()) < 1e-13
    with pytest.raises(Exception):
        st.superop2pauli_liouville_batch(torch.zeros((2, 4, 4), dtype=torch.complex128, device="cuda"), variant="dense_mma")


@pytest.mark.parametrize("n,batch", [(4, 61), (5, 5)])
def test_ptm_fused_two_pass_path(torch, n, batch):
    """n = 4, 5 with enough matrices for the single-launch path (both passes through an L2-resident ring with
    producer / consumer counters): every matrix against the two-kernel path taken by small batches (bit-identical: the
    same arithmetic), spot checks against the oracle, ring re-use (batch > ring slots), both directions."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rng = np.random.default_rng(1300 + n)
    m = 4 ** n
    x = torch.from_numpy(rng.standard_normal((batch, m, m)) + 1j * rng.standard_normal((batch, m, m))).cuda()
    for fn, ref in ((st.superop2pauli_liouville_batch, orc.superop2pauli_liouville),
                    (st.pauli_liouville2superop_batch, orc.pauli_liouville2superop)):
        for rep in range(3):  # repeated launches re-use the workspace tail (queue + counters are reset per call)
            got = fn(x)
        small = torch.cat([fn(x[i:i + 2].contiguous()) for i in range(0, batch, 2)])[:batch]
        assert torch.equal(got, small)
        for b in ((0, batch // 2, batch - 1) if n == 4 else (batch - 1,)):
            assert relerr(got[b].cpu().numpy(), ref(x[b].cpu().numpy())) < 1e-13


def test_empty_batches_everywhere(torch):
    """B = 0 through every batched entry point (the reference's list-based API degenerates to empty lists): no launch, an
    empty result of the right shape."""
    from forest_benchmarking_b200 import tomography as tm, distance_measures as dm
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st, project_superoperators as ps
    from forest_benchmarking_b200.operator_tools.project_state_matrix import project_state_matrix_to_physical_batch
    z = lambda *shape: torch.empty(shape, dtype=torch.complex128, device="cuda")
    for n in (1, 3, 4):
        m, d = 4 ** n, 2 ** n
        assert st.kraus2choi_batch(z(0, 2, d, d)).shape == (0, m, m)
        assert st.reshuffle_batch(z(0, m, m)).shape == (0, m, m)
        assert st.superop2pauli_liouville_batch(z(0, m, m)).shape == (0, m, m)
        assert st.pauli_liouville2superop_batch(z(0, m, m)).shape == (0, m, m)
        k, c, e = st.choi2kraus_batch(z(0, m, m))
        assert k.shape == (0, m, d, d) and c.shape == (0,) and e.shape == (0, m)
        assert ps.proj_choi_to_completely_positive_batch(z(0, m, m)).shape == (0, m, m)
        assert ps.proj_choi_to_trace_preserving_batch(z(0, m, m)).shape == (0, m, m)
        assert ps.proj_choi_to_physical_batch(z(0, m, m)).shape == (0, m, m)
        assert dm.fidelity_batch(z(0, d, d), z(0, d, d)).shape == (0,)
        assert dm.trace_distance_batch(z(0, d, d), z(0, d, d)).shape == (0,)
        assert project_state_matrix_to_physical_batch(z(0, d, d)).shape == (0, d, d)
    plan = tm.PgdbPlan.complete(1)
    e0 = torch.empty((0, plan.S), dtype=torch.float64, device="cuda")
    assert tm.pgdb_process_estimate_batch(plan, e0, e0).shape == (0, 4, 4)
    assert tm.linear_inv_process_estimate_batch(plan, e0).shape == (0, 4, 4)
    torch.cuda.synchronize()
