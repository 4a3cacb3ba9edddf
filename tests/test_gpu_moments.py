"""GPU parity: raw shots -> observable moments ("next" row 4) vs the oracle; counts are bit-exact."""
import numpy as np
import pytest

from oracle import ref_numpy as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _bits(rng, b, s, q):
    p = rng.uniform(.05, .95, size=(b, 1, q))
    return (rng.random((b, s, q)) < p).astype(np.uint8)


@pytest.mark.parametrize("q", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 13, 16, 17])
@pytest.mark.parametrize("s", [1, 37, 500, 1000])
def test_shots_to_obs_moments_batch(torch, q, s):
    """Every width (SWAR kernels for 1/2/4/8 columns, compacting stream kernel for the other widths up to 16, byte kernel beyond), shot counts that leave the settings
    unaligned to the 16-byte words, all column subsets incl. the identity term, both estimators."""
    from forest_benchmarking_b200 import observable_estimation as oe
    rng = np.random.default_rng(100 * q + s)
    b = 67
    bits = _bits(rng, b, s, q)
    masks = rng.integers(0, 2 ** q, size=b).astype(np.int32)
    masks[0], masks[1] = 0, 2 ** q - 1
    coeffs = rng.choice([1.0, -1.0, 0.5, 2.0], size=b)
    for prior in (False, True):
        mean, var = oe.shots_to_obs_moments_batch(torch.from_numpy(bits).cuda(), torch.from_numpy(masks).cuda(),
                                                  torch.from_numpy(coeffs).cuda(), prior)
        mean, var = mean.cpu().numpy(), var.cpu().numpy()
        for i in range(b):
            idxs = [c for c in range(q) if (masks[i] >> c) & 1]
            wm, wv = orc.shots_to_obs_moments(bits[i], idxs, coeffs[i], prior)
            assert abs(mean[i] - wm) <= 1e-13 * max(1.0, abs(wm)), (i, mean[i], wm)
            assert abs(var[i] - wv) <= 1e-12 * max(abs(wv), 1e-300) + 1e-18, (i, var[i], wv)
            if not prior and idxs:
                # integer bookkeeping is exact: the mean encodes n_plus - n_minus
                vals = np.prod(1 - 2 * bits[i][:, idxs].astype(np.int64), axis=1)
                assert round(mean[i] / coeffs[i] * s) == int(vals.sum())


def test_unaligned_view_and_dropin(torch):
    from forest_benchmarking_b200 import observable_estimation as oe
    from forest_benchmarking_b200.paulis import PauliTerm
    rng = np.random.default_rng(3)
    # a view that starts 3 bytes into an allocation: the flat-stream kernel must not care
    raw = torch.from_numpy(rng.integers(0, 2, size=3 + 5 * 123 * 2, dtype=np.uint8)).cuda()
    view = raw[3:].view(5, 123, 2)
    bits = view.cpu().numpy()
    masks = np.array([3, 1, 2, 3, 0], dtype=np.int32)
    mean, var = oe.shots_to_obs_moments_batch(view, torch.from_numpy(masks).cuda())
    for i in range(5):
        idxs = [c for c in range(2) if (masks[i] >> c) & 1]
        wm, wv = orc.shots_to_obs_moments(bits[i], idxs)
        assert abs(mean[i].item() - wm) < 1e-14 and abs(var[i].item() - wv) < 1e-15
    # drop-in signature, pyquil-style observable on a subset of the measured qubits
    qubits = [4, 7, 9]
    bits = (rng.random((300, 3)) < [.2, .5, .8]).astype(np.int64)
    term = PauliTerm.from_list([("X", 7), ("Z", 9)], coefficient=-0.5)
    for prior in (False, True):
        got = oe.shots_to_obs_moments(bits, qubits, term, prior)
        want = orc.shots_to_obs_moments(bits, [1, 2], -0.5, prior)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-16)
    ident = PauliTerm("I", 0, 0.7)
    assert oe.shots_to_obs_moments(bits, qubits, ident) == (0.7, 0)


def test_calibrate_estimates_batch(torch):
    from forest_benchmarking_b200 import observable_estimation as oe
    rng = np.random.default_rng(4)
    a, va = rng.uniform(-1, 1, 1000), rng.uniform(1e-4, 1e-2, 1000)
    b, vb = rng.uniform(.7, 1, 1000), rng.uniform(1e-5, 1e-3, 1000)
    cm, cv = oe.calibrate_estimates_batch(*(torch.from_numpy(x).cuda() for x in (a, va, b, vb)))
    assert np.allclose(cm.cpu().numpy(), a / b, rtol=1e-15)
    assert np.allclose(cv.cpu().numpy(), orc.ratio_variance(a, va, b, vb), rtol=1e-14)
    assert np.allclose(oe.ratio_variance(a, va, b, vb), orc.ratio_variance(a, va, b, vb), rtol=1e-15)


def test_next_rows_golden(torch):
    """Every SURVEY 8(f) entry point against outputs of the reference's own functions (tests/golden/next_rows.npz)."""
    from util import golden, relerr
    from forest_benchmarking_b200 import observable_estimation as oe, tomography as tm
    from forest_benchmarking_b200.operator_tools import project_superoperators as pj
    g = golden("next_rows")
    for tag, n in (("lip_1q_pauli", 1), ("lip_1q_sic", 1), ("lip_2q_sic", 2)):
        plan = tm.PgdbPlan(n, g[tag + "_codes"], g[tag + "_pidx"])
        got = tm.linear_inv_process_estimate_batch(plan, torch.from_numpy(np.ascontiguousarray(g[tag + "_ex"])).cuda())
        for b in range(3):
            assert relerr(got[b].cpu().numpy(), g[tag + "_choi"][b]) < 1e-10
    for n in (1, 2, 3):
        got = pj.proj_choi_to_unitary_batch(torch.from_numpy(g[f"unitary_n{n}_in"]).cuda()).cpu().numpy()
        for b in range(4):
            assert relerr(got[b], g[f"unitary_n{n}_out"][b]) < 1e-9
    plan = tm.MlePlan(2, g["ll_pidx"])
    ll = tm.state_log_likelihood_batch(plan, torch.from_numpy(g["ll_rho"]).cuda(), torch.from_numpy(g["ll_ex"]).cuda(),
                                       torch.from_numpy(g["ll_cnt"].astype(np.float64)).cuda()).cpu().numpy()
    assert np.max(np.abs(ll - g["ll_value"]) / np.abs(g["ll_value"])) < 1e-11
    for prior in (0, 1):
        mean, var = oe.shots_to_obs_moments_batch(torch.from_numpy(g["mom_bits"]).cuda(),
                                                  torch.from_numpy(g["mom_masks"]).cuda(),
                                                  torch.from_numpy(g["mom_coeffs"]).cuda(), bool(prior))
        assert np.allclose(mean.cpu().numpy(), g["mom_out"][prior, :, 0], rtol=1e-13, atol=1e-16)
        assert np.allclose(var.cpu().numpy(), g["mom_out"][prior, :, 1], rtol=1e-12, atol=1e-18)
    a, va, b, vb = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in g["rv_in"])
    cm, cv = oe.calibrate_estimates_batch(a, va, b, vb)
    assert np.allclose(cv.cpu().numpy(), g["rv_out"], rtol=1e-14)
    assert np.allclose(cm.cpu().numpy(), g["rv_in"][0] / g["rv_in"][2], rtol=1e-15)
