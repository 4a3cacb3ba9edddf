"""Prototype (CPU, numpy): warm-started cyclic Jacobi stopped EARLY at a relative off-diagonal norm tau, with the
first-order (Loewner / divided-difference) correction of the PSD projection

    P_+(D + R) ~= D_+ + L o R,   L_ij = (max(d_i,0) - max(d_j,0)) / (d_i - d_j)  in [0, 1]

run inside the oracle's PGDB loop, compared with the exact run: final estimate, eigh / outer / cost counts, sweeps.
Decides whether the kernel may stop Jacobi at 1e-5..1e-6 instead of 1e-8.  Test infrastructure only.
usage: python scripts/proto/early_stop_cp.py <golden name> <tau> [correct=1] [items]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_numpy as orc


def rr_pairs(m, s):
    m1 = m - 1
    out = []
    for i in range(m // 2):
        p = (s + i) % m1
        q = (s - i) % m1
        if i == 0:
            q = m1
        out.append((p, q))
    return out


class WarmJacobi:
    def __init__(self, tau, correct):
        self.v = None
        self.tau, self.correct = tau, correct
        self.sweeps = 0
        self.calls = 0

    def eigh(self, x):
        m = x.shape[0]
        if self.v is None:
            self.v = np.eye(m, dtype=complex)
            tau = 1e-15  # cold start: converge fully
        else:
            tau = self.tau
        v = self.v
        a = v.conj().T @ x @ v
        a = (a + a.conj().T) / 2
        nsweep = 0
        while True:
            off = np.linalg.norm(a - np.diag(np.diag(a)))
            if off <= tau * np.linalg.norm(a) or nsweep >= 30:
                break
            for s in range(m - 1):
                j = np.eye(m, dtype=complex)
                for p, q in rr_pairs(m, s):
                    al, ga, be = a[p, p].real, a[q, q].real, a[p, q]
                    ab = abs(be)
                    if ab <= 1e-18 * (abs(al) + abs(ga)) or ab == 0:
                        continue
                    dlt = ga - al
                    r = np.sqrt(dlt * dlt + 4 * ab * ab)
                    c2 = 0.5 + 0.5 * abs(dlt) / r
                    c = np.sqrt(c2)
                    sg = (1.0 if dlt >= 0 else -1.0) / (r * c)
                    sn = be * sg
                    j[p, p] = c; j[q, q] = c; j[p, q] = sn; j[q, p] = -np.conj(sn)
                a = j.conj().T @ a @ j
                a = (a + a.conj().T) / 2
                v = v @ j
            nsweep += 1
        self.v = v
        self.sweeps += nsweep
        self.calls += 1
        return a, v

    def proj_cp(self, x):
        h = (x + x.conj().T) / 2
        a, v = self.eigh(h)
        d = np.diag(a).real
        dp = np.maximum(d, 0)
        core = np.diag(dp).astype(complex)
        if self.correct:
            r = a - np.diag(np.diag(a))
            di, dj = d[:, None], d[None, :]
            den = di - dj
            with np.errstate(divide="ignore", invalid="ignore"):
                l = (dp[:, None] - dp[None, :]) / den
            same = np.abs(den) < 1e-300
            l[same] = ((di > 0) & (dj > 0))[same].astype(float) if True else 0
            np.fill_diagonal(l, 0)
            core = core + l * r
        return v @ core @ v.conj().T


def run(name, tau, correct, items):
    from util import golden
    g = golden(name)
    n = int(g["n"])
    settings = [(tuple(int(c) for c in s), int(k)) for s, k in zip(g["state_codes"], g["pauli_idx"])]
    orig = orc.proj_choi_to_completely_positive
    for b in range(min(items, len(g["expectations"]))):
        wj = WarmJacobi(tau, correct)
        orc.proj_choi_to_completely_positive = wj.proj_cp
        t0 = time.time()
        try:
            est, c = orc.pgdb_process_estimate(settings, np.ones(len(settings)), g["expectations"][b], g["counts"][b], n,
                                               trace_preserving=bool(g["trace_preserving"]), return_counters=True)
        finally:
            orc.proj_choi_to_completely_positive = orig
        err = np.linalg.norm(est - g["choi_ref"][b]) / np.linalg.norm(g["choi_ref"][b])
        print(f"{name}[{b}] tau={tau:g} correct={correct}: err={err:.2e} eigh={c['eighs']} (ref {g['counters_ref'][b][0]}) "
              f"cost={c['cost_evals']} (ref {g['counters_ref'][b][1]}) outer={c['outer']} "
              f"sweeps/eigh={wj.sweeps / wj.calls:.2f} [{time.time() - t0:.0f}s]", flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    name, tau = sys.argv[1], float(sys.argv[2])
    correct = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    items = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    run(name, tau, correct, items)
