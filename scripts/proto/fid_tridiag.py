"""Numpy prototype of the fidelity_tri_kernel arithmetic (dev tool): Cholesky of rho, Y = L^dagger sigma L, Householder
reduction of Y to a real symmetric tridiagonal (only d and e^2 kept), square-root-free QL (Pal-Walker-Kahan) for the
eigenvalues, fidelity = (sum sqrt(max(ev, 0)))^2.  Checks the formulas the CUDA kernel uses against numpy's eigvalsh."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ref_numpy as orc

EPS2 = (2.0 ** -53) ** 2


def tridiag(Y):
    D = Y.shape[0]
    A = Y.copy()
    e2 = np.zeros(D - 1)
    for k in range(D - 2):
        x = A[:, k].copy()
        x[:k + 1] = 0
        n2 = np.sum(np.abs(x) ** 2)
        e2[k] = n2
        alpha = x[k + 1]
        aa = abs(alpha) ** 2
        xn, an = np.sqrt(n2), np.sqrt(aa)
        ph = alpha / an if aa > 0 else 1.0
        u = x.copy()
        u[k + 1] = ph * (an + xn)
        beta = 1.0 / (xn * (xn + an)) if n2 > 0 else 0.0
        p = beta * (A @ u)
        K = 0.5 * beta * np.real(np.vdot(u, p))
        w = p - K * u
        A = A - np.outer(u, w.conj()) - np.outer(w, u.conj())
    e2[D - 2] = abs(A[D - 1, D - 2]) ** 2
    return np.real(np.diag(A)).copy(), e2


def pwk_ql(d, e2):
    d, e = d.copy(), np.append(e2, 0.0)
    D = len(d)
    its = 0
    for l in range(D):
        it = 0
        while True:
            m = l
            while m < D - 1 and not (e[m] <= EPS2 * abs(d[m] * d[m + 1])):
                m += 1
            if m == l:
                break
            it += 1
            its += 1
            assert it < 40
            p = d[l]
            rte = np.sqrt(e[l])
            sg = (d[l + 1] - p) / (2 * rte)
            rr = np.sqrt(sg * sg + 1)
            sigma = p - rte / (sg + np.copysign(rr, sg))
            c, s, gamma = 1.0, 0.0, d[m] - sigma
            p = gamma * gamma
            for i in range(m - 1, l - 1, -1):
                bb = e[i]
                r = p + bb
                if i != m - 1:
                    e[i + 1] = s * r
                oldc = c
                c, s = p / r, bb / r
                oldgam = gamma
                al = d[i]
                gamma = c * (al - sigma) - s * oldgam
                d[i + 1] = oldgam + (al - gamma)
                p = gamma * gamma / c if c != 0 else oldc * bb
            e[l] = s * p
            d[l] = sigma + gamma
    return d, its


def fid(rho, sig):
    L = np.linalg.cholesky(rho)
    Y = L.conj().T @ sig @ L
    Y = (Y + Y.conj().T) / 2
    d, e2 = tridiag(Y)
    ev, its = pwk_ql(d, e2)
    ref = np.linalg.eigvalsh(Y)
    return np.sum(np.sqrt(np.maximum(ev, 0))) ** 2, np.max(np.abs(np.sort(ev) - ref)), its


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for D in (4, 8, 16):
        worst, worst_ev, worst_r1, its_all = 0, 0, 0, []
        for b in range(300):
            rho = orc.ginibre_state(rng, D)
            r1 = b % 3 == 0
            sig = orc.ginibre_state(rng, D, rank=1 if r1 else (2 if b % 3 == 1 else None))
            f, dev, its = fid(rho, sig)
            want = np.real(orc.fidelity(rho, sig))
            its_all.append(its)
            err = abs(f - want) / max(want, 1e-3)
            if b % 3 == 2:
                worst = max(worst, err)
            else:
                worst_r1 = max(worst_r1, err)
            worst_ev = max(worst_ev, dev)
        print(f"D={D}: rel err full rank {worst:.2e}, rank-deficient sigma {worst_r1:.2e}, eigenvalue err {worst_ev:.2e}, "
              f"QL iterations mean {np.mean(its_all):.1f} max {np.max(its_all)}")
