"""CPU test of the host mirror's control flow with the library calls stubbed out (the GPU tests cover the real thing):
catches wiring errors in psacore / psa_compute (argument plumbing, matrix cache, per-batch vectors) without a device."""
import numpy as np
import pytest

import psa_oracle as po


class _FakeCtx:
    """Stands in for PsaGpu: computes with the ORACLE (test infrastructure only -- this file is a test)."""

    def __init__(self):
        self.shape, self.calls, self._Z = None, [], None

    def set_thresholds(self, a, b):
        self.calls.append(("thr", a, b))

    def set_matrix(self, T, upper_wrapper=False):
        self.T = self.Tfull = np.array(T)
        self.shape = T.shape
        self.calls.append(("set", T.shape))

    def grid(self, x, y, q0, tol=1e-5, maxit=99, bigsigma=None, want_iters=True, cancel=None, row_batch=None, want_Z=True):
        import c_oracle as co
        q0 = np.atleast_2d(q0)
        Z = np.zeros((len(y), len(x)))
        for j in range(len(y)):
            b = 0 if row_batch is None else int(row_batch[j])
            Z[j] = co.grid(self.T, x, y[j:j + 1], q0[b], tol=tol, maxit=maxit)[0][0]
        self._Z = Z
        self.calls.append(("grid", len(x), len(y)))
        algo = "sq_lanc" if self.T.shape[0] == self.T.shape[1] else "HessQR"
        return (Z if want_Z else None), None, np.zeros(4, dtype=np.int64), algo

    def project(self, selection, want_T=True):
        from pseudospectra_jl_b200 import compute as hc
        self.calls.append(("project", len(selection)))
        if len(selection) == self.Tfull.shape[0]:
            self.T, self.shape = self.Tfull, self.Tfull.shape
            return None
        eig = np.diagonal(self.Tfull)
        class Th:  # selects exactly `selection`
            minnev = 0
        Tp = self.Tfull
        # reuse the host restatement of the swaps
        sel = np.asarray(selection)
        Tp = np.triu(np.array(self.Tfull, dtype=np.complex128))
        for i in range(len(sel)):
            for k in range(sel[i] - 1, i - 1, -1):
                c, sn = hc._givens(np.conj(Tp[k, k + 1]), np.conj(Tp[k, k] - Tp[k + 1, k + 1]))
                sn = -np.conj(sn)
                a1, a2 = Tp[:, k].copy(), Tp[:, k + 1].copy()
                Tp[:, k] = a1 * c + a2 * np.conj(sn)
                Tp[:, k + 1] = -a1 * sn + a2 * c
                r1, r2 = Tp[k, :].copy(), Tp[k + 1, :].copy()
                Tp[k, :] = c * r1 + sn * r2
                Tp[k + 1, :] = -np.conj(sn) * r1 + c * r2
        k = len(sel)
        self.T = np.asfortranarray(np.triu(Tp[:k, :k]))
        self.shape = (k, k)
        return self.T if want_T else None

    def postprocess(self, y, lx, n_mirror, ps_tiny, small_sigma=1e-150):
        Z, yf = po.unmirror(self._Z, np.asarray(y), n_mirror)
        Z = np.clip(Z, ps_tiny, np.inf)
        return Z, yf, (Z.min(), Z.max(), Z[Z >= small_sigma].min())


def test_psa_compute_plumbing_matches_oracle(pkg):
    T, eigA = po.complex_schur(po.grcar(70))
    ax = po.vec2ax(eigA)
    ctx = _FakeCtx()
    from pseudospectra_jl_b200 import compute as hc
    rb, nb = hc.row_batches(70, 12)
    qs = np.stack([po.start_vector(70, 20 + b) for b in range(nb)])
    out = pkg.psa_compute(T, 12, ax, eigA, {"real_matrix": True}, q0=qs, ctx=ctx)
    ref = po.psa_compute(T, 12, ax, eigA, {"real_matrix": True}, q_for_batch=lambda b, n: qs[b])
    assert out[7] == "sq_lanc" and np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2])
    ok = ref[0] > 1e-9
    assert np.allclose(out[0][ok], ref[0][ok], rtol=1e-9) and np.allclose(out[3], ref[3])
    n_set = sum(c[0] == "set" for c in ctx.calls)
    pkg.psa_compute(T, 6, [ax[0], ax[1], 0.5, ax[3]], eigA, {}, q0=qs[0], ctx=ctx)  # zoom: same matrix, no upload
    assert sum(c[0] == "set" for c in ctx.calls) == n_set


def test_psacore_hess_fast_path_plumbing(pkg):
    H = np.asfortranarray(po.grcar(230)[:, :-1]).astype(complex)
    q0 = po.start_vector(229, 1)
    x, y = np.linspace(-1, 3, 4), np.linspace(-2, 2, 3)
    c1, c2 = _FakeCtx(), _FakeCtx()
    Zg, algo_g, _ = pkg.psacore(H, None, q0, x, y, 2, ctx=c1)
    Zf, algo_f, _ = pkg.psacore(H, None, q0, x, y, 2, ctx=c2, hess_via_schur=True)
    assert algo_g == algo_f == "HessQR" and c2.shape == (229, 229) and c1.shape == (230, 229)
    ok = Zg > 1e-9
    assert np.allclose(Zf[ok], Zg[ok], rtol=1e-6)


def test_psa_compute_device_projection_plumbing(pkg):
    """proj_lev finite: the full matrix is uploaded once, projected on the `device`, the grid runs on the projected
    matrix with the TRUNCATED (not renormalised) start vector; a second call without projection restores the full one."""
    T, eigA = po.complex_schur(po.random_complex(200, 6))
    ax = [-4.0, 4.0, -4.0, 4.0]
    q0 = po.start_vector(200, 0)
    ctx = _FakeCtx()
    out = pkg.psa_compute(T, 6, ax, eigA, {"proj_lev": 1.0}, q0=q0, ctx=ctx)
    from pseudospectra_jl_b200 import compute as hc
    Tp, ep = hc._maybe_project(T, 1.0, ax, eigA)
    k = Tp.shape[0]
    assert 55 <= k < 200 and out[5].shape == (k, k) and np.allclose(out[5], Tp) and np.array_equal(out[6], ep)
    import c_oracle as co
    Zc = np.clip(co.grid(Tp, out[1], out[2], q0[:k])[0], po.PS_TINY, np.inf)
    assert np.allclose(out[0], Zc, rtol=1e-9)
    assert [c[0] for c in ctx.calls].count("set") == 1
    out2 = pkg.psa_compute(T, 5, ax, eigA, {}, q0=q0, ctx=ctx)   # no projection: full matrix again, no re-upload
    assert [c[0] for c in ctx.calls].count("set") == 1 and ctx.shape == (200, 200) and out2[5] is T
