"""CPU tests: the oracle against dense SVD and the committed golden vectors, the C restatement against the numpy
one, and the reference's grid-construction rules (src/compute.jl:122-145,270-279; src/utils.jl:112-141,175-191)."""
import math
import numpy as np
import pytest

import c_oracle as co
import psa_oracle as po


def test_step_size_matches_reference_table():
    # src/compute.jl:270-279; SURVEY §8a quotes step=2 (C1), 1 (C2, C3, C5), 9 (C4: m=201 -> 4*500//201)
    assert po.get_step_size(150, 100) == 2
    assert po.get_step_size(1000, 103) == 1
    assert po.get_step_size(4000, 400) == 1
    assert po.get_step_size(201, 500) == 9
    assert po.get_step_size(60, 100) == 12  # n < 100 branch: ly // 8
    assert po.get_step_size(32, 4) == 1


def test_vec2ax_readme_matrix():
    T, eigA = po.complex_schur(po.readme_matrix(150))
    ax = po.vec2ax(eigA)
    assert np.allclose(ax, [-12.5, 11.0, -9.2, 11.6], atol=1e-12)  # SURVEY §8d probe


def test_tidyaxes():
    lo, hi = po.tidyaxes(-1.2345, 2.3456, 0.02)
    assert lo <= -1.2345 and hi >= 2.3456 and hi - lo < 3.8


@pytest.mark.parametrize("ax,npts", [([-1.44, 3.24, -3.8, 3.8], 20), ([-1, 2, -3, 4], 21), ([-1, 2, -4, 3], 20),
                                      ([-2, 2, -2, 2], 15), ([-1, 1, -0.5, 3], 16)])
def test_shift_axes_mirror_reconstructs_symmetric_grid(ax, npts):
    """The half grid + un-mirroring (compute.jl:222-235) must give a y axis symmetric about 0 where mirrored and a Z
    equal to evaluating f(|y|)."""
    y, nm = po.shift_axes(ax, npts)
    assert np.all(np.diff(y) > 0)
    Z = np.abs(y)[:, None] * np.ones((1, 3))
    Zf, yf = po.unmirror(Z, y, nm)
    assert np.all(np.diff(yf) > 0)
    assert np.allclose(Zf[:, 0], np.abs(yf))
    assert yf[0] >= ax[2] - 1e-9 * (1 + abs(ax[2])) or nm >= 0
    k = abs(nm)
    assert len(yf) == len(y) + (k if nm < 0 or y[0] != 0 else min(k, len(y) - 1))


def test_make_grid_complex_matrix_is_plain_linspace():
    x, y, nm = po.make_grid([-1, 2, -3, 4], 17, False)
    assert nm == 0 and len(x) == 17 and len(y) == 17 and x[0] == -1 and y[-1] == 4


def test_oracle_matches_dense_svd_readme(golden):
    g = golden("readme150")
    Z, svd = g["Z"], g["svd"]
    rel = np.abs(Z - svd) / svd
    assert np.median(rel) < 5e-6 and rel.max() < 1e-4  # Lanczos stops at tol 1e-5 (median 5e-7, max 7e-6)


def test_oracle_reproduces_golden(golden):
    g = golden("readme150")
    Z, algo, warned, it = po.psacore(g["T"], g["q0"], g["x"], g["y"], 1, want_iters=True)
    assert algo == "sq_lanc" == str(g["algo"])
    assert np.allclose(Z, g["Z"], rtol=1e-9) and np.array_equal(it, g["iters"])


def test_c_oracle_matches_numpy_oracle_square(golden):
    g = golden("readme150")
    Zc, itc, fl = co.grid(g["T"], g["x"], g["y"], g["q0"], nthreads=4)
    assert np.allclose(Zc, g["Z"], rtol=1e-10)
    assert np.array_equal(itc, g["iters"]) and fl.sum() == 0


def test_c_oracle_matches_numpy_oracle_hessqr(golden):
    g = golden("hessqr210")
    assert str(g["algo"]) == "HessQR"  # test/medrect.jl:16
    Zc, itc, fl = co.grid(g["H"], g["x"], g["y"], g["q0"], nthreads=4)
    ok = g["Z"] > 1e-13
    assert np.allclose(Zc[ok], g["Z"][ok], rtol=1e-8)
    assert np.mean(itc == g["iters"]) > 0.97


def test_hessqr_drops_last_subdiagonal_entry():
    """Reference quirk (compute.jl:581,592): the sweep never rotates H[n+1,n] in, so changing it changes nothing."""
    H = np.asfortranarray(po.grcar(210)[:, :-1]).astype(np.complex128)
    H2 = H.copy()
    H2[209, 208] = 123.0
    q0 = po.start_vector(209, 1)
    x, y = np.array([0.5]), np.array([1.0])
    assert po.psacore(H, q0, x, y, 2)[0][0, 0] == po.psacore(H2, q0, x, y, 2)[0][0, 0]


def test_svd_branch_and_rect_qr_algo_symbols():
    rng = np.random.default_rng(0)
    T = np.triu(rng.random((32, 32)))  # test/runtests.jl:8-13 smoke case
    Z, algo, _ = po.psacore(T, po.start_vector(32), np.linspace(0, 1, 3), np.linspace(0, 1, 3), 1)
    assert algo == "SVD" and np.all(np.isfinite(Z))
    A = po.grcar(190)[:, :-1]  # test/smrect.jl: m <= 200 -> rect_qr
    Z, algo, _ = po.psacore(A, po.start_vector(189), np.array([0.5]), np.array([1.0]), 2)
    assert algo == "rect_qr"


def test_nonfinite_recurrence_maps_to_bigsigma():
    """Defined behaviour where the reference has none (SURVEY §7 'Overflow'): z on an eigenvalue -> bigsigma."""
    T = np.asfortranarray(np.triu(np.ones((60, 60), dtype=np.complex128)))
    sig, conv, fail, it = po.psa_lanczos(T - np.eye(60), po.start_vector(60), 1e-5, 99, po.BIGSIGMA)
    assert fail and sig == po.BIGSIGMA
    Zc, itc, fl = co.grid(T, np.array([1.0]), np.array([0.0]), po.start_vector(60))
    assert fl[0, 0] & 2 and Zc[0, 0] < po.PS_TINY


def test_psa_compute_real_matrix_mirror_golden(golden):
    g = golden("grcar120_real")
    out = po.psa_compute(g["T"], 24, list(g["ax"]), np.diagonal(g["T"]), {"real_matrix": True},
                         q_for_batch=lambda b, n: g["q0"])
    ok = g["Z"] > 1e-10
    assert np.allclose(out[0][ok], g["Z"][ok], rtol=1e-8) and np.allclose(out[2], g["y"])
    assert np.allclose(out[2], -out[2][::-1])  # symmetric axes here: mirrored y is antisymmetric
    assert np.allclose(out[0][ok], out[0][::-1][ok], rtol=1e-12)


def test_recalc_levels_basic():
    Z = np.logspace(-6, 0, 49).reshape(7, 7)
    lev, err = po.recalc_levels(Z, [-1, 1, -1, 1])
    assert err == 0 and len(lev) >= 2 and np.all(np.diff(lev) > 0) and lev[-1] <= np.log10(0.03 * 2) + 1e-9
    lev, err = po.recalc_levels(np.full((3, 3), 1e-200), [-1, 1, -1, 1])
    assert err == -2


@pytest.mark.parametrize("name", ["grcar", "toeplitz", "random"])
def test_oracle_no_worse_than_tolerance_vs_svd_across_matrix_classes(name):
    """Pins the oracle (both restatements) against dense svdvals on sampled points for the BASELINE matrix families;
    the Lanczos stopping rule (tol 1e-5 on sigma = 1/sigma_min^2) bounds the error wherever sigma_min is above the
    rounding floor n*eps*||T||."""
    if name == "grcar":
        A = po.grcar(80)
    elif name == "toeplitz":
        A = po.toeplitz_from_symbol([0.0, 1.0], [0.0, 0.0, 0.25], 90)  # examples/toeplitz.jl 'triangle'
    else:
        A = po.random_complex(70, 11)
    T, eigA = po.complex_schur(A)
    n = T.shape[0]
    x, y, _ = po.make_grid(po.vec2ax(eigA), 9, False)
    q0 = po.start_vector(n, 3)
    Z, algo, warned, it = po.psacore(T, q0, x, y, 1, want_iters=True)
    Zc, itc, fl = co.grid(T, x, y, q0, nthreads=4)
    floor = 1e-9 * np.linalg.norm(T)
    worst = 0.0
    for j in range(len(y)):
        for k in range(len(x)):
            s = po.sigmin_svd(T, x[k] + 1j * y[j])
            if s > floor:
                worst = max(worst, abs(Z[j, k] - s) / s, abs(Zc[j, k] - s) / s)
    assert worst < 2e-4
    ok = Z > floor
    assert np.allclose(Zc[ok], Z[ok], rtol=1e-9)


def test_shift_axes_property_based():
    """hypothesis: for any window straddling the real axis the mirrored y axis is strictly increasing, symmetric where
    it was mirrored and inside the window."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(lo=st.floats(-50.0, -0.01), hi=st.floats(0.01, 50.0), npts=st.integers(6, 300))
    def check(lo, hi, npts):
        ax = [-1.0, 1.0, lo, hi]
        y, nm = po.shift_axes(ax, npts)
        if len(y) < 2:
            return
        assert np.all(np.diff(y) > 0)
        Z = np.abs(y)[:, None]
        Zf, yf = po.unmirror(Z, y, nm)
        assert np.all(np.diff(yf) > -1e-12 * (abs(hi) + abs(lo)))
        assert np.allclose(Zf[:, 0], np.abs(yf))
        assert yf[0] >= lo - 1e-9 * (hi - lo) and yf[-1] <= hi + 1e-9 * (hi - lo)

    check()


def test_reference_medcplx_input_selects_sq_lanc():
    """test/medcplx.jl:7-14 restated: landau_fox_li(225, 16) is complex dense 225 x 225 -> `algo == :sq_lanc`; its
    spectrum lies in the unit disc (Landau 1977), which pins the generator."""
    A = po.landau_fox_li(225, 16)
    assert A.shape == (225, 225) and np.iscomplexobj(A)
    T, eigA = po.complex_schur(A)
    assert np.abs(eigA).max() < 1.0 + 1e-10
    Z, algo, _ = po.psacore(T, po.start_vector(225), np.array([0.0, 0.6]), np.array([1.0]), 1)
    assert algo == "sq_lanc" and np.all(np.isfinite(Z))


@pytest.mark.parametrize("n", [20, 60])
def test_reference_threads_jl_self_consistency(n):
    """test/threads.jl:7-19 restated: two psa_compute runs on the same Schur factor (fresh random start vectors each,
    compute.jl:207) agree to norm(Z0 - Z1)/norm(Z0) < 1e-4 -- the ONLY numerical assertion the reference's test-suite
    makes at this boundary.  Here: numpy restatement (per-batch vectors) vs C restatement (another vector)."""
    rng = np.random.default_rng(n)
    T, eigA = po.complex_schur(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    ax, npts = [-3, 3, -3, 3], 100 if n == 20 else 40  # (n = 60 on a coarser grid: the numpy loop is slow)
    Z0 = po.psa_compute(T, npts, ax, eigA, {})[0]
    x, y, _ = po.make_grid(ax, npts, False)
    if n < 55:
        Z1 = po.psacore(T, po.start_vector(n, 99), x, y, 1)[0]
    else:
        Z1 = co.grid(T, x, y, po.start_vector(n, 99), nthreads=4)[0]
    Z1 = np.clip(Z1, po.PS_TINY, np.inf)
    assert np.linalg.norm(Z0 - Z1) / np.linalg.norm(Z0) < 1e-4


# ---- known answers the reference's own test-suite holds (test/radius_abscissa.jl:20-28) -----------------------------
# Grcar(100): top eigenvalue 2.2617668 (atol 1e-6); eps = 1e-4 pseudospectral abscissa 2.41276 and radius 2.85216
# (Guglielmi & Overton 2012, atol 1e-5).  By definition alpha_eps = max{Re z : sigma_min(zI - A) <= eps}, so on the
# vertical line Re z = alpha the minimum of sigma_min is exactly eps, and moving the line by the test's own tolerance
# must cross the level: min sigma(alpha - 1e-5) < eps < min sigma(alpha + 1e-5).  Same for the circle |z| = rho.
# These are published numbers about the very quantity this path computes -- the only ones the reference holds.
GO_EPS, GO_ALPHA, GO_RHO, GO_TOPEIG = 1e-4, 2.41276, 2.85216, 2.2617668


def go_points():
    """Points on the abscissa line (the level set touches it on the real axis) and on the radius circle (around the
    touching angle, located with dense SVDs), each at the published value and at +-1e-5 (the reference's atol)."""
    A = po.grcar(100)
    th = np.linspace(1.3272 - 0.002, 1.3272 + 0.002, 41)
    smin = [po.sigmin_svd(A, GO_RHO * np.exp(1j * t)) for t in th]
    t0 = th[int(np.argmin(smin))]
    assert 0 < int(np.argmin(smin)) < 40  # an interior minimum: the touching point is inside the bracket
    line = {d: np.array([GO_ALPHA + d + 1j * yy for yy in np.linspace(0.0, 0.004, 5)]) for d in (-1e-5, 0.0, 1e-5)}
    circ = {d: np.array([(GO_RHO + d) * np.exp(1j * t) for t in np.linspace(t0 - 2e-4, t0 + 2e-4, 5)]) for d in (-1e-5, 0.0, 1e-5)}
    return A, line, circ


def check_go(sig_of):
    """sig_of(zs) -> sigma_min at each z (any implementation under test)."""
    A, line, circ = go_points()
    for name, fam in (("abscissa", line), ("radius", circ)):
        lo, mid, hi = (float(np.min(sig_of(fam[d]))) for d in (-1e-5, 0.0, 1e-5))
        assert lo < GO_EPS < hi, (name, lo, hi)                 # the published value is right to its stated 1e-5
        assert abs(mid / GO_EPS - 1) < 3e-4, (name, mid)         # and sigma_min there is eps to 3e-4 relative


def test_grcar100_published_values_pin_the_svd_checker():
    A = po.grcar(100)
    assert abs(np.linalg.eigvals(A).imag.max() - GO_TOPEIG) < 1e-6  # test/radius_abscissa.jl:20
    check_go(lambda zs: np.array([po.sigmin_svd(A, z) for z in zs]))


def test_grcar100_published_values_pin_the_lanczos_oracles():
    """Both restatements of _psa_lanczos! (numpy + LAPACK, plain C) on the Schur factor, as the reference computes it."""
    T, eigA = po.complex_schur(po.grcar(100))
    q0 = po.start_vector(100, 0)

    def numpy_oracle(zs):
        return np.array([po.psacore(T, q0, np.array([z.real]), np.array([z.imag]), 1)[0][0, 0] for z in zs])

    def c_oracle(zs):
        return np.array([co.grid(T, np.array([z.real]), np.array([z.imag]), q0)[0][0, 0] for z in zs])

    check_go(numpy_oracle)
    check_go(c_oracle)


# ---- closed-form known answers (independent of any implementation) --------------------------------------------------
def closed_form_cases(n):
    """(name, upper-triangular T, exact sigma_min(T - zI) as a function of z) for two families with a closed form:
    a diagonal (normal) matrix -- sigma_min = distance to the spectrum -- and lambda*I plus one corner entry alpha,
    whose only non-trivial singular pair comes from the 2x2 block [[mu, alpha], [0, mu]], mu = lambda - z:
    sigma_min^2 = (2|mu|^2 + |alpha|^2 - sqrt((2|mu|^2 + |alpha|^2)^2 - 4|mu|^4)) / 2."""
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    lam, alpha = 0.3 - 0.2j, 2.5 + 1.0j
    C = lam * np.eye(n, dtype=complex)
    C[0, n - 1] = alpha

    def corner(z):
        m2, a2 = abs(lam - z) ** 2, abs(alpha) ** 2
        s = 2 * m2 + a2
        return math.sqrt(2 * m2 * m2 / (s + math.sqrt(s * s - 4 * m2 * m2)))  # the same root, without cancellation

    return [("diagonal", np.diag(d), lambda z: float(np.min(np.abs(d - z)))), ("corner", C, corner)]


CLOSED_FORM_Z = [0.9 + 0.4j, -1.3 + 0.2j, 0.1 - 2.0j, 2.2 + 1.7j, 0.31 - 0.18j]


@pytest.mark.parametrize("n", [60, 130])
def test_oracles_reproduce_closed_form_sigma_min(n):
    """Lanczos stops on a 1e-5 relative change of sigma_max of the inverse (compute.jl:658-661), so sigma_min is good to
    that order; both restatements must land on the closed form."""
    q0 = po.start_vector(n, 3)
    for name, T, exact in closed_form_cases(n):
        T = np.asfortranarray(T)
        for z in CLOSED_FORM_Z:
            ref = exact(z)
            zn = po.psacore(T, q0, np.array([z.real]), np.array([z.imag]), 1)[0][0, 0]
            zc = co.grid(T, np.array([z.real]), np.array([z.imag]), q0)[0][0, 0]
            assert abs(zn / ref - 1) < 1e-4 and abs(zc / ref - 1) < 1e-4, (name, z, zn, zc, ref)
            assert abs(po.sigmin_svd(T, z) / ref - 1) < 1e-10, (name, z)
