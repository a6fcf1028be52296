"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle on
identical (T, x, y, q0) and against the committed golden vectors.  Tolerance: sigma_min within 1e-6 relative of the
oracle (north_star); iteration counts identical up to rounding-level ties at the stopping threshold."""
import math
import os

import numpy as np
import pytest
import scipy.linalg as sla

import c_oracle as co
import psa_oracle as po

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # north_star: sigma_min within 1e-6 relative


@pytest.fixture(scope="module")
def gpu(pkg):
    g = pkg.PsaGpu(1)
    yield g
    g.close()


def _floor(T):
    """sigma_min is only defined to about n*eps*||T||; the 1e-6 relative bar applies above this floor."""
    return 1e-9 * float(np.linalg.norm(T))


def _rel(Z, Zref, floor=1e-13):
    ok = Zref > floor
    return float(np.max(np.abs(Z[ok] - Zref[ok]) / Zref[ok])) if ok.any() else 0.0


@pytest.mark.parametrize("n,ns", [(64, 1), (100, 64), (150, 70), (192, 5), (333, 130)])
def test_solve_kernel_matches_triangular_solves(gpu, n, ns):
    """One application of v = F1 \\ (F1' \\ q) (compute.jl:630) for ns shifts at once vs LAPACK ztrtrs."""
    rng = np.random.default_rng(n + ns)
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + np.diag(3 * rng.standard_normal(n))
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    gpu.set_matrix(T)
    V = gpu.apply_resolvent(z, Q)
    for s in range(ns):
        F = T - z[s] * np.eye(n)
        v = sla.solve_triangular(F, sla.solve_triangular(F, Q[s], trans="C"))
        assert np.linalg.norm(V[s] - v) <= 1e-11 * np.linalg.norm(v)


def test_solve_kernel_is_linear_and_ignores_strict_lower_part(gpu):
    rng = np.random.default_rng(7)
    n, ns = 200, 16
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))  # full matrix: i > j must be ignored
    z = 4 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q1 = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    Q2 = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    gpu.set_matrix(A, upper_wrapper=True)  # parent array of an UpperTriangular wrapper: i > j ignored, explicitly
    V1, V2, V12 = gpu.apply_resolvent(z, Q1), gpu.apply_resolvent(z, Q2), gpu.apply_resolvent(z, Q1 + 2j * Q2)
    assert np.allclose(V12, V1 + 2j * V2, rtol=1e-10, atol=1e-12 * np.abs(V12).max())
    gpu.set_matrix(np.triu(A))
    assert np.array_equal(gpu.apply_resolvent(z, Q1), V1)  # UpperTriangular semantics, bit-exact


def test_readme150_full_grid_vs_oracle(gpu):
    """BASELINE config 1: README 150x150 banded matrix, default 100x100 grid, sq_lanc."""
    T, eigA = po.complex_schur(po.readme_matrix(150))
    x, y, _ = po.make_grid(po.vec2ax(eigA), 100, False)
    q0 = po.start_vector(150, 0)
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, flc = co.grid(T, x, y, q0, nthreads=8)
    assert algo == "sq_lanc"
    assert _rel(Z, Zc) < RTOL
    assert np.mean(it == itc) > 0.999
    assert counts[3] == 10000 and counts[2] == it.sum() and counts[0] == 0 and counts[1] == 0


def test_golden_readme_and_svd(gpu, golden):
    g = golden("readme150")
    gpu.set_matrix(g["T"])
    Z, it, counts, algo = gpu.grid(g["x"], g["y"], g["q0"])
    assert _rel(Z, g["Z"]) < RTOL and np.array_equal(it, g["iters"])
    # no worse than the reference algorithm versus svdvals on sampled points
    err_gpu = np.abs(Z - g["svd"]) / g["svd"]
    err_ref = np.abs(g["Z"] - g["svd"]) / g["svd"]
    assert err_gpu.max() <= err_ref.max() * 1.01 + 1e-9


def test_real_matrix_mirror_through_psa_compute(pkg, golden):
    """Real matrix -> shift_axes half grid + un-mirror + clamp; same tuple as the reference's psa_compute."""
    g = golden("grcar120_real")
    out = pkg.psa_compute(g["T"], 24, list(g["ax"]), np.diagonal(g["T"]), {"real_matrix": True}, q0=g["q0"])
    Z, x, y, levels, err, Tproj, eigAproj, algo = out
    assert algo == "sq_lanc" and err == 0 and Z.shape == g["Z"].shape
    assert np.array_equal(x, g["x"]) and np.array_equal(y, g["y"])  # grid coordinates exact
    assert _rel(Z, g["Z"], _floor(g["T"])) < RTOL
    assert np.allclose(levels, g["levels"])
    assert Z.min() >= 10 * math.sqrt(np.finfo(float).tiny)


def test_projection_option_through_psa_compute(pkg):
    """opts[:proj_lev] (compute.jl:116,147-153): the host-side Schur reordering shrinks T, the GPU path then runs on
    the projected matrix; compared with the oracle's Lanczos on the same projected T."""
    T, eigA = po.complex_schur(po.random_complex(200, 6))
    ax = [-4.0, 4.0, -4.0, 4.0]
    q0 = po.start_vector(200, 0)
    out = pkg.psa_compute(T, 12, ax, eigA, {"proj_lev": 1.0}, q0=q0)
    Z, x, y, levels, err, Tproj, eigAproj, algo = out
    k = Tproj.shape[0]
    assert 55 <= k < 200 and len(eigAproj) == k and algo == "sq_lanc"
    qk = q0[:k]  # `q = q0[1:n]` (compute.jl:622): truncated, not renormalised
    Zc, _, _ = co.grid(Tproj, x, y, qk, nthreads=4)
    assert _rel(Z, np.clip(Zc, po.PS_TINY, np.inf), _floor(Tproj)) < RTOL


def test_hessqr_golden_medrect(gpu, golden):
    """test/medrect.jl: grcar(210)[:,1:end-1] -> :HessQR."""
    g = golden("hessqr210")
    gpu.set_matrix(g["H"])
    Z, it, counts, algo = gpu.grid(g["x"], g["y"], g["q0"])
    assert algo == "HessQR"
    assert _rel(Z, g["Z"], _floor(g["H"])) < RTOL
    assert np.mean(it == g["iters"]) > 0.97


def test_hessqr_arnoldi_projection_vs_oracle(gpu):
    """BASELINE config 4 in miniature: convdiff_fd operator projected by Arnoldi -> (m+1) x m Hessenberg."""
    A = po.convdiff_fd(6)
    H = po.arnoldi_hessenberg(A, 208, seed=0)
    ev = np.linalg.eigvals(H[:208, :])
    x, y, _ = po.make_grid(po.vec2ax(ev), 14, False)
    q0 = po.start_vector(208, 2)
    gpu.set_matrix(H)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, flc = co.grid(np.asarray(H, dtype=complex), x, y, q0, nthreads=8)
    assert algo == "HessQR" and _rel(Z, Zc, _floor(H)) < RTOL and np.mean(it == itc) > 0.97


@pytest.mark.parametrize("n", [257, 300, 600, 1000, 1024])
def test_hessqr_large_n_streaming_kernels(gpu, n, monkeypatch):
    """256 < n <= 1024: warp-per-slot streaming kernel with 16 / 32 vector elements per lane (hess_solve_big_kernel);
    PSA_HESS_CTA=1 selects the CTA-per-slot kernel (different summation order: agreement to rounding, not bit for bit)."""
    rng = np.random.default_rng(11)
    H = np.triu(rng.standard_normal((n + 1, n)) + 1j * rng.standard_normal((n + 1, n)), -1) / np.sqrt(n)
    H[np.arange(1, n + 1), np.arange(n)] = 1.0           # spectrum of the square part ~ unit disc: the window sees real sigma_min
    x, y = np.linspace(-1.5, 1.5, 5), np.linspace(-1.5, 1.5, 4)
    q0 = po.start_vector(n, 1)
    gpu.set_matrix(np.asfortranarray(H))
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, _ = co.grid(H, x, y, q0, nthreads=8)
    assert np.sum(Zc > _floor(H)) >= 10
    assert algo == "HessQR" and _rel(Z, Zc, _floor(H)) < RTOL and np.mean(it == itc) > 0.9
    monkeypatch.setenv("PSA_HESS_CTA", "1")
    Z2, it2, _, _ = gpu.grid(x, y, q0)
    assert _rel(Z2, Z, _floor(H)) < 1e-9 and np.mean(it2 == it) > 0.9


def test_random_complex_600_vs_oracle_and_svd(gpu):
    T, eigA = po.complex_schur(po.random_complex(600, 1))
    x, y, _ = po.make_grid(po.vec2ax(eigA), 40, False)
    q0 = po.start_vector(600, 0)
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, _ = co.grid(T, x, y, q0, nthreads=8)
    assert _rel(Z, Zc) < RTOL and np.mean(it == itc) > 0.999
    for (j, k) in [(0, 0), (20, 20), (39, 5), (7, 31)]:
        s = po.sigmin_svd(T, x[k] + 1j * y[j])
        assert abs(Z[j, k] - s) / s <= abs(Zc[j, k] - s) / s * 1.01 + 1e-9


def test_overflow_points_use_bigsigma_like_the_oracle(gpu):
    """Strongly non-normal input: the solves overflow near the spectrum; both sides map that to bigsigma (then the
    clamp to ps_tiny), and agree to 1e-6 wherever sigma_min is above rounding level."""
    T, eigA = po.complex_schur(po.grcar(400))
    x = np.linspace(0.2, 2.0, 10)
    y = np.linspace(0.1, 2.5, 9)
    q0 = po.start_vector(400, 0)
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, flc = co.grid(T, x, y, q0, nthreads=8)
    assert _rel(Z, Zc, _floor(T)) < RTOL
    fb_gpu, fb_cpu = Z < po.PS_TINY, (flc & 2) != 0
    assert np.array_equal(fb_gpu, fb_cpu)
    assert counts[1] == fb_cpu.sum()


def test_edge_shapes_and_reuse(gpu):
    """1x1 grid, ragged panel counts, n a multiple of the tile and not, matrix reuse across calls (zoom)."""
    for n in (55, 64, 65, 128, 191):
        T, _ = po.complex_schur(po.random_complex(n, n))
        q0 = po.start_vector(n, 1)
        gpu.set_matrix(T)
        for lx, ly in ((1, 1), (65, 1), (3, 43)):
            x, y = np.linspace(-1.5, 2.0, lx), np.linspace(-0.7, 1.1, ly)
            Z, it, counts, algo = gpu.grid(x, y, q0)
            Zc, itc, _ = co.grid(T, x, y, q0, nthreads=4)
            assert Z.shape == (ly, lx) and _rel(Z, Zc) < RTOL and counts[3] == lx * ly
        # zoom: second grid on the resident matrix
        Z2, _, _, _ = gpu.grid(np.linspace(0.0, 0.5, 7), np.linspace(0.0, 0.5, 5), q0)
        Zc2, _, _ = co.grid(T, np.linspace(0.0, 0.5, 7), np.linspace(0.0, 0.5, 5), q0)
        assert _rel(Z2, Zc2) < RTOL


def test_svd_branch_small_matrices(gpu):
    """n < 55 -> :SVD, Z = sigma_min directly (compute.jl:510-518); incl. the reference's own smoke input
    (test/runtests.jl:8-13: 32x32 real triu(rand)), a complex dense, an odd n and a rectangular case."""
    rng = np.random.default_rng(3)
    cases = [np.triu(rng.random((32, 32))), rng.standard_normal((40, 40)) + 1j * rng.standard_normal((40, 40)),
             rng.standard_normal((17, 17)) + 0j, rng.standard_normal((60, 50)) + 1j * rng.standard_normal((60, 50)),
             po.grcar(54)]
    x, y = np.linspace(-1.5, 2.5, 9), np.linspace(-2.0, 2.0, 7)
    for A in cases:
        gpu.set_matrix(A)
        Z, it, counts, algo = gpu.grid(x, y, np.ones(A.shape[1]))
        Zo, algo_o, _ = po.psacore(A, np.ones(A.shape[1]), x, y, A.shape[0] - A.shape[1] + 1)
        assert algo == "SVD" == algo_o and counts[3] == 63 and it.max() == 0
        nrm = np.linalg.norm(A, 2)
        assert np.max(np.abs(Z - Zo)) <= 1e-12 * nrm          # absolute accuracy of any backward-stable SVD
        ok = Zo > 1e-8 * nrm
        assert _rel(Z[ok], Zo[ok], 0.0) < RTOL


def test_rect_qr_small_hessenberg(gpu):
    """test/smrect.jl: grcar(190)[:,1:end-1] (m <= 200, Hessenberg) -> :rect_qr (compute.jl:588-590)."""
    A = np.asfortranarray(po.grcar(190)[:, :-1]).astype(complex)
    x, y = np.linspace(-1.0, 3.0, 8), np.linspace(-3.0, 3.0, 7)
    q0 = po.start_vector(189, 5)
    gpu.set_matrix(A)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zo, algo_o, _, ito = po.psacore(A, q0, x, y, 2, want_iters=True)
    assert algo == "rect_qr" == algo_o
    assert _rel(Z, Zo, _floor(A)) < RTOL and np.mean(it == ito) > 0.95
    # ARPACK-default-sized projection (ncv = 60 -> 61 x 60 Hessenberg), test/swmethod.jl:24
    H = po.arnoldi_hessenberg(po.convdiff_fd(4), 60, seed=1)
    ev = np.linalg.eigvals(H[:60, :])
    xs, ys, _ = po.make_grid(po.vec2ax(ev), 10, False)
    q1 = po.start_vector(60, 2)
    gpu.set_matrix(H)
    Z, it, counts, algo = gpu.grid(xs, ys, q1)
    Zo, algo_o, _, ito = po.psacore(H, q1, xs, ys, 2, want_iters=True)
    assert algo == "rect_qr" == algo_o and _rel(Z, Zo, _floor(H)) < RTOL


def test_plain_c_driver_matches_oracle(pkg, tmp_path):
    """The C ABI driven from plain C (tests/abi_driver.c), the way a ccall host uses it."""
    import subprocess
    from test_host import _build_abi_driver
    n, lx, ly = 96, 4, 3
    r = subprocess.run([_build_abi_driver(pkg, tmp_path), str(n), str(lx), str(ly)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    T = np.where(i < j, 0.3 * np.sin(1.0 + i + 2.0 * j) + 0.3j * np.cos(2.0 + 3.0 * i + j), 0)
    T = T + np.diag(0.05 * np.arange(n) - 1.0 + 0.02j * np.arange(n))
    x = np.array([-2.0 + 4.0 * k / (lx - 1) for k in range(lx)])
    y = np.array([-1.0 + 3.0 * jj / (ly - 1) for jj in range(ly)])
    q0 = np.cos(0.7 * np.arange(n)) + 1j * np.sin(1.3 * np.arange(n) + 0.2)
    q0 = q0 / np.linalg.norm(q0)
    Zc, itc, _ = co.grid(T, x, y, q0)
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("Z ")]
    assert len(lines) == lx * ly and r.stdout.startswith("algo 1 ")
    for _, jj, kk, val, its in lines:
        assert abs(float(val) - Zc[int(jj), int(kk)]) <= RTOL * Zc[int(jj), int(kk)] and int(its) == itc[int(jj), int(kk)]


def test_pseudomode_inverse_lanczos(pkg):
    """psmode_inv_lanczos (modes.jl:185-274): sigma_min and the pseudomode at one point through the
    reference-shaped Python mirror; compared with the oracle's restatement and with the defining residual."""
    T, eigA = po.complex_schur(po.random_complex(320, 9))
    q0 = po.start_vector(320, 4)
    for z in (2.0 + 1.5j, -25.0 + 3.0j):
        Zg, qg = pkg.psmode_inv_lanczos(T, q0, z, 1e-10, 60)
        Zo, qo, l = po.psmode_inv_lanczos(T, q0, z, 1e-10, 60)
        assert abs(Zg - Zo) <= 1e-8 * Zo
        ph = np.vdot(qo, qg) / abs(np.vdot(qo, qg))
        assert np.linalg.norm(qg - ph * qo) <= 1e-6 * np.linalg.norm(qo)
        s = po.sigmin_svd(T, z)
        assert abs(Zg - s) <= 1e-6 * s
        assert abs(np.linalg.norm((T - z * np.eye(320)) @ qg) / np.linalg.norm(qg) - s) <= 1e-4 * s
    Zs, qs = pkg.psmode_inv_lanczos(T[:150, :150], q0[:150], 1.0 + 1.0j, 1e-8, 30)  # n < 200: host SVD (modes.jl:193)
    assert abs(Zs - po.sigmin_svd(T[:150, :150], 1.0 + 1.0j)) <= 1e-12 * Zs


def test_unsupported_shapes_raise(pkg, gpu):
    A = np.ones((190, 189))
    with pytest.raises(pkg.PsaGpuUnsupported):
        gpu.set_matrix(A)  # rectangular but not Hessenberg: rect_fact makes it a generalised problem (host path)
    with pytest.raises(pkg.PsaGpuUnsupported):
        gpu.set_matrix(np.ones((300, 250)))  # bw > 2
    with pytest.raises(pkg.PsaGpuUnsupported):
        gpu.set_matrix(np.ones((301, 300)))  # m == n+1 > 200 but dense below the sub-diagonal: not :HessQR input
    with pytest.raises(pkg.PsaGpuUnsupported):
        # square but not upper triangular (new_matrix hands A itself over when schur fails): the reference
        # factorises it (`if !istriu(T1) F1 = factorize(T1)`, compute.jl:597-601) -- host path, never silently wrong
        gpu.set_matrix(np.triu(np.ones((80, 80))) + np.eye(80, k=-3))
    with pytest.raises(pkg.PsaGpuUnsupported):
        pkg.psacore(np.triu(np.ones((301, 300)), -1), None, np.ones(300), [0.0], [0.0], 3, ctx=gpu)  # bw != 2
    assert gpu.shape is None  # a rejected matrix leaves no stale shape behind
    with pytest.raises(pkg.PsaGpuError):
        pkg.PsaGpu(1).grid([0.0], [0.0], np.ones(4))  # no matrix set


def test_cancel_flag_returns_none(pkg):
    T, eigA = po.complex_schur(po.random_complex(300, 2))
    flag = np.ones(1, dtype=np.int32)  # already cancelled: compute.jl:202-204 -> nothing
    assert pkg.psa_compute(T, 30, [-20, 20, -20, 20], eigA, {}, ctrlflag=flag, q0=po.start_vector(300)) is None
    flag[0] = 0
    out = pkg.psa_compute(T, 6, [-20, 20, -20, 20], eigA, {}, ctrlflag=flag, q0=po.start_vector(300))
    assert out is not None and out[0].shape == (6, 6)


def test_maxit_nonconvergence_counted(gpu):
    T, _ = po.complex_schur(po.random_complex(200, 3))
    q0 = po.start_vector(200, 0)
    gpu.set_matrix(T)
    x, y = np.linspace(-10, 10, 9), np.linspace(-10, 10, 9)
    Z, it, counts, _ = gpu.grid(x, y, q0, maxit=3)
    Zc, itc, flc = co.grid(T, x, y, q0, maxit=3)
    assert it.max() <= 3 and counts[0] == (flc & 1).sum() and _rel(Z, Zc) < RTOL


def test_conjugate_symmetry_large_real_matrix(gpu):
    """Size-independent property at a larger size: for a real matrix sigma_min(conj z) == sigma_min(z)."""
    T, eigA = po.complex_schur(po.toeplitz_from_symbol([0.0, 1.0], [0.0, 0.0, 0.25], 1024))  # toeplitz.jl 'triangle'
    x = np.linspace(-0.2, 1.4, 12)
    y = np.array([0.31, 0.77, 1.3])
    q0 = po.start_vector(1024, 0)
    gpu.set_matrix(T)
    Zp, _, _, _ = gpu.grid(x, y, q0)
    Zm, _, _, _ = gpu.grid(x, -y, q0)
    ok = Zp > 1e-6
    # two independent Lanczos runs stopped at tol 1e-5 (up to 70 iterations here): the reference's own
    # self-consistency test allows 1e-4 in norm (test/threads.jl:18); pointwise we allow 1e-3
    assert ok.sum() >= 24 and np.allclose(Zp[ok], Zm[ok], rtol=1e-3)
    Zc, _, _ = co.grid(T, x, y, q0, nthreads=8)
    assert _rel(Zp, Zc, _floor(T)) < RTOL


# ---- round 2: the rest of the boundary ---------------------------------------------------------------------------

def test_thresholds_are_mutable_like_the_reference(gpu):
    """compute.jl:18-30: raising minlancs4psa forces the SVD branch ("reference solutions"), lowering it sends a small
    matrix through the Lanczos path; maxstdqr4hess moves the HessQR / rect_qr boundary."""
    T, eigA = po.complex_schur(po.random_complex(60, 8))
    x, y = np.linspace(-6, 6, 7), np.linspace(-5, 5, 6)
    q0 = po.start_vector(60, 3)
    try:
        gpu.set_thresholds(100, 200)
        gpu.set_matrix(T)
        Zs, its, _, algo = gpu.grid(x, y, q0)
        assert algo == "SVD" and its.max() == 0
        svd = np.array([[po.sigmin_svd(T, xx + 1j * yy) for xx in x] for yy in y])
        assert np.max(np.abs(Zs - svd) / svd) < 1e-10
        gpu.set_thresholds(55, 200)
        gpu.set_matrix(T)
        Zl, itl, _, algo = gpu.grid(x, y, q0)
        assert algo == "sq_lanc" and np.max(np.abs(Zl - svd) / svd) < 1e-4  # Lanczos at tol 1e-5 vs its own SVD solution
        gpu.set_thresholds(20, 200)
        T30 = np.asfortranarray(T[:30, :30])
        gpu.set_matrix(T30)
        Z30, it30, _, algo = gpu.grid(x, y, q0[:30] / np.linalg.norm(q0[:30]))
        Zc, itc, _ = co.grid(T30, x, y, q0[:30] / np.linalg.norm(q0[:30]))
        assert algo == "sq_lanc" and _rel(Z30, Zc) < RTOL and np.array_equal(it30, itc)
        H = np.asfortranarray(po.grcar(190)[:, :-1]).astype(complex)  # 190 x 189
        gpu.set_thresholds(55, 150)
        gpu.set_matrix(H)
        Zh, _, _, algo = gpu.grid(x[:3], y[:2], po.start_vector(189, 1))
        Zhc, _, _ = co.grid(H, x[:3], y[:2], po.start_vector(189, 1))  # the C oracle's m != n branch is the Givens sweep
        assert algo == "HessQR" and _rel(Zh, Zhc, _floor(H)) < RTOL
    finally:
        gpu.set_thresholds(55, 200)


def test_per_batch_start_vectors_match_reference_loop(pkg, gpu):
    """The reference draws a new q per row batch (compute.jl:199-208).  With the same per-batch vectors the GPU path
    reproduces the oracle's row-batched psa_compute point for point."""
    T, eigA = po.complex_schur(po.readme_matrix(150))
    ax = po.vec2ax(eigA)
    npts = 22
    from pseudospectra_jl_b200 import compute as hc
    rb, nb = hc.row_batches(150, npts)
    assert po.get_step_size(150, npts) == 1 and nb == 22
    qs = np.stack([po.start_vector(150, 100 + b) for b in range(nb)])
    out = pkg.psa_compute(T, npts, ax, eigA, {}, q0=qs, ctx=gpu)
    ref = po.psa_compute(T, npts, ax, eigA, {}, q_for_batch=lambda b, n: qs[b])
    assert out[7] == ref[7] == "sq_lanc"
    assert np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2])
    assert _rel(out[0], ref[0]) < RTOL and np.allclose(out[3], ref[3])
    # n < 100 rule: ly // 8 rows per batch
    T60, e60 = po.complex_schur(po.random_complex(60, 2))
    rb60, nb60 = hc.row_batches(60, 20)
    assert nb60 == 10 and list(rb60[-2:]) == [0, 0] and rb60[0] == 9
    qs60 = np.stack([po.start_vector(60, 7 + b) for b in range(nb60)])
    out = pkg.psa_compute(T60, 20, [-8, 8, -8, 8], e60, {}, q0=qs60, ctx=gpu)
    ref = po.psa_compute(T60, 20, [-8, 8, -8, 8], e60, {}, q_for_batch=lambda b, n: qs60[b])
    assert _rel(out[0], ref[0]) < RTOL


@pytest.mark.parametrize("ax,npts", [([-1.44, 3.24, -3.8, 3.8], 20), ([-1, 2, -3, 4], 21), ([-1, 2, -4, 3], 20),
                                      ([-2, 2, -2, 2], 15), ([-2, 2, -2, 2], 16), ([-1, 2, 0.5, 4], 9)])
def test_device_postprocess_matches_reference_rules(gpu, ax, npts):
    """psa_gpu_postprocess = un-mirror (compute.jl:222-235) + clamp (236-238) + min/max for recalc_levels, for every
    branch of shift_axes (top master, bottom master, symmetric with / without a row at y = 0, no symmetry)."""
    T, eigA = po.complex_schur(po.grcar(64))
    x, y, nm = po.make_grid(ax, npts, True)
    q0 = po.start_vector(64, 0)
    gpu.set_matrix(T)
    Z, _, _, _ = gpu.grid(x, y, q0)
    Z2, _, _, _ = gpu.grid(x, y, q0, want_Z=False)
    assert Z2 is None
    tiny = float(np.median(Z))  # a clamp level that actually bites
    Zf, yf, zs = gpu.postprocess(y, len(x), nm, tiny, 1e-150)
    Zo, yo = po.unmirror(Z, y, nm)
    Zo = np.clip(Zo, tiny, np.inf)
    assert np.array_equal(yf, yo) and np.array_equal(Zf, Zo)
    assert zs[0] == Zo.min() and zs[1] == Zo.max() and zs[2] == Zo[Zo >= 1e-150].min()


def test_pseudomode_through_the_c_abi(pkg, gpu):
    """psa_gpu_pseudomode (modes.jl:185-274): basis kept on the device; against the oracle's restatement, the dense SVD
    and the defining residual."""
    T, eigA = po.complex_schur(po.random_complex(320, 9))
    q0 = po.start_vector(320, 4)
    gpu.set_matrix(T)
    for z in (2.0 + 1.5j, -25.0 + 3.0j):
        Zg, qg, l, info = gpu.pseudomode(z, q0, 1e-10, 60)
        Zo, qo, lo = po.psmode_inv_lanczos(T, q0, z, 1e-10, 60)
        assert info == 0 and l == lo and abs(Zg - Zo) <= 1e-8 * Zo
        ph = np.vdot(qo, qg) / abs(np.vdot(qo, qg))
        assert np.linalg.norm(qg - ph * qo) <= 1e-6 * np.linalg.norm(qo)
        s = po.sigmin_svd(T, z)
        assert abs(Zg - s) <= 1e-6 * s
        assert abs(np.linalg.norm((T - z * np.eye(320)) @ qg) / np.linalg.norm(qg) - s) <= 1e-4 * s
    Zn, qn, l, info = gpu.pseudomode(2.0 + 1.5j, q0, 1e-14, 3)  # cannot converge in 3 steps: info = 1 (modes.jl:262)
    assert info == 1 and l == 3
    assert np.isnan(pkg.psmode_inv_lanczos(T, q0, 2.0 + 1.5j, 1e-14, 3, ctx=gpu)[0])
    with pytest.raises(pkg.PsaGpuUnsupported):
        gpu.set_matrix(T[:150, :150])
        gpu.pseudomode(1.0, q0[:150], 1e-8, 30)  # n < 200: the reference's plain-SVD branch stays on the host


def test_maxit_above_126_and_fallback_counts(gpu):
    """maxit_lancs is mutable (compute.jl:27): no silent clamp.  Fallback points count as non-converged too
    (compute.jl:654-656 `break`s with lancz_converged false, so both warnings fire)."""
    T, _ = po.complex_schur(po.random_complex(400, 5))  # n > maxit: no exact Krylov breakdown before the limit
    q0 = po.start_vector(400, 0)
    gpu.set_matrix(T)
    x, y = np.linspace(-9, 9, 6), np.linspace(-9, 9, 5)
    Z, it, counts, _ = gpu.grid(x, y, q0, tol=0.0, maxit=150)  # `resid < 0` never holds: every point runs 150 iterations
    Zc, itc, flc = co.grid(T, x, y, q0, tol=0.0, maxit=150)
    assert it.min() == 150 and np.array_equal(it, itc) and _rel(Z, Zc) < RTOL and counts[0] == (flc & 1).sum() == 30
    To = np.asfortranarray(np.triu(np.ones((60, 60), dtype=np.complex128)))  # z = 1 sits exactly on an eigenvalue
    gpu.set_matrix(To)
    xs, ys = np.array([1.0, 3.0]), np.array([0.0, 2.0])
    Z, it, counts, _ = gpu.grid(xs, ys, po.start_vector(60, 0))
    Zc, itc, flc = co.grid(To, xs, ys, po.start_vector(60, 0))
    assert counts[1] == ((flc & 2) != 0).sum() == 1 and counts[0] == ((flc & 1) != 0).sum() >= 1
    assert Z[0, 0] < po.PS_TINY and Zc[0, 0] < po.PS_TINY


def test_zoom_reuses_the_resident_matrix(pkg, gpu):
    """A zoom re-enters psa_compute with the same Schur factor (zooming.jl:24-99): the upload + packing are skipped."""
    T, eigA = po.complex_schur(po.random_complex(700, 12))
    q0 = po.start_vector(700, 0)
    gpu._matrix_key = None
    calls = []
    orig = gpu.set_matrix
    gpu.set_matrix = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        o1 = pkg.psa_compute(T, 10, [-30, 30, -30, 30], eigA, {}, q0=q0, ctx=gpu)
        o2 = pkg.psa_compute(T, 10, [-5, 5, -5, 5], eigA, {}, q0=q0, ctx=gpu)  # zoom
        assert len(calls) == 1
        T2 = T.copy()
        T2[0, 5] += 1.0
        pkg.psa_compute(T2, 4, [-5, 5, -5, 5], eigA, {}, q0=q0, ctx=gpu)      # a different matrix is uploaded
        assert len(calls) == 2
    finally:
        gpu.set_matrix = orig
    x, y, _ = po.make_grid([-5, 5, -5, 5], 10, False)
    Zc, _, _ = co.grid(T, x, y, q0, nthreads=8)
    assert _rel(o2[0], np.clip(Zc, po.PS_TINY, np.inf)) < RTOL


def test_plain_c_driver_full_boundary(pkg, tmp_path):
    """thresholds / grid_batched / postprocess / pseudomode driven from plain C, checked against the oracle."""
    import subprocess
    from test_host import _build_abi_driver
    n, lx, ly = 208, 5, 4
    r = subprocess.run([_build_abi_driver(pkg, tmp_path), str(n), str(lx), str(ly), "full"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    T = np.where(i < j, 0.3 * np.sin(1.0 + i + 2.0 * j) + 0.3j * np.cos(2.0 + 3.0 * i + j), 0)
    T = np.asfortranarray(T + np.diag(0.05 * np.arange(n) - 1.0 + 0.02j * np.arange(n)))
    x = np.array([-2.0 + 4.0 * k / (lx - 1) for k in range(lx)])
    y = np.array([-1.0 + 3.0 * jj / (ly - 1) for jj in range(ly)])
    q0 = np.cos(0.7 * np.arange(n)) + 1j * np.sin(1.3 * np.arange(n) + 0.2)
    q0 = q0 / np.linalg.norm(q0)
    q1 = np.conj(q0[::-1])
    Zb = np.zeros((ly, lx))
    itb = np.zeros((ly, lx), dtype=int)
    for jj in range(ly):
        Zr, itr, _ = co.grid(T, x, y[jj:jj + 1], q1 if jj & 1 else q0, maxit=150)
        Zb[jj], itb[jj] = Zr[0], itr[0]
    tok = [l.split() for l in r.stdout.splitlines()]
    B = {(int(t[1]), int(t[2])): int(t[3]) for t in tok if t[0] == "B"}
    assert len(B) == lx * ly and all(B[(jj, k)] == itb[jj, k] for jj in range(ly) for k in range(lx))
    Zo, yo = po.unmirror(Zb, y, 2)
    Zo = np.clip(Zo, 1e-2, np.inf)
    Pm = np.zeros_like(Zo)
    for t in tok:
        if t[0] == "P":
            Pm[int(t[1]), int(t[2])] = float(t[3])
    assert np.allclose(Pm, Zo, rtol=RTOL, atol=0) and np.array_equal(Pm == 1e-2, Zo == 1e-2)
    Y = np.array([float(t[2]) for t in tok if t[0] == "Y"])
    assert np.array_equal(Y, yo)
    S = [t for t in tok if t[0] == "S"][0]
    assert float(S[1]) == Pm.min() and float(S[2]) == Pm.max()
    M = [t for t in tok if t[0] == "M"][0]
    Zm, qm, lm = po.psmode_inv_lanczos(T, q0, 0.3 + 0.4j, 1e-10, 60)
    V = np.array([float(t[2]) + 1j * float(t[3]) for t in tok if t[0] == "V"])
    assert int(M[3]) == 0 and int(M[2]) == lm and abs(float(M[1]) - Zm) <= 1e-8 * Zm
    ph = np.vdot(qm, V) / abs(np.vdot(qm, V))
    assert np.linalg.norm(V - ph * qm) <= 1e-6 * np.linalg.norm(qm)


# ---- round 2: parity at the sizes of the BASELINE configs -----------------------------------------------------------

def test_c2_grcar1000_mirror_vs_oracle(pkg, gpu):
    """BASELINE config 2: Grcar n = 1000 Schur factor, 200x200 grid with the real-matrix mirror, through psa_compute;
    every 9th computed row (about 2400 points) against the oracle, incl. the overflow points (bigsigma -> ps_tiny)."""
    T, eigA = po.complex_schur(po.grcar(1000))
    ax = po.vec2ax(eigA)
    q0 = po.start_vector(1000, 0)
    Z, x, y, levels, err, Tp, ep, algo = pkg.psa_compute(T, 200, ax, eigA, {"real_matrix": True}, q0=q0, ctx=gpu)
    xo, yo, nm = po.make_grid(ax, 200, True)
    assert algo == "sq_lanc" and Z.shape == (200, 200) and np.array_equal(x, xo) and nm != 0
    rows = np.arange(0, len(yo), 9)
    Zc, itc, flc = co.grid(T, xo, yo[rows], q0, nthreads=os.cpu_count() or 8)
    Zfull_o, yfull_o = po.unmirror(np.zeros((len(yo), len(xo))), yo, nm)
    assert np.array_equal(y, yfull_o)
    off = len(yfull_o) - len(yo) if nm > 0 else 0  # computed rows sit after the mirrored ones (top half is master)
    Zg = Z[off + rows] if nm > 0 else Z[rows]
    Zc = np.clip(Zc, po.PS_TINY, np.inf)
    floor = _floor(T)
    assert (Zc > floor).sum() > 500
    assert _rel(Zg, Zc, floor) < RTOL
    # overflow points (bigsigma -> ps_tiny): whether 1/sigma_min^2 crosses floatmax is decided in the last bits, so the
    # two sides need not flag the very same points -- but a point flagged by one is far below the floor on the other
    assert np.all(Zg[(flc & 2) != 0] < floor) and np.all(Zc[Zg <= po.PS_TINY] < floor)


def test_n4000_slice_vs_oracle_and_svd(gpu):
    """The headline size (BASELINE config 3, n = 4000).  T is a synthetic Schur factor with the statistics of the
    Ginibre ensemble's (strictly upper part iid complex Gaussian, eigenvalues in the disc of radius sqrt(2n)), so no
    O(n^3) host reduction is needed inside the test; bench.py checks the real Schur factor the same way every run.
    80 points against the oracle, 8 of them against dense svdvals (the reference's SVD-branch quantity)."""
    n = 4000
    rng = np.random.default_rng(42)
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), 1)
    rad = np.sqrt(2 * n) * np.sqrt(rng.random(n))
    T[np.arange(n), np.arange(n)] = rad * np.exp(2j * np.pi * rng.random(n))
    T = np.asfortranarray(T)
    q0 = po.start_vector(n, 0)
    x = np.linspace(-140.0, 140.0, 10)
    y = np.linspace(-135.0, 135.0, 8)
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    Zc, itc, flc = co.grid(T, x, y, q0, nthreads=os.cpu_count() or 8)
    floor = _floor(T)
    assert algo == "sq_lanc" and counts[3] == 80
    assert (Zc > floor).sum() >= 64 and _rel(Z, Zc, floor) < RTOL
    assert np.mean(it == itc) >= 0.97
    checked = 0
    for (j, k) in [(0, 0), (7, 9), (0, 9), (7, 0), (3, 0), (4, 9), (0, 4), (7, 5)]:  # far-field points: well-conditioned
        s = po.sigmin_svd(T, x[k] + 1j * y[j])
        if s > floor:
            assert abs(Z[j, k] - s) / s <= abs(Zc[j, k] - s) / s * 1.01 + 1e-9  # no worse than the reference algorithm
            assert abs(Z[j, k] - s) / s < 1e-3
            checked += 1
    assert checked >= 8


# ---- the reference's own tests at this boundary, restated ----------------------------------------------------------

@pytest.mark.parametrize("n", [20, 60])
def test_reference_threads_jl(pkg, gpu, n):
    """test/threads.jl:7-19: complex random n in {20 (SVD), 60 (sq_lanc)}, 100x100 grid on [-3,3]^2; two runs with
    fresh random start vectors agree to norm(Z0 - Z1)/norm(Z0) < 1e-4.  Z0 = GPU psa_compute (vectors drawn per row
    batch like compute.jl:207), Z1 = the oracle with another vector."""
    rng = np.random.default_rng(n)
    T, eigA = po.complex_schur(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    ax = [-3, 3, -3, 3]
    out = pkg.psa_compute(T, 100, ax, eigA, {}, rng=np.random.default_rng(5), ctx=gpu)
    assert out[7] == ("SVD" if n < 55 else "sq_lanc")
    x, y, _ = po.make_grid(ax, 100, False)
    if n < 55:
        Z1 = np.array([[po.sigmin_svd(T, xx + 1j * yy) for xx in x] for yy in y])
    else:
        Z1 = co.grid(T, x, y, po.start_vector(n, 99), nthreads=os.cpu_count() or 8)[0]
    Z1 = np.clip(Z1, po.PS_TINY, np.inf)
    assert np.linalg.norm(out[0] - Z1) / np.linalg.norm(out[0]) < 1e-4


def test_reference_medcplx(pkg, gpu):
    """test/medcplx.jl:7-14: landau_fox_li(225, 16), 60x60 grid on [-1.1,1.2]x[-1.1,1.1] -> :sq_lanc; then the
    pseudomode at z = 1im the way plotmode asks for it (tol 1e-15, 20 iterations,
    ext/PseudospectraPlots/PseudospectraPlots.jl:221-223)."""
    T, eigA = po.complex_schur(po.landau_fox_li(225, 16))
    ax = [-1.1, 1.2, -1.1, 1.1]
    q0 = po.start_vector(225, 0)
    Z, x, y, levels, err, Tp, ep, algo = pkg.psa_compute(T, 60, ax, eigA, {"levels": np.arange(-5, -0.9, 0.5)}, q0=q0,
                                                         ctx=gpu)
    assert algo == "sq_lanc" and err == 0 and Z.shape == (60, 60)
    Zc, itc, flc = co.grid(T, x, y, q0, nthreads=os.cpu_count() or 8)
    assert _rel(Z, np.clip(Zc, po.PS_TINY, np.inf), _floor(T)) < RTOL
    Zg, qg, l, info = gpu.pseudomode(1j, q0, 1e-15, 20)
    Zo, qo, lo = po.psmode_inv_lanczos(T, q0, 1j, 1e-15, 20)
    # tol = 1e-15 asks for agreement of successive Ritz values to the last bit, so WHEN the test fires is decided by
    # rounding: the iteration counts may differ by a step or two, the converged value may not
    assert abs(l - lo) <= 2 and abs(Zg - Zo) <= 1e-8 * Zo
    if l < 20 and lo < 20:
        ph = np.vdot(qo, qg) / abs(np.vdot(qo, qg))
        assert info == 0 and np.linalg.norm(qg - ph * qo) <= 1e-6
    elif l >= 20:
        assert info == 1  # `l >= num_its` -> the reference returns NaN (modes.jl:262-272)


def test_hessqr_shared_schur_fast_path(pkg, gpu):
    """:HessQR via one Schur factorisation of the square part (R^-1 R^-H == (Hs - z)^-1 (Hs - z)^-H because the
    reference's sweep never touches row n+1, compute.jl:581-592): agrees with the literal Givens path."""
    A = po.convdiff_fd(6)
    H = np.asfortranarray(po.arnoldi_hessenberg(A, 208, seed=0), dtype=complex)
    ev = np.linalg.eigvals(H[:208, :])
    x, y, _ = po.make_grid(po.vec2ax(ev), 16, False)
    q0 = po.start_vector(208, 2)
    Zg, algo_g, _, itg = pkg.psacore(H, None, q0, x, y, 2, ctx=gpu, want_iters=True)
    Zf, algo_f, _, itf = pkg.psacore(H, None, q0, x, y, 2, ctx=gpu, want_iters=True, hess_via_schur=True)
    assert algo_g == algo_f == "HessQR"
    assert _rel(Zf, Zg, _floor(H)) < RTOL and np.mean(itf == itg) >= 0.97


def test_device_projection_matches_host_restatement(pkg, gpu):
    """psa_gpu_project (the Givens reordering of _maybe_project, compute.jl:339-357, as a device wavefront) against the
    host restatement of the reference's sequential swaps: same projected matrix to rounding, same eigenvalue order,
    and the grid on it agrees with the oracle's Lanczos on the host-projected matrix."""
    from pseudospectra_jl_b200 import compute as hc
    T, eigA = po.complex_schur(po.random_complex(300, 6))
    ax = [-5.0, 5.0, -5.0, 5.0]
    q0 = po.start_vector(300, 0)
    Th, eh = hc._maybe_project(T, 0.6, ax, eigA)                      # host: sequential swaps, as the reference
    gpu._matrix_key = None
    out = pkg.psa_compute(T, 14, ax, eigA, {"proj_lev": 0.6}, q0=q0, ctx=gpu)
    Z, x, y, levels, err, Tp, ep, algo = out
    k = Th.shape[0]
    assert 55 <= k < 300 and Tp.shape == (k, k) and algo == "sq_lanc" and np.array_equal(ep, eh)
    assert np.abs(np.tril(Tp, -1)).max() == 0
    assert np.allclose(np.diagonal(Tp), eh, rtol=0, atol=1e-11 * np.abs(eigA).max())
    assert np.linalg.norm(Tp - Th) <= 1e-11 * np.linalg.norm(Th)
    Zc, _, _ = co.grid(Th, x, y, q0[:k], nthreads=os.cpu_count() or 8)
    assert _rel(Z, np.clip(Zc, po.PS_TINY, np.inf), _floor(Th)) < RTOL
    # zoom with another window: projected again from the resident FULL matrix, no upload
    calls = []
    orig = gpu.set_matrix
    gpu.set_matrix = lambda *a, **kw: (calls.append(1), orig(*a, **kw))[1]
    try:
        ax2 = [-2.0, 3.0, -1.0, 4.0]
        out2 = pkg.psa_compute(T, 10, ax2, eigA, {"proj_lev": 0.4}, q0=q0, ctx=gpu)
        Th2, eh2 = hc._maybe_project(T, 0.4, ax2, eigA)
        assert calls == [] and out2[5].shape == Th2.shape and np.array_equal(out2[6], eh2)
        if Th2.shape[0] >= 55:
            Zc2, _, _ = co.grid(Th2, out2[1], out2[2], q0[:Th2.shape[0]], nthreads=os.cpu_count() or 8)
            assert _rel(out2[0], np.clip(Zc2, po.PS_TINY, np.inf), _floor(Th2)) < RTOL
        out3 = pkg.psa_compute(T, 6, ax, eigA, {}, q0=q0, ctx=gpu)   # no projection: the full matrix is current again
        assert calls == [] and out3[5] is T and gpu.shape == (300, 300)
        Zc3, _, _ = co.grid(T, out3[1], out3[2], q0, nthreads=os.cpu_count() or 8)
        assert _rel(out3[0], np.clip(Zc3, po.PS_TINY, np.inf), _floor(T)) < RTOL
    finally:
        gpu.set_matrix = orig


def test_projection_below_minlancs_uses_svd_branch(pkg, gpu):
    """A projection can shrink T below minlancs4psa (minnev = 20 <= np): the reference then takes the SVD branch on the
    projected matrix; so does the device path."""
    T, eigA = po.complex_schur(po.random_complex(120, 6))
    ax = [-3.0, 3.0, -3.0, 3.0]
    gpu._matrix_key = None
    out = pkg.psa_compute(T, 8, ax, eigA, {"proj_lev": 0.5}, q0=po.start_vector(120, 0), ctx=gpu)
    k = out[5].shape[0]
    assert 20 <= k < 55 and out[7] == "SVD"
    svd = np.array([[po.sigmin_svd(out[5], xx + 1j * yy) for xx in out[1]] for yy in out[2]])
    assert np.max(np.abs(out[0] - svd) / svd) < 1e-9


def test_grcar100_published_abscissa_and_radius(gpu):
    """The one published known answer about sigma_min the reference's test-suite holds (test/radius_abscissa.jl:25-28,
    Guglielmi & Overton 2012): Grcar(100), eps = 1e-4, abscissa 2.41276 and radius 2.85216 to 1e-5 -- checked on the GPU
    path through the definition of the level set (see tests/test_oracle.py: check_go)."""
    from test_oracle import check_go
    T, eigA = po.complex_schur(po.grcar(100))
    q0 = po.start_vector(100, 0)
    gpu.set_matrix(T)

    def gpu_sig(zs):
        return np.array([gpu.grid(np.array([z.real]), np.array([z.imag]), q0)[0][0, 0] for z in zs])

    check_go(gpu_sig)


def test_kernel_variants_are_bit_identical(gpu, monkeypatch):
    """The four choreographies of the solve (round-1 kernel, deep ring, sparse, cluster pipeline over 2 / 4 SMs) and the
    three panel widths execute the same arithmetic in the same order: one operator application, bit for bit."""
    rng = np.random.default_rng(5)
    n, ns = 900, 45
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + np.diag(3 * rng.standard_normal(n))
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    gpu.set_matrix(T)
    monkeypatch.setenv("PSA_SOLVE_FLAVOR", "shallow")
    V0 = gpu.apply_resolvent(z, Q)
    F = T - z[0] * np.eye(n)
    v = sla.solve_triangular(F, sla.solve_triangular(F, Q[0], trans="C"))
    assert np.linalg.norm(V0[0] - v) <= 1e-11 * np.linalg.norm(v)
    monkeypatch.setenv("PSA_SOLVE_FLAVOR", "deep")
    for width in ("8", "16", "32"):
        monkeypatch.setenv("PSA_PANEL_WIDTH", width)
        for cluster in ("1", "2", "4"):
            monkeypatch.setenv("PSA_CLUSTER", cluster)
            assert np.array_equal(gpu.apply_resolvent(z, Q), V0), (width, cluster)
        # the full-launch instantiation (2 CTAs per SM, lane-per-shift diagonal steps) and, 32-wide, two panels per CTA
        monkeypatch.setenv("PSA_CLUSTER", "1")
        monkeypatch.setenv("PSA_DEEP_CW", "4")
        assert np.array_equal(gpu.apply_resolvent(z, Q), V0), (width, "full")
        monkeypatch.setenv("PSA_MULTI", "2")
        assert np.array_equal(gpu.apply_resolvent(z, Q), V0), (width, "multi")
        monkeypatch.delenv("PSA_MULTI")
        monkeypatch.delenv("PSA_DEEP_CW")


@pytest.mark.parametrize("n", [55, 64, 97, 150, 163])
def test_point_resident_kernel_small_matrices(gpu, n, monkeypatch):
    """Small square matrices run `_psa_lanczos!` (compute.jl:619-665) one warp per grid point in a single launch
    (csrc/psa_point.cuh) instead of ticks of the DMMA kernels.  Same contract: sigma_min within 1e-6 of the oracle,
    iteration counts equal, the same points non-converged -- and the tick engine (PSA_NO_POINT_KERNEL=1) agrees.
    n covers the register-block boundaries of the kernel and the largest size that fits beside the matrix."""
    T, eigA = po.complex_schur(po.random_complex(n, n))
    q0 = po.start_vector(n, 1)
    ax = po.vec2ax(eigA)
    x, y = np.linspace(ax[0], ax[1], 23), np.linspace(ax[2], ax[3], 19)
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, q0)
    st = gpu.last_stats()
    assert algo == "sq_lanc" and st["ticks"] == 0 and st["total_launches"] <= 4  # one kernel (+ queue-order launches)
    Zc, itc, flc = co.grid(T, x, y, q0, nthreads=8)
    assert _rel(Z, Zc, _floor(T)) < RTOL and np.mean(it == itc) >= 0.995
    assert counts[3] == Z.size and counts[2] == it.sum() and counts[0] == (flc & 1).sum()
    monkeypatch.setenv("PSA_NO_POINT_KERNEL", "1")
    Zt, itt, ct, _ = gpu.grid(x, y, q0)
    assert gpu.last_stats()["ticks"] > 0
    assert _rel(Z, Zt, _floor(T)) < 1e-9 and np.mean(it == itt) >= 0.995
    monkeypatch.delenv("PSA_NO_POINT_KERNEL")
    Z3, it3, c3, _ = gpu.grid(x, y, q0, maxit=3)  # non-convergence is counted (compute.jl:603-606)
    Zc3, itc3, flc3 = co.grid(T, x, y, q0, maxit=3, nthreads=8)
    assert it3.max() <= 3 and c3[0] == (flc3 & 1).sum() and _rel(Z3, Zc3, _floor(T)) < RTOL
    Zl, _, _, _ = gpu.grid(x, y, q0, maxit=2000)  # long histories do not fit beside the matrix: tick engine, same answer
    assert gpu.last_stats()["ticks"] > 0 and _rel(Zl, Z, _floor(T)) < 1e-9


def test_point_resident_kernel_batched_start_vectors_and_order(pkg, gpu):
    """Per-batch start vectors (compute.jl:199-208) and more points than warps in flight (queue order) through the
    point kernel: equal to the oracle run batch by batch; a second call returns the same bits (no scheduling effects)."""
    n = 120
    T, eigA = po.complex_schur(po.random_complex(n, 9))
    ax = po.vec2ax(eigA)
    x, y = np.linspace(ax[0], ax[1], 70), np.linspace(ax[2], ax[3], 64)  # 4480 points > 148 x 16 warps
    rb = (np.arange(len(y)) // 20).astype(np.int32)
    qs = np.stack([po.start_vector(n, s) for s in range(rb.max() + 1)])
    gpu.set_matrix(T)
    Z, it, counts, algo = gpu.grid(x, y, qs, row_batch=rb)
    assert gpu.last_stats()["ticks"] == 0
    for b in range(rb.max() + 1):
        rows = np.flatnonzero(rb == b)[::5]
        Zc, itc, _ = co.grid(T, x, y[rows], qs[b], nthreads=8)
        assert _rel(Z[rows], Zc, _floor(T)) < RTOL and np.mean(it[rows] == itc) >= 0.995
    Z2, it2, _, _ = gpu.grid(x, y, qs, row_batch=rb)
    assert np.array_equal(Z, Z2) and np.array_equal(it, it2)


@pytest.mark.parametrize("n", [60, 130, 300])
def test_closed_form_sigma_min(gpu, n):
    """Known answers that need no oracle: a diagonal matrix (sigma_min = distance to the spectrum) and lambda*I plus one
    corner entry (2x2 closed form), see tests/test_oracle.py: closed_form_cases.  n = 60, 130 run the point-resident
    kernel, n = 300 the tensor-core tick engine; the bar is the Lanczos stopping tolerance (compute.jl:658-661)."""
    from test_oracle import CLOSED_FORM_Z, closed_form_cases
    q0 = po.start_vector(n, 3)
    for name, T, exact in closed_form_cases(n):
        gpu.set_matrix(np.asfortranarray(T))
        for z in CLOSED_FORM_Z:
            Z, it, counts, algo = gpu.grid(np.array([z.real]), np.array([z.imag]), q0)
            assert algo == "sq_lanc" and counts[0] == 0 and abs(Z[0, 0] / exact(z) - 1) < 1e-4, (name, z, Z[0, 0], exact(z))
