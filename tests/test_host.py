"""CPU tests of the host side: C-ABI exports, the Python mirror's grid/post-processing logic against the oracle,
error behaviour without a GPU, and the world_size-2 row sharding on gloo."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import psa_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "psa_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(psa_gpu_\w+)\s*\(", hdr))
    assert declared == set(pkg.EXPORTS)
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in ctypes.cast(lib.psa_gpu_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def _build_abi_driver(pkg, tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_driver.c"),
                           "-o", exe, "-L", libdir, "-lpsagpu", "-lm", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device exit of the plain-C driver")
def test_plain_c_driver_links_and_fails_loudly_without_gpu(pkg, tmp_path):
    """A C program that only includes include/psa_gpu.h links against libpsagpu.so (no Python, no torch)."""
    import subprocess
    r = subprocess.run([_build_abi_driver(pkg, tmp_path), "64", "3", "2"], capture_output=True, text=True)
    assert r.returncode == 10 and "no usable sm_100" in r.stdout


def test_strerror_and_argument_errors_without_gpu(pkg):
    lib = pkg.load_library()
    assert lib.psa_gpu_strerror(0) == b"ok"
    assert b"cancel" in lib.psa_gpu_strerror(-3)
    ctx = ctypes.c_void_p()
    assert lib.psa_gpu_init(0, None, ctypes.byref(ctx)) == -1  # PSA_ERR_ARG
    assert lib.psa_gpu_free(None) == -1
    assert lib.psa_gpu_set_matrix(None, None, 1, 1, 1) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a GPU-less host")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.PsaGpuError) as e:
        pkg.psa_compute(np.triu(np.ones((60, 60))), 8, [-1, 2, -1, 1], [], {})
    assert e.value.status == -5  # PSA_ERR_NODEVICE: the product path never computes on the CPU


@pytest.mark.parametrize("ax,npts,real", [([-1.44, 3.24, -3.8, 3.8], 20, True), ([-1, 2, -3, 4], 21, True),
                                           ([-1, 2, -4, 3], 20, True), ([-2, 2, -2, 2], 15, True),
                                           ([-1, 2, -3, 4], 16, False), ([-1, 2, 0.5, 4], 9, True)])
def test_mirror_grid_logic_matches_oracle(pkg, ax, npts, real):
    from pseudospectra_jl_b200 import compute as hc
    x, y, nm = hc._grid([float(a) for a in ax], npts, real, False)
    xo, yo, nmo = po.make_grid(ax, npts, real)
    assert nm == nmo and np.array_equal(x, xo) and np.array_equal(y, yo)
    Z = np.random.default_rng(0).random((len(y), len(x)))
    Zm, ym = hc._unmirror(Z, y, nm)
    Zo, yo2 = po.unmirror(Z, yo, nmo)
    assert np.array_equal(Zm, Zo) and np.array_equal(ym, yo2)


def test_scale_equal_and_levels_match_oracle(pkg):
    from pseudospectra_jl_b200 import compute as hc
    x, y, nm = hc._grid([-4.0, 4.0, -1.0, 1.0], 40, False, True)
    xo, yo, _ = po.make_grid([-4, 4, -1, 1], 40, False, True)
    assert len(x) == 40 and len(y) == 10 and np.array_equal(y, yo)
    Z = np.logspace(-7, -0.5, 100).reshape(10, 10)
    lev, err = hc.recalc_levels(Z, [-4.0, 4.0, -1.0, 1.0])
    levo, erro = po.recalc_levels(Z, [-4.0, 4.0, -1.0, 1.0])
    assert err == erro and np.allclose(lev, levo)


def test_maybe_project_reorders_schur_form(pkg):
    """_maybe_project (compute.jl:284-361): unitary similarity that moves the selected eigenvalues to the top."""
    from pseudospectra_jl_b200 import compute as hc
    T, eigA = po.complex_schur(po.random_complex(120, 3))
    ax = [-3.0, 3.0, -3.0, 3.0]
    Tn, en = hc._maybe_project(T, np.inf, ax, eigA)      # default: every eigenvalue selected -> untouched
    assert Tn is T and np.array_equal(en, eigA)
    Tp, ep = hc._maybe_project(T, 0.5, ax, eigA)
    k = len(ep)
    assert 20 <= k < 120 and Tp.shape == (k, k) and np.abs(np.tril(Tp, -1)).max() == 0
    assert np.allclose(np.diagonal(Tp), ep, atol=1e-11)  # selected eigenvalues, in their original order
    inside = (eigA.real > -3 - 6 * 0.5) & (eigA.real < 3 + 6 * 0.5)
    assert set(np.round(ep, 9)) <= set(np.round(eigA[inside], 9))
    for z in (0.3 + 0.2j, -1.0 + 2.0j):                  # restriction to an invariant subspace: resolvent norm shrinks
        assert po.sigmin_svd(Tp, z) >= po.sigmin_svd(T, z) * (1 - 1e-10)
    Tm, em = hc._maybe_project(T, -0.2, ax, eigA)        # negative factor: keep |lambda| > 1/|factor| (halving until >= 20)
    assert len(em) >= 20 and np.abs(em).min() > 5.0 / 2 ** 8
    small, es = po.complex_schur(po.random_complex(15, 1))
    assert hc._maybe_project(small, 0.1, ax, es)[0] is small  # m <= minnev: no projection


def test_psacore_argument_errors(pkg):
    with pytest.raises(ValueError):
        pkg.psacore(np.ones((3, 5)), None, np.ones(5), [0.0], [0.0], 1)  # m < n, compute.jl:459-461
    with pytest.raises(NotImplementedError):
        pkg.psacore(np.ones((70, 70)), np.eye(70), np.ones(70), [0.0], [0.0], 1)  # generalised


def _shard_worker(rank, world, port, ly, lx, out):
    import torch.distributed as dist
    import psa_b200_loader
    psa_b200_loader.load_package()
    from pseudospectra_jl_b200 import sharding
    import c_oracle as co
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    T = torch.zeros((64, 64), dtype=torch.complex128)
    Tn, _ = po.complex_schur(po.random_complex(64, 5))
    if rank == 0:
        T.copy_(torch.from_numpy(np.ascontiguousarray(Tn)))
    sharding.broadcast_matrix(T, 0)
    x, y = np.linspace(-3, 3, lx), np.linspace(-2, 2, ly)
    rows = sharding.rows_for_rank(ly, rank, world)
    Zl, itl, _ = co.grid(T.numpy(), x, y[rows], po.start_vector(64, 0))  # compute stand-in for the GPU call
    Z = sharding.gather_rows(torch.from_numpy(np.ascontiguousarray(Zl)), ly, lx, 0)
    it = sharding.gather_rows(torch.from_numpy(np.ascontiguousarray(itl)), ly, lx, 0)
    if rank == 0:
        np.save(out, np.stack([Z.numpy(), it.numpy().astype(np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_cyclic_sharding_world2_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    import c_oracle as co
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ly, lx = 7, 5  # odd row count: ranks own 4 and 3 rows
    out = str(tmp_path / "z.npy")
    mp.spawn(_shard_worker, args=(2, port, ly, lx, out), nprocs=2, join=True)
    got = np.load(out)
    Tn, _ = po.complex_schur(po.random_complex(64, 5))
    Zref, itref, _ = co.grid(Tn, np.linspace(-3, 3, lx), np.linspace(-2, 2, ly), po.start_vector(64, 0))
    assert np.allclose(got[0], Zref, rtol=1e-12) and np.array_equal(got[1].astype(int), itref)


def test_rows_for_rank_partition(pkg):
    from pseudospectra_jl_b200 import sharding
    for ly in (1, 7, 400):
        for world in (1, 2, 4, 8):
            rows = np.concatenate([sharding.rows_for_rank(ly, r, world) for r in range(world)])
            assert sorted(rows.tolist()) == list(range(ly))


def test_plan_selects_the_machine_from_the_shape_only(pkg):
    """psa_gpu_plan: the branch of psacore (compute.jl:478, 534, 565, 579, 588-590) and whether a square matrix takes the
    single-launch point-resident kernel -- host logic, no device.  The documented range: n <= 163 at maxit = 99 (the
    packed triangle + >= 8 warps' tridiagonals must fit one SM's shared memory), the tick engine above or when the
    alpha / beta histories of a large maxit no longer fit."""
    from pseudospectra_jl_b200._lib import plan
    assert plan(54, 54) == ("SVD", 0) and plan(40, 30) == ("SVD", 0)          # n < minlancs4psa
    assert plan(55, 55) == ("sq_lanc", 16) and plan(150, 150) == ("sq_lanc", 16)
    assert plan(155, 155)[1] == 16 and 8 <= plan(160, 160)[1] < 16
    assert plan(163, 163) == ("sq_lanc", 8) and plan(164, 164) == ("sq_lanc", 0)
    assert plan(4000, 4000) == ("sq_lanc", 0)
    assert plan(150, 150, maxit=3)[1] == 16 and plan(150, 150, maxit=400)[1] == 0 and plan(55, 55, maxit=2000)[1] == 0
    assert plan(201, 200) == ("HessQR", 0) and plan(101, 100) == ("rect_qr", 0)  # m > maxstdqr4hess or not
    assert plan(300, 200) == (None, 0) and plan(0, 5) == (None, 0) and plan(100, 100, maxit=5000) == (None, 0)
    # the shared-memory layout the kernel relies on: eight consecutive triangular numbers are distinct mod 8, so a
    # quarter-warp reading one matrix row out of the packed upper triangle hits eight different 16-byte bank groups
    tri = [i * (i + 1) // 2 for i in range(512)]
    assert all(len({t % 8 for t in tri[a:a + 8]}) == 8 for a in range(0, 504, 8))
