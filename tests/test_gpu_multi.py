"""Multi-device contexts (psa_gpu_init(ngpus > 1)): rows dealt cyclically over the devices of ONE process, T replicated
by NCCL broadcast, sigma_min tiles gathered on device 0 over NCCL.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a
single-GPU box.  Results must be bitwise identical to the single-device context (points are independent)."""
import numpy as np
import pytest

import psa_oracle as po

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("ly,lx", [(7, 5), (40, 33)])
def test_two_device_context_matches_single_device(pkg, ly, lx):
    T, eigA = po.complex_schur(po.random_complex(300, 4))
    x, y = np.linspace(-25, 25, lx), np.linspace(-20, 20, ly)
    q0 = po.start_vector(300, 0)
    with pkg.PsaGpu(1) as g1:
        g1.set_matrix(T)
        Z1, it1, c1, a1 = g1.grid(x, y, q0)
    ng = min(_ngpus(), 4)
    with pkg.PsaGpu(ng) as g2:
        g2.set_matrix(T)
        Z2, it2, c2, a2 = g2.grid(x, y, q0)
        Z3, _, _, _ = g2.grid(x[:3], y[:1], q0)  # fewer rows than devices
    assert a1 == a2 == "sq_lanc"
    assert np.array_equal(Z1, Z2) and np.array_equal(it1, it2) and np.array_equal(c1, c2)
    assert np.array_equal(Z3, Z1[:1, :3])


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_two_device_hessqr(pkg):
    H = np.asfortranarray(po.grcar(210)[:, :-1]).astype(complex)
    x, y = np.linspace(-1, 3, 6), np.linspace(-3, 3, 5)
    q0 = po.start_vector(209, 1)
    with pkg.PsaGpu(1) as g1:
        g1.set_matrix(H)
        Z1, it1, _, _ = g1.grid(x, y, q0)
    with pkg.PsaGpu(2) as g2:
        g2.set_matrix(H)
        Z2, it2, _, algo = g2.grid(x, y, q0)
    assert algo == "HessQR" and np.array_equal(Z1, Z2) and np.array_equal(it1, it2)


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_multi_device_batched_q_and_device_postprocess(pkg):
    """Per-batch start vectors and the device-side assembly + post-processing on a multi-device context: the tiles are
    interleaved on device 0 (no host loop), un-mirrored and clamped there; bitwise equal to the single-device result."""
    T, eigA = po.complex_schur(po.grcar(96))
    x, y, nm = po.make_grid([-1.0, 3.0, -3.0, 4.0], 21, True)
    from pseudospectra_jl_b200 import compute as hc
    rb, nb = hc.row_batches(96, len(y))
    qs = np.stack([po.start_vector(96, 40 + b) for b in range(nb)])
    res = []
    for ng in (1, min(_ngpus(), 4)):
        with pkg.PsaGpu(ng) as g:
            g.set_matrix(T)
            Z, it, c, _ = g.grid(x, y, qs, row_batch=rb)
            Zf, yf, zs = g.postprocess(y, len(x), nm, po.PS_TINY)
            res.append((Z, it, c, Zf, yf, zs))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    Zo, yo = po.unmirror(res[0][0], y, nm)
    assert np.array_equal(res[0][3], np.clip(Zo, po.PS_TINY, np.inf)) and np.array_equal(res[0][4], yo)


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_multi_device_projection(pkg):
    """psa_gpu_project on a multi-device context: reordered on device 0, the projected matrix broadcast over NCCL."""
    T, eigA = po.complex_schur(po.random_complex(260, 8))
    ax = [-5.0, 5.0, -5.0, 5.0]
    q0 = po.start_vector(260, 0)
    res = []
    for ng in (1, 2):
        with pkg.PsaGpu(ng) as g:
            out = pkg.psa_compute(T, 9, ax, eigA, {"proj_lev": 0.6}, q0=q0, ctx=g)
            res.append(out)
    assert res[0][5].shape == res[1][5].shape and 55 <= res[0][5].shape[0] < 260
    assert np.array_equal(res[0][5], res[1][5]) and np.array_equal(res[0][0], res[1][0])


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("ly,lx", [(3, 5), (90, 64)])
def test_multi_device_point_resident_kernel(pkg, ly, lx):
    """Small matrices (n <= 163) run the point-resident kernel (csrc/psa_point.cuh) on every device: one launch per
    device over its rows, each with its own queue; bitwise equal to one device (a point's arithmetic does not depend on
    which warp of which device picks it up)."""
    T, eigA = po.complex_schur(po.readme_matrix(150))
    ax = po.vec2ax(eigA)
    x, y = np.linspace(ax[0], ax[1], lx), np.linspace(ax[2], ax[3], ly)
    q0 = po.start_vector(150, 0)
    with pkg.PsaGpu(1) as g1:
        g1.set_matrix(T)
        Z1, it1, c1, a1 = g1.grid(x, y, q0)
        assert g1.last_stats()["ticks"] == 0
    with pkg.PsaGpu(min(_ngpus(), 4)) as g2:
        g2.set_matrix(T)
        Z2, it2, c2, a2 = g2.grid(x, y, q0)
        assert g2.last_stats()["ticks"] == 0
    assert a1 == a2 == "sq_lanc"
    assert np.array_equal(Z1, Z2) and np.array_equal(it1, it2) and np.array_equal(c1, c2)
