"""Known-answer tests for the grid / post-processing rules, derived BY HAND from the reference's Julia source
(/root/reference/src/utils.jl:21-73, 112-141, 175-191, 237-248; src/compute.jl:222-235, 270-279) and applied to BOTH
restatements -- the oracle (oracle/psa_oracle.py) and the product's host mirror (pseudospectra.jl_b200/compute.py) --
so that neither copy is only ever checked against the other."""
import numpy as np
import pytest

import psa_oracle as po


@pytest.fixture(scope="module")
def impls(pkg):
    from pseudospectra_jl_b200 import compute as hc
    return [("oracle", po.shift_axes, po.unmirror, po.recalc_levels, po.vec2ax, po.tidyaxes),
            ("mirror", hc.shift_axes, hc._unmirror, hc.recalc_levels, hc.vec2ax, hc.tidyaxes)]


def test_shift_axes_top_master(impls):
    # ax = [-1,2,-3,4], npts = 21: ax[4] > -ax[3].  tpts = floor(4/7*21) = 12; range(4/23, 4, 12) has step 8/23:
    # y_i = (4+8i)/23.  n = #{y < 3} = 9 (i <= 8).  y[min(12,10)] = 76/23 != 3 -> insert 3 after the 9th entry.
    # num_mirror = n + 1 = 10.
    want = np.array([(4 + 8 * i) / 23 for i in range(9)] + [3.0] + [(4 + 8 * i) / 23 for i in range(9, 12)])
    for name, shift_axes, *_ in impls:
        y, nm = shift_axes([-1, 2, -3, 4], 21)
        assert nm == 10 and len(y) == 13, name
        assert np.allclose(y, want, rtol=0, atol=2e-15), name


def test_shift_axes_bottom_master(impls):
    # ax = [-1,2,-4,3], npts = 20: ax[4] < -ax[3].  bpts = floor(4/7*20) = 11; range(-4, -4/21, 11) has step 8/21:
    # y_i = (-84+8i)/21.  n = #{y < -3} = 3.  y[max(1,3)] = -68/21 != -3 -> insert -3 after the 3rd entry (12 points).
    # num_mirror = -(12 - (3+1) + 1) = -9.
    want = np.array([(-84 + 8 * i) / 21 for i in range(3)] + [-3.0] + [(-84 + 8 * i) / 21 for i in range(3, 11)])
    for name, shift_axes, *_ in impls:
        y, nm = shift_axes([-1, 2, -4, 3], 20)
        assert nm == -9 and len(y) == 12, name
        assert np.allclose(y, want, rtol=0, atol=2e-15), name


def test_shift_axes_symmetric(impls):
    for name, shift_axes, *_ in impls:
        y, nm = shift_axes([-2, 2, -2, 2], 15)   # odd: range(0, 2, 8), num_mirror = 7
        assert nm == 7 and np.allclose(y, np.arange(8) * 2 / 7, rtol=0, atol=2e-15), name
        y, nm = shift_axes([-2, 2, -2, 2], 16)   # even: step = 0.25, range(0.125, 2, 8), num_mirror = 8
        assert nm == 8 and np.allclose(y, 0.125 + np.arange(8) * 1.875 / 7, rtol=0, atol=2e-15), name


def test_unmirror_three_branches(impls):
    Z = np.arange(12.0).reshape(4, 3)
    for name, _, unmirror, *_ in impls:
        # bottom master, n_mirror = -2: [Z; reverse(Z[end-1:end])], y extended by -reverse(y[end-1:end])
        y = np.array([-4.0, -3.0, -2.0, -1.0])
        Zf, yf = unmirror(Z, y, -2)
        assert np.array_equal(Zf, np.vstack([Z, Z[[3, 2]]])) and np.array_equal(yf, [-4, -3, -2, -1, 1, 2]), name
        # top master, y[1] != 0, n_mirror = 2: [reverse(Z[1:2]); Z]
        y = np.array([0.5, 1.5, 2.5, 3.5])
        Zf, yf = unmirror(Z, y, 2)
        assert np.array_equal(Zf, np.vstack([Z[[1, 0]], Z])) and np.array_equal(yf, [-1.5, -0.5, 0.5, 1.5, 2.5, 3.5]), name
        # y[1] == 0, n_mirror = 2: [reverse(Z[2:3]); Z] -- the row at y = 0 is not duplicated
        y = np.array([0.0, 1.0, 2.0, 3.0])
        Zf, yf = unmirror(Z, y, 2)
        assert np.array_equal(Zf, np.vstack([Z[[2, 1]], Z])) and np.array_equal(yf, [-2, -1, 0, 1, 2, 3]), name


def test_recalc_levels_hand_derived(impls):
    # sigmin spans 1e-7 .. 10^-0.5 on ax = [-4,4,-1,1]: scale = 2, last_lev = log10(0.06) = -1.2218 < max_val = -0.5
    # -> num_lines = ceil(8 * 5.7782/6.5) = 8, max_val = -1.2218.  stepsize: round(5.7782/8*4)/4 = 0.75 -> 1.0;
    # 5.78/1 < 8 -> round(5.7782/8*10)/10 = 0.7 -> 0.5; 11.6 >= 4 -> min_val = ceil(-14)*0.5 = -7, max_val = floor(-2.44)*0.5
    # = -1.5; levels = -7:0.5:-1.5 (12 values), stride max(12 div 9, 1) = 1.
    Z = np.logspace(-7, -0.5, 100).reshape(10, 10)
    for name, _, _, recalc_levels, *_ in impls:
        lev, err = recalc_levels(Z, [-4.0, 4.0, -1.0, 1.0])
        assert err == 0 and np.allclose(lev, np.arange(-7.0, -1.49, 0.5)), name
        # everything below smallsigma = 1e-150 -> err = -2, no levels (utils.jl:24-28)
        lev, err = recalc_levels(np.full((3, 3), 1e-200), [-1, 1, -1, 1])
        assert err == -2 and len(lev) == 0, name
        # 20 levels -> stride floor(20/9) = 2, taken from the TOP: smax = 1e-1 (< last_lev), smin = 1e-20, 8 lines:
        # stepsize round(19/8*4)/4 = 2.5 -> 1.0; 19 >= 8; min -20, max -1: -20:1:-1 (20 values) -> every 2nd from the end
        Zw = np.logspace(-20, -1, 50).reshape(5, 10)
        lev, err = recalc_levels(Zw, [-100.0, 100.0, -100.0, 100.0])
        assert err == 0 and np.allclose(lev, np.arange(-19.0, -0.9, 2.0)), name


def test_vec2ax_and_tidyaxes_hand_derived(impls):
    # points spanning [-1, 2] x [-0.5, 1.5]: ext = max(3, 2)/3 = 1 -> [-2, 3, -1.5, 2.5]; tidyaxes(tol 0.02):
    # x: round_to = 0.1 -> [-2, 3]; y: span 4 -> 0.08 -> mantissa round(8) * 10^-2 = 0.08 -> -1.52 (19 steps), 2.56 (32 steps)
    pts = np.array([-1 - 0.5j, 2 + 1.5j, 0.3 + 0.1j])
    for name, *_, vec2ax, tidyaxes in impls:
        ax = vec2ax(pts)
        assert np.allclose(ax, [-2.0, 3.0, -1.52, 2.56], atol=1e-12), (name, ax)
        lo, hi = tidyaxes(-1.2345, 2.3456, 0.02)  # span 3.5801 -> 0.0716 -> 0.07: floor(-17.6)*0.07, ceil(33.5)*0.07
        assert np.allclose([lo, hi], [-1.26, 2.38], atol=1e-12), name


def test_step_size_rule(pkg):
    from pseudospectra_jl_b200 import compute as hc
    for n, ly, want in ((150, 100, 2), (1000, 103, 1), (4000, 400, 1), (201, 500, 9), (60, 100, 12), (32, 4, 1)):
        assert po.get_step_size(n, ly) == want == hc._get_step_size(n, ly)  # compute.jl:270-279
    rb, nb = hc.row_batches(150, 7)   # step 1 -> rows 7,6,...,1 are batches 0..6: row j (0-based) -> 6 - j
    assert nb == 7 and list(rb) == [6, 5, 4, 3, 2, 1, 0]
    rb, nb = hc.row_batches(60, 20)   # step 2 -> rows (20,19) batch 0, ..., (2,1) batch 9
    assert nb == 10 and list(rb[:4]) == [9, 9, 8, 8] and list(rb[-2:]) == [0, 0]
