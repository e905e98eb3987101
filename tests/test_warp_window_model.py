"""CPU: the candidate window of the gather adjoint (fiery_b200/csrc/warp_sample.cuh: inverse_map, adjoint_scan_begin) restated in
numpy with fp32 arithmetic, checked against brute force: for random rigid maps (any rotation, translations up to the map's size,
square and rectangular grids) every output pixel whose bilinear / nearest sample touches a source pixel must lie inside that
pixel's window -- a candidate outside the window would be a silently missing term of the gradient.  The GPU tests compare the
kernel with autograd on a handful of maps; this covers the enumeration logic on thousands of pixels x hundreds of maps."""
import numpy as np
import pytest

F = np.float32


def _coords(th, W, H):
    """sample_coords for every output pixel (i along W, j along H), fp32 like the device code (fma emulated in fp64)."""
    i = np.arange(W, dtype=F)
    j = np.arange(H, dtype=F)
    xs = ((F(2.0) * i + F(1.0)) / F(W) - F(1.0)).astype(F)
    ys = ((F(2.0) * j + F(1.0)) / F(H) - F(1.0)).astype(F)
    X, Y = np.meshgrid(xs.astype(np.float64), ys.astype(np.float64))            # (H, W)
    gx = (th[0] * X + (th[1] * Y + th[2]).astype(F).astype(np.float64)).astype(F)
    gy = (th[3] * X + (th[4] * Y + th[5]).astype(F).astype(np.float64)).astype(F)
    ix = ((gx + F(1.0)) * F(W) - F(1.0)) * F(0.5)
    iy = ((gy + F(1.0)) * F(H) - F(1.0)) * F(0.5)
    return ix.astype(F), iy.astype(F)


def _inverse_map(th, W, H):
    a, b, c, d = F(th[0]), F(th[1]) * F(W) / F(H), F(th[3]) * F(H) / F(W), F(th[4])
    ix, iy = _coords(th, W, H)
    det = F(a * d - b * c)
    return dict(ia=F(d / det), ib=F(-b / det), ic=F(-c / det), id=F(a / det), ix0=ix[0, 0], iy0=iy[0, 0],
                windowed=bool(abs(det) >= 0.25 and abs(d / det) + abs(b / det) <= 4 and abs(c / det) + abs(a / det) <= 4))


def _windows(m, W, H, nearest):
    """adjoint_scan_begin for every source pixel: inclusive [i_lo, i_hi] x [j_lo, j_hi]."""
    sx, sy = np.meshgrid(np.arange(W, dtype=F), np.arange(H, dtype=F))
    dx, dy = (sx - m["ix0"]).astype(F), (sy - m["iy0"]).astype(F)
    ci = (m["ia"] * dx + m["ib"] * dy).astype(F)
    cj = (m["ic"] * dx + m["id"] * dy).astype(F)
    r = F(0.5 if nearest else 1.0)
    slack = (F(0.05) + F(1e-4) * (np.abs(ci) + np.abs(cj))).astype(F)
    ei = r * (abs(m["ia"]) + abs(m["ib"])) + slack
    ej = r * (abs(m["ic"]) + abs(m["id"])) + slack
    return (np.maximum(np.ceil(ci - ei), 0), np.minimum(np.floor(ci + ei), W - 1),
            np.maximum(np.ceil(cj - ej), 0), np.minimum(np.floor(cj + ej), H - 1))


def _theta(rng, W, H):
    ang = rng.uniform(-np.pi, np.pi)
    ex, ey = rng.uniform(20, 60), rng.uniform(20, 60)
    tx, ty = rng.uniform(-0.8, 0.8) * ex, rng.uniform(-0.8, 0.8) * ey
    cs, sn = np.cos(ang), np.sin(ang)                                     # write_theta, warp.cu
    return np.array([cs, -sn, ty / ey, sn, cs, -(tx / ex)], dtype=F)


@pytest.mark.parametrize("W,H", [(40, 24), (24, 40), (32, 32), (200, 200)])
@pytest.mark.parametrize("nearest", [False, True])
def test_every_contributing_output_pixel_is_inside_the_window(W, H, nearest):
    rng = np.random.default_rng(W * 1000 + H + int(nearest))
    n_maps = 12 if W == 200 else 150
    checked = 0
    for _ in range(n_maps):
        th = _theta(rng, W, H)
        m = _inverse_map(th, W, H)
        assert m["windowed"], "every map warp_features builds is a rotation: the window path must take it"
        i_lo, i_hi, j_lo, j_hi = _windows(m, W, H, nearest)
        ix, iy = _coords(th, W, H)
        oi, oj = np.meshgrid(np.arange(W), np.arange(H))                      # output pixel (i, j) at [j, i]
        if nearest:
            taps = [(np.rint(ix), np.rint(iy))]
        else:
            x0, y0 = np.floor(ix), np.floor(iy)
            fx, fy = ix - x0, iy - y0
            taps = [(x0 + kx, y0 + ky) for ky in (0, 1) for kx in (0, 1)]
            wts = [((1 - fx) if kx == 0 else fx) * ((1 - fy) if ky == 0 else fy) for ky in (0, 1) for kx in (0, 1)]
        for t, (tx, ty) in enumerate(taps):
            ok = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
            if not nearest:
                ok &= wts[t] != 0
            sxi, syi = tx[ok].astype(int), ty[ok].astype(int)                    # the source pixel this output pixel reads
            ci, cj = oi[ok], oj[ok]
            inside = (ci >= i_lo[syi, sxi]) & (ci <= i_hi[syi, sxi]) & (cj >= j_lo[syi, sxi]) & (cj <= j_hi[syi, sxi])
            assert inside.all(), (th, int((~inside).sum()))
            checked += int(ok.sum())
    assert checked > 1000
    # the windows stay small: the point of the gather
    live = (i_hi >= i_lo) & (j_hi >= j_lo)                                   # windows clipped away entirely are empty
    area = ((i_hi - i_lo + 1) * (j_hi - j_lo + 1))[live]
    assert float(area.max()) <= (16 if W == H else 30)


def test_non_rigid_maps_leave_the_window_path():
    for th in ([0.5, 0, 0.3, 0, 0.45, 0], [0, 0, 0.2, 0, 0, -0.3], [np.nan, 0, 0, 0, 1, 0], [0.05, 0, 0, 0, 1, 0]):
        with np.errstate(all="ignore"):
            assert not _inverse_map(np.array(th, dtype=F), 40, 24)["windowed"]
