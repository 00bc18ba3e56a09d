"""Known-answer and finite-difference tests of the CPU oracle itself: the
rasterizer part of the oracle has no reference fixture (parity unpinned), so it is pinned to
closed-form cases and to its own float64 finite differences instead."""
import math

import torch

import helpers
from oracle import splat_oracle as so
from rodygs_b200 import synthetic


def _cam(H, W):
    cam = synthetic.make_camera(0, 1, H, W, 1)
    return cam._replace(world_view_transform=torch.eye(4))


def _render(cam, xyz, scales, rots, ops, colors, bg=None, dtype=torch.float32):
    bg = torch.zeros(3) if bg is None else bg
    st = helpers.oracle_settings(cam, bg.to(dtype), 0)
    return so.rasterize(xyz.to(dtype), None, None, colors.to(dtype), ops.to(dtype), scales.to(dtype), rots.to(dtype),
                        torch.eye(4, dtype=dtype), st)


def test_single_gaussian_closed_form():
    H = W = 32
    cam = _cam(H, W)
    z, s, o = 4.0, 0.2, 0.8
    out = _render(cam, torch.tensor([[0.0, 0.0, z]]), torch.full((1, 3), s), torch.tensor([[1.0, 0, 0, 0]]),
                  torch.tensor([[o]]), torch.tensor([[0.9, 0.5, 0.1]]))
    f = W / (2 * cam.tanfovx)
    var = (f * s / z) ** 2 + 0.3
    centre = (W - 1) / 2.0
    for (px, py) in ((15, 15), (16, 15), (10, 20)):
        d2 = (px - centre) ** 2 + (py - centre) ** 2
        a_ref = o * math.exp(-0.5 * d2 / var)
        if a_ref < 1 / 255:
            a_ref = 0.0
        assert abs(out.alpha[0, py, px].item() - a_ref) < 1e-6
        assert abs(out.depth[0, py, px].item() - a_ref * z) < 1e-5
        assert abs(out.color[0, py, px].item() - 0.9 * a_ref) < 1e-6
    assert out.radii.item() == math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))
    # the key of its first duplicate: tile 0, depth bits of 4.0f
    assert int(out.bn.keys[0]) == (0 << 32) | 0x40800000
    assert out.bn.ranges[0].tolist() == [0, 1]


def test_near_plane_low_opacity_and_order():
    H = W = 32
    cam = _cam(H, W)
    xyz = torch.tensor([[0.0, 0.0, 0.2],      # exactly on the near plane: culled (z <= 0.2)
                        [0.0, 0.0, 3.0],      # opacity below 1/255: binned but never blended
                        [0.0, 0.0, 6.0],      # far, red
                        [0.0, 0.0, 5.0]])     # near, green  -> must be blended first
    scales = torch.full((4, 3), 0.3)
    rots = torch.tensor([[1.0, 0, 0, 0]]).repeat(4, 1)
    ops = torch.tensor([[0.9], [0.003], [0.6], [0.5]])
    cols = torch.tensor([[1.0, 1, 1], [1.0, 1, 1], [1.0, 0, 0], [0.0, 1, 0]])
    out = _render(cam, xyz, scales, rots, ops, cols)
    assert out.radii[0].item() == 0 and out.pp.tiles_touched[0].item() == 0
    assert out.radii[1].item() > 0
    # per tile, sorted values are in depth order: 1 (z=3), 3 (z=5), 2 (z=6)
    lo, hi = out.bn.ranges[0].tolist()
    assert out.bn.vals[lo:hi].tolist() == [1, 3, 2]
    # centre pixel: green in front of red, the low-opacity one contributes nothing
    f = W / (2 * cam.tanfovx)
    c = (W - 1) / 2.0
    d2 = 2 * (15 - c) ** 2

    def alpha(o, z):
        return o * math.exp(-0.5 * d2 / ((f * 0.3 / z) ** 2 + 0.3))

    ag, ar = alpha(0.5, 5.0), alpha(0.6, 6.0)
    assert abs(out.color[1, 15, 15].item() - ag) < 1e-6
    assert abs(out.color[0, 15, 15].item() - ar * (1 - ag)) < 1e-6
    assert abs(out.alpha[0, 15, 15].item() - (ag + ar * (1 - ag))) < 1e-6
    assert out.bl.n_contrib[15, 15].item() == 3


def test_saturating_stack_terminates_and_background():
    H = W = 16
    cam = _cam(H, W)
    n = 12
    xyz = torch.tensor([[0.0, 0.0, 2.0 + 0.1 * i] for i in range(n)])
    out = _render(cam, xyz, torch.full((n, 3), 5.0), torch.tensor([[1.0, 0, 0, 0]]).repeat(n, 1),
                  torch.full((n, 1), 0.9), torch.ones(n, 3), bg=torch.tensor([0.0, 0.0, 1.0]))
    # huge Gaussians: alpha ~ 0.9 at every pixel, T = 0.1, 0.01, 0.001, 1.0e-4 (>= 1e-4: blended);
    # the fifth would push T to 1e-5 < 1e-4 -> the loop stops before blending it (App. A.4)
    f = W / (2 * cam.tanfovx)
    T, cnt = 1.0, 0
    for i in range(n):
        var = (f * 5.0 / (2.0 + 0.1 * i)) ** 2 + 0.3
        a = min(0.99, 0.9 * math.exp(-0.5 * 0.5 / var))   # pixel (8,8) is 0.5 px off the centre in x and y
        if T * (1 - a) < 1e-4:
            break
        T *= 1 - a
        cnt += 1
    assert cnt == 4
    assert out.bl.n_contrib[8, 8].item() == cnt
    assert abs(out.bl.final_T[8, 8].item() - T) < 1e-7
    assert abs(out.alpha[0, 8, 8].item() - (1 - T)) < 1e-6
    assert abs(out.color[2, 8, 8].item() - ((1 - T) + T * 1.0)) < 1e-6   # blue: blended white + T * bg
    assert abs(out.color[0, 8, 8].item() - (1 - T)) < 1e-6               # red: no background contribution


def test_tile_border_rectangle():
    """A Gaussian centred exactly on a tile corner touches the 4 surrounding tiles (and more with its radius)."""
    H = W = 64
    cam = _cam(H, W)
    f = W / (2 * cam.tanfovx)
    z = 4.0
    # pixel x = ((ndc+1) W - 1)/2 = 31.5 at the optical axis; shift to land on pixel 32.0 (tile border at 32)
    x = (32.0 - 31.5) * z / f
    out = _render(cam, torch.tensor([[x, x, z]]), torch.full((1, 3), 0.02), torch.tensor([[1.0, 0, 0, 0]]),
                  torch.tensor([[0.9]]), torch.ones(1, 3))
    r = out.radii.item()
    var = (f * 0.02 / z) ** 2 + 0.3
    assert r == math.ceil(3 * math.sqrt(var + math.sqrt(0.1))) == 3 and abs(out.pp.xy[0, 0].item() - 32.0) < 1e-4
    assert out.pp.rect_min[0].tolist() == [1, 1] and out.pp.rect_max[0].tolist() == [3, 3]
    assert out.pp.tiles_touched[0].item() == 4
    tiles = sorted((int(k) >> 32) for k in out.bn.keys.tolist())
    assert tiles == [1 * 4 + 1, 1 * 4 + 2, 2 * 4 + 1, 2 * 4 + 2]


def test_covariance_convention_matches_reference_helper():
    """Sigma = R diag(s^2) R^T with R(q) of /root/reference/src/utils/general_utils.py:92-115 (unit quaternion)."""
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(5, 4, generator=g))
    s = torch.rand(5, 3, generator=g) + 0.1
    r, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(5, 3, 3)
    L = R @ torch.diag_embed(s)
    sigma = L @ L.transpose(1, 2)
    # push points through the oracle with an identity camera looking at them and compare the 2D covariance
    H = W = 64
    cam = _cam(H, W)
    xyz = torch.tensor([[0.0, 0.0, 5.0]]).repeat(5, 1)
    pp = so.preprocess(xyz, s, q, torch.ones(5, 1), None, torch.ones(5, 3), torch.eye(4), helpers.oracle_settings(cam, torch.zeros(3), 0))
    f = W / (2 * cam.tanfovx)
    J = torch.tensor([[f / 5.0, 0, 0], [0, f / 5.0, 0]])
    cov2d = J @ sigma @ J.t()
    ref = torch.stack([cov2d[:, 0, 0] + 0.3, cov2d[:, 0, 1], cov2d[:, 1, 1] + 0.3], 1)
    assert torch.allclose(pp.cov2d, ref, rtol=1e-4, atol=1e-5)


def test_oracle_gradients_against_float64_finite_differences():
    """Autograd of the oracle (with the straight-through conventions) vs central differences in float64."""
    H, W, n = 32, 32, 24
    sc = synthetic.make_scene(n, H, W, 3, seed=5, radius_px=3.0)
    cam = synthetic.make_camera(1, 5, H, W, 3)
    acts = [t.double() for t in helpers.activated_concat(sc, cam)]
    xyz, op, scl, rot, feat = acts
    op = op.clamp(0.05, 0.9)   # keep away from the 0.99 clamp / 1/255 threshold, which are deliberately non-smooth
    vm = cam.world_view_transform.t().contiguous().double()
    g = torch.Generator().manual_seed(3)
    wc, wd, wa = torch.randn(3, H, W, generator=g).double(), torch.randn(1, H, W, generator=g).double(), torch.randn(1, H, W, generator=g).double()
    st = so.Settings(H, W, cam.tanfovx, cam.tanfovy, torch.tensor([0.2, 0.3, 0.1]).double(), 1.0,
                     cam.projection_matrix.t().contiguous().double(), 2)

    def f(xyz_, op_, scl_, rot_, feat_, vm_):
        out = so.rasterize(xyz_, None, feat_, None, op_, scl_, rot_, vm_, st)
        return (out.color * wc).sum() + (out.depth * wd).sum() + (out.alpha * wa).sum()

    inputs = [t.clone().requires_grad_(True) for t in (xyz, op, scl, rot, feat, vm)]
    f(*inputs).backward()
    eps = 1e-6
    gen = torch.Generator().manual_seed(11)
    for k, t in enumerate(inputs):
        if k == 5:
            idxs = [(0, 0), (1, 2), (3, 0), (3, 2), (2, 1)]      # V^T entries (rotation and translation parts)
        else:
            flat = torch.randint(0, t.numel(), (6,), generator=gen).tolist()
            idxs = [tuple(int(v) for v in torch.unravel_index(torch.tensor(i), t.shape)) for i in flat]
        for idx in idxs:
            base = [u.detach().clone() for u in inputs]
            base[k][idx] += eps
            fp = f(*base).item()
            base[k][idx] -= 2 * eps
            fm = f(*base).item()
            fd = (fp - fm) / (2 * eps)
            an = t.grad[idx].item()
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(fd), abs(an)), (k, idx, fd, an)
