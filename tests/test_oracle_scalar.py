"""The two CPU restatements of the rasterizer spec against each other: the vectorised
PyTorch oracle (oracle/splat_oracle.py, autograd backward) and the scalar-loop one (oracle/splat_scalar.py, blend
backward written out by hand).  Integer outputs bit-exact, images 1e-5, per-Gaussian screen-space gradients 1e-4."""
import numpy as np
import pytest
import torch

import helpers
from oracle import splat_oracle as so
from oracle import splat_scalar as ss


def _scene(n, H, W, seed, radius_px):
    sc, cam = helpers.small_scene(n, H, W, 4, seed=seed, radius_px=radius_px)
    acts = helpers.activated_concat(sc, cam)
    return acts, cam


@pytest.mark.parametrize("n,H,W,seed,radius,deg", [(60, 40, 56, 1, 6.0, 3), (48, 33, 47, 2, 9.0, 1), (64, 48, 48, 3, 3.0, 0)])
def test_scalar_loops_agree_with_the_vectorised_oracle(n, H, W, seed, radius, deg):
    acts, cam = _scene(n, H, W, seed, radius)
    xyz, op, scl, rot, feat = [t.detach().clone().requires_grad_(True) for t in acts]
    bg = torch.tensor([0.3, 0.1, 0.6])
    vm = cam.world_view_transform.t().contiguous()
    m2 = torch.zeros(n, 3, requires_grad=True)
    st = helpers.oracle_settings(cam, bg, deg)
    out = so.rasterize(xyz, m2, feat, None, op, scl, rot, vm, st)
    g = torch.Generator().manual_seed(7)
    gc, gd, ga = torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g), torch.randn(1, H, W, generator=g)
    loss = (out.color * gc).sum() + (out.depth * gd).sum() + (out.alpha * ga).sum()
    pp = out.pp
    loss.backward()
    # gradients of the blend alone w.r.t. its per-Gaussian inputs (leaf copies: tz also feeds the projection upstream)
    leaf = {k: getattr(pp, k).detach().clone().requires_grad_(True) for k in ("xy", "conic", "opacity", "rgb", "depth")}
    bl = so.blend(pp._replace(**leaf), out.bn, bg, H, W)
    loss_b = (bl.color * gc).sum() + (bl.depth * gd).sum() + (bl.alpha * ga).sum()
    g_xy, g_conic, g_op, g_rgb, g_dep = torch.autograd.grad(loss_b, [leaf[k] for k in ("xy", "conic", "opacity", "rgb", "depth")])

    # ---- scalar restatement ----
    spp = ss.preprocess(xyz.detach().numpy(), scl.detach().numpy(), rot.detach().numpy(), op.detach().numpy().reshape(-1),
                        feat.detach().numpy(), vm.numpy(), cam.projection_matrix.t().contiguous().numpy(), H, W,
                        cam.tanfovx, cam.tanfovy, 1.0, deg)
    keys, vals, ranges = ss.bin_tiles(spp, H, W)
    color, depth, alpha, final_T, n_contrib = ss.blend(spp, vals, ranges, bg.numpy(), H, W)

    # integer contract: bit-exact
    assert [g_["radius"] for g_ in spp] == out.radii.tolist()
    assert [g_["tiles"] for g_ in spp] == pp.tiles_touched.tolist()
    assert keys == out.bn.keys.tolist()
    assert vals == out.bn.vals.tolist()
    assert ranges == out.bn.ranges.tolist()
    assert len(keys) > 2 * n, "scene too sparse to say anything"
    # depth bits and the float outputs of preprocess
    vis = [i for i, g_ in enumerate(spp) if g_["visible"]]
    sub = {int(i): k for k, i in enumerate(pp.idx.tolist())}
    for i in vis:
        k = sub[i]
        assert np.float32(spp[i]["depth"]).view(np.uint32) == pp.depth[k].detach().numpy().view(np.uint32)
        assert np.allclose([float(v) for v in spp[i]["xy"]], pp.xy[k].detach().numpy(), rtol=0, atol=0)
        assert np.allclose([float(v) for v in spp[i]["conic"]], pp.conic[k].detach().numpy(), rtol=1e-6, atol=0)
        assert np.allclose(spp[i]["rgb"], pp.rgb[k].detach().numpy(), atol=2e-6)
    # images
    assert np.abs(color - out.color.detach().numpy()).max() <= 1e-5
    assert np.abs(depth - out.depth[0].detach().numpy()).max() <= 1e-5
    assert np.abs(alpha - out.alpha[0].detach().numpy()).max() <= 1e-5
    assert np.abs(final_T - out.bl.final_T.numpy()).max() <= 1e-6
    assert (n_contrib == out.bl.n_contrib.numpy()).mean() >= 0.999      # exp() ulps may flip a threshold on a stray pixel

    # ---- App. A.5 by hand vs autograd ----
    gr = ss.blend_backward(spp, vals, ranges, bg.numpy(), H, W, final_T, n_contrib, gc.numpy(), gd[0].numpy(), ga[0].numpy())

    def chk(name, hand, auto):
        hand, auto = np.asarray(hand, np.float64), np.asarray(auto, np.float64)
        scale = max(np.abs(auto).max(), 1e-30)
        err = np.abs(hand - auto).max() / scale
        assert err <= 1e-4, f"{name}: {err:.3e}"

    ids = pp.idx.tolist()
    chk("dL/dxy", gr["xy"][ids], g_xy.numpy())
    chk("dL/dconic", gr["conic"][ids], g_conic.numpy())
    chk("dL/dopacity", gr["opacity"][ids], g_op.numpy())
    chk("dL/drgb", gr["rgb"][ids], g_rgb.numpy())
    # depth also feeds nothing else in the blend; the sort key is not differentiable
    chk("dL/ddepth", gr["depth"][ids], g_dep.numpy())
    chk("dL/dmeans2D (per NDC unit)", gr["mean2D"], m2.grad[:, :2].numpy())
