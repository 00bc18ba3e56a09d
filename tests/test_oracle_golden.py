"""Pin the oracle against the reference's own Python (fixtures made by
tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import deform_oracle as do
from oracle import loss_oracle as lo
from oracle import splat_oracle as so
from rodygs_b200 import synthetic


def _load(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLDEN, name)).items()}


def test_sh_basis_matches_reference_eval_sh():
    g = _load("sh_eval.npz")
    sh_ref_layout = g["sh"]                    # [M,3,16] (reference eval_sh layout)
    sh = sh_ref_layout.permute(0, 2, 1).contiguous()   # ours: [M,16,3] like get_features
    for deg in range(4):
        out = so.sh_to_rgb(deg, sh, g["dirs"])
        assert torch.allclose(out, g[f"deg{deg}"], atol=2e-6, rtol=1e-6), deg


def test_losses_match_reference():
    g = _load("losses.npz")
    a = g["a"].clone().requires_grad_(True)
    l1 = lo.l1(a, g["b"])
    assert abs(l1.item() - g["l1"].item()) < 1e-7
    (gl1,) = torch.autograd.grad(l1, a)
    assert torch.equal(gl1, g["l1_grad"])
    s = lo.ssim(a, g["b"])
    assert abs(s.item() - g["ssim"].item()) < 2e-6
    (gs,) = torch.autograd.grad(s, a)
    assert torch.allclose(gs, g["ssim_grad"], atol=1e-8, rtol=1e-4)
    d1 = g["d1"].clone().requires_grad_(True)
    p = lo.pearson_depth(d1, g["d2"])
    assert abs(p.item() - g["pearson"].item()) < 1e-6
    (gp,) = torch.autograd.grad(p, d1)
    assert torch.allclose(gp, g["pearson_grad"], atol=1e-9, rtol=1e-4)


def test_deformation_matches_reference():
    g = _load("deform.npz")
    state = {k[len("state/"):]: v for k, v in g.items() if k.startswith("state/")}
    emb = do.time_embedding(g["t_query"], 26, False)
    assert torch.allclose(emb, g["emb"], atol=1e-6)
    table = do.motion_basis(state, do.time_embedding(g["times"], 26, False))
    assert torch.allclose(table, g["table"], atol=1e-6, rtol=1e-5)
    basis_t = do.motion_basis(state, emb)
    trans, rot = do.gaussian_deformation(g["coeff"].squeeze(1), basis_t, table, g["time_ind"].long(),
                                         float(g["spatial_lr_scale"]))
    assert torch.allclose(trans, g["trans"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(rot, g["rot"], atol=1e-6, rtol=1e-5)


def test_product_motion_network_matches_reference():
    """rodygs_b200.deform.MotionBasisNetwork (batched heads) loads the reference's state_dict
    and reproduces its outputs."""
    from rodygs_b200.deform import MotionBasisNetwork
    g = _load("deform.npz")
    state = {k[len("state/"):]: v for k, v in g.items() if k.startswith("state/")}
    net = MotionBasisNetwork(int(g["netwidth"]), 16, 26, False)
    missing = net.load_state_dict(state, strict=True)
    emb = net.t_embedder(g["t_query"])
    assert torch.allclose(emb, g["emb"], atol=1e-6)
    basis_t, table = net.query_and_table(g["t_query"], net.batch_embedding(g["times"]))
    assert torch.allclose(table, g["table"], atol=1e-6, rtol=1e-5)
    tr, ro = net(g["coeff"], g["t_query"])
    trans, rot = do.gaussian_deformation(g["coeff"].squeeze(1), basis_t, table, g["time_ind"].long(),
                                         float(g["spatial_lr_scale"]))
    assert torch.allclose(trans, g["trans"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(rot, g["rot"], atol=1e-6, rtol=1e-5)


def test_camera_conventions():
    g = _load("camera.npz")
    P = synthetic.projection_matrix(0.01, 100.0, float(g["fovx"]), float(g["fovy"]))
    assert torch.allclose(P, g["P"], atol=1e-6)
    # the synthetic cameras are rigid transforms looking down +z
    cam = synthetic.make_camera(1, 8, 64, 96, 10)
    R = cam.world_view_transform[:3, :3]
    assert torch.allclose(R @ R.t(), torch.eye(3), atol=1e-6)
    assert abs(torch.linalg.det(R).item() - 1.0) < 1e-5
    p = cam.world_view_transform @ torch.tensor([0.0, 0.0, 4.0, 1.0])
    assert p[2] > 0 and abs(p[0]) < 1e-5 and abs(p[1]) < 1e-5


def test_rasterizer_oracle_is_frozen():
    """oracle/splat_oracle.py reproduces tests/golden/raster_small.npz (its own outputs, frozen by
    tests/golden/make_golden_raster.py): integer outputs bit for bit, floats to summation-order noise.  The
    rasterizer's reference source is absent, so this pins the oracle against drift, not against upstream."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden_raster as mg
    frozen = np.load(os.path.join(GOLDEN, "raster_small.npz"))
    data, meta = mg.build()
    assert [meta["n"], meta["H"], meta["W"], meta["deg"]] == frozen["meta"].tolist()
    for k in ("means3D", "opacities", "scales", "rotations", "shs", "viewmatrix", "projmatrix", "dL_dcolor"):
        assert np.array_equal(data[k], frozen[k]), f"input {k} changed: the scene generator moved"
    for k in ("radii", "tiles_touched", "keys", "vals", "ranges"):
        assert np.array_equal(data[k], frozen[k]), k
    for k in ("color", "depth", "alpha", "final_T"):
        assert np.abs(data[k] - frozen[k]).max() <= 2e-6, k
    for k in ("g_means3D", "g_means2D", "g_shs", "g_opacities", "g_scales", "g_rotations", "g_viewmatrix"):
        ref = frozen[k]
        assert np.abs(data[k] - ref).max() <= 1e-5 * np.abs(ref).max(), k
