"""Parity of the CUDA path (through the C ABI / drop-in API) against the CPU oracle on
identical seeded inputs.

Bars (BASELINE.json north_star): radii, sort keys, sorted order and tile ranges
bit-exact; colour / depth / alpha <= 1e-4 max-abs; gradients <= 1e-3 relative
(max|a-b| / max|b| per tensor; float atomics make the accumulation order - hence the
last bits - nondeterministic).
"""
import math

import pytest
import torch

import helpers
from oracle import deform_oracle as do
from oracle import splat_oracle as so
from rodygs_b200 import engine, synthetic
from rodygs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


def _settings(cam, bg, deg=3, mod=1.0, cov=True, sh=True, dev="cuda"):
    return GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, bg.to(dev), mod,
                                         cam.projection_matrix.t().contiguous().to(dev), deg, False, False, cov, sh)


def _run_boundary(acts, cam, bg, deg=3, mod=1.0, colors=None, grads=None, cov=True, sh=True):
    xyz, op, sc, rot, feat = [t.detach().clone().cuda().requires_grad_(True) for t in acts]
    vm = cam.world_view_transform.t().contiguous().cuda().requires_grad_(True)
    m2 = torch.zeros(xyz.shape[0], 3, device="cuda", requires_grad=True)
    col = None if colors is None else colors.detach().clone().cuda().requires_grad_(True)
    rast = GaussianRasterizer(_settings(cam, bg, deg, mod, cov, sh))
    out = rast(means3D=xyz, means2D=m2, shs=None if col is not None else feat, colors_precomp=col, opacities=op,
               scales=sc, rotations=rot, cov3Ds_precomp=None, viewmatrix=vm)
    res = {"out": out}
    if grads is not None:
        gc, gd, ga = [g.cuda() for g in grads]
        (out[0] * gc).sum().add((out[1] * gd).sum()).add((out[3] * ga).sum()).backward()
        res["grads"] = {"means3D": xyz.grad, "means2D": m2.grad, "shs": feat.grad if col is None else col.grad,
                        "opacities": op.grad, "scales": sc.grad, "rotations": rot.grad, "viewmatrix": vm.grad}
    return res


def _run_oracle(acts, cam, bg, deg=3, mod=1.0, colors=None, grads=None, cov=True, sh=True):
    xyz, op, sc, rot, feat = [t.detach().clone().requires_grad_(True) for t in acts]
    vm = cam.world_view_transform.t().contiguous().requires_grad_(True)
    m2 = torch.zeros(xyz.shape[0], 3, requires_grad=True)
    col = None if colors is None else colors.detach().clone().requires_grad_(True)
    st = helpers.oracle_settings(cam, bg, deg, mod, cov, sh)
    out = so.rasterize(xyz, m2, None if col is not None else feat, col, op, sc, rot, vm, st)
    res = {"out": out}
    if grads is not None:
        gc, gd, ga = grads
        ((out.color * gc).sum() + (out.depth * gd).sum() + (out.alpha * ga).sum()).backward()
        z = lambda t, ref: t.grad if t.grad is not None else torch.zeros_like(ref)
        res["grads"] = {"means3D": z(xyz, xyz), "means2D": z(m2, m2), "shs": z(feat, feat) if col is None else z(col, col),
                        "opacities": z(op, op), "scales": z(sc, sc), "rotations": z(rot, rot), "viewmatrix": z(vm, vm)}
    return res


def _upstream_grads(H, W, seed=5):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(3, H, W, generator=g), 0.3 * torch.randn(1, H, W, generator=g), 0.2 * torch.randn(1, H, W, generator=g))


def _check_images(cu, orc, tol=IMG_TOL):
    for name, a, b in (("color", cu[0], orc.color), ("depth", cu[1], orc.depth), ("alpha", cu[3], orc.alpha)):
        err = (a.cpu() - b).abs().max().item()
        assert err <= tol, f"{name}: max-abs {err:.3e} > {tol}"


@pytest.fixture(autouse=True)
def _debug_off():
    yield
    engine.config.debug_keep_unsorted = False
    engine.config.debug_activated = False
    engine.config.sync_free = False
    engine.config.binning = "tiles"


@pytest.mark.parametrize("binning", ["tiles", "lsd"])
@pytest.mark.parametrize("H,W,n,deg", [(96, 128, 3000, 3), (100, 75, 2000, 1), (64, 64, 1500, 0), (160, 256, 12000, 2)])
def test_boundary_forward_bitexact_and_images(H, W, n, deg, binning):
    """radii / keys / order / ranges bit-exact, images within 1e-4 (ragged sizes included), for both
    binning implementations: "tiles" (per-tile shared-memory sort) and "lsd" (scan + duplicateWithKeys +
    global radix sort, which also exposes the emission-order arrays)."""
    sc, cam = helpers.small_scene(n, H, W, 6, seed=n)
    acts = helpers.activated_concat(sc, cam)
    bg = torch.tensor([0.1, 0.3, 0.2])
    orc = _run_oracle(acts, cam, bg, deg)["out"]
    engine.config.debug_keep_unsorted = binning == "lsd"
    xyz, op, scl, rot, feat = [t.cuda() for t in acts]
    scene = engine.SceneArgs(st=engine.SetArgs(xyz=xyz, scaling=scl, rotation=rot, opacity=op, sh_dc=feat, sh_rest=feat,
                                               sh_dc_stride=48, sh_rest_stride=48, sh_rest_offset=3), raw=False)
    view = engine.ViewArgs(H, W, cam.tanfovx, cam.tanfovy, 1.0, deg, cam.world_view_transform.t().contiguous().cuda(),
                           cam.projection_matrix.t().contiguous().cuda(), bg.cuda())
    color, depth, alpha, radii, state = engine.render_forward(scene, view)
    D = int(state.num_rendered[0].item())
    assert int(state.num_rendered[1].item()) == 0
    # --- integer contract: bit-exact ---
    assert torch.equal(radii.cpu(), orc.radii), "radii differ"
    assert torch.equal(state.geom["tiles_touched"].cpu(), orc.pp.tiles_touched), "tiles_touched differ"
    assert D == orc.bn.keys.numel(), "duplicate count differs"
    if binning == "lsd":
        assert torch.equal(state.extras["point_offsets"].cpu().long(), orc.bn.point_offsets), "scan differs"
        assert torch.equal(state.extras["keys_unsorted"][:D].cpu(), orc.bn.keys_unsorted), "duplicateWithKeys keys differ"
        assert torch.equal(state.extras["vals_unsorted"][:D].cpu(), orc.bn.vals_unsorted), "duplicateWithKeys values differ"
    assert torch.equal(state.extras["keys_sorted"][:D].cpu(), orc.bn.keys), "sorted keys differ"
    assert torch.equal(state.vals_sorted[:D].cpu(), orc.bn.vals), "sorted order differs"
    assert torch.equal(state.ranges.cpu(), orc.bn.ranges), "tile ranges differ"
    # depth bits of visible Gaussians
    vis = orc.pp.idx[orc.pp.visible]
    assert torch.equal(state.geom["p2"][:, 1].cpu()[vis].view(torch.int32), orc.pp.depth.detach()[orc.pp.visible].view(torch.int32))
    # --- float contract ---
    _check_images((color, depth, None, alpha), orc)
    assert (state.final_T.cpu() - orc.bl.final_T).abs().max().item() <= IMG_TOL


@pytest.mark.parametrize("deg,cov,sh", [(3, True, True), (0, True, True), (2, False, False), (1, True, False)])
def test_boundary_backward(deg, cov, sh):
    H, W, n = 96, 128, 2500
    sc, cam = helpers.small_scene(n, H, W, 6, seed=11 + deg)
    acts = helpers.activated_concat(sc, cam)
    bg = torch.tensor([0.2, 0.1, 0.4])
    up = _upstream_grads(H, W)
    orc = _run_oracle(acts, cam, bg, deg, grads=up, cov=cov, sh=sh)
    cu = _run_boundary(acts, cam, bg, deg, grads=up, cov=cov, sh=sh)
    _check_images(cu["out"], orc["out"])
    assert torch.equal(cu["out"][4].cpu(), orc["out"].radii)
    for k, ref in orc["grads"].items():
        got = cu["grads"][k].cpu()
        assert got.shape == ref.shape, k
        err = helpers.rel_err(got, ref)
        assert err <= GRAD_TOL, f"grad {k}: rel err {err:.3e}"
    ok, ratio = helpers.elementwise_ok(cu["grads"]["viewmatrix"], orc["grads"]["viewmatrix"])
    assert ok, f"dL/dV: element-wise {ratio:.2f} x (1e-3 |b| + 1e-5 max|b|)"


def test_boundary_colors_precomp_and_scale_modifier():
    H, W, n = 80, 112, 1500
    sc, cam = helpers.small_scene(n, H, W, 6, seed=3)
    acts = helpers.activated_concat(sc, cam)
    colors = torch.rand(n, 3, generator=torch.Generator().manual_seed(9))
    bg = torch.zeros(3)
    up = _upstream_grads(H, W)
    orc = _run_oracle(acts, cam, bg, 0, mod=0.7, colors=colors, grads=up)
    cu = _run_boundary(acts, cam, bg, 0, mod=0.7, colors=colors, grads=up)
    _check_images(cu["out"], orc["out"])
    assert torch.equal(cu["out"][4].cpu(), orc["out"].radii)
    for k, ref in orc["grads"].items():
        err = helpers.rel_err(cu["grads"][k].cpu(), ref)
        assert err <= GRAD_TOL, f"grad {k}: rel err {err:.3e}"


def test_edge_cases_empty_culled_and_errors():
    cam = synthetic.make_camera(0, 1, 48, 64, 4)
    bg = torch.tensor([0.5, 0.25, 0.75])
    st = _settings(cam, bg, 0)
    rast = GaussianRasterizer(st)
    vm = cam.world_view_transform.t().contiguous().cuda()
    # all Gaussians behind the near plane: image == background, radii == 0
    n = 64
    xyz = torch.randn(n, 3, device="cuda") * 0.01
    xyz[:, 2] = -1.0
    args = dict(means3D=xyz, means2D=torch.zeros(n, 3, device="cuda"), shs=torch.randn(n, 16, 3, device="cuda"),
                colors_precomp=None, opacities=torch.rand(n, 1, device="cuda"), scales=torch.rand(n, 3, device="cuda") * 0.1,
                rotations=torch.randn(n, 4, device="cuda"), cov3Ds_precomp=None, viewmatrix=vm)
    color, depth, normal, alpha, radii, extra = rast(**args)
    assert (radii == 0).all() and extra is None and normal.abs().sum() == 0
    assert torch.allclose(color, bg.cuda().view(3, 1, 1).expand_as(color)) and depth.abs().sum() == 0 and alpha.abs().sum() == 0
    # N == 0
    e = dict(args)
    for k in ("means3D", "means2D", "shs", "opacities", "scales", "rotations"):
        e[k] = args[k][:0]
    color0 = rast(**e)[0]
    assert torch.allclose(color0, color)
    # error behaviour of the upstream wrapper
    with pytest.raises(Exception):
        rast(**{**args, "colors_precomp": torch.rand(n, 3, device="cuda")})
    with pytest.raises(Exception):
        rast(**{**args, "shs": None})
    with pytest.raises(Exception):
        rast(**{**args, "scales": None})
    with pytest.raises(RuntimeError):
        rast(**{k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in args.items()})


def test_known_answer_single_gaussian():
    """One isotropic Gaussian on the optical axis: centre pixel alpha = opacity * exp(-0.5 d^T conic d)."""
    H = W = 32
    cam = synthetic.make_camera(0, 1, H, W, 1)
    V = torch.eye(4)
    cam = cam._replace(world_view_transform=V)
    z, s, o = 4.0, 0.2, 0.8
    xyz = torch.tensor([[0.0, 0.0, z]])
    st = _settings(cam, torch.zeros(3), 0)
    rast = GaussianRasterizer(st)
    col = torch.tensor([[0.9, 0.5, 0.1]])
    out = rast(means3D=xyz.cuda(), means2D=torch.zeros(1, 3, device="cuda"), shs=None, colors_precomp=col.cuda(),
               opacities=torch.tensor([[o]]).cuda(), scales=torch.full((1, 3), s).cuda(),
               rotations=torch.tensor([[1.0, 0, 0, 0]]).cuda(), cov3Ds_precomp=None, viewmatrix=V.t().contiguous().cuda())
    color, depth, _, alpha, radii, _ = out
    f = W / (2 * cam.tanfovx)
    var = (f * s / z) ** 2 + 0.3
    centre = (W - 1) / 2.0            # pixel coords of the principal point: ((0+1)*W-1)/2
    d2 = 2 * (15 - centre) ** 2       # pixel (15,15)
    a_ref = o * math.exp(-0.5 * d2 / var)
    assert abs(alpha[0, 15, 15].item() - a_ref) < 1e-5
    assert abs(depth[0, 15, 15].item() - a_ref * z) < 1e-4
    assert torch.allclose(color[:, 15, 15].cpu(), col[0] * a_ref, atol=1e-5)
    # isotropic footprint: lambda_max = var + sqrt(max(0.1, 0)) (App. A.2 step 6)
    assert radii.item() == math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))


@pytest.fixture
def _tunables():
    from rodygs_b200 import _lib
    yield _lib.set_tunable
    for name, v in (("pre_grid_cap", 0), ("dtable_v1", 0), ("diff_smem", 1)):
        _lib.set_tunable(name, v)


@pytest.mark.parametrize("n,cap,dt1,diff", [(3000, 2, 0, 1), (2602, 3, 0, 0), (2602, 2, 1, 1), (3000, 0, 1, 0)])
def test_fused_path_multi_chunk_staging(_tunables, n, cap, dt1, diff):
    """The persistent preprocess kernels with several chunks per CTA (grid capped), chunks whose row count is not a
    multiple of 4 (2602 -> 1301 per model -> a 21-row last chunk that cannot be bulk-copied, in the middle of a CTA's
    chunk sequence), and both dL/dtable reductions."""
    _tunables("pre_grid_cap", cap)
    _tunables("dtable_v1", dt1)
    _tunables("diff_smem", diff)     # B(t) - table rows from shared memory / gathered from global memory
    test_fused_dynamic_path_matches_oracle_chain(n)
    test_fused_path_bitexact_given_its_own_activations(n)


def test_fused_path_with_more_frames_than_the_shared_memory_table_holds():
    """num_times = 150 > 140: the launchers fall back from the staged B(t) - table rows to per-Gaussian gathers."""
    test_fused_dynamic_path_matches_oracle_chain(2400, T=150)


def test_fused_dynamic_path_matches_oracle_chain(n=3000, T=6):
    """Raw parameters + deformation fused in the kernel vs the reference chain on the CPU
    (activations -> deformation -> concat -> rasterize), forward and all gradients."""
    from rodygs_b200.dynamic import GaussianParams, render_dynamic
    H, W = 96, 128
    sc, cam = helpers.small_scene(n, H, W, T, seed=21)
    bg = torch.tensor([0.0, 0.0, 0.0])
    up = _upstream_grads(H, W)
    # ---- oracle chain with autograd down to the raw parameters ----
    leaves = {}
    def leaf(t):
        return t.detach().clone().requires_grad_(True)
    st = do.RawGaussians(**{k: leaf(v) for k, v in sc["static"].items()})
    dy = do.RawGaussians(**{k: leaf(v) for k, v in sc["dynamic"].items()})
    coeff, table = leaf(sc["motion_coeff"]), leaf(sc["table"])
    basis_t = leaf(sc["table"][cam.time_index])
    acts = do.assemble(st, dy, coeff.squeeze(1), basis_t, table, sc["time_ind"].long(), sc["spatial_lr_scale"], True)
    vm = leaf(cam.world_view_transform.t().contiguous())
    m2 = torch.zeros(n, 3, requires_grad=True)
    xyz, op, scl, rot, feat = acts
    orc = so.rasterize(xyz, m2, feat, None, op, scl, rot, vm, helpers.oracle_settings(cam, bg, 3))
    ((orc.color * up[0]).sum() + (orc.depth * up[1]).sum() + (orc.alpha * up[2]).sum()).backward()
    # ---- fused CUDA path ----
    engine.config.debug_activated = True
    cst = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["static"].items()})
    cdy = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["dynamic"].items()})
    ccoeff = sc["motion_coeff"].cuda().requires_grad_(True)
    ctable = sc["table"].cuda().requires_grad_(True)
    cbasis = sc["table"][cam.time_index].cuda().requires_grad_(True)
    cvm = cam.world_view_transform.t().contiguous().cuda().requires_grad_(True)
    pkg = render_dynamic(cst, cdy, _settings(cam, bg, 3), cvm, ccoeff, cbasis, ctable, sc["time_ind"].cuda(),
                         sc["spatial_lr_scale"], True)
    ((pkg["rendered_image"] * up[0].cuda()).sum() + (pkg["rendered_depth"] * up[1].cuda()).sum()
     + (pkg["rendered_alpha"] * up[2].cuda()).sum()).backward()
    # activated parameters the kernel used vs the reference chain (exp/sigmoid differ by ulps)
    ns = sc["static"]["xyz"].shape[0]
    state_dbg = None
    _check_images((pkg["rendered_image"], pkg["rendered_depth"], None, pkg["rendered_alpha"]), orc)
    mism = (pkg["radii"].cpu() != orc.radii).float().mean().item()
    assert mism <= 2e-3, f"radii mismatch fraction {mism}"   # exp() ulps can flip a ceil(); bit-exactness is checked below
    pairs = [("static.xyz", cst.xyz, st.xyz), ("static.features_dc", cst.features_dc, st.features_dc),
             ("static.features_rest", cst.features_rest, st.features_rest), ("static.scaling", cst.scaling, st.scaling),
             ("static.rotation", cst.rotation, st.rotation), ("static.opacity", cst.opacity, st.opacity),
             ("dynamic.xyz", cdy.xyz, dy.xyz), ("dynamic.features_dc", cdy.features_dc, dy.features_dc),
             ("dynamic.features_rest", cdy.features_rest, dy.features_rest), ("dynamic.scaling", cdy.scaling, dy.scaling),
             ("dynamic.rotation", cdy.rotation, dy.rotation), ("dynamic.opacity", cdy.opacity, dy.opacity),
             ("motion_coeff", ccoeff, coeff), ("table", ctable, table), ("basis_t", cbasis, basis_t),
             ("viewmatrix", cvm, vm), ("means2D", pkg["viewspace_points"], m2)]
    for name, a, b in pairs:
        assert a.grad is not None, name
        err = helpers.rel_err(a.grad.cpu(), b.grad)
        assert err <= GRAD_TOL, f"grad {name}: rel err {err:.3e}"
        if name in ("motion_coeff", "table", "basis_t", "viewmatrix"):     # small sums of many paths: element-wise as well
            ok, ratio = helpers.elementwise_ok(a.grad, b.grad)
            assert ok, f"grad {name}: element-wise {ratio:.2f} x (1e-3 |b| + 1e-5 max|b|)"


def test_fused_path_bitexact_given_its_own_activations(n=2000):
    """The fused kernel's integer outputs are bit-exact w.r.t. the oracle evaluated on the
    activated values the kernel itself produced (isolates exp()/sigmoid ulp differences)."""
    H, W, T = 64, 96, 5
    sc, cam = helpers.small_scene(n, H, W, T, seed=33)
    engine.config.debug_activated = True
    dev = "cuda"
    def mk(d):
        return engine.SetArgs(xyz=d["xyz"].to(dev), scaling=d["scaling"].to(dev), rotation=d["rotation"].to(dev),
                              opacity=d["opacity"].to(dev), sh_dc=d["features_dc"].to(dev), sh_rest=d["features_rest"].to(dev))
    scene = engine.SceneArgs(st=mk(sc["static"]), dy=mk(sc["dynamic"]), raw=True, use_deform=True,
                             motion_coeff=sc["motion_coeff"].squeeze(1).contiguous().to(dev), time_ind=sc["time_ind"].to(dev),
                             basis_t=sc["table"][cam.time_index].contiguous().to(dev), table=sc["table"].to(dev),
                             spatial_lr_scale=sc["spatial_lr_scale"])
    bg = torch.zeros(3)
    view = engine.ViewArgs(H, W, cam.tanfovx, cam.tanfovy, 1.0, 3, cam.world_view_transform.t().contiguous().to(dev),
                           cam.projection_matrix.t().contiguous().to(dev), bg.to(dev))
    color, depth, alpha, radii, state = engine.render_forward(scene, view)
    act = state.geom["dbg_activated"].cpu()
    # activated values agree with the reference chain to a few ulps
    ref = helpers.activated_concat(sc, cam)
    assert torch.allclose(act[:, 0:3], ref[0], atol=1e-6, rtol=1e-6)
    assert torch.allclose(act[:, 3:6], ref[2], rtol=2e-6)
    assert torch.allclose(act[:, 6:10], ref[3], atol=1e-6)
    assert torch.allclose(act[:, 10:11], ref[1], atol=1e-6)
    feat = torch.cat([torch.cat((sc[k]["features_dc"], sc[k]["features_rest"]), 1) for k in ("static", "dynamic")], 0)
    orc = so.rasterize(act[:, 0:3].contiguous(), None, feat, None, act[:, 10:11].contiguous(), act[:, 3:6].contiguous(),
                       act[:, 6:10].contiguous(), cam.world_view_transform.t().contiguous(), helpers.oracle_settings(cam, bg, 3))
    D = int(state.num_rendered[0].item())
    assert torch.equal(radii.cpu(), orc.radii)
    assert D == orc.bn.keys.numel()
    assert torch.equal(state.extras["keys_sorted"][:D].cpu(), orc.bn.keys)
    assert torch.equal(state.vals_sorted[:D].cpu(), orc.bn.vals)
    assert torch.equal(state.ranges.cpu(), orc.bn.ranges)
    _check_images((color, depth, None, alpha), orc)


def test_losses_match_oracle():
    from oracle import loss_oracle as lo
    from rodygs_b200 import losses
    g = torch.Generator().manual_seed(4)
    for (H, W) in ((40, 56), (97, 131)):
        a = torch.rand(3, H, W, generator=g)
        b = (a + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
        ar = a.clone().requires_grad_(True)
        ref = lo.photometric(ar, b)
        ref.backward()
        ac = a.cuda().requires_grad_(True)
        loss, parts = losses.photometric_loss(ac, b.cuda(), 0.8, 0.2, return_parts=True)
        loss.backward()
        assert abs(loss.item() - ref.item()) < 2e-6
        assert abs(parts[1].item() - lo.l1(a, b).item()) < 1e-6 and abs(parts[2].item() - lo.ssim(a, b).item()) < 2e-6
        assert helpers.rel_err(ac.grad.cpu(), ar.grad) < GRAD_TOL
        assert abs(losses.ssim(a.cuda(), b.cuda()).item() - lo.ssim(a, b).item()) < 2e-6
        assert abs(losses.l1_loss(a.cuda(), b.cuda()).item() - lo.l1(a, b).item()) < 1e-6
        d1 = (torch.rand(1, H, W, generator=g) * 5 + 1)
        d2 = d1 * 0.7 + torch.rand(1, H, W, generator=g)
        dr = d1.clone().requires_grad_(True)
        pr = lo.pearson_depth(dr, d2)
        pr.backward()
        dc = d1.cuda().requires_grad_(True)
        pc = losses.pearson_depth_loss(dc, d2.cuda())
        pc.backward()
        assert abs(pc.item() - pr.item()) < 2e-6
        assert helpers.rel_err(dc.grad.cpu(), dr.grad) < GRAD_TOL
    # local Pearson over boxes (one batched launch) vs the reference's per-box loop
    H, W = 300, 280
    d1 = (torch.rand(1, H, W, generator=g) * 5 + 1)
    d2 = d1 * 0.7 + torch.rand(1, H, W, generator=g)
    x0 = torch.tensor([3, 100, 60, 171]); y0 = torch.tensor([7, 20, 150, 99])
    dr = d1.clone().requires_grad_(True)
    lr = lo.local_pearson_depth(dr, d2, x0, y0, 128)
    lr.backward()
    dc = d1.cuda().requires_grad_(True)
    lc = losses.local_pearson_depth_loss(dc, d2.cuda(), 128, origins=torch.stack([x0, y0], 1))
    lc.backward()
    assert abs(lc.item() - lr.item()) < 2e-6
    assert helpers.rel_err(dc.grad.cpu(), dr.grad) < GRAD_TOL

@pytest.mark.gpu
def test_fused_loss_stage_matches_oracle():
    """rdg_losses: photometric + global Pearson + alpha regulariser in three launches (the trainer's loss stage)."""
    import ctypes as C
    from oracle import loss_oracle as lo
    from rodygs_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    for (H, W) in ((33, 70), (128, 96), (97, 131)):
        a = torch.rand(3, H, W, generator=g)
        b = (a + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
        d1 = torch.rand(1, H, W, generator=g) * 5 + 1
        d2 = d1 * 0.7 + torch.rand(1, H, W, generator=g)
        al = torch.rand(1, H, W, generator=g)
        w_p, w_a = 0.05, 0.01
        ar, dr, alr = a.clone().requires_grad_(True), d1.clone().requires_grad_(True), al.clone().requires_grad_(True)
        photo = lo.photometric(ar, b)
        pear = lo.pearson_depth(dr, d2)
        areg = w_a * (1.0 - alr).mean()
        (photo + w_p * pear + areg).backward()
        dev = "cuda"
        ac, bc, dc, d2c, alc = a.to(dev), b.to(dev), d1.to(dev), d2.to(dev), al.to(dev)
        out = torch.zeros(8, device=dev)
        gcol = torch.empty(3, H, W, device=dev)
        gdep = torch.full((1, H, W), 7.0, device=dev)
        galp = torch.full((1, H, W), 7.0, device=dev)
        ws_bytes = int(lib.rdg_l1_dssim_workspace_bytes(3, H, W))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        terms = _lib.RdgLossTerms()
        terms.depth, terms.gt_depth, terms.dL_ddepth = dc.data_ptr(), d2c.data_ptr(), gdep.data_ptr()
        terms.w_pearson, terms.pearson_eps = w_p, 1e-6
        terms.alpha, terms.dL_dalpha, terms.w_alpha = alc.data_ptr(), galp.data_ptr(), w_a
        for _ in range(2):   # twice: the workspace header is reset by the call itself
            _lib.check(lib.rdg_losses(ac.data_ptr(), bc.data_ptr(), 3, H, W, 0.8, 0.2, C.byref(terms), out.data_ptr(),
                                      gcol.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
        o = out.cpu()
        assert abs(o[0].item() - photo.item()) < 2e-6
        assert abs(o[1].item() - lo.l1(a, b).item()) < 1e-6 and abs(o[2].item() - lo.ssim(a, b).item()) < 2e-6
        assert abs(o[3].item() - pear.item()) < 2e-6
        assert abs(o[4].item() - areg.item()) < 1e-7
        assert helpers.rel_err(gcol.cpu(), ar.grad) < GRAD_TOL
        assert helpers.rel_err(gdep.cpu(), dr.grad) < GRAD_TOL
        assert helpers.rel_err(galp.cpu(), alr.grad) < 1e-6
        # terms switched off: only the photometric entries are written
        out2 = torch.full((8,), -1.0, device=dev)
        _lib.check(lib.rdg_losses(ac.data_ptr(), bc.data_ptr(), 3, H, W, 0.8, 0.2, None, out2.data_ptr(),
                                  gcol.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
        assert abs(out2[0].item() - photo.item()) < 2e-6 and out2[3].item() == -1.0 and out2[4].item() == -1.0


def test_full_reference_loss_in_one_loss_stage():
    """Every reference config's objective on the rendered maps - 0.8 L1 + 0.2 D-SSIM (train_kubric_mrig.yaml:135-144) + 0.05
    global Pearson (:145-150) + 0.15 local Pearson over random 128x128 boxes (:151-158, losses.py:132-182) - from ONE
    rdg_losses call, value and gradients, against loss_oracle (pinned to the reference's loss_utils by losses.npz)."""
    import ctypes as C
    from oracle import loss_oracle as lo
    from rodygs_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(12)
    for (H, W, box_p, n_box) in ((300, 420, 128, 3), (270, 480, 64, 14), (160, 160, 128, 1)):
        a = torch.rand(3, H, W, generator=g)
        b = (a + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
        d1 = torch.rand(1, H, W, generator=g) * 5 + torch.linspace(0, 3, W).reshape(1, 1, W)
        d2 = (d1 * 0.7 + 0.5 * torch.rand(1, H, W, generator=g))
        x0 = torch.randint(0, H - box_p, (n_box,), generator=g)
        y0 = torch.randint(0, W - box_p, (n_box,), generator=g)
        for w_p in (0.05, 0.0):
            ar, dr = a.clone().requires_grad_(True), d1.clone().requires_grad_(True)
            photo = lo.photometric(ar, b)
            pear = lo.pearson_depth(dr, d2)
            local = lo.local_pearson_depth(dr, d2, x0, y0, box_p)
            (photo + w_p * pear + 0.15 * local).backward()
            dev = "cuda"
            ac, bc, dc, d2c = a.to(dev), b.to(dev), d1.to(dev), d2.to(dev)
            boxes = torch.stack([x0, y0, torch.full_like(x0, box_p), torch.full_like(x0, box_p)], 1).to(torch.int32).to(dev).contiguous()
            out = torch.full((8,), -1.0, device=dev)
            gcol = torch.empty(3, H, W, device=dev)
            gdep = torch.full((1, H, W), 7.0, device=dev)
            ws_bytes = int(lib.rdg_l1_dssim_workspace_bytes(3, H, W))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            terms = _lib.RdgLossTerms()
            terms.depth, terms.gt_depth, terms.dL_ddepth = dc.data_ptr(), d2c.data_ptr(), gdep.data_ptr()
            terms.w_pearson, terms.pearson_eps = w_p, 1e-6
            terms.local_boxes, terms.n_local_boxes, terms.w_local = boxes.data_ptr(), n_box, 0.15
            for _ in range(2):   # twice: the call resets its own scratch
                _lib.check(lib.rdg_losses(ac.data_ptr(), bc.data_ptr(), 3, H, W, 0.8, 0.2, C.byref(terms), out.data_ptr(),
                                          gcol.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
            o = out.cpu()
            assert abs(o[0].item() - photo.item()) < 2e-6
            if w_p:
                assert abs(o[3].item() - pear.item()) < 2e-6
            assert abs(o[5].item() - local.item()) < 2e-6, (o[5].item(), local.item())
            assert helpers.rel_err(gcol.cpu(), ar.grad) < GRAD_TOL
            assert helpers.rel_err(gdep.cpu(), dr.grad) < GRAD_TOL, (H, W, w_p)
            ok, ratio = helpers.elementwise_ok(gdep, dr.grad, rtol=1e-3, floor=1e-4)
            assert ok, f"dL/ddepth element-wise {ratio:.2f}"
            # the stand-alone batched kernel agrees with the fused term
            from rodygs_b200 import losses
            dd = d1.to(dev).requires_grad_(True)
            lp = losses.local_pearson_depth_loss(dd, d2c, box_p, origins=torch.stack([x0, y0], 1))
            assert abs(lp.item() - local.item()) < 2e-6


def test_sync_free_mode_matches_and_reports_overflow():
    H, W, n = 64, 64, 1500
    sc, cam = helpers.small_scene(n, H, W, 4, seed=2)
    acts = helpers.activated_concat(sc, cam)
    bg = torch.zeros(3)
    a = _run_boundary(acts, cam, bg, 1)["out"][0]
    engine.config.sync_free = True
    b = _run_boundary(acts, cam, bg, 1)["out"][0]
    assert torch.equal(a, b)
    torch.cuda.synchronize()
    engine._capacity[0]["pending"] = None     # (the lazy check would otherwise re-grow the capacity)
    engine._capacity[0]["cap"] = 128          # force an overflow on the next sync-free frame
    _run_boundary(acts, cam, bg, 1)
    with pytest.raises(RuntimeError, match="truncated"):
        _run_boundary(acts, cam, bg, 1)
    c = _run_boundary(acts, cam, bg, 1)["out"][0]   # capacity has been raised
    assert torch.equal(a, c)


def test_render_wrapper_with_the_reference_signature():
    """rodygs_b200.renderer.render (mirror of /root/reference/src/trainer/renderer.py:17-114): same keywords, same return
    dict, camera object with FoVx / FoVy / image_height / image_width / projection_matrix / world_view_transform; checked
    against the oracle (forward + gradients incl. the retained viewspace_points.grad and the pose gradient through
    world_view_transform) and against the fused render_dynamic on the same scene."""
    from types import SimpleNamespace
    from rodygs_b200.dynamic import GaussianParams, render_dynamic
    from rodygs_b200.renderer import render
    H, W, n, T = 80, 112, 2000, 5
    sc, cam = helpers.small_scene(n, H, W, T, seed=44)
    acts = helpers.activated_concat(sc, cam)
    bg = torch.tensor([0.0, 0.0, 0.0])
    up = _upstream_grads(H, W, seed=8)
    orc = _run_oracle(acts, cam, bg, 3, grads=up)
    xyz, op, scl, rot, feat = [t.detach().clone().cuda().requires_grad_(True) for t in acts]
    wvt = cam.world_view_transform.clone().cuda().requires_grad_(True)          # mathematical V; render() transposes it
    vp = SimpleNamespace(FoVx=cam.FoVx, FoVy=cam.FoVy, image_height=H, image_width=W,
                         projection_matrix=cam.projection_matrix.cuda(), world_view_transform=wvt)
    pkg = render(xyz, 3, op, scl, rot, feat, vp, bg.cuda(), enable_sh_grad=True, enable_cov_grad=True)
    assert set(pkg) == {"rendered_image", "rendered_depth", "rendered_normal", "rendered_alpha", "viewspace_points",
                        "visibility_filter", "radii", "extra"}                   # renderer.py:103-114
    assert pkg["rendered_depth"].shape == (1, H, W) and pkg["rendered_normal"].shape == (3, H, W) and pkg["extra"] is None
    _check_images((pkg["rendered_image"], pkg["rendered_depth"], None, pkg["rendered_alpha"]), orc["out"])
    assert torch.equal(pkg["radii"].cpu(), orc["out"].radii)
    assert torch.equal(pkg["visibility_filter"].cpu(), orc["out"].radii > 0)
    ((pkg["rendered_image"] * up[0].cuda()).sum() + (pkg["rendered_depth"] * up[1].cuda()).sum()
     + (pkg["rendered_alpha"] * up[2].cuda()).sum()).backward()
    got = {"means3D": xyz.grad, "means2D": pkg["viewspace_points"].grad, "shs": feat.grad, "opacities": op.grad,
           "scales": scl.grad, "rotations": rot.grad, "viewmatrix": wvt.grad.t()}
    for k, ref in orc["grads"].items():
        assert got[k] is not None, k
        err = helpers.rel_err(got[k].cpu(), ref)
        assert err <= GRAD_TOL, f"render(): grad {k}: rel err {err:.3e}"
    # densification statistic the trainer reads (rodygs.py:322-331): norm of viewspace_points.grad[:, :2]
    stat = pkg["viewspace_points"].grad[:, :2].norm(dim=-1)
    assert torch.allclose(stat.cpu(), orc["grads"]["means2D"][:, :2].norm(dim=-1), rtol=2e-3, atol=1e-6 * float(stat.max()))
    # the fused path renders the same scene from the raw parameters
    cst = GaussianParams(**{k: v.cuda() for k, v in sc["static"].items()})
    cdy = GaussianParams(**{k: v.cuda() for k, v in sc["dynamic"].items()})
    fused = render_dynamic(cst, cdy, _settings(cam, bg, 3), cam.world_view_transform.t().contiguous().cuda(),
                           sc["motion_coeff"].cuda(), sc["table"][cam.time_index].cuda(), sc["table"].cuda(),
                           sc["time_ind"].cuda(), sc["spatial_lr_scale"], True)
    assert set(fused) == set(pkg)
    for k in ("rendered_image", "rendered_depth", "rendered_alpha"):
        assert (fused[k] - pkg[k]).abs().max().item() <= 2e-3, k               # activations differ by exp() / sigmoid ulps
    assert (fused["radii"] != pkg["radii"]).float().mean().item() <= 2e-3
