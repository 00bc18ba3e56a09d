"""GPU parity of the motion regularisers through the C ABI:
against the reference's own classes (tests/golden/motion.npz) and against the oracle on larger seeded inputs.
Tolerances: neighbour indices and squared distances bit-exact; losses 1e-5 relative; gradients 1e-3 relative
(max|a-b| / max|b| per tensor; float atomics make their last bits run-to-run nondeterministic)."""
import random

import pytest
import torch

from helpers import rel_err
from test_oracle_motion import load_motion
from oracle import motion_oracle as mo

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3


def _dev(g, *names):
    return [g[k].cuda().clone().requires_grad_(True) for k in names]


def _check(tag, g, loss, leaves):
    ref = g[f"{tag}/loss"].item()
    assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref)), (tag, loss.item(), ref)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    for k, gr in zip(leaves, grads):
        key = f"{tag}/d_{k}"
        if key not in g:
            assert gr is None or gr.abs().max().item() == 0, (tag, k)
            continue
        assert rel_err(gr.cpu(), g[key]) < GRAD_TOL, (tag, k, rel_err(gr.cpu(), g[key]))


def test_coefficient_regularisers_match_reference():
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    (coeff,) = _dev(g, "coeff")
    _check("motion_l1", g, mr.motion_l1_loss(coeff), {"coeff": coeff})
    (coeff,) = _dev(g, "coeff")
    _check("motion_sparsity", g, mr.motion_sparsity_loss(coeff), {"coeff": coeff})


def test_coefficient_regularisers_in_place_accumulate():
    from rodygs_b200 import motion_reg as mr
    gen = torch.Generator().manual_seed(3)
    c = (torch.randn(50_001, 1, 16, generator=gen) * 0.2)
    c[5] = 0                                               # an all-zero row: |c| has a zero sub-gradient
    cc = c.clone().requires_grad_(True)
    ref = 0.01 * mo.motion_l1(cc) + 0.002 * mo.motion_sparsity(cc)
    (gref,) = torch.autograd.grad(ref, cc)
    d0 = torch.randn(c.shape, generator=gen) * gref.abs().max()      # same magnitude as the gradient, so the sum keeps its bits
    d = d0.clone().cuda()
    parts = mr.motion_coeff_reg_(c.cuda(), 0.01, 0.002, d, accumulate=True)
    assert abs(parts[0].item() - mo.motion_l1(c).item()) < 1e-6
    assert abs(parts[1].item() - mo.motion_sparsity(c).item()) < 1e-6
    assert rel_err(d.cpu() - d0, gref) < GRAD_TOL


@pytest.mark.parametrize("tag,fmode", [("basis_cum_exponential", "cum_exponential"), ("basis_vanilla", "vanilla")])
def test_basis_regulariser_matches_reference(tag, fmode):
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    (table,) = _dev(g, "table")

    class M:
        def get_total_motion_table(self):
            return table
    reg = mr.MotionBasisRegularizaiton(0, 0, fmode)
    assert torch.allclose(reg.reg_coeff.cpu(), g[f"{tag}/reg_coeff"], rtol=1e-6)
    _check(tag, g, reg(M()), {"table": table})


def test_basis_regulariser_disabled_terms():
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    reg_c = mr.basis_reg_coeff("laplacian")
    for td, rd in ((0, -1), (-1, 0)):
        t_ref = g["table"].clone().requires_grad_(True)
        ref = mo.motion_basis_reg(t_ref, reg_c, td, rd)
        (gref,) = torch.autograd.grad(ref, t_ref)
        (table,) = _dev(g, "table")
        loss = mr._BasisRegFn.apply(table, reg_c.cuda(), td, rd)
        (gr,) = torch.autograd.grad(loss, table)
        assert abs(loss.item() - ref.item()) < 1e-5
        assert rel_err(gr.cpu(), gref) < GRAD_TOL


def _clouds():
    gen = torch.Generator().manual_seed(11)
    uni = torch.rand(6000, 3, generator=gen) * 4 - 2
    centres = torch.randn(12, 3, generator=gen) * 3
    clustered = centres[torch.randint(0, 12, (5000,), generator=gen)] + torch.randn(5000, 3, generator=gen) * \
        torch.rand(5000, 1, generator=gen) * 0.3
    planar = torch.rand(3000, 3, generator=gen)
    planar[:, 2] = 0.25
    dup = torch.rand(2000, 3, generator=gen)
    dup[1000:1200] = dup[:200]                              # clones: exact ties at distance 0
    line = torch.zeros(1500, 3)
    line[:, 0] = torch.linspace(0, 1, 1500)                 # equal spacing: exact ties between left and right
    same = torch.ones(64, 3) * 0.7                          # every distance is 0
    return {"uniform": uni, "clustered": clustered, "planar": planar, "duplicates": dup, "line": line, "same": same}


@pytest.mark.parametrize("name", ["uniform", "clustered", "planar", "duplicates", "line", "same"])
@pytest.mark.parametrize("K", [8, 3, 16])
def test_knn_bit_exact(name, K):
    from rodygs_b200 import motion_reg as mr
    p = _clouds()[name]
    d_ref, i_ref = mo.knn_points(p, K)
    d, i = mr.knn_points(p.cuda(), K)
    assert torch.equal(i.cpu().long(), i_ref), name
    assert torch.equal(d.cpu(), d_ref), name


def test_knn_rejects_bad_arguments():
    from rodygs_b200 import motion_reg as mr
    with pytest.raises(RuntimeError):
        mr.knn_points(torch.rand(4, 3).cuda(), 8)           # fewer points than neighbours
    with pytest.raises(RuntimeError):
        mr.knn_points(torch.rand(40, 3).cuda(), 17)
    with pytest.raises(RuntimeError):
        mr.knn_points(torch.rand(40, 3), 8)                 # CPU tensor: no fallback


def test_knn_full_size_against_exhaustive_subset():
    """BASELINE config 4: 500 K sampled dynamic Gaussians in the frustum slab; 512 random queries are checked
    against an exhaustive scan, all rows for sortedness and self-first."""
    from rodygs_b200 import motion_reg as mr
    gen = torch.Generator().manual_seed(2)
    n = 500_000
    z = torch.rand(n, generator=gen) * 6 + 2
    p = torch.stack([(torch.rand(n, generator=gen) * 2 - 1) * 0.6 * z, (torch.rand(n, generator=gen) * 2 - 1) * 0.35 * z, z], 1)
    pc = p.cuda()
    d, i = mr.knn_points(pc, 8)
    assert (d[:, 1:] >= d[:, :-1]).all()
    assert (d[:, 0] == 0).all() and (i[:, 0].long() == torch.arange(n, device="cuda")).all()
    q = torch.randint(0, n, (512,), generator=gen).cuda()
    diff = pc[q][:, None, :] - pc[None, :, :]
    full = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    ref_d, ref_i = full.sort(dim=1, stable=True)
    assert torch.equal(i[q].long(), ref_i[:, :8])
    assert torch.allclose(d[q], ref_d[:, :8], rtol=1e-6, atol=0)


@pytest.mark.parametrize("tag,mode", [("rigid_cfg", ("distance_preserving", "surface")), ("rigid_surface", ("surface",)),
                                      ("rigid_dp", ("distance_preserving",))])
def test_rigidity_matches_reference(tag, mode):
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    xyz, coeff, table, pred_tr = _dev(g, "xyz", "coeff", "table", "pred_tr")
    ti = g.get(f"{tag}/time_indices")
    loss = mr.rigidity_loss(xyz, coeff, pred_tr, table, g[f"{tag}/indice"], ti, K=int(g["K"]), mode=mode)
    _check(tag, g, loss, {"xyz": xyz, "coeff": coeff, "table": table, "pred_tr": pred_tr})


def test_rigidity_mode_coeff_is_refused():
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    xyz, coeff, table, pred_tr = _dev(g, "xyz", "coeff", "table", "pred_tr")
    with pytest.raises(NotImplementedError):
        mr.rigidity_loss(xyz, coeff, pred_tr, table, g["rigid_coeff/indice"], None, mode=("coeff",))


def test_rigidity_against_oracle_larger():
    from rodygs_b200 import motion_reg as mr
    gen = torch.Generator().manual_seed(21)
    N, T, B, K = 9000, 40, 16, 8
    xyz = torch.randn(N, 3, generator=gen)
    coeff = torch.randn(N, 1, B, generator=gen) * 0.2
    table = torch.randn(T, B, 7, generator=gen) * 0.1
    pred = torch.randn(N, 3, generator=gen) * 0.02
    indice = torch.randperm(N, generator=gen)[: N // 2]
    ti = torch.randint(0, T - 1, (T // 4,), generator=gen)
    lv = [t.clone().requires_grad_(True) for t in (xyz, coeff, table, pred)]
    ref, _ = mo.rigidity(lv[0], lv[1], torch.zeros(N, 1, 3), lv[3], lv[2], indice, ti, K=K)
    gref = torch.autograd.grad(ref, lv)
    dv = [t.cuda().requires_grad_(True) for t in (xyz, coeff, table, pred)]
    loss = mr.rigidity_loss(dv[0], dv[1], dv[3], dv[2], indice, ti, K=K)
    gd = torch.autograd.grad(loss, dv)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    for name, a, b in zip(("xyz", "coeff", "table", "pred_translation"), gd, gref):
        assert rel_err(a.cpu(), b) < GRAD_TOL, (name, rel_err(a.cpu(), b))


def test_rigidity_class_draws_like_the_reference():
    """RigidityLoss.forward makes the reference's two draws (random.sample, torch.randint on the CPU generator)."""
    from rodygs_b200 import motion_reg as mr
    g = load_motion()
    xyz, coeff, table, pred_tr = _dev(g, "xyz", "coeff", "table", "pred_tr")

    class M:
        _xyz, _motion_coeff = xyz, coeff
        unique_times = list(range(table.shape[0]))

        def get_total_motion_table(self):
            return table
    random.seed(99)
    torch.manual_seed(99)
    loss = mr.RigidityLoss(K=int(g["K"]), mode=["distance_preserving", "surface"])(M(), pred_tr)
    _check("rigid_cfg", g, loss, {"xyz": xyz, "coeff": coeff, "table": table, "pred_tr": pred_tr})


def test_trainer_motion_losses_add_reference_gradients():
    """SplatTrainStep.motion_losses_backward on the flat buffers == the reference's weighted sum of the four
    regularisers (configs/train/train_kubric_mrig.yaml:198-237) differentiated by autograd on the CPU, with
    pred_translation = spatial_lr_scale * c . (B(t) - B(t_i))[:3] (rodygs_dynamic.py:122-138)."""
    from rodygs_b200 import synthetic, trainer, motion_reg as mr
    H, W, T = 64, 96, 12
    sc = synthetic.make_scene(6000, H, W, T, seed=4, radius_px=4.0)
    sc["spatial_lr_scale"] = 1.7
    step = trainer.SplatTrainStep(sc, H, W)
    nd = step.nd
    gen = torch.Generator().manual_seed(8)
    basis_t = (torch.randn(16, 7, generator=gen) * 0.05)
    indice = torch.randperm(nd, generator=gen)[: nd // 2]
    fi = torch.randint(0, T - 1, (T // 4,), generator=gen)
    w = dict(w_motion_l1=0.01, w_sparsity=0.002, w_basis=0.1, w_rigidity=0.5)
    # --- reference composition on the CPU ---
    xyz = sc["dynamic"]["xyz"].clone().requires_grad_(True)
    coeff = sc["motion_coeff"].clone().requires_grad_(True)          # [nd, 1, 16]
    table = sc["table"].clone().requires_grad_(True)
    bt = basis_t.clone().requires_grad_(True)
    ti = sc["time_ind"].long()
    delta = (coeff.squeeze(1)[:, :, None] * (bt[None] - table[ti])).sum(1)         # [nd, 7]
    pred = delta[:, :3] * sc["spatial_lr_scale"]
    rig, rparts = mo.rigidity(xyz, coeff, torch.zeros(nd, 1, 3), pred, table, indice, fi, K=8)
    reg = mr.basis_reg_coeff("cum_exponential")
    total = (w["w_motion_l1"] * mo.motion_l1(coeff) + w["w_sparsity"] * mo.motion_sparsity(coeff)
             + w["w_basis"] * mo.motion_basis_reg(table, reg) + w["w_rigidity"] * rig)
    gref = torch.autograd.grad(total, [xyz, coeff, table, bt])
    # --- CUDA path: gradients are added to what the buffers hold ---
    step.grads.fill_(0.0)
    parts = step.motion_losses_backward(5, basis_t.cuda(), indice=indice, frame_indices=fi, **w)
    assert abs(parts[0].item() - mo.motion_l1(coeff).item()) < 1e-6
    assert abs(parts[4].item() - rparts["surface"].item()) < 1e-5
    assert abs(parts[5].item() - rparts["distance_preserving"].item()) < 1e-5
    got = [step.g("dynamic.xyz"), step.g("motion_coeff"), step.g("table"), step.g("basis_t")]
    for name, a, b in zip(("xyz", "coeff", "table", "basis_t"), got, gref):
        assert rel_err(a.cpu().reshape(b.shape), b) < GRAD_TOL, (name, rel_err(a.cpu().reshape(b.shape), b))
    # rigidity is skipped off its frequency (MultiLoss gating), the other three still run
    step.grads.fill_(0.0)
    parts = step.motion_losses_backward(6, basis_t.cuda(), indice=indice, frame_indices=fi, **w)
    assert parts[4].item() == 0 and parts[5].item() == 0 and parts[0].item() > 0
    assert step.g("dynamic.xyz").abs().max().item() == 0
