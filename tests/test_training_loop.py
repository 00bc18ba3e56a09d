"""Optimiser + densification rows around the hot path: the per-group Adam step against
torch.optim.Adam built the reference's way, the learning-rate schedule against the reference's function, and a
short training loop with statistics, densify_and_prune and an opacity reset on the flat buffers."""
import math

import numpy as np
import pytest
import torch

from rodygs_b200 import optim as ro


def test_expon_lr_matches_reference_formula():
    """get_expon_lr_func (general_utils.py:40-73), restated here with numpy exactly as the reference writes it."""
    def ref(step, lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)

    for step in (-1, 0, 1, 500, 19999, 20000, 50000):
        for kw in ({}, {"lr_delay_steps": 100, "lr_delay_mult": 0.01}):
            a = ro.expon_lr(step, 1.6e-4 * 2.5, 1.6e-6 * 2.5, max_steps=20000, **kw)
            b = ref(step, 1.6e-4 * 2.5, 1.6e-6 * 2.5, max_steps=20000, **kw)
            assert math.isclose(a, float(b), rel_tol=1e-12, abs_tol=0.0), (step, kw)
    assert ro.expon_lr(10, 0.0, 0.0) == 0.0


def test_group_lrs_follow_the_reference_setup():
    """rodygs_static.py:106-141: f_rest = feature_lr / 20, xyz scaled by spatial_lr_scale and scheduled."""
    lrs = ro.GaussianLRs(motion_coeff_lr=1.6e-4).group_lrs(0, 2.0)
    assert math.isclose(lrs["xyz"], 1.6e-4 * 2.0, rel_tol=1e-12)
    assert lrs["f_rest"] == 0.0025 / 20.0 and lrs["opacity"] == 0.05 and lrs["motion_coeff"] == 1.6e-4
    assert math.isclose(ro.GaussianLRs().group_lrs(20000, 2.0)["xyz"], 1.6e-6 * 2.0, rel_tol=1e-9)
    assert "motion_coeff" not in ro.GaussianLRs().group_lrs(0, 1.0)


# ---------------------------------------------------------------------------------------- GPU

def _make_step(N=30_000, H=160, W=224, T=6, seed=3):
    from rodygs_b200 import synthetic
    from rodygs_b200.trainer import SplatTrainStep
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=seed), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.0)
    cam = synthetic.make_camera(1, 8, H, W, T)
    vm = cam.world_view_transform.t().contiguous().cuda()
    pm = cam.projection_matrix.t().contiguous().cuda()
    gt = torch.rand(3, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.5 + 0.25
    return step, cam, vm, pm, gt


@pytest.mark.gpu
def test_group_adam_matches_torch_adam_groups():
    """One rdg_adam_groups launch == torch.optim.Adam over the reference's groups, over several steps of a moving schedule."""
    step, cam, vm, pm, gt = _make_step(N=8_000)
    lrs = {"static": ro.GaussianLRs(position_lr_max_steps=30), "dynamic": ro.GaussianLRs(position_lr_max_steps=30, scaling_lr=0.001,
                                                                                       motion_coeff_lr=1.6e-4)}
    for tag in ("static", "dynamic"):
        step.attach_optimizer(tag, lrs[tag])
    # torch side: the same tensors as nn.Parameters, groups as in optim_setup / append_motion_optim
    ref_params, ref_opts = {}, {}
    for tag in ("static", "dynamic"):
        groups = []
        for field, group in ro.GROUP_OF.items():
            p = torch.nn.Parameter(step.p(f"{tag}.{field}").clone())
            ref_params[(tag, group)] = p
            groups.append({"params": [p], "lr": 0.0, "name": group})
        if tag == "dynamic":
            p = torch.nn.Parameter(step.p("motion_coeff").clone())
            ref_params[(tag, "motion_coeff")] = p
            groups.append({"params": [p], "lr": 0.0, "name": "motion_coeff"})
        ref_opts[tag] = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    gen = torch.Generator(device="cuda").manual_seed(1)
    inv = {g: f for f, g in ro.GROUP_OF.items()}
    for it in range(1, 6):
        step.grads.copy_(torch.randn(step.grads.shape, device="cuda", generator=gen) * 1e-3)
        for tag in ("static", "dynamic"):
            lr_now = lrs[tag].group_lrs(it, step.spatial_lr_scale)
            for grp in ref_opts[tag].param_groups:
                grp["lr"] = lr_now[grp["name"]]
                name = "motion_coeff" if grp["name"] == "motion_coeff" else f"{tag}.{inv[grp['name']]}"
                grp["params"][0].grad = step.g(name).clone()
            ref_opts[tag].step()
            step.optimizer_step(tag, it)
    for (tag, group), p in ref_params.items():
        name = "motion_coeff" if group == "motion_coeff" else f"{tag}.{inv[group]}"
        got = step.p(name)
        # m / sqrt(v) with eps = 1e-15 carries a few ulp; the absolute error scales with the group's step size
        atol = 1e-7 + 1e-5 * 5 * lrs[tag].group_lrs(1, step.spatial_lr_scale)[group]
        assert torch.allclose(got, p.detach(), rtol=2e-5, atol=atol), (tag, group, float((got - p.detach()).abs().max()))
    # untouched ranges (table, padding) stay as they were
    assert float(step.optim["static"].exp_avg[step.layout["table"][0]:step.layout["table"][0] + 10].abs().sum()) == 0.0


@pytest.mark.gpu
def test_training_loop_with_densification_and_opacity_reset():
    """rodygs.py:198-369 on the flat buffers: forward/backward, statistics, Adam, densify_and_prune (N changes; moments of the
    survivors are kept, the other model is untouched), opacity reset; the loss keeps falling afterwards."""
    step, cam, vm, pm, gt = _make_step()
    step.attach_optimizer("static", ro.GaussianLRs(feature_lr=0.02))
    step.attach_optimizer("dynamic", ro.GaussianLRs(feature_lr=0.02, scaling_lr=0.001, motion_coeff_lr=1.6e-4))
    step.enable_densification("static")
    step.enable_densification("dynamic")
    losses = []

    def one_iter(it, tag):
        bt = step.p("table")[cam.time_index].clone()
        lp = step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, None)
        losses.append(float(lp[0]))
        step.add_densification_stats(tag)
        step.optimizer_step(tag, it)

    for it in range(1, 7):
        one_iter(it, "static")
        one_iter(it, "dynamic")
    st = step.stats["dynamic"]
    vis = int((step.last_outputs[3][step.ns:] > 0).sum())
    assert float(st.denom.max()) == 6.0 and int((st.denom > 0).sum()) >= vis > 0
    assert float(st.max_radii2D.max()) >= 1.0 and float(st.grad_accum.sum()) > 0.0

    ns0, nd0 = step.ns, step.nd
    static_before = {k: step.p(f"static.{k}").clone() for k in ("xyz", "features_rest")}
    static_m_before = step.optim["static"].moments("xyz")[0].clone()
    dyn_xyz, dyn_m = step.p("dynamic.xyz").clone(), step.optim["dynamic"].moments("xyz")[0].clone().view(-1, 3)
    thr = float(torch.quantile((st.grad_accum / st.denom.clamp(min=1))[st.denom > 0], 0.7))
    info = step.densify_and_prune("dynamic", thr, 0.005, 4.0, None, generator=torch.Generator(device="cuda").manual_seed(0))
    assert info["clones"] + info["split_selected"] > 0 and step.nd == info["rows"] != nd0 and step.ns == ns0
    assert step.time_ind.shape[0] == step.nd and step.stats["dynamic"].n == step.nd and step.stats["static"].n == ns0
    # the static model and its optimiser state moved to the new flat buffers unchanged
    for k, t in static_before.items():
        assert torch.equal(step.p(f"static.{k}"), t), k
    assert torch.equal(step.optim["static"].moments("xyz")[0], static_m_before)
    assert step.optim["dynamic"].steps == 6
    # survivors come first and in order: their rows and moments are copies, new rows start with zero moments
    a = info["survivors"]
    new_m = step.optim["dynamic"].moments("xyz")[0].view(-1, 3)
    assert float(new_m[a:].abs().sum()) == 0.0
    keep = torch.isin(dyn_xyz.view(-1, 3)[:, 0], step.p("dynamic.xyz")[:a, 0])
    assert int(keep.sum()) >= a and torch.equal(step.p("dynamic.xyz")[:a], dyn_xyz[keep][:a]) and torch.equal(new_m[:a], dyn_m[keep][:a])

    step.reset_opacity("static")
    assert float(torch.sigmoid(step.p("static.opacity")).max()) <= 0.01 + 1e-6
    assert float(step.optim["static"].moments("opacity")[0].abs().sum()) == 0.0
    first_after = None
    for it in range(7, 19):
        one_iter(it, "static")
        one_iter(it, "dynamic")
        if first_after is None:
            first_after = losses[-2]
            # the reference's optimizer.step() right after densify_and_prune is a no-op for that model (fresh Parameters have
            # no .grad): no update, no step count
            assert step.optim["dynamic"].steps == 6 and step.optim["static"].steps == 7
    assert step.optim["dynamic"].steps == 6 + 11
    assert all(math.isfinite(x) for x in losses)
    assert losses[-1] < first_after, (first_after, losses[-1])


def _dp_densify_worker(rank, world, port, ret):
    """Two ranks on ONE GPU (gloo carries the CUDA tensors through the host): each renders its own views, then both densify."""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rodygs_b200 import synthetic
        step, cam, vm, pm, gt = _make_step(N=12_000, H=96, W=128)
        step.attach_optimizer("dynamic", ro.GaussianLRs(scaling_lr=0.001, motion_coeff_lr=1.6e-4))
        step.enable_densification("dynamic")                       # built WITHOUT a process_group: the default group is used
        for it in range(3):
            c = synthetic.make_camera(2 * it + rank, 8, 96, 128, 6)   # every rank sees different views
            bt = step.p("table")[c.time_index].clone()
            step.forward_backward(c.world_view_transform.t().contiguous().cuda(), c.projection_matrix.t().contiguous().cuda(),
                                  c.tanfovx, c.tanfovy, bt, gt, None)
            step.add_densification_stats("dynamic")
        st = step.stats["dynamic"]
        local = float(st.grad_accum.sum())
        thr = 0.5 * float((st.grad_accum / st.denom.clamp(min=1))[st.denom > 0].mean())
        thr_t = torch.tensor([thr], dtype=torch.float64)
        dist.broadcast(thr_t, 0)                                   # same threshold everywhere; the statistics do the rest
        info = step.densify_and_prune("dynamic", float(thr_t), 0.005, 4.0, None,
                                      generator=torch.Generator(device="cuda").manual_seed(0))
        ret.put((rank, step.nd, info["clones"], info["split_selected"], local, float(step.p("dynamic.xyz").double().sum()),
                 float(step.p("motion_coeff").double().abs().sum()), int(step.time_ind.long().sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_data_parallel_ranks_densify_identically_on_the_default_group():
    """without the all-reduce of the statistics the ranks clone / split / prune different rows and
    their Gaussian counts diverge (the next collective then hangs or corrupts)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_dp_densify_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([ret.get(timeout=300), ret.get(timeout=300)])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = res
    assert a[4] != b[4], "the two ranks were supposed to see different views"
    assert a[1:4] == b[1:4] and a[5:] == b[5:], f"ranks diverged: {a} vs {b}"
    assert a[2] + a[3] > 0
