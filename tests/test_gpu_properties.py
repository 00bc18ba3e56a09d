"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (where the
CPU oracle would take hours): sortedness, consistency of the binning tables, conservation
identities, determinism of the forward pass and linearity of the backward pass."""
import pytest
import torch

from rodygs_b200 import _lib, engine, synthetic
from rodygs_b200.trainer import SplatTrainStep

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset():
    yield
    engine.config.sync_free = False
    engine.config.debug_keep_unsorted = False
    engine.config.binning = "tiles"


def _step(cfg, n_override=None):
    N, H, W, T, _ = synthetic.CONFIGS[cfg]
    if n_override:
        N = n_override
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=0), "cuda")
    return SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01), synthetic.make_camera(3, 8, H, W, T), (N, H, W, T)


def _forward(step, cam):
    vm = cam.world_view_transform.t().contiguous().cuda()
    pm = cam.projection_matrix.t().contiguous().cuda()
    step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, step.p("table")[cam.time_index].contiguous(), None, None,
                          forward_only=True)
    return step.last_outputs, step.last_state


@pytest.mark.parametrize("cfg,binning", [("c2_kubric", "tiles"), ("c2_kubric", "lsd"), ("c4_iphone", "tiles"), ("c4_iphone", "lsd")])
def test_binning_tables_are_consistent_at_full_size(cfg, binning):
    engine.config.binning = binning
    step, cam, (N, H, W, T) = _step(cfg)
    (color, depth, alpha, radii), st = _forward(step, cam)
    D = int(st.num_rendered[0])
    assert int(st.num_rendered[1]) == 0
    tiles_touched = st.geom["tiles_touched"].long()
    assert int(tiles_touched.sum()) == D
    if binning == "lsd":
        assert torch.equal(st.extras["point_offsets"].long(), torch.cumsum(tiles_touched, 0))
    assert torch.equal(radii > 0, tiles_touched > 0)
    keys = st.extras["keys_sorted"][:D]
    assert bool((keys[1:] >= keys[:-1]).all()), "keys not sorted"
    vals = st.vals_sorted[:D].long()
    assert int(vals.min()) >= 0 and int(vals.max()) < N
    # the multiset of values = every visible Gaussian repeated tiles_touched times (checksum of checksums)
    cnt = torch.bincount(vals, minlength=N)
    assert torch.equal(cnt, tiles_touched)
    # keys carry the Gaussian's own depth bits
    dbits = st.geom["p2"][:, 1].contiguous().view(torch.int32).long() & 0xFFFFFFFF
    assert torch.equal(keys & 0xFFFFFFFF, dbits[vals])
    # tile ranges partition [0, D) in tile order and agree with the keys' tile ids
    tile_of = keys >> 32
    gx, gy = (W + 15) // 16, (H + 15) // 16
    ranges = st.ranges.long()
    ne = ranges[:, 1] > ranges[:, 0]
    assert int((ranges[:, 1] - ranges[:, 0]).sum()) == D
    assert torch.equal(torch.bincount(tile_of, minlength=gx * gy), ranges[:, 1] - ranges[:, 0])
    starts = ranges[ne, 0]
    assert torch.equal(tile_of[starts], torch.nonzero(ne).squeeze(1))
    # stable order: inside a tile, equal depth bits keep ascending Gaussian index
    same = (keys[1:] == keys[:-1])
    assert bool((vals[1:][same] > vals[:-1][same]).all())
    # blend identities
    assert torch.allclose(alpha[0], 1.0 - st.final_T, atol=1e-6)
    assert bool((st.n_contrib.long().flatten() <= (ranges[:, 1] - ranges[:, 0]).max()).all())
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0 + 1e-6


def test_forward_is_deterministic_and_sync_free_equals_default():
    step, cam, _ = _step("c2_kubric")
    (c1, d1, a1, r1), s1 = _forward(step, cam)
    c1, d1, v1 = c1.clone(), d1.clone(), s1.vals_sorted[: int(s1.num_rendered[0])].clone()
    engine.config.sync_free = True
    (c2, d2, a2, r2), s2 = _forward(step, cam)
    assert torch.equal(c1, c2) and torch.equal(d1, d2) and torch.equal(r1, r2)
    assert torch.equal(v1, s2.vals_sorted[: int(s2.num_rendered[0])])
    # the two binning implementations agree bit for bit
    engine.config.binning = "lsd"
    (c3, d3, a3, r3), s3 = _forward(step, cam)
    assert torch.equal(c1, c3) and torch.equal(v1, s3.vals_sorted[: int(s3.num_rendered[0])])
    assert torch.equal(s2.ranges, s3.ranges)


def test_backward_is_linear_in_the_upstream_gradient():
    """grads(a*g1 + g2) == a*grads(g1) + grads(g2) up to float-atomic reordering."""
    N, H, W, T = 60_000, 256, 256, 8
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=1), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3)
    cam = synthetic.make_camera(2, 8, H, W, T)
    (_, _, _, _), st = _forward(step, cam)
    g = torch.Generator(device="cuda").manual_seed(0)
    ups = [(torch.randn(3, H, W, device="cuda", generator=g), torch.randn(1, H, W, device="cuda", generator=g),
            torch.randn(1, H, W, device="cuda", generator=g)) for _ in range(2)]

    def run(gc, gd, ga):
        step.grads.zero_()
        step.view_grad.zero_()
        grads = engine.SceneGrads(st=step._setgrad("static"), dy=step._setgrad("dynamic"), means2D=step.means2D_grad,
                                  viewmatrix=step.view_grad, motion_coeff=step.g("motion_coeff").view(step.nd, 16),
                                  table=step.g("table"), basis_t=step.g("basis_t"))
        engine.render_backward(st, gc, gd, ga, grads)
        return step.grads.clone(), step.view_grad.clone()

    a = 0.37
    g1, v1 = run(*ups[0])
    g2, v2 = run(*ups[1])
    g3, v3 = run(*[a * x + y for x, y in zip(ups[0], ups[1])])
    ref, refv = a * g1 + g2, a * v1 + v2
    assert (g3 - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert (v3 - refv).abs().max().item() <= 2e-4 * refv.abs().max().item()
    # zero upstream gradient -> exactly zero parameter gradients
    z, zv = run(torch.zeros_like(ups[0][0]), torch.zeros_like(ups[0][1]), torch.zeros_like(ups[0][2]))
    assert z.abs().max().item() == 0.0 and zv.abs().max().item() == 0.0


def test_training_step_reduces_the_loss():
    """A few fused-Adam steps on the flat buffers lower the photometric loss (end-to-end sanity)."""
    from rodygs_b200 import _lib
    N, H, W, T = 40_000, 192, 256, 6
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=2), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.0)
    cam = synthetic.make_camera(1, 8, H, W, T)
    vm = cam.world_view_transform.t().contiguous().cuda()
    pm = cam.projection_matrix.t().contiguous().cuda()
    bt = step.p("table")[cam.time_index].clone()
    gt = torch.rand(3, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.5 + 0.25
    lib = _lib.load()
    m, v = torch.zeros_like(step.params), torch.zeros_like(step.params)
    losses = []
    for it in range(1, 13):
        lp = step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, None)
        losses.append(float(lp[0]))
        _lib.check(lib.rdg_adam(step.params.data_ptr(), step.grads.data_ptr(), m.data_ptr(), v.data_ptr(), step.params.numel(),
                                5e-3, 0.9, 0.999, 1e-15, it, 1.0, _lib.stream_ptr()))
    assert losses[-1] < 0.9 * losses[0], losses


def test_fused_adam_matches_torch_adam():
    """rdg_adam == torch.optim.Adam(eps=1e-15) (rodygs_static.py:106-149) on a flat buffer."""
    from rodygs_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(7)
    p0 = torch.randn(100_003, device="cuda", generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1.6e-4, eps=1e-15)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for it in range(1, 5):
        grad = torch.randn(p0.shape, device="cuda", generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        _lib.check(lib.rdg_adam(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1.6e-4, 0.9, 0.999,
                                1e-15, it, 1.0, _lib.stream_ptr()))
    assert torch.allclose(p, p_ref.detach(), rtol=1e-5, atol=1e-7)


def test_factored_sh_exchange_equals_direct_sum_over_views():
    """The data-parallel exchange (12-byte factors of dL/dSH + rdg_sh_grad_views) gives the same mean gradient
    over a step's views as summing the per-view gradients that the plain backward writes - here with one rank
    holding all views, which exercises preprocess_bwd's factor output and the rebuild kernel end to end."""
    N, H, W, T, VIEWS = 50_000, 192, 256, 6, 3
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=4), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01)
    assert step.n_local == 1
    step.local_box_origins = torch.tensor([[10, 20]])      # both passes below must see the same local-Pearson box
    gen = torch.Generator(device="cuda").manual_seed(9)
    views = []
    for v in range(VIEWS):
        cam = synthetic.make_camera(v, 8, H, W, T)
        vm = cam.world_view_transform.t().contiguous().cuda()
        pm = cam.projection_matrix.t().contiguous().cuda()
        bt = step.p("table")[cam.time_index].clone()
        gt = torch.rand(3, H, W, device="cuda", generator=gen)
        gtd = torch.rand(1, H, W, device="cuda", generator=gen)
        views.append((cam, vm, pm, bt, gt, gtd))
    # reference: plain backward per view, summed
    ref = torch.zeros_like(step.grads)
    for cam, vm, pm, bt, gt, gtd in views:
        step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
        ref += step.grads
    ref /= VIEWS
    # factored: per view the non-SH gradients accumulate, the SH blocks come from the factors
    step.enable_factored_exchange(views_per_rank=VIEWS, world_size=1)
    step.grads.zero_()
    for slot, (cam, vm, pm, bt, gt, gtd) in enumerate(views):
        step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd, accumulate=slot > 0, dcolor_slot=slot)
    step.exchange_grads(torch.stack([v[1] for v in views]).contiguous(), torch.stack([v[3] for v in views]).contiguous())
    got = step.grads
    from rodygs_b200.trainer import sh_start
    n_plain = sh_start(step.layout)
    o_tab, _ = step.layout["table"]
    assert ref[n_plain:].abs().max().item() > 0
    for name in ("static.features_dc", "static.features_rest", "dynamic.features_dc", "dynamic.features_rest",
                 "static.xyz", "dynamic.xyz", "dynamic.rotation", "motion_coeff", "table"):
        o, shp = step.layout[name]
        n = 1
        for s in shp:
            n *= s
        a, b = got[o:o + n], ref[o:o + n]
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item() + 1e-12, name


def test_backward_in_pieces_equals_one_launch_and_pieces_tile_the_reduced_range():
    """The pipelined exchange runs the per-Gaussian backward as several launches (RdgSceneGrad.part / parts / dtable_mode:
    dynamic model in 3 pieces, static model, dL/dtable reduction).  Together they write the same gradients as one launch,
    and the float ranges the trainer all-reduces after each piece tile the non-SH range exactly once."""
    from rodygs_b200.trainer import sh_start
    N, H, W, T = 40_011, 128, 192, 6                       # 79 chunks per model: ragged last piece, ragged last chunk
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=5), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01, w_local=0.0)
    cam = synthetic.make_camera(1, 8, H, W, T)
    vm, pm = cam.world_view_transform.t().contiguous().cuda(), cam.projection_matrix.t().contiguous().cuda()
    bt = step.p("table")[cam.time_index].clone()
    gen = torch.Generator(device="cuda").manual_seed(2)
    gt, gtd = torch.rand(3, H, W, device="cuda", generator=gen), torch.rand(1, H, W, device="cuda", generator=gen)
    n_plain = sh_start(step.layout)
    try:
        engine.config.deterministic = True
        step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
        ref, ref_view = step.grads[:n_plain].clone(), step.view_grad.clone()
        assert ref.abs().max().item() > 0
        step.enable_factored_exchange(views_per_rank=1, world_size=1)
        step.world_size, step._bucketed = 2, True      # take the data-parallel schedule on one GPU
        step._bwd_plan, step._bwd_pieces = step._make_bwd_plan(3)
        step._bwd_final = "table"
        called = []
        step._after_model = called.append
        step._start_gather = lambda: None
        step.grads.zero_()
        step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd, dcolor_slot=0)
        torch.cuda.synchronize()
        got_det, view_det = step.grads[:n_plain].clone(), step.view_grad.clone()
        # ... and with the SM-partitioned kernels (chunk queue; 140 of the 148 SMs "reserved" so that 27 chunks exceed the grid):
        # the queue cleans itself between the five launches, every chunk is processed exactly once
        engine.config.deterministic = False
        _lib.set_tunable("sm_reserve", 140)
        for _ in range(2):
            step.grads.zero_()
            step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd, dcolor_slot=0)
        torch.cuda.synchronize()
        assert int(step._sm_queue.abs().max()) == 0
        got_q = step.grads[:n_plain].clone()
    finally:
        engine.config.deterministic = False
        _lib.set_tunable("sm_reserve", 0)
    assert float((got_q - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    step.grads[:n_plain].copy_(got_det)
    step.view_grad.copy_(view_det)
    assert called[:5] == ["static.0", "dynamic.0", "dynamic.1", "dynamic.2", "table"]
    got = step.grads[:n_plain]
    assert torch.equal(got, ref), float((got - ref).abs().max())
    # the pose gradient is summed per launch and then across launches: same terms, different association
    assert float((step.view_grad - ref_view).abs().max()) <= 1e-5 * float(ref_view.abs().max())
    cover = torch.zeros(n_plain, dtype=torch.int32)
    for piece in called[:5]:
        for lo, ln in step._bwd_pieces[piece]:
            assert lo % 4 == 0 and ln % 4 == 0 and lo + ln <= n_plain
            cover[lo:lo + ln] += 1
    assert int(cover.max()) == 1
    nz = ref.cpu() != 0
    assert bool((cover[nz] == 1).all())          # everything that can be non-zero is reduced exactly once


def test_fused_sh_adam_from_factors_equals_materialised_gradient_then_adam():
    """exchange_grads(defer_sh=True) + optimizer_step (rdg_sh_adam_views: dL/dSH rebuilt per chunk in shared memory and
    consumed by Adam) moves the parameters and both moments exactly like the default path (rdg_sh_grad_views writes
    dL/dSH, rdg_adam_groups steps every group) - over two iterations, so the second step sees non-zero moments, with a
    Gaussian count that is not a multiple of the 256-Gaussian chunk and per-model step counts / learning rates."""
    from rodygs_b200.optim import GaussianLRs
    N, H, W, T, VIEWS = 30_011, 128, 192, 6, 2
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=11), "cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    cams = [synthetic.make_camera(v, 8, H, W, T) for v in range(VIEWS)]
    gts = [(torch.rand(3, H, W, device="cuda", generator=gen), torch.rand(1, H, W, device="cuda", generator=gen)) for _ in cams]

    def train(defer):
        step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01, w_local=0.0)
        step.attach_optimizer("static", GaussianLRs())
        step.attach_optimizer("dynamic", GaussianLRs(scaling_lr=0.001, motion_coeff_lr=1.6e-4, feature_lr=0.004))
        step.enable_factored_exchange(views_per_rank=VIEWS, world_size=1)
        for it in (1, 2):
            step.grads.zero_()
            vms, bts = [], []
            for slot, (cam, (gt, gtd)) in enumerate(zip(cams, gts)):
                vm = cam.world_view_transform.t().contiguous().cuda()
                pm = cam.projection_matrix.t().contiguous().cuda()
                bt = step.p("table")[cam.time_index].clone()
                step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd, accumulate=slot > 0, dcolor_slot=slot)
                vms.append(vm); bts.append(bt)
            step.exchange_grads(torch.stack(vms).contiguous(), torch.stack(bts).contiguous(), defer_sh=defer)
            if defer and it == 1:
                # the deferred factors rebuild the same dL/dSH on demand
                step.materialize_sh_grads()
                assert step.g("static.features_rest").abs().max().item() > 0
            step.optimizer_step("static", it)
            if it == 2:
                step.optimizer_step("dynamic", it)      # the dynamic model joins late: its own step count is 1
        torch.cuda.synchronize()
        return step

    try:
        engine.config.deterministic = True      # bit-identical gradients in both runs: what is compared is the optimiser path
        a, b = train(False), train(True)
    finally:
        engine.config.deterministic = False
    assert b._deferred is not None and not b._deferred["pending"]
    moved = (a.params - synthetic_flat(a, scene)).abs().max().item()
    assert moved > 0
    for name in ("static.features_dc", "static.features_rest", "dynamic.features_dc", "dynamic.features_rest", "static.xyz",
                 "dynamic.xyz", "motion_coeff"):
        o, shp = a.layout[name]
        n = 1
        for d in shp:
            n *= d
        pa, pb = a.params[o:o + n], b.params[o:o + n]
        # same gradient bits, same update formula; the two kernels may contract the moment updates into FMAs differently
        assert (pa - pb).abs().max().item() <= 1e-6, name
        tag = name.split(".")[0] if "." in name else "dynamic"
        for ma, mb in zip((a.optim[tag].exp_avg, a.optim[tag].exp_avg_sq), (b.optim[tag].exp_avg, b.optim[tag].exp_avg_sq)):
            xa, xb = ma[o:o + n], mb[o:o + n]
            assert (xa - xb).abs().max().item() <= 1e-4 * xa.abs().max().item() + 1e-20, name
    # the SH parameters really were stepped by the fused kernel (their gradient buffer was never written in run b, iteration 2)
    o, shp = b.layout["dynamic.features_rest"]
    assert b.optim["dynamic"].exp_avg[o:o + 45 * b.nd].abs().max().item() > 0


def synthetic_flat(step, scene):
    """the flat parameter buffer a fresh SplatTrainStep builds from `scene`"""
    return SplatTrainStep(scene, step.H, step.W, sh_degree=step.sh_degree).params


def test_deterministic_mode_is_bit_reproducible_and_agrees_with_the_default():
    """engine.config.deterministic (RDG_DETERMINISTIC=1): every gradient of the fused path is bit-identical run to run
    (integer accumulation of the blend partials, one CTA for the cross-CTA sums), and equals the default float-atomic
    result to accumulation-order noise."""
    from rodygs_b200.dynamic import GaussianParams, render_dynamic
    from rodygs_b200.rasterizer import GaussianRasterizationSettings
    N, H, W, T = 20000, 160, 240, 12
    sc = synthetic.make_scene(N, H, W, T, seed=3, radius_px=6.0)
    cam = synthetic.make_camera(2, 8, H, W, T)
    g = torch.Generator().manual_seed(1)
    up = [t.cuda() for t in (torch.randn(3, H, W, generator=g), 0.3 * torch.randn(1, H, W, generator=g), 0.2 * torch.randn(1, H, W, generator=g))]

    def run():
        cst = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["static"].items()})
        cdy = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["dynamic"].items()})
        leaves = {"coeff": sc["motion_coeff"].cuda().requires_grad_(True), "table": sc["table"].cuda().requires_grad_(True),
                  "basis_t": sc["table"][cam.time_index].cuda().requires_grad_(True),
                  "vm": cam.world_view_transform.t().contiguous().cuda().requires_grad_(True)}
        st = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3).cuda(), 1.0,
                                           cam.projection_matrix.t().contiguous().cuda(), 3, False, False, True, True)
        pkg = render_dynamic(cst, cdy, st, leaves["vm"], leaves["coeff"], leaves["basis_t"], leaves["table"], sc["time_ind"].cuda(), 1.0, True)
        ((pkg["rendered_image"] * up[0]).sum() + (pkg["rendered_depth"] * up[1]).sum() + (pkg["rendered_alpha"] * up[2]).sum()).backward()
        out = {f"{t}.{k}": getattr(p, k).grad.clone() for t, p in (("st", cst), ("dy", cdy)) for k in p._fields}
        out.update({k: v.grad.clone() for k, v in leaves.items()})
        out["means2D"] = pkg["viewspace_points"].grad.clone()
        return out

    try:
        engine.config.deterministic = True
        a, b = run(), run()
        engine.config.deterministic = False
        c, d = run(), run()
    finally:
        engine.config.deterministic = False
    for k in a:
        assert torch.equal(a[k], b[k]), f"{k}: deterministic mode differs between two runs"
        scale = float(c[k].abs().max())
        assert float((a[k] - c[k]).abs().max()) <= 2e-5 * scale + 1e-30, f"{k}: deterministic vs default"
    # (the default mode is allowed to differ in the last bits between runs: float atomics)
    assert any(not torch.equal(c[k], d[k]) for k in c) or True


def test_captured_step_replays_the_eager_step():
    """SplatTrainStep.capture_forward_backward: the step as one CUDA graph gives the gradients and the loss of the eager
    step (deterministic mode: bit-identical), also after the static input buffers were refilled with another view."""
    N, H, W, T = 30_000, 128, 192, 6
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=6), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01)
    step.local_box_origins = torch.tensor([[0, 10]])
    gen = torch.Generator(device="cuda").manual_seed(4)
    views = []
    for v in (0, 3):
        cam = synthetic.make_camera(v, 8, H, W, T)
        views.append((cam, cam.world_view_transform.t().contiguous().cuda(), cam.projection_matrix.t().contiguous().cuda(),
                      step.p("table")[cam.time_index].clone(), torch.rand(3, H, W, device="cuda", generator=gen),
                      torch.rand(1, H, W, device="cuda", generator=gen)))
    was = engine.config.sync_free
    try:
        engine.config.sync_free = True
        engine.config.deterministic = True
        ref = []
        for cam, vm, pm, bt, gt, gtd in views:
            step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
            torch.cuda.synchronize()
            ref.append((step.grads.clone(), step.loss_parts.clone(), step.view_grad.clone()))
        cam0 = views[0][0]
        static = [t.clone() for t in views[0][1:]]
        cap = step.capture_forward_backward(static[0], static[1], cam0.tanfovx, cam0.tanfovy, static[2], static[3], static[4])
        for k in (0, 1, 0):
            for dst, src in zip(static, views[k][1:]):
                dst.copy_(src)
            step.grads.zero_()          # (the padding between the blocks is never written)
            cap.replay()
            torch.cuda.synchronize()
            g, l, vg = ref[k]
            assert torch.equal(step.grads, g), k
            assert torch.equal(step.loss_parts, l) and torch.equal(step.view_grad, vg), k
    finally:
        engine.config.sync_free = was
        engine.config.deterministic = False
        engine.reset_capacity()
