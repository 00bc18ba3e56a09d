"""World-size-2 test of the view-sharded data-parallel step on CPU (gloo stands in for NCCL):
two ranks each take their round-robin share of a step's views, write per-view gradients of
every parameter into the flat buffer layout the CUDA trainer uses, and all-reduce it; the result
must equal the single-process sum over all views.  The per-view gradients come from the oracle
(the CUDA engine cannot run here), so this covers exactly the host logic of the N>1 path:
shard_views, flat_layout and allreduce_flat."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from oracle import deform_oracle as do
from oracle import loss_oracle as lo
from oracle import splat_oracle as so
from rodygs_b200 import synthetic
from rodygs_b200.trainer import PARAM_ORDER, allgather_rows, allreduce_flat, flat_layout, sh_start, shard_views

N, H, W, T, VIEWS = 600, 48, 64, 4, 4


def _view_grads(sc, view, layout, total, factors=False):
    cam = synthetic.make_camera(view, VIEWS, H, W, T)
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    st = do.RawGaussians(**{k: leaf(v) for k, v in sc["static"].items()})
    dy = do.RawGaussians(**{k: leaf(v) for k, v in sc["dynamic"].items()})
    coeff, table = leaf(sc["motion_coeff"]), leaf(sc["table"])
    basis_t = leaf(sc["table"][cam.time_index])
    xyz, op, scl, rot, feat = do.assemble(st, dy, coeff.squeeze(1), basis_t, table, sc["time_ind"].long(), 1.0, True)
    out = so.rasterize(xyz, None, feat, None, op, scl, rot, cam.world_view_transform.t().contiguous(),
                       helpers.oracle_settings(cam, torch.zeros(3), 3))
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1000 + view))
    out.pp.rgb.retain_grad()
    lo.photometric(out.color, gt).backward()
    flat = torch.zeros(total)
    def put(name, t):
        off, shp = layout[name]
        g = t.grad if t.grad is not None else torch.zeros_like(t)
        flat[off:off + g.numel()] = g.reshape(-1)
    for tag, model in (("static", st), ("dynamic", dy)):
        for k in PARAM_ORDER:
            put(f"{tag}.{k}", getattr(model, k))
    put("motion_coeff", coeff); put("table", table); put("basis_t", basis_t)
    if factors:
        # dL/d(rgb) after the SH clamp mask, zero for Gaussians this view does not see; the deformed means; campos
        dcolor = torch.zeros(xyz.shape[0], 3)
        dcolor[out.pp.idx] = out.pp.rgb.grad * (out.pp.rgb.detach() > 0)
        V = cam.world_view_transform
        campos = -(V[:3, :3].t() @ V[:3, 3])
        return flat, dcolor, xyz.detach(), campos
    return flat


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synthetic.make_scene(N, H, W, T, seed=3)
        layout, total = flat_layout(sc["static"]["xyz"].shape[0], sc["dynamic"]["xyz"].shape[0], 16, T)
        flat = torch.zeros(total)
        for v in shard_views(VIEWS, world, rank):
            flat += _view_grads(sc, v, layout, total)
        allreduce_flat(flat, 1.0 / VIEWS)
        if rank == 0:
            ret.put(flat.numpy().copy())   # by value: a torch tensor travels as a shared-memory handle that dies with this process
    finally:
        dist.destroy_process_group()


def test_view_sharded_allreduce_matches_sequential():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.from_numpy(ret.get(timeout=240))
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    sc = synthetic.make_scene(N, H, W, T, seed=3)
    layout, total = flat_layout(sc["static"]["xyz"].shape[0], sc["dynamic"]["xyz"].shape[0], 16, T)
    ref = torch.zeros(total)
    for v in range(VIEWS):
        ref += _view_grads(sc, v, layout, total)
    ref /= VIEWS
    assert ref.abs().max() > 0
    # two ranks sum two views each and then add across ranks; the reference sums the four views in order (and the workers run
    # with 2 threads, this process with all cores): equal up to float32 summation order
    assert helpers.rel_err(got, ref) < 1e-5
    assert torch.allclose(got, ref, rtol=1e-3, atol=1e-7)


def _worker_factored(rank, world, port, ret):
    """The exchange of SplatTrainStep.exchange_grads on CPU: all-gather the 12-byte factors of dL/dSH,
    all-reduce only the non-SH range, rebuild dL/dSH of all views from the factors."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synthetic.make_scene(N, H, W, T, seed=3)
        ns, nd = sc["static"]["xyz"].shape[0], sc["dynamic"]["xyz"].shape[0]
        layout, total = flat_layout(ns, nd, 16, T)
        n_plain = sh_start(layout)
        per_rank = VIEWS // world
        flat = torch.zeros(total)
        local = torch.zeros(per_rank, 3, ns + nd, 3)      # per view: factors, deformed means, (campos in row 0)
        for slot in range(per_rank):
            v = rank * per_rank + slot                      # rank-major view order, as exchange_grads expects
            f, dcolor, xyz, campos = _view_grads(sc, v, layout, total, factors=True)
            flat += f
            local[slot, 0], local[slot, 1] = dcolor, xyz
            local[slot, 2, 0] = campos
        flat[n_plain:] = 0.0                                # the SH blocks are NOT exchanged
        gathered = torch.zeros(VIEWS, 3, ns + nd, 3)
        allgather_rows(local, gathered)
        allreduce_flat(flat[:n_plain], 1.0 / VIEWS)
        g = so.sh_grad_from_factors(3, [gathered[v, 1] for v in range(VIEWS)], [gathered[v, 2, 0] for v in range(VIEWS)],
                                    [gathered[v, 0] for v in range(VIEWS)]) / VIEWS
        for tag, lo_, hi_ in (("static", 0, ns), ("dynamic", ns, ns + nd)):
            o, _ = layout[f"{tag}.features_dc"]
            flat[o:o + (hi_ - lo_) * 3] = g[lo_:hi_, :1].reshape(-1)
            o, _ = layout[f"{tag}.features_rest"]
            flat[o:o + (hi_ - lo_) * 45] = g[lo_:hi_, 1:].reshape(-1)
        if rank == 0:
            ret.put(flat.numpy().copy())   # by value: a torch tensor travels as a shared-memory handle that dies with this process
    finally:
        dist.destroy_process_group()


def test_factored_sh_exchange_matches_sequential():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker_factored, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.from_numpy(ret.get(timeout=240))
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    sc = synthetic.make_scene(N, H, W, T, seed=3)
    layout, total = flat_layout(sc["static"]["xyz"].shape[0], sc["dynamic"]["xyz"].shape[0], 16, T)
    ref = torch.zeros(total)
    for v in range(VIEWS):
        ref += _view_grads(sc, v, layout, total)
    ref /= VIEWS
    o, _ = layout["static.features_rest"]
    assert ref[o:].abs().max() > 0
    assert helpers.rel_err(got, ref) < 1e-5
    assert torch.allclose(got, ref, rtol=1e-3, atol=1e-7)


def _worker_stats(rank, world, port, ret):
    """DensifyStats.all_reduce on the DEFAULT group (what SplatTrainStep.densify_and_prune calls when it was built without a
    process_group, as bench.py and every data-parallel user does): sum / sum / max over the ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rodygs_b200.densify import DensifyStats
        g = torch.Generator().manual_seed(100 + rank)
        st = DensifyStats(50, "cpu")
        st.grad_accum.copy_(torch.rand(50, generator=g))
        st.denom.copy_(torch.randint(0, 3, (50,), generator=g).float())
        st.max_radii2D.copy_(torch.randint(0, 40, (50,), generator=g).float())
        st.all_reduce(None)
        ret.put((rank, st.grad_accum.numpy().copy(), st.denom.numpy().copy(), st.max_radii2D.numpy().copy()))   # by value
    finally:
        dist.destroy_process_group()


def test_densification_statistics_are_rank_consistent_on_the_default_group():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker_stats, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r[0], tuple(torch.from_numpy(a) for a in r[1:])) for r in (ret.get(timeout=120), ret.get(timeout=120)))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    per_rank = []
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        per_rank.append((torch.rand(50, generator=g), torch.randint(0, 3, (50,), generator=g).float(),
                         torch.randint(0, 40, (50,), generator=g).float()))
    want = (per_rank[0][0] + per_rank[1][0], per_rank[0][1] + per_rank[1][1], torch.maximum(per_rank[0][2], per_rank[1][2]))
    for rank in range(2):
        for a, b in zip(got[rank], want):
            assert torch.equal(a, b)


def test_trainer_densify_always_reduces_its_statistics():
    """the all-reduce must not be gated on an explicit process_group."""
    import inspect
    from rodygs_b200.trainer import SplatTrainStep
    src = inspect.getsource(SplatTrainStep.densify_and_prune)
    assert "self.stats[tag].all_reduce(self.pg)" in src and "self.pg is not None" not in src
