"""N-rank correctness check of the data-parallel exchange on real GPUs: every rank renders its own view, exchanges
(factored SH + bucketed all-reduce); the result must equal the mean over the views of the single-rank gradients (each rank
recomputes all views locally with the plain path) and be identical on all ranks.  Run under torchrun."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rodygs_b200 import synthetic
from rodygs_b200.trainer import SplatTrainStep
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N, H, W, T = 120_000, 256, 384, 10
scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=1), "cuda")
step = SplatTrainStep(scene, H, W, w_pearson=0.05, w_alpha=0.01, w_local=0.0)
gen = torch.Generator(device="cuda").manual_seed(5)
gts = [(torch.rand(3, H, W, device="cuda", generator=gen), torch.rand(1, H, W, device="cuda", generator=gen)) for _ in range(world)]
cams = [synthetic.make_camera(v, 8, H, W, T) for v in range(world)]
def run_view(v, **kw):
    c = cams[v]
    return step.forward_backward(c.world_view_transform.t().contiguous().cuda(), c.projection_matrix.t().contiguous().cuda(), c.tanfovx,
                                 c.tanfovy, step.p("table")[c.time_index].contiguous(), gts[v][0], gts[v][1], **kw)
# reference: mean over all views with the plain path on this rank
ref = torch.zeros_like(step.grads)
for v in range(world):
    run_view(v)
    ref += step.grads
ref /= world
for mode in ("multicast", "copy-engine"):
    step.enable_factored_exchange(1, world, multicast=(mode == "multicast"))
    vm_all = torch.stack([c.world_view_transform.t().contiguous() for c in cams]).cuda().contiguous()
    bt_all = torch.stack([step.p("table")[c.time_index] for c in cams]).contiguous()
    for rep in range(3):   # several steps: exercises the double-buffered factors
        run_view(rank, dcolor_slot=0)
        step.exchange_grads(vm_all, bt_all)
        torch.cuda.synchronize()
        got = step.grads.clone()
        from rodygs_b200.trainer import sh_start
        # basis_t is the rank's own B(t): its gradient is rank-local by construction, leave it out of the comparison
        o, shp = step.layout["basis_t"]
        got[o:o + 112] = 0; r2 = ref.clone(); r2[o:o + 112] = 0
        err = float((got - r2).abs().max() / r2.abs().max())
        g0 = got.clone(); dist.broadcast(g0, 0)
        same = float((got - g0).abs().max() / g0.abs().max())
        if rank == 0 or err > 1e-4:
            print(f"[rank {rank}] {mode} step {rep}: max rel err vs sequential mean {err:.2e}; vs rank 0 {same:.2e}; mc={bool(step._mc_base)}", flush=True)
        assert err < 2e-4 and same < 1e-5, (mode, rep, err, same)
dist.barrier()
if rank == 0:
    print("dp_check OK")
dist.destroy_process_group()
