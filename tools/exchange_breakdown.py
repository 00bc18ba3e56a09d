"""Time the three pieces of the data-parallel exchange in isolation (run under torchrun, one rank per GPU)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rodygs_b200 import synthetic
from rodygs_b200.trainer import SplatTrainStep, sh_start

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N, H, W, T, _ = synthetic.CONFIGS["c4_iphone"]
scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=0), f"cuda:{lr}")
step = SplatTrainStep(scene, H, W, device=f"cuda:{lr}")
step.enable_factored_exchange(1, world)
step.dcolor_local.normal_()
cams = [synthetic.make_camera(r, 8, H, W, T) for r in range(world)]
vm = torch.stack([c.world_view_transform.t().contiguous() for c in cams]).cuda().contiguous()
bt = torch.stack([step.p("table")[c.time_index] for c in cams]).contiguous()
n_plain = sh_start(step.layout)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
tmp_all = torch.empty_like(step.dcolor_all)
res["NCCL all_gather factors (%.0f MB/rank)" % (step.dcolor_local.numel() * 4 / 1e6)] = timeit(lambda: dist.all_gather_into_tensor(tmp_all, step.dcolor_local))
def gather_only():
    step._start_gather()
    torch.cuda.current_stream().wait_event(step._ev_gathered)
for ns in (1, 2, 4, 7):
    step._gather_streams = [torch.cuda.Stream() for _ in range(ns)]
    step._ev_push = [torch.cuda.Event() for _ in range(ns)]
    res["copy-engine push gather, %d streams" % ns] = timeit(gather_only)
step._gather_streams = [torch.cuda.Stream() for _ in range(4)]
step._ev_push = [torch.cuda.Event() for _ in range(4)]
res["all_reduce AVG non-SH (%.0f MB)" % (n_plain * 4 / 1e6)] = timeit(lambda: dist.all_reduce(step.grads[:n_plain], op=dist.ReduceOp.AVG))
res["all_reduce SUM whole buffer (%.0f MB)" % (step.grads.numel() * 4 / 1e6)] = timeit(lambda: dist.all_reduce(step.grads))
def sh_only():
    step._gather_started = True
    step._ev_gathered.record()
    # exchange without collectives: world-size-1 semantics on this rank
    pg, step.pg = step.pg, None
    ws, step.world_size = step.world_size, step.world_size
    step.exchange_grads(vm, bt)
    step.pg = pg
res["exchange_grads (all pieces, overlapped)"] = timeit(lambda: (step._start_gather(), step.exchange_grads(vm, bt)))
import ctypes as C
from rodygs_b200 import _lib, engine
from rodygs_b200.engine import SceneArgs
lib = _lib.load()
def sh_kernel():
    scene_a = SceneArgs(st=step._set("static"), dy=step._set("dynamic"), raw=True, use_deform=True,
                        motion_coeff=step.p("motion_coeff").view(step.nd, step.num_basis), time_ind=step.time_ind, basis_t=bt[0],
                        table=step.p("table"), spatial_lr_scale=1.0, frame_order=step.frame_order, frame_offsets=step.frame_offsets)
    sc_s = engine._scene_struct(scene_a)
    gst, gdy = engine._setgrad_struct(step._setgrad("static")), engine._setgrad_struct(step._setgrad("dynamic"))
    _lib.check(lib.rdg_sh_grad_views(C.byref(sc_s), 3, world, vm.data_ptr(), bt.data_ptr(), step.dcolor_all.data_ptr(), 1.0 / world,
                                     C.byref(gst), C.byref(gdy), None, _lib.stream_ptr()))
res["rdg_sh_grad_views (%d views)" % world] = timeit(sh_kernel)
# timeline of two overlapped exchanges (kernel start/end per stream)
from torch.profiler import profile, ProfilerActivity
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step._start_gather(); step.exchange_grads(vm, bt)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in evs)
    for e in sorted(evs, key=lambda e: e.time_range.start):
        print(f"  t={e.time_range.start - t0:9.1f} us  dur={e.time_range.end - e.time_range.start:8.1f} us  {e.name[:70]}")
if rank == 0:
    for k, v in res.items():
        print(f"{k:50s} {v:8.3f} ms")
dist.destroy_process_group()
