"""All-reduce of the non-SH gradient ranges (44 / 108 / 152 MB): NCCL vs the symmetric-memory (NVLS multimem) collectives.
Run under torchrun, one rank per GPU."""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sizes = [11_000_000, 27_000_000, 38_000_000]      # floats: static range, dynamic range, both
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
res = []
for n in sizes:
    x = torch.randn(n, device="cuda")
    res.append((f"NCCL all_reduce AVG {n * 4 / 1e6:.0f} MB [{os.environ.get('NCCL_ALGO', 'default')}/{os.environ.get('NCCL_PROTO', 'default')}]", timeit(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))))
    res.append((f"NCCL all_reduce SUM {n * 4 / 1e6:.0f} MB", timeit(lambda: dist.all_reduce(x, op=dist.ReduceOp.SUM))))
if os.environ.get("SYMM", "1") == "1":
    group = dist.group.WORLD
    try:
        buf = symm_mem.empty(sizes[-1], dtype=torch.float32, device="cuda")
        hdl = symm_mem.rendezvous(buf, group)
        buf.normal_()
        for n in sizes:
            v = buf[:n]
            for name, fn in (("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(v, "sum", group.group_name)),):
                try:
                    res.append((f"symm_mem {name} {n * 4 / 1e6:.0f} MB", timeit(fn)))
                except Exception as e:
                    res.append((f"symm_mem {name} {n * 4 / 1e6:.0f} MB: {type(e).__name__}: {str(e)[:100]}", float("nan")))
        # the repo's own in-switch kernel (rdg_allreduce_multimem) over the same buffer, bracketed by the barriers the trainer uses
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from rodygs_b200 import _lib
        lib = _lib.load()
        for unroll in (4, 8):
            _lib.set_tunable("ar_unroll", unroll)
            for ctas in (16, 32, 64, 128):
                for n in sizes:
                    def own():
                        hdl.barrier(channel=0)
                        _lib.check(lib.rdg_allreduce_multimem(int(hdl.multicast_ptr), n, rank, world, 1.0 / world, ctas, _lib.stream_ptr()))
                        hdl.barrier(channel=0)
                    res.append((f"rdg_allreduce_multimem {n * 4 / 1e6:.0f} MB, {ctas} CTAs x 512, unroll {unroll} (+2 barriers)", timeit(own)))
        if rank == 0:
            print("multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "signal pads", len(hdl.signal_pad_ptrs))
    except Exception as e:
        res.append((f"symmetric memory unavailable: {type(e).__name__}: {str(e)[:200]}", float("nan")))
if rank == 0:
    for k, v in res:
        print(f"{k:70s} {v:.3f} ms")
dist.destroy_process_group()
