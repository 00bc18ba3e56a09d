"""Host-side cost of one SplatTrainStep.forward_backward call (python + ctypes + allocator), measured on a tiny scene so
that the GPU is never the bottleneck: wall time per call without synchronisation, and a cProfile of 200 calls."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from rodygs_b200 import engine, synthetic
from rodygs_b200.trainer import SplatTrainStep

N, H, W, T = 2000, 64, 64, 100
scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=0), "cuda")
step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.05, w_alpha=0.01)
cam = synthetic.make_camera(1, 8, H, W, T)
vm = cam.world_view_transform.t().contiguous().cuda()
pm = cam.projection_matrix.t().contiguous().cuda()
gt = torch.rand(3, H, W, device="cuda")
gtd = torch.rand(1, H, W, device="cuda")
bt = step.p("table")[cam.time_index].clone()
engine.config.sync_free = True
for _ in range(20):
    step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per forward_backward call: {(t1 - t0) / 300 * 1e6:.1f} us (drain {1e3 * (t2 - t1):.2f} ms)")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, bt, gt, gtd)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
