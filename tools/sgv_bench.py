"""Single-GPU timing of rdg_sh_grad_views at the bench workload (V views gathered on one rank)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rodygs_b200 import synthetic, _lib, engine
from rodygs_b200.engine import SceneArgs
from rodygs_b200.trainer import SplatTrainStep
V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N, H, W, T, _ = synthetic.CONFIGS["c4_iphone"]
scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=0), "cuda")
step = SplatTrainStep(scene, H, W)
step.enable_factored_exchange(V, 1)
step.dcolor_all.normal_()
cams = [synthetic.make_camera(r, 8, H, W, T) for r in range(V)]
vm = torch.stack([c.world_view_transform.t().contiguous() for c in cams]).cuda().contiguous()
bt = torch.stack([step.p("table")[c.time_index] for c in cams]).contiguous()
lib = _lib.load()
sc = SceneArgs(st=step._set("static"), dy=step._set("dynamic"), raw=True, use_deform=True,
               motion_coeff=step.p("motion_coeff").view(step.nd, step.num_basis), time_ind=step.time_ind, basis_t=bt[0],
               table=step.p("table"), spatial_lr_scale=1.0, frame_order=step.frame_order, frame_offsets=step.frame_offsets)
sc_s = engine._scene_struct(sc)
gst, gdy = engine._setgrad_struct(step._setgrad("static")), engine._setgrad_struct(step._setgrad("dynamic"))
def run():
    _lib.check(lib.rdg_sh_grad_views(C.byref(sc_s), 3, V, vm.data_ptr(), bt.data_ptr(), step.dcolor_all.data_ptr(), 1.0 / V,
                                     C.byref(gst), C.byref(gdy), None, _lib.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"rdg_sh_grad_views V={V}: {e0.elapsed_time(e1) / 20:.3f} ms")
