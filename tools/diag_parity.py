"""Diagnostic: CUDA forward vs the oracle in float32 and float64 on the two scenes that exceeded 1e-4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
import importlib.util
from oracle import splat_oracle as so, deform_oracle as do
from rodygs_b200 import synthetic
from rodygs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
spec = importlib.util.spec_from_file_location('tg', os.path.join(ROOT, 'tests/test_gpu_golden.py')); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)

def run(acts, cam, H, W, bg, mod, tag):
    xyz, op, scl, rot, feat = acts
    n = xyz.shape[0]
    vm = cam.world_view_transform.t().contiguous()
    o32 = so.rasterize(xyz, torch.zeros(n, 3), feat, None, op, scl, rot, vm, helpers.oracle_settings(cam, bg, 3, mod))
    d = lambda t: t.double()
    st64 = so.Settings(H, W, cam.tanfovx, cam.tanfovy, d(bg), mod, d(cam.projection_matrix.t().contiguous()), 3)
    o64 = so.rasterize(d(xyz), torch.zeros(n, 3).double(), d(feat), None, d(op), d(scl), d(rot), d(vm), st64)
    st = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg.cuda(), mod, cam.projection_matrix.t().contiguous().cuda(), 3, False, False, True, True)
    out = GaussianRasterizer(st)(means3D=xyz.cuda(), means2D=torch.zeros(n, 3, device="cuda"), shs=feat.cuda(), colors_precomp=None,
                                 opacities=op.cuda(), scales=scl.cuda(), rotations=rot.cuda(), cov3Ds_precomp=None, viewmatrix=vm.cuda())
    same_radii = torch.equal(o32.radii, o64.radii)
    print(f"== {tag}: D={o32.bn.keys.numel()} radii32==radii64 {same_radii}")
    for name, cu, a32, a64 in (("color", out[0], o32.color, o64.color), ("depth", out[1], o32.depth, o64.depth), ("alpha", out[3], o32.alpha, o64.alpha)):
        cu = cu.cpu()
        e_cu32 = (cu - a32).abs(); e_cu64 = (cu.double() - a64).abs(); e_3264 = (a32.double() - a64).abs()
        idx = e_cu32.flatten().argmax().item()
        print(f"  {name:6s} |cuda-o32| {e_cu32.max():.2e}  |cuda-o64| {e_cu64.max():.2e}  |o32-o64| {e_3264.max():.2e}   at worst: cuda {cu.flatten()[idx]:.6f} o32 {a32.flatten()[idx]:.6f} o64 {a64.flatten()[idx]:.6f} ncontrib {o32.bl.n_contrib.flatten()[idx % (H*W)]}")
        print(f"         pixels with |cuda-o32| > 1e-4: {(e_cu32 > 1e-4).sum().item()}, > 1e-4 vs o64: {(e_cu64 > 1e-4).sum().item()}, o32 vs o64 > 1e-4: {(e_3264 > 1e-4).sum().item()}")

Nf, Hf, Wf, T, _ = synthetic.CONFIGS["c4_iphone"]
N = 62745
sc = synthetic.make_scene(N, 256, 256, T, seed=0)
cam = synthetic.make_camera(0, 8, 256, 256, T)
acts = helpers.activated_concat(sc, cam)
run([a.detach() for a in acts], cam, 256, 256, torch.zeros(3), 1.0, "window")
acts, cam = tg._adversarial("needle", 300, 112, 176, seed=len("needle") * 1000 + 300)
run(acts, cam, 112, 176, torch.tensor([0.1, 0.2, 0.3]), 3.0, "needle x3")
