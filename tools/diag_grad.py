"""Diagnostic: element-wise gradient error of the fused path on the density-matched window, against the float32 and the
float64 CPU chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from oracle import splat_oracle as so, deform_oracle as do
from rodygs_b200 import synthetic
from rodygs_b200.dynamic import GaussianParams, render_dynamic
from rodygs_b200.rasterizer import GaussianRasterizationSettings

def chain(sc, cam, H, W, up, dt):
    c = lambda t: t.detach().clone().to(dt).requires_grad_(True)
    st = do.RawGaussians(**{k: c(v) for k, v in sc["static"].items()})
    dy = do.RawGaussians(**{k: c(v) for k, v in sc["dynamic"].items()})
    coeff, table = c(sc["motion_coeff"]), c(sc["table"])
    basis_t = c(sc["table"][cam.time_index])
    acts = do.assemble(st, dy, coeff.squeeze(1), basis_t, table, sc["time_ind"].long(), sc["spatial_lr_scale"], True)
    vm = c(cam.world_view_transform.t().contiguous())
    n = acts[0].shape[0]
    xyz, op, scl, rot, feat = acts
    s = so.Settings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, dtype=dt), 1.0, cam.projection_matrix.t().contiguous().to(dt), 3)
    orc = so.rasterize(xyz, torch.zeros(n, 3, dtype=dt, requires_grad=True), feat, None, op, scl, rot, vm, s)
    ((orc.color * up[0].to(dt)).sum() + (orc.depth * up[1].to(dt)).sum() + (orc.alpha * up[2].to(dt)).sum()).backward()
    return {"viewmatrix": vm.grad, "table": table.grad, "basis_t": basis_t.grad, "motion_coeff": coeff.grad, "static.xyz": st.xyz.grad,
            "dynamic.xyz": dy.xyz.grad, "dynamic.rotation": dy.rotation.grad}

N, H, W, T = 62745, 256, 256, 100
sc = synthetic.make_scene(N, H, W, T, seed=0)
cam = synthetic.make_camera(0, 8, H, W, T)
g = torch.Generator().manual_seed(5)
up = (torch.randn(3, H, W, generator=g), 0.3 * torch.randn(1, H, W, generator=g), 0.2 * torch.randn(1, H, W, generator=g))
r32 = chain(sc, cam, H, W, up, torch.float32)
r64 = chain(sc, cam, H, W, up, torch.float64)
for rep in range(2):
    cst = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["static"].items()})
    cdy = GaussianParams(**{k: v.cuda().requires_grad_(True) for k, v in sc["dynamic"].items()})
    ccoeff = sc["motion_coeff"].cuda().requires_grad_(True); ctable = sc["table"].cuda().requires_grad_(True)
    cbasis = sc["table"][cam.time_index].cuda().requires_grad_(True); cvm = cam.world_view_transform.t().contiguous().cuda().requires_grad_(True)
    settings = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3).cuda(), 1.0, cam.projection_matrix.t().contiguous().cuda(), 3, False, False, True, True)
    pkg = render_dynamic(cst, cdy, settings, cvm, ccoeff, cbasis, ctable, sc["time_ind"].cuda(), sc["spatial_lr_scale"], True)
    ((pkg["rendered_image"] * up[0].cuda()).sum() + (pkg["rendered_depth"] * up[1].cuda()).sum() + (pkg["rendered_alpha"] * up[2].cuda()).sum()).backward()
    cu = {"viewmatrix": cvm.grad, "table": ctable.grad, "basis_t": cbasis.grad, "motion_coeff": ccoeff.grad, "static.xyz": cst.xyz.grad, "dynamic.xyz": cdy.xyz.grad, "dynamic.rotation": cdy.rotation.grad}
    for k in cu:
        a, b32, b64 = cu[k].cpu().double(), r32[k].double(), r64[k]
        mx = b64.abs().max().item()
        def worst(x, y):
            bound = 1e-3 * y.abs() + 1e-5 * y.abs().max()
            return ((x - y).abs() / bound).max().item()
        print(f"run {rep} {k:18s} max|b| {mx:.3e}  normwise cu-o32 {(a-b32).abs().max().item()/mx:.2e} cu-o64 {(a-b64).abs().max().item()/mx:.2e} o32-o64 {(b32-b64).abs().max().item()/mx:.2e}"
              f"   elementwise ratio cu-o32 {worst(a,b32):.2f} cu-o64 {worst(a,b64):.2f} o32-o64 {worst(b32,b64):.2f}")
