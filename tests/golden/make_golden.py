"""Generate the golden fixtures in this directory by importing the reference's
own Python from /root/reference (only possible in the build container; the
GPU box has no /root/reference, which is why the outputs are committed).

    python tests/golden/make_golden.py

Pinned here:
  sh_eval.npz   <- src/utils/sh_utils.py::eval_sh (deg 0..3)
  losses.npz    <- src/utils/loss_utils.py::{l1_loss, ssim, pearson_depth_loss} values + input gradients
  deform.npz    <- src/model/rodygs_dynamic.py::{TimestepEmbedder, MLPBasisNetwork, DynRoDyGS.get_gaussian_deformation}
  camera.npz    <- src/utils/graphic_utils.py::{getProjectionMatrix, quaternion_to_matrix},
                   src/data/utils.py::FixedCameraTorch.world_view_transform (formula at :161-170)
The rasterizer itself (diff_gauss_pose) is not in the reference snapshot, so it
has no fixture: parity for it is unpinned (see oracle/__init__.py).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def main():
    sys.path.insert(0, REF)
    # un-vendored native modules that the reference imports at module scope
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=None)
    _stub("diff_gauss_pose", GaussianRasterizationSettings=None, GaussianRasterizer=None)
    _stub("plyfile", PlyData=None, PlyElement=None)

    g = torch.Generator().manual_seed(1234)

    # ---- SH ----------------------------------------------------------------
    from src.utils.sh_utils import eval_sh
    M = 64
    sh = torch.randn(M, 3, 16, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g))
    out = {"sh": sh.numpy(), "dirs": dirs.numpy()}
    for deg in range(4):
        out[f"deg{deg}"] = eval_sh(deg, sh, dirs).numpy()
    np.savez_compressed(os.path.join(OUT, "sh_eval.npz"), **out)

    # ---- losses --------------------------------------------------------------
    from src.utils.loss_utils import l1_loss, ssim, pearson_depth_loss
    a = torch.rand(3, 40, 56, generator=g).requires_grad_(True)
    b = (a.detach() + 0.1 * torch.randn(3, 40, 56, generator=g)).clamp(0, 1)
    l1v = l1_loss(a, b)
    (gl1,) = torch.autograd.grad(l1v, a)
    sv = ssim(a, b)
    (gs,) = torch.autograd.grad(sv, a)
    d1 = (torch.rand(1, 40, 56, generator=g) * 5 + 1).requires_grad_(True)
    d2 = d1.detach() * 0.7 + torch.rand(1, 40, 56, generator=g)
    pv = pearson_depth_loss(d1, d2, 1e-6)
    (gp,) = torch.autograd.grad(pv, d1)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), a=a.detach().numpy(), b=b.numpy(),
                        l1=l1v.item(), l1_grad=gl1.numpy(), ssim=sv.item(), ssim_grad=gs.numpy(),
                        d1=d1.detach().numpy(), d2=d2.numpy(), pearson=pv.item(), pearson_grad=gp.numpy())

    # ---- deformation -------------------------------------------------------------
    from src.model.rodygs_dynamic import DynRoDyGS, MLPBasisNetwork
    torch.manual_seed(7)
    netw, nb, mr = 32, 16, 26
    net = MLPBasisNetwork(netw, nb, mr, False)
    with torch.no_grad():  # the reference inits with std 1e-2 (outputs ~1e-9); widen so the test bites
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.05))
    T, Nd = 6, 40
    times = torch.arange(T, dtype=torch.float32) / T
    model = DynRoDyGS(sh_degree=3, deform_netwidth=netw, deform_t_emb_multires=mr,
                      deform_t_log_sampling=False, num_basis=nb, inverse_motion=True)
    model._deform_network = net
    model._motion_coeff = torch.nn.Parameter(torch.randn(Nd, 1, nb, generator=g) * 0.2)
    model.gaussian_to_time_ind = torch.randint(0, T, (Nd,), generator=g)
    model._time_batch_embeddings = torch.stack([net.t_embedder(t) for t in times]).squeeze()
    model.spatial_lr_scale = 2.5
    t_query = torch.tensor(0.37)
    trans, rot = model.get_gaussian_deformation(t_query)
    table = model.get_total_motion_table()
    emb = net.t_embedder(t_query)
    state = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "deform.npz"),
                        times=times.numpy(), t_query=t_query.numpy(), emb=emb.detach().numpy(),
                        coeff=model._motion_coeff.detach().numpy(), time_ind=model.gaussian_to_time_ind.numpy(),
                        table=table.detach().numpy(), trans=trans.detach().numpy(), rot=rot.detach().numpy(),
                        spatial_lr_scale=2.5, netwidth=netw, **{"state/" + k: v for k, v in state.items()})

    # ---- camera conventions ----------------------------------------------------------
    from src.utils.graphic_utils import getProjectionMatrix, quaternion_to_matrix
    fovx, fovy = 0.9, 0.6
    P = getProjectionMatrix(0.01, 100.0, fovx, fovy)
    quat = torch.nn.functional.normalize(torch.randn(4, generator=g), dim=0)
    T_c2w = torch.randn(3, generator=g)
    R_c2w = quaternion_to_matrix(quat)
    R_w2c = R_c2w.transpose(0, 1)                       # src/data/utils.py:161-170
    T_w2c = -torch.einsum("ij, j -> i", R_w2c, T_c2w)
    Vm = torch.eye(4)
    Vm[:3, :3] = R_w2c
    Vm[:3, 3] = T_w2c
    np.savez_compressed(os.path.join(OUT, "camera.npz"), fovx=fovx, fovy=fovy, P=P.numpy(),
                        quat=quat.numpy(), T_c2w=T_c2w.numpy(), V=Vm.numpy())
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
