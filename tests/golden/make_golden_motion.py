"""Pin oracle/motion_oracle.py to the reference's own motion regularisers
(/root/reference/src/trainer/losses.py:185-525), executed on CPU in the build container.

    python tests/golden/make_golden_motion.py      ->  tests/golden/motion.npz

The reference classes are imported unmodified.  What is stubbed, and why:
* ``pytorch3d.ops`` (un-vendored third-party, README.md:35): ``knn_points`` is bound to the exhaustive-search
  restatement in oracle/motion_oracle.py, ``knn_gather`` to plain indexing (its documented behaviour);
* ``omegaconf`` (only used by the config reader, not by the losses);
* ``torch.Tensor.cuda`` is the identity while ``MotionBasisRegularizaiton.__init__`` runs (losses.py:489 moves its
  16 weights to the GPU; there is none here).
``random.sample`` and ``torch.randint`` are wrapped to record the draws so that the oracle and the CUDA path can be
fed the same ones.
"""
import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _KNN:
    def __init__(self, dists, idx):
        self.dists, self.idx = dists, idx


def main():
    from oracle import motion_oracle as mo

    def knn_points(p1, p2, K):
        assert p1.shape[0] == 1 and p1.data_ptr() == p2.data_ptr()
        d, i = mo.knn_points(p1[0], K)
        return _KNN(d[None], i[None])

    def knn_gather(x, idx):
        return x[0][idx[0]][None]                      # [1, n, K, C]

    sys.path.insert(0, REF)
    _stub("pytorch3d")
    _stub("pytorch3d.ops", knn_points=knn_points, knn_gather=knn_gather)
    _stub("omegaconf", DictConfig=dict, OmegaConf=None)
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=None)
    _stub("diff_gauss_pose", GaussianRasterizationSettings=None, GaussianRasterizer=None)
    _stub("plyfile", PlyData=None, PlyElement=None)
    from src.trainer import losses as L

    g = torch.Generator().manual_seed(4321)
    N, B, T, K = 600, 16, 24, 8
    xyz = (torch.rand(N, 3, generator=g) * 2 - 1).requires_grad_(True)
    coeff = (torch.randn(N, 1, B, generator=g) * 0.3).requires_grad_(True)
    fdc = torch.randn(N, 1, 3, generator=g).requires_grad_(True)
    table = (torch.randn(T, B, 7, generator=g) * 0.2).requires_grad_(True)
    pred_tr = (torch.randn(N, 3, generator=g) * 0.05).requires_grad_(True)
    # a few exact duplicates (densification clones are exact copies) to exercise distance ties
    with torch.no_grad():
        xyz[10] = xyz[3]
        pred_tr[10] = pred_tr[3]
        xyz[200] = xyz[77]
        pred_tr[200] = pred_tr[77]

    class Model:
        pass
    model = Model()
    model._xyz, model._motion_coeff, model._features_dc = xyz, coeff, fdc
    model.unique_times = list(range(T))
    model.get_total_motion_table = lambda: table
    model.get_motion_for_times = lambda timesteps, time_indices=None: table[time_indices]

    out = dict(xyz=xyz.detach().numpy(), coeff=coeff.detach().numpy(), fdc=fdc.detach().numpy(),
               table=table.detach().numpy(), pred_tr=pred_tr.detach().numpy(), K=K)
    leaves = dict(xyz=xyz, coeff=coeff, table=table, pred_tr=pred_tr)

    def grads(loss, tag):
        gs = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
        out[f"{tag}/loss"] = loss.item()
        for k, gr in zip(leaves, gs):
            if gr is not None:
                out[f"{tag}/d_{k}"] = gr.numpy()

    # ---- MotionL1Loss / MotionSparsityLoss ----
    grads(L.MotionL1Loss()(model), "motion_l1")
    grads(L.MotionSparsityLoss()(model), "motion_sparsity")

    # ---- MotionBasisRegularizaiton ----
    cuda_attr = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for mode in ("cum_exponential", "vanilla"):
            reg = L.MotionBasisRegularizaiton(transl_degree=0, rot_degree=0, freq_div_mode=mode)
            out[f"basis_{mode}/reg_coeff"] = reg.reg_coeff.numpy()
            grads(reg(model), f"basis_{mode}")
        reg = L.MotionBasisRegularizaiton(transl_degree=1, rot_degree=1, freq_div_mode="gaussian")
        out["basis_deg1/reg_coeff"] = reg.reg_coeff.numpy()
        grads(reg(model), "basis_deg1")
    finally:
        torch.Tensor.cuda = cuda_attr

    # ---- RigidityLoss ----
    draws = {}
    real_sample, real_randint = random.sample, torch.randint

    def rec_sample(pop, k):
        r = real_sample(pop, k)
        draws["indice"] = list(r)
        return r

    def rec_randint(*a, **k):
        r = real_randint(*a, **k)
        draws["time_indices"] = r.clone()
        return r

    for tag, mode in (("rigid_cfg", ["distance_preserving", "surface"]), ("rigid_surface", ["surface"]),
                      ("rigid_dp", ["distance_preserving"]), ("rigid_coeff", ["coeff"])):
        random.seed(99)
        torch.manual_seed(99)
        random.sample, torch.randint = rec_sample, rec_randint
        try:
            loss = L.RigidityLoss(K=K, mode=mode)(model, pred_tr)
        finally:
            random.sample, torch.randint = real_sample, real_randint
        out[f"{tag}/indice"] = np.asarray(draws["indice"], dtype=np.int64)
        if "time_indices" in draws:
            out[f"{tag}/time_indices"] = draws["time_indices"].numpy()
        grads(loss, tag)
        draws.clear()

    np.savez_compressed(os.path.join(OUT, "motion.npz"), **out)
    print("wrote", os.path.join(OUT, "motion.npz"), {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()
                                                      if k.endswith("loss")})


if __name__ == "__main__":
    main()
