"""Golden fixture for the densification row: runs the REFERENCE's own
`DynTrainer.densify_and_prune`, `add_densification_stats` and `reset_opacity`
(/root/reference/src/trainer/rodygs_static.py, rodygs_dynamic.py, utils.py) on CPU and stores
inputs + outputs in tests/golden/densify.npz.  Only possible in the build container
(/root/reference is not on the GPU box), which is why the output is committed.

    python tests/golden/make_golden_densify.py

The reference hard-codes device="cuda" and draws the split offsets with torch.normal; this
script redirects "cuda" allocations to the CPU and replaces torch.normal(mean, std) by
mean + noise * std with a recorded unit-normal `noise`, nothing else is touched.
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


def _redirect_cuda():
    real_zeros, real_ones = torch.zeros, torch.ones

    def zeros(*a, **k):
        if str(k.get("device", "")) == "cuda":
            k.pop("device")
        return real_zeros(*a, **k)

    def ones(*a, **k):
        if str(k.get("device", "")) == "cuda":
            k.pop("device")
        return real_ones(*a, **k)

    torch.zeros, torch.ones = zeros, ones
    torch.Tensor.cuda = lambda self, *a, **k: self


def make_case(seed, n, use_size, extent, isotropic=False):
    from src.model.rodygs_dynamic import DynRoDyGS
    from src.trainer.rodygs_dynamic import DynTrainer

    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)   # noqa: E731
    model = DynRoDyGS.__new__(DynRoDyGS)
    model.isotropic = isotropic
    P = torch.nn.Parameter
    model._xyz = P(rn(n, 3) * 2)
    model._features_dc = P(rn(n, 1, 3))
    model._features_rest = P(rn(n, 15, 3) * 0.1)
    model._opacity = P(rn(n, 1) * 2.0 - 3.0)
    model._scaling = P(rn(n, 1 if isotropic else 3) * 1.0 + float(np.log(0.01 * extent)))
    model._rotation = P(rn(n, 4))
    model._motion_coeff = P(rn(n, 1, 16) * 0.1)
    model.gaussian_to_time_ind = torch.randint(0, 12, (n,), generator=g)
    model.gaussian_to_time = model.gaussian_to_time_ind.float() / 12
    net = [P(rn(4, 4)), P(rn(4))]

    tr = DynTrainer.__new__(DynTrainer)
    tr.model = model
    tr.percent_dense = 0.01
    groups = [
        {"params": [model._xyz], "lr": 1e-3, "name": "xyz"},
        {"params": [model._features_dc], "lr": 1e-3, "name": "f_dc"},
        {"params": [model._features_rest], "lr": 1e-3, "name": "f_rest"},
        {"params": [model._opacity], "lr": 1e-3, "name": "opacity"},
        {"params": [model._scaling], "lr": 1e-3, "name": "scaling"},
        {"params": [model._rotation], "lr": 1e-3, "name": "rotation"},
        {"params": net, "lr": 1e-3, "name": "deform_network"},
        {"params": [model._motion_coeff], "lr": 1e-3, "name": "motion_coeff"},
    ]
    tr.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for grp in groups:                      # one real Adam step so that both moments are populated
        for p in grp["params"]:
            p.grad = rn(*p.shape) * 0.01
    tr.optimizer.step()

    # statistics: a few iterations of add_densification_stats + the max-radii lines (rodygs.py:334-341)
    tr.xyz_gradient_accum = torch.zeros(n, 1)
    tr.denom = torch.zeros(n, 1)
    tr.max_radii2D = torch.zeros(n)
    stats_in = []
    for _ in range(3):
        radii = torch.randint(-2, 40, (n,), generator=g).clamp(min=0).int()
        m2g = rn(n, 3) * 1.6e-4
        vis = radii > 0
        tr.max_radii2D[vis] = torch.max(tr.max_radii2D[[vis]], radii[vis])
        tr.add_densification_stats(vis, torch.norm(m2g[:, :2], dim=-1, keepdim=True))
        stats_in.append((radii.numpy(), m2g.numpy()))

    names = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "motion_coeff"]

    def snapshot(prefix, out, only=None):
        for grp in tr.optimizer.param_groups:
            if grp["name"] not in (only or names):
                continue
            p = grp["params"][0]
            st = tr.optimizer.state[p]
            out[f"{prefix}/{grp['name']}"] = p.detach().numpy().copy()
            out[f"{prefix}/{grp['name']}.m"] = st["exp_avg"].numpy().copy()
            out[f"{prefix}/{grp['name']}.v"] = st["exp_avg_sq"].numpy().copy()
        if only:
            return
        out[f"{prefix}/time"] = model.gaussian_to_time.numpy().copy()
        out[f"{prefix}/time_ind"] = model.gaussian_to_time_ind.numpy().copy()
        out[f"{prefix}/grad_accum"] = tr.xyz_gradient_accum.numpy().copy()
        out[f"{prefix}/denom"] = tr.denom.numpy().copy()
        out[f"{prefix}/max_radii2D"] = tr.max_radii2D.numpy().copy()

    out = {}
    for k, (r, m) in enumerate(stats_in):
        out[f"stats/radii{k}"], out[f"stats/m2g{k}"] = r, m
    snapshot("in", out)

    noise_log = []
    real_normal = torch.normal

    def normal(mean, std, **kw):
        z = torch.randn(std.shape, generator=g)
        noise_log.append(z)
        return mean + z * std

    torch.normal = normal
    try:
        tr.densify_and_prune(0.0002, 0.005, extent, 20 if use_size else None)
    finally:
        torch.normal = real_normal
    snapshot("out", out)
    out["noise"] = noise_log[0].numpy()
    out["extent"], out["use_size"], out["isotropic"] = extent, int(use_size), int(isotropic)

    tr.reset_opacity()
    snapshot("reset", out, only=["opacity"])
    return out


def main():
    sys.path.insert(0, REF)
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=None)
    _stub("diff_gauss_pose", GaussianRasterizationSettings=None, GaussianRasterizer=None)
    _stub("plyfile", PlyData=None, PlyElement=None)
    _stub("omegaconf", DictConfig=dict, OmegaConf=None)
    _redirect_cuda()
    cases = {"a": make_case(11, 240, True, 5.0), "b": make_case(12, 160, False, 3.0),
             "iso": make_case(13, 120, True, 4.0, isotropic=True)}
    flat = {f"{c}/{k}": v for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "densify.npz"), **flat)
    for c, d in cases.items():
        print(c, "rows", d["in/xyz"].shape[0], "->", d["out/xyz"].shape[0], "split-selected", d["noise"].shape[0] // 2)


if __name__ == "__main__":
    main()
