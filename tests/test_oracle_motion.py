"""Pin oracle/motion_oracle.py against the reference's own motion regularisers
(fixtures made by tests/golden/make_golden_motion.py from /root/reference/src/trainer/losses.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import motion_oracle as mo


def load_motion():
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLDEN, "motion.npz")).items()}


def leaves(g):
    return {k: g[k].clone().requires_grad_(True) for k in ("xyz", "coeff", "table", "pred_tr")}


def check_grads(g, tag, loss, lv, atol=1e-7, rtol=2e-5):
    assert abs(loss.item() - g[f"{tag}/loss"].item()) <= 2e-6 * max(1.0, abs(g[f"{tag}/loss"].item())), tag
    grads = torch.autograd.grad(loss, list(lv.values()), allow_unused=True)
    for k, gr in zip(lv, grads):
        key = f"{tag}/d_{k}"
        if key not in g:
            assert gr is None or gr.abs().max() == 0, (tag, k)
            continue
        ref = g[key]
        assert gr is not None, (tag, k)
        err = (gr - ref).abs().max().item()
        assert err <= atol + rtol * ref.abs().max().item(), (tag, k, err)


def test_coefficient_regularisers_match_reference():
    g = load_motion()
    lv = leaves(g)
    check_grads(g, "motion_l1", mo.motion_l1(lv["coeff"]), lv)
    lv = leaves(g)
    check_grads(g, "motion_sparsity", mo.motion_sparsity(lv["coeff"]), lv)


@pytest.mark.parametrize("tag,td,rd", [("basis_cum_exponential", 0, 0), ("basis_vanilla", 0, 0), ("basis_deg1", 1, 1)])
def test_basis_regulariser_matches_reference(tag, td, rd):
    g = load_motion()
    lv = leaves(g)
    check_grads(g, tag, mo.motion_basis_reg(lv["table"], g[f"{tag}/reg_coeff"], td, rd), lv)


@pytest.mark.parametrize("tag,mode", [("rigid_cfg", ("distance_preserving", "surface")), ("rigid_surface", ("surface",)),
                                      ("rigid_dp", ("distance_preserving",)), ("rigid_coeff", ("coeff",))])
def test_rigidity_matches_reference(tag, mode):
    g = load_motion()
    lv = leaves(g)
    ti = g.get(f"{tag}/time_indices")
    loss, _ = mo.rigidity(lv["xyz"], lv["coeff"], g["fdc"], lv["pred_tr"], lv["table"], g[f"{tag}/indice"],
                          ti, K=int(g["K"]), mode=mode)
    check_grads(g, tag, loss, lv)


def test_knn_restatement_properties():
    """Exhaustive-search contract: ascending squared distances, self first, ties towards the lower index."""
    gen = torch.Generator().manual_seed(5)
    p = torch.rand(300, 3, generator=gen)
    p[17] = p[4]                                          # exact duplicate
    d, i = mo.knn_points(p, 8)
    assert (d[:, 1:] >= d[:, :-1]).all()
    assert (d[:, 0] == 0).all()
    assert i[4, 0] == 4 and i[4, 1] == 17 and i[17, 0] == 4 and i[17, 1] == 17
    full = ((p[:, None] - p[None]) ** 2).sum(-1)
    assert torch.allclose(d, full.sort(dim=1).values[:, :8], atol=1e-7)
