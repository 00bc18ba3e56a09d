"""Densification row.

CPU: the oracle restatement against the fixture made by running the reference's own
DynTrainer.densify_and_prune / add_densification_stats / reset_opacity (tests/golden/make_golden_densify.py).
GPU: the CUDA stream compaction against that fixture and, at a larger size, against the oracle.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import densify_oracle as dor

NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "motion_coeff"]
CASES = ["a", "b", "iso"]


def _golden(case):
    z = np.load(os.path.join(GOLDEN, "densify.npz"))
    pre = case + "/"
    return {k[len(pre):]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(pre)}


def _oracle_run(g):
    state = {k: (g[f"in/{k}"], g[f"in/{k}.m"], g[f"in/{k}.v"]) for k in NAMES}
    noise = g["noise"]
    return dor.densify_and_prune(state, g["in/time"], g["in/time_ind"], g["in/grad_accum"].clone(), g["in/denom"].clone(),
                                 0.0002, 0.005, float(g["extent"]), 20 if int(g["use_size"]) else None, 0.01,
                                 lambda S: noise)


@pytest.mark.parametrize("case", CASES)
def test_oracle_densify_matches_reference(case):
    g = _golden(case)
    state, time, time_ind, info = _oracle_run(g)
    assert info["rows"] == g["out/xyz"].shape[0] and info["split"] * 2 == g["noise"].shape[0]
    for k in NAMES:
        p, m, v = state[k]
        assert torch.equal(p, g[f"out/{k}"]), k
        assert torch.equal(m, g[f"out/{k}.m"]) and torch.equal(v, g[f"out/{k}.v"]), k
    assert torch.equal(time, g["out/time"]) and torch.equal(time_ind, g["out/time_ind"])
    # densification_postfix: every statistic of the new model is zero
    assert g["out/grad_accum"].abs().sum() == 0 and g["out/max_radii2D"].abs().sum() == 0
    assert g["out/grad_accum"].shape[0] == info["rows"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_stats_and_reset_match_reference(case):
    g = _golden(case)
    n = g["in/xyz"].shape[0]
    mr, ga, dn = torch.zeros(n), torch.zeros(n, 1), torch.zeros(n, 1)
    for k in range(3):
        mr, ga, dn = dor.add_densification_stats(g[f"stats/radii{k}"], g[f"stats/m2g{k}"], mr, ga, dn)
    assert torch.equal(mr, g["in/max_radii2D"]) and torch.equal(ga, g["in/grad_accum"]) and torch.equal(dn, g["in/denom"])
    assert torch.equal(dor.reset_opacity(g["out/opacity"]), g["reset/opacity"])
    assert g["reset/opacity.m"].abs().sum() == 0 and g["reset/opacity.v"].abs().sum() == 0


def test_densify_schedule_gates():
    """rodygs.py:317,343-350 with train_kubric_mrig.yaml:168-173."""
    from rodygs_b200 import densify as dn
    assert not dn.should_densify(500, 500, 20000, 100)
    assert dn.should_densify(600, 500, 20000, 100)
    assert not dn.should_densify(650, 500, 20000, 100)
    assert not dn.should_densify(20000, 500, 20000, 100)
    assert not dn.should_densify(600, 500, 20000, 0)
    assert dn.size_threshold(100, 5000000) is None and dn.size_threshold(5000001, 5000000) == 20


def test_densify_refuses_cpu_tensors():
    from rodygs_b200 import densify as dn
    st = dn.DensifyStats(4, device="cpu")
    with pytest.raises(RuntimeError):
        st.add(torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        dn.reset_opacity(torch.zeros(4, 1))


# ---------------------------------------------------------------------------------------- GPU

def _cuda_run(g, dev="cuda"):
    from rodygs_b200 import densify as dn
    n = g["in/xyz"].shape[0]
    params = {k: g[f"in/{k}"].to(dev).contiguous() for k in NAMES}
    moments = {k: (g[f"in/{k}.m"].to(dev).contiguous(), g[f"in/{k}.v"].to(dev).contiguous()) for k in NAMES}
    extras = {"time": g["in/time"].to(dev), "time_ind": g["in/time_ind"].to(torch.int32).to(dev)}
    stats = dn.DensifyStats(n, dev)
    stats.grad_accum.copy_(g["in/grad_accum"].view(-1))
    stats.denom.copy_(g["in/denom"].view(-1))
    stats.max_radii2D.copy_(g["in/max_radii2D"])
    return dn.densify_and_prune(params, moments, extras, stats, 0.0002, 0.005, float(g["extent"]),
                                20 if int(g["use_size"]) else None, 0.01, noise=g["noise"].to(dev).contiguous())


def _compare(new_p, new_m, new_e, want, want_time, want_time_ind, n_copy_rows):
    """Rows [0, n_copy_rows) are survivors + clones: bit-exact copies.  The split children differ from the CPU
    only through expf / logf / the 3-term dot product: a few ulp."""
    for k in NAMES:
        p, m, v = want[k]
        got = new_p[k].cpu()
        assert got.shape == p.shape, k
        assert torch.equal(got[:n_copy_rows], p[:n_copy_rows]), k
        if k in ("xyz", "scaling"):
            assert torch.allclose(got[n_copy_rows:], p[n_copy_rows:], rtol=2e-6, atol=2e-6), k
        else:
            assert torch.equal(got, p), k
        assert torch.equal(new_m[k][0].cpu(), m) and torch.equal(new_m[k][1].cpu(), v), k
    assert torch.equal(new_e["time"].cpu(), want_time)
    assert torch.equal(new_e["time_ind"].cpu().long(), want_time_ind.long())


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_densify_matches_reference_fixture(case):
    g = _golden(case)
    new_p, new_m, new_e, new_stats, info = _cuda_run(g)
    want = {k: (g[f"out/{k}"], g[f"out/{k}.m"], g[f"out/{k}.v"]) for k in NAMES}
    assert info["rows"] == g["out/xyz"].shape[0] and info["split_selected"] * 2 == g["noise"].shape[0]
    _compare(new_p, new_m, new_e, want, g["out/time"], g["out/time_ind"], info["survivors"] + info["clones"])
    assert new_stats.n == info["rows"] and float(new_stats.grad_accum.abs().sum()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_stats_and_reset_match_reference_fixture(case):
    from rodygs_b200 import densify as dn
    g = _golden(case)
    n = g["in/xyz"].shape[0]
    stats = dn.DensifyStats(n)
    pad = 7                                     # the model's rows sit at an offset inside the concatenated scene
    for k in range(3):
        radii = torch.cat((torch.full((pad,), 5, dtype=torch.int32), g[f"stats/radii{k}"])).cuda()
        m2g = torch.cat((torch.ones(pad, 3), g[f"stats/m2g{k}"])).cuda()
        stats.add(radii, m2g, offset=pad)
    assert torch.equal(stats.max_radii2D.cpu(), g["in/max_radii2D"])
    assert torch.allclose(stats.grad_accum.cpu(), g["in/grad_accum"].view(-1), rtol=1e-6, atol=0)
    assert torch.equal(stats.denom.cpu(), g["in/denom"].view(-1))
    op = g["out/opacity"].cuda().contiguous()
    m, v = torch.ones_like(op), torch.ones_like(op)
    dn.reset_opacity(op, (m, v))
    assert torch.allclose(op.cpu(), g["reset/opacity"], rtol=1e-5, atol=1e-5)
    assert float(m.abs().sum()) == 0.0 and float(v.abs().sum()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("n,use_size", [(200_000, True), (50_001, False), (3, True)])
def test_cuda_densify_matches_oracle_at_size(n, use_size):
    """Larger than one scan chunk (many CTAs, ragged last block); thresholds are compared away from ties."""
    from rodygs_b200 import densify as dn
    gen = torch.Generator().manual_seed(n)
    rn = lambda *s: torch.randn(*s, generator=gen)   # noqa: E731
    extent = 4.0
    state = {"xyz": rn(n, 3), "f_dc": rn(n, 1, 3), "f_rest": rn(n, 15, 3), "opacity": rn(n, 1) * 2 - 3,
             "scaling": rn(n, 3) + float(np.log(0.01 * extent)), "rotation": rn(n, 4), "motion_coeff": rn(n, 1, 16)}
    state = {k: (p, rn(*p.shape), rn(*p.shape).abs()) for k, p in state.items()}
    time_ind = torch.randint(0, 100, (n,), generator=gen)
    time = time_ind.float() / 100
    denom = torch.randint(0, 4, (n, 1), generator=gen).float()
    accum = torch.rand(n, 1, generator=gen) * 4e-4 * denom
    # keep every predicate away from its threshold so that 1-ulp differences of expf / the division cannot flip it
    ms = torch.exp(state["scaling"][0]).max(1).values
    tie = ((ms / (0.01 * extent) - 1).abs() < 1e-4) | ((ms / (0.1 * extent) - 1).abs() < 1e-4) | \
          ((ms / 1.6 / (0.1 * extent) - 1).abs() < 1e-4) | ((torch.sigmoid(state["opacity"][0]).view(-1) / 0.005 - 1).abs() < 1e-4)
    state["opacity"][0][tie] = 0.0
    state["scaling"][0][tie] = float(np.log(0.02 * extent))
    g = accum / denom
    accum[((g / 0.0002 - 1).abs() < 1e-4)] *= 1.01
    box = {}

    def noise_fn(S):
        box["z"] = torch.randn(2 * S, 3, generator=gen)
        return box["z"]

    want, w_time, w_ind, info_o = dor.densify_and_prune({k: tuple(t.clone() for t in v) for k, v in state.items()}, time, time_ind,
                                                        accum.clone(), denom.clone(), 0.0002, 0.005, extent,
                                                        20 if use_size else None, 0.01, noise_fn)
    params = {k: v[0].cuda().contiguous() for k, v in state.items()}
    moments = {k: (v[1].cuda().contiguous(), v[2].cuda().contiguous()) for k, v in state.items()}
    extras = {"time": time.cuda(), "time_ind": time_ind.to(torch.int32).cuda()}
    stats = dn.DensifyStats(n)
    stats.grad_accum.copy_(accum.view(-1))
    stats.denom.copy_(denom.view(-1))
    new_p, new_m, new_e, new_stats, info = dn.densify_and_prune(params, moments, extras, stats, 0.0002, 0.005, extent,
                                                                20 if use_size else None, 0.01, noise=box["z"].cuda().contiguous())
    assert info["rows"] == info_o["rows"] and info["split_selected"] == info_o["split"] and info["clones"] <= info_o["clones"]
    _compare(new_p, new_m, new_e, want, w_time, w_ind, info["survivors"] + info["clones"])
