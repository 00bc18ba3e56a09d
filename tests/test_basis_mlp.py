"""Motion-basis MLP: oracle and host logic on CPU, `rdg_basis_mlp_fwd/_bwd` on the GPU.

Pinned to the REFERENCE's own MLPBasisNetwork forward + autograd backward through tests/golden/basis_mlp.npz
(tests/golden/make_golden_basis.py).  Tolerances (float32): outputs 1e-5 relative to max|B|, parameter gradients
1e-4 relative (max|a-b| / max|b| per tensor); the backward is atomics-free, so two runs are bit-identical."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import rel_err
from oracle import deform_oracle as do
from rodygs_b200 import deform

CASES = {"gelu64": (64, "gelu"), "relu32": (32, "relu")}


def load_case(tag):
    z = np.load(os.path.join(GOLDEN, "basis_mlp.npz"))
    g = {k[len(tag) + 1:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(tag + "/")}
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    grads = {k[len("grad/"):]: v for k, v in g.items() if k.startswith("grad/")}
    return g, state, grads


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_forward_and_backward_match_reference(tag):
    width, act = CASES[tag]
    g, state, grads = load_case(tag)
    emb = do.time_embedding(g["times"], 26, False)
    assert torch.allclose(emb, g["emb"], atol=1e-6)
    leaves = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    basis = do.motion_basis(leaves, g["emb"], 16, gelu=(act == "gelu"))
    assert rel_err(basis.detach(), g["basis"]) < 1e-5
    (basis * g["d_basis"]).sum().backward()
    for k, ref in grads.items():
        assert rel_err(leaves[k].grad, ref) < 1e-4, k


@pytest.mark.parametrize("tag", list(CASES))
def test_torch_twin_matches_reference(tag):
    width, act = CASES[tag]
    g, state, grads = load_case(tag)
    net = deform.MotionBasisNetwork(width, 16, 26, False, activation=act)
    net.load_state_dict(state)
    b_t, table = net.query_and_table(g["times"][0], net.batch_embedding(g["times"][1:]))
    basis = torch.cat((b_t.unsqueeze(0), table))
    assert rel_err(basis.detach(), g["basis"]) < 1e-5
    (basis * g["d_basis"]).sum().backward()
    for k, p in net.named_parameters():
        assert rel_err(p.grad, grads[k]) < 1e-4, k


def test_packed_layout_round_trip_and_counts():
    from rodygs_b200 import _lib
    lib = _lib.load()
    for width in (32, 64, 128):
        net = deform.MotionBasisNetwork(width)
        sd = net.state_dict()
        flat = deform.pack_state_dict(sd, 53, width, 16)
        assert flat.numel() == sum(p.numel() for p in net.parameters())
        assert flat.numel() == lib.rdg_basis_mlp_param_count(53, width, 16, 7)
        back = deform.unpack_to_state_dict(flat, 53, width, 16)
        assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
        assert lib.rdg_basis_mlp_saved_floats(53, width, 16) == 53 + 2 * width + width // 2 + 16 * (width // 4)
        assert lib.rdg_basis_mlp_bwd_workspace_bytes(101, 53, width, 16) == 2 * 101 * 4 * lib.rdg_basis_mlp_saved_floats(53, width, 16)
    assert lib.rdg_basis_mlp_param_count(53, 128, 16, 7) == 68656
    assert lib.rdg_basis_mlp_param_count(53, 130, 16, 7) == -1             # width must be a multiple of 4


def test_basis_mlp_refuses_cpu():
    with pytest.raises(RuntimeError, match="CUDA only"):
        deform.BasisMLP(32, device="cpu")


# ---- GPU -------------------------------------------------------------------------------------------------------------

def _native(width, act, state):
    net = deform.BasisMLP(width, 16, 26, False, activation=act)
    net.load_state_dict(state)
    return net


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
@pytest.mark.parametrize("embed_in_kernel", [False, True])
def test_kernels_match_reference_golden(tag, embed_in_kernel):
    width, act = CASES[tag]
    g, state, grads = load_case(tag)
    net = _native(width, act, state)
    if embed_in_kernel:
        basis = net.forward_rows(times=g["times"].cuda())
        emb_used = net._saved[:, :53].cpu()
        # sin/cos of arguments up to 2^25 pi: a few ulp of a value <= 1 between CUDA and the CPU libm
        assert (emb_used - g["emb"]).abs().max().item() < 1e-6
    else:
        basis = net.forward_rows(emb=g["emb"].cuda())
    tol_f = 1e-5 if not embed_in_kernel else 2e-4     # the 1e-7 embedding differences are amplified by the widened weights
    assert rel_err(basis.cpu(), g["basis"]) < tol_f, rel_err(basis.cpu(), g["basis"])
    net.backward_rows(g["d_basis"].cuda())
    got = net.grad_state_dict()
    tol_g = 1e-4 if not embed_in_kernel else 2e-3
    for k, ref in grads.items():
        assert rel_err(got[k].cpu(), ref) < tol_g, (k, rel_err(got[k].cpu(), ref))


@pytest.mark.gpu
def test_kernels_match_torch_twin_at_reference_size_and_are_deterministic():
    torch.manual_seed(3)
    T = 100
    twin = deform.MotionBasisNetwork(128, 16, 26, False)
    with torch.no_grad():
        for p in twin.parameters():
            p.copy_(torch.randn_like(p) * (1.5 / p.shape[1] ** 0.5 if p.dim() > 1 else 0.1))
    times = torch.arange(T, dtype=torch.float32) / T
    t = torch.tensor(0.4321)
    emb = twin.batch_embedding(torch.cat((t.reshape(1), times)))
    b_t, table = twin.query_and_table(t, emb[1:])
    d_bt, d_table = torch.randn_like(b_t), torch.randn_like(table)
    ((b_t * d_bt).sum() + (table * d_table).sum()).backward()

    net = _native(128, "gelu", twin.state_dict())
    basis = net.forward_rows(emb=emb.cuda())
    assert rel_err(basis[0].cpu(), b_t.detach()) < 1e-5
    assert rel_err(basis[1:].cpu(), table.detach()) < 1e-5
    d_basis = torch.cat((d_bt.unsqueeze(0), d_table)).cuda()
    g1 = net.backward_rows(d_basis).clone()
    for k, p in twin.named_parameters():
        assert rel_err(net.grad_state_dict()[k].cpu(), p.grad) < 1e-4, k
    # accumulate adds; a second run is bit-identical (no atomics)
    g2 = net.backward_rows(d_basis, accumulate=True).clone()
    assert torch.equal(g2, g1 + g1)
    g3 = net.backward_rows(d_basis).clone()
    assert torch.equal(g3, g1)

    # the reference's call pattern with autograd into the flat parameter buffer
    bt2, table2 = net.query_and_table(t, times)
    assert rel_err(bt2.detach().cpu(), b_t.detach()) < 5e-4 and rel_err(table2.detach().cpu(), table.detach()) < 5e-4
    ((bt2 * d_bt.cuda()).sum() + (table2 * d_table.cuda()).sum()).backward()
    assert rel_err(net.params.grad, g1) < 5e-3


@pytest.mark.gpu
def test_relu_and_error_paths():
    from rodygs_b200 import _lib
    net = deform.BasisMLP(32, 16, 26, False, activation="relu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.forward_rows(times=torch.zeros(3))
    with pytest.raises(RuntimeError, match="before forward_rows"):
        net.backward_rows(torch.zeros(1, 16, 7, device="cuda"))
    out = net.forward_rows(times=torch.zeros(0, device="cuda"))          # empty batch
    assert out.shape == (0, 16, 7)
    import ctypes as C
    bad = _lib.RdgBasisMlp(53, 30, 16, 7, 0, 1, 0, 0, 0, 0, 0, 0)
    assert _lib.load().rdg_basis_mlp_fwd(C.byref(bad), None) == -1
    assert b"width" in _lib.load().rdg_last_error()


@pytest.mark.gpu
def test_trainer_flow_feeds_table_and_returns_network_gradient():
    """SplatTrainStep.basis_forward / basis_backward: B(t) and the table are written straight into the flat parameter
    buffer (row-0 indirection of the kernels), dL/dB(t) and dL/dtable are read straight from the flat gradient buffer;
    the packed network gradient equals torch autograd of the twin network fed with the same upstream gradients."""
    from rodygs_b200 import synthetic
    from rodygs_b200.trainer import SplatTrainStep
    N, H, W, T = 6000, 96, 128, 7           # T * 112 is not a multiple of 64: table and B(t) slices are not adjacent
    scene = synthetic.to_device(synthetic.make_scene(N, H, W, T, seed=4), "cuda")
    step = SplatTrainStep(scene, H, W, sh_degree=3, w_pearson=0.0)
    cam = synthetic.make_camera(1, 8, H, W, T)
    vm = cam.world_view_transform.t().contiguous().cuda()
    pm = cam.projection_matrix.t().contiguous().cuda()
    gt = torch.rand(3, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.5 + 0.25

    torch.manual_seed(9)
    twin = deform.MotionBasisNetwork(128, 16, 26, False)
    with torch.no_grad():
        for p in twin.parameters():
            p.copy_(torch.randn_like(p) * (0.6 / p.shape[1] ** 0.5 if p.dim() > 1 else 0.02))
    times = torch.arange(T, dtype=torch.float32) / T
    t = 0.4321
    mlp = step.attach_basis_mlp(_native(128, "gelu", twin.state_dict()), times, lr=1e-3)
    basis_t = step.basis_forward(t)
    emb = twin.batch_embedding(torch.cat((torch.tensor([t]), times)))
    b_ref, table_ref = twin.query_and_table(torch.tensor(t), emb[1:])
    assert rel_err(step.p("basis_t").cpu(), b_ref.detach()) < 5e-4
    assert rel_err(step.p("table").cpu(), table_ref.detach()) < 5e-4

    step.forward_backward(vm, pm, cam.tanfovx, cam.tanfovy, basis_t, gt, None)
    d_bt, d_table = step.g("basis_t").cpu().clone(), step.g("table").cpu().clone()
    assert d_bt.abs().max() > 0 and d_table.abs().max() > 0
    step.basis_backward()
    ((b_ref * d_bt).sum() + (table_ref * d_table).sum()).backward()
    # compared per packed block (the 16 heads stacked): the upstream gradients of a real render are tiny and of mixed
    # sign, so a single head's bias gradient can sit at the float32 cancellation noise of its own sum
    # The deformation only sees B(t) - B(t_i), so dL/dB(t) = -sum_i dL/dtable[i] and e.g. the gradient of the last
    # biases is analytically ZERO: both sides hold float32 cancellation noise there.  Hence an absolute floor scaled by
    # the upstream gradient (rows x max|dL/dB| x a few ulp) next to the relative bound.
    want = deform.pack_state_dict({k: p.grad for k, p in twin.named_parameters()}, 53, 128, 16)
    got = mlp.grad.cpu()
    floor = 4e-6 * (T + 1) * max(d_bt.abs().max().item(), d_table.abs().max().item())
    for name, (off, shape) in mlp.layout.items():
        n_el = 1
        for d in shape:
            n_el *= d
        a, b = got[off:off + n_el], want[off:off + n_el]
        err = (a - b).abs().max().item()
        assert err < 5e-3 * b.abs().max().item() + floor, (name, err, b.abs().max().item(), floor)

    # one Adam step on the packed buffer == torch.optim.Adam on the twin (eps 1e-15, constant lr)
    before = mlp.params.detach().clone()
    step.basis_optimizer_step(1)
    opt = torch.optim.Adam(twin.parameters(), lr=1e-3, eps=1e-15)
    opt.step()
    after = mlp.state_dict()
    assert not torch.equal(before, mlp.params.detach())
    for k, p in twin.named_parameters():
        # first Adam step = -lr * sign(g): compare where the gradient is well above its own parity error
        sure = p.grad.abs() > max(2e-2 * p.grad.abs().max().item(), 50 * floor)
        assert ((after[k].cpu() - p.detach()).abs() * sure).max().item() < 2e-5, k
