"""GPU parity of the reverse sweep (a10) against the reference's own gradients
(tests/golden, produced by loss.backward() through the reference classes with the restated
torchdiffeq adjoint) and against the oracle on fresh seeds.
Tolerance: fp32 gradients are long reductions (over T*M*H terms) whose summation order differs
between implementations; 2e-4 * max|grad| per tensor (the oracle's own fp32-vs-reference noise is
~2e-6, tests/test_oracle_golden.py), and 1e-3 against the float64 gradients."""
import pytest
import torch

from _util import LARGE_CASES, STRICT_CASES, Golden
from oracle import gnode_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_CASES = STRICT_CASES


@pytest.fixture(scope="module")
def gn():
    import gn_ode_sir_b200 as g
    g.build_library()
    return g


def build_block(gn, g, mode):
    if g.variant == "sim":
        of = gn.ode_sim.ODEfunc(g.adjs[0], 0.2, 0.1, g.H, DEV)
        blk = gn.ode_sim.ODEBlock(g.maxTime, g.deltaT, g.adjs[0].shape[0], [0, 1], g.H, of, DEV)
    else:
        of = gn.ode_ngraphs.ODEfunc(g.adjs, g.H, DEV)
        blk = gn.ode_ngraphs.ODEBlock(g.maxTime, g.deltaT, g.H, of, DEV)
    blk.load_state_dict(g.params)
    blk.to(DEV)
    blk.grad_mode = mode
    return blk


def cuda_grads(gn, g, mode):
    blk = build_block(gn, g, mode)
    blk.train()
    S, I, R = blk(g.x_as_model_input().to(DEV))
    probs = torch.cat((S, I, R), -1)
    (probs * g.weight().to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return blk, probs.detach().cpu()


@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_gradients_match_reference(gn, name, mode):
    g = Golden(name)
    blk, probs = cuda_grads(gn, g, mode)
    assert (probs - g.probs32).abs().max().item() < 1e-5       # training-mode forward (trajectory stored)
    got = {k: p.grad.detach().cpu() for k, p in blk.named_parameters() if p.grad is not None}
    for k in orc.GRAD_KEYS:
        ref32, ref64 = g.grads[("g32", mode)][k], g.grads[("g64", mode)][k]
        # floor of 1: linearS2.bias has an exactly-zero true gradient (softmax is shift invariant)
        scale = max(ref64.abs().max().item(), 1.0)
        e32 = (got[k] - ref32).abs().max().item() / scale
        e64 = (got[k].double() - ref64).abs().max().item() / scale
        assert e32 < 2e-4 and e64 < 1e-3, (k, e32, e64)
    if mode == "adjoint":
        # unused LayerNorm of ODEfunc: zero gradients (not None), like torchdiffeq's adjoint params
        assert float(got["odefunc.ln.weight"].abs().max()) == 0.0
        assert float(got["odefunc.ln.bias"].abs().max()) == 0.0
    assert blk.ln.weight.grad is None


@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
@pytest.mark.parametrize("name", [c for c in LARGE_CASES if Golden(c).has_grads])
def test_gradients_large_graphs_against_fp64(gn, name, mode):
    """openflights, wiki-vote (hub rows in the A^T gather of bwd_gz_kernel, ragged tiles at scale) and the five-graph
    training batch. The reference's own fp32 gradients are 2e-3 (train5) .. ~1 (wiki-vote, scale-relative) away from its
    float64 gradients on these graphs, so the yardstick is the float64 gradient with the reference's fp32 error as bar."""
    g = Golden(name)
    blk, probs = cuda_grads(gn, g, mode)
    got = {k: p.grad.detach().cpu() for k, p in blk.named_parameters() if p.grad is not None}
    for k in orc.GRAD_KEYS:
        ref32, ref64 = g.grads[("g32", mode)][k], g.grads[("g64", mode)][k]
        scale = max(ref64.abs().max().item(), 1.0)
        e_ref = (ref32.double() - ref64).abs().max().item() / scale
        e64 = (got[k].double() - ref64).abs().max().item() / scale
        print("%s %s %s: ours %.3e reference fp32 %.3e" % (name, mode, k, e64, e_ref))
        # (wiki-vote: the reference's own fp32 gradient is 20 .. 100 % away from its float64 one -- rounding noise
        # amplified by the dynamics -- so this only says "same order of noise"; the sharp checks on that graph are the
        # short-horizon test below and the per-kernel aggregation / teacher-forced tests)
        assert e64 <= max(1e-3, 4.0 * e_ref), (k, e64, e_ref)


@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
def test_gradients_wikivote_short_horizon(gn, mode):
    """wiki-vote (max degree 1065) with a 6-point grid: the reverse sweep's A^T hub gather, decoder and weight-gradient
    reductions against the CPU oracle in float64. Even at this horizon the fp32 oracle (= the reference's arithmetic) is
    1.5e-3 .. 2.6e-3 scale-relative away from float64 on the [64,64] weight (sums over 14k rows with 1e3-term hub rows),
    so the bar is max(1e-3, 2 x the fp32 oracle's own error), per tensor."""
    g = Golden("sim_wikivote_b2")
    A, N, B = g.adjs[0], g.adjs[0].shape[0], 2
    x = g.x
    t = orc.time_grid(3, 0.5)
    w = torch.randn(len(t), B * N, 3, generator=torch.Generator().manual_seed(11))
    coo = orc.batch_coo([A], [0] * B)
    _, ref32 = orc.loss_and_grads(x, g.params, coo, t, w, mode)
    torch.set_default_dtype(torch.float64)
    try:
        _, want = orc.loss_and_grads(x.double(), {k: v.double() for k, v in g.params.items()}, coo, t, w.double(), mode)
    finally:
        torch.set_default_dtype(torch.float32)
    of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, DEV)
    blk = gn.ode_sim.ODEBlock(3, 0.5, N, [0, 1], 64, of, DEV)
    blk.load_state_dict(g.params)
    blk.to(DEV)
    blk.grad_mode = mode
    S, I, R = blk(x.view(B, N, -1).to(DEV))
    (torch.cat((S, I, R), -1) * w.to(DEV)).sum().backward()
    for k in orc.GRAD_KEYS:
        got = dict(blk.named_parameters())[k].grad.cpu().double()
        scale = max(want[k].abs().max().item(), 1.0)
        err = (got - want[k]).abs().max().item() / scale
        e_ref = (ref32[k].double() - want[k]).abs().max().item() / scale
        print("wiki-vote T=6 %s %s: ours %.3e, fp32 oracle %.3e" % (mode, k, err, e_ref))
        assert err <= max(1e-3, 2.0 * e_ref), (mode, k, err, e_ref)


def test_backward_is_deterministic(gn):
    g = Golden("ng_mixed_b5")
    b1, _ = cuda_grads(gn, g, "adjoint")
    b2, _ = cuda_grads(gn, g, "adjoint")
    for (k, p1), (_, p2) in zip(b1.named_parameters(), b2.named_parameters()):
        if p1.grad is not None:
            assert torch.equal(p1.grad, p2.grad), k


def test_gradients_fresh_seed_uneven_tiles(gn):
    """M = 7*62 = 434 rows (not a tile multiple), non-default grid, both modes vs the oracle in fp64."""
    g = Golden("sim_dolphins_b4")
    A, N, B = g.adjs[0], g.adjs[0].shape[0], 7
    params = orc.default_params(64, seed=77)
    x = torch.cat([orc.synthetic_trial(N, 64, 300 + b) for b in range(B)])
    t = orc.time_grid(6, 0.25)
    w = torch.randn(len(t), B * N, 3, generator=torch.Generator().manual_seed(8))
    coo = orc.batch_coo([A], [0] * B)
    for mode in ("adjoint", "discrete"):
        torch.set_default_dtype(torch.float64)
        try:
            _, want = orc.loss_and_grads(x.double(), {k: v.double() for k, v in params.items()}, coo, t, w.double(), mode)
        finally:
            torch.set_default_dtype(torch.float32)
        of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, DEV)
        blk = gn.ode_sim.ODEBlock(6, 0.25, N, [0, 1], 64, of, DEV)
        blk.load_state_dict(params)
        blk.to(DEV)
        blk.grad_mode = mode
        S, I, R = blk(x.view(B, N, -1).to(DEV))
        (torch.cat((S, I, R), -1) * w.to(DEV)).sum().backward()
        for k in orc.GRAD_KEYS:
            got = dict(blk.named_parameters())[k].grad.cpu().double()
            scale = max(want[k].abs().max().item(), 1.0)
            assert (got - want[k]).abs().max().item() / scale < 1e-3, (mode, k)


def test_training_step_reduces_loss(gn):
    """Three Adam steps on karate through the drop-in module (what train() does,
    ode_nn_ngraph_sim.py:217-246) lower the L1 loss on a fixed target."""
    g = Golden("sim_karate_b8")
    blk = build_block(gn, g, "adjoint")
    opt = torch.optim.Adam(blk.parameters(), lr=1e-2)
    x = g.x_as_model_input().to(DEV)
    target = g.probs64.float().to(DEV).roll(1, dims=-1)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        S, I, R = blk(x)
        loss = torch.nn.functional.l1_loss(torch.cat((S, I, R), -1), target)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


@pytest.fixture
def no_aux(gn):
    """Trajectory-only reverse sweep (four launches, two neighbour gathers per step): the path that also runs when the
    forward could not fill the auxiliary I' / A I' storage."""
    prev = gn.rollout.AUX_STORAGE
    gn.rollout.AUX_STORAGE = False
    yield
    gn.rollout.AUX_STORAGE = prev


@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
@pytest.mark.parametrize("name", ["sim_karate_b8", "ng_mixed_b5"])
def test_gradients_without_auxiliary_storage(gn, no_aux, name, mode):
    """The goldens' gradients through the trajectory-only sweep (the default tests above run with the forward's
    auxiliary storage)."""
    g = Golden(name)
    blk, _ = cuda_grads(gn, g, mode)
    got = {k: p.grad.detach().cpu() for k, p in blk.named_parameters() if p.grad is not None}
    for k in orc.GRAD_KEYS:
        ref32 = g.grads[("g32", mode)][k]
        scale = max(g.grads[("g64", mode)][k].abs().max().item(), 1.0)
        assert (got[k] - ref32).abs().max().item() / scale < 2e-4, k


@pytest.mark.parametrize("persistent", [0, 1], ids=["launch-per-step", "cooperative"])
@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
def test_auxiliary_storage_sweep_equals_trajectory_sweep(gn, mode, persistent):
    """Same batch, same cotangent: the sweep that reads I'_j and A I'_j kept by the forward (three launches, one gather
    per step; dense cotangent in adjoint mode: grid point T-1 has no stored A I' and takes the four-launch step) against
    the trajectory-only sweep. The two differ only in the evaluation order of a few products: 2e-5 of the largest entry.
    M = 5 * 620 + ragged tail: partial last tile, rows past M in the padded I' planes."""
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    g = Golden("sim_fbfood_b2")
    A, N, B = g.adjs[0], g.adjs[0].shape[0], 5
    params = {k: v.to(DEV).requires_grad_(k in orc.GRAD_KEYS) for k, v in orc.default_params(64, seed=5).items()}
    order = gn.rollout.PARAM_ORDER
    x = torch.cat([orc.synthetic_trial(N, 64, 60 + b) for b in range(B)]).to(DEV)
    t = orc.time_grid(8, 0.5)
    dt = gn.rollout.dt_array(t)
    batch = gn.DeviceBatch([gn.DeviceGraph(A)] * B)
    w = torch.randn(len(t), B * N, 3, generator=torch.Generator().manual_seed(3)).to(DEV)
    prev_p = L.gnode_get_persistent()
    grads, probs = {}, {}
    try:
        _lib.check(L.gnode_set_persistent(persistent), "set_persistent")
        for aux in (True, False):
            gn.rollout.AUX_STORAGE = aux
            for p in params.values():
                p.grad = None
            out = gn.rollout.rollout(x, batch, dt, [params[k] for k in order], grad_mode=mode)
            (out * w).sum().backward()
            grads[aux] = {k: params[k].grad.detach().clone() for k in order}
            probs[aux] = out.detach().clone()
    finally:
        gn.rollout.AUX_STORAGE = True
        L.gnode_set_persistent(prev_p)
    assert torch.equal(probs[True], probs[False])            # the forward's arithmetic does not depend on the storage
    for k in order:
        scale = max(grads[False][k].abs().max().item(), 1e-3)
        err = (grads[True][k] - grads[False][k]).abs().max().item() / scale
        assert err < 2e-5, (k, err)


@pytest.mark.parametrize("mode", ["adjoint", "discrete"])
@pytest.mark.parametrize("name", ["sim_karate_b8", "ng_mixed_b5", "sim_fbfood_b2"])
def test_tile_kernels_of_the_reverse_sweep_agree(gn, name, mode):
    """The two tile kernels of the reverse sweep (2 = two CTAs per SM with the next unit prefetched, 1 = round 1's
    one-CTA kernel): every gradient within 2e-5 of the largest entry, and kernel 1 against the reference's own gradients
    at the suite's tolerance (kernel 2 is the default every other test runs)."""
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    g = Golden(name)
    prev = L.gnode_get_bwd_kernel()
    got = {}
    try:
        for k in (2, 1):
            _lib.check(L.gnode_set_bwd_kernel(k), "gnode_set_bwd_kernel")
            blk, _ = cuda_grads(gn, g, mode)
            got[k] = {n: p.grad.detach().cpu() for n, p in blk.named_parameters() if p.grad is not None}
    finally:
        L.gnode_set_bwd_kernel(prev)
    for k in orc.GRAD_KEYS:
        scale = max(got[2][k].abs().max().item(), 1e-3)
        for other in (1,):
            err = (got[other][k] - got[2][k]).abs().max().item() / scale
            assert err < 2e-5, (other, k, err)
        ref32 = g.grads[("g32", mode)][k]
        rscale = max(g.grads[("g64", mode)][k].abs().max().item(), 1.0)
        assert (got[1][k] - ref32).abs().max().item() / rscale < 2e-4, k
