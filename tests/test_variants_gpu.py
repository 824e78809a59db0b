"""Parity of every forward kernel variant (FFMA / tcgen05 3xTF32, accurate / MUFU sigmoid)
against the reference's outputs, and their agreement with one another."""
import pytest
import torch

from _util import LARGE_CASES, STRICT_CASES, Golden
from oracle import gnode_oracle as orc
from test_parity_gpu import DEV, dev_params, make_batch, run_cuda

pytestmark = pytest.mark.gpu
VARIANTS = {0: "ffma+expf", 1: "tcgen05+expf", 2: "ffma+mufu", 3: "tcgen05+mufu"}


@pytest.fixture(scope="module")
def gn():
    import gn_ode_sir_b200 as g
    g.build_library()
    return g


@pytest.fixture
def variant(gn, request):
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    prev = L.gnode_get_variant()
    _lib.check(L.gnode_set_variant(request.param), "gnode_set_variant")
    yield request.param
    L.gnode_set_variant(prev)


@pytest.mark.parametrize("variant", list(VARIANTS), indirect=True, ids=list(VARIANTS.values()))
@pytest.mark.parametrize("name", STRICT_CASES)
def test_variant_rollout_matches_reference(gn, variant, name):
    g = Golden(name)
    probs = run_cuda(gn, g)
    err = (probs[:: g.tstride] - g.probs32).abs().max().item()
    print("variant %d %s: max|cuda - reference| = %.3e" % (variant, name, err))
    assert err < 1e-5, err


@pytest.mark.parametrize("variant", list(VARIANTS), indirect=True, ids=list(VARIANTS.values()))
def test_variant_rhs_teacher_forced(gn, variant):
    """One f(t,y) on oracle states. fp32 FFMA path: 2e-6 scale-relative (summation-order noise).
    Tensor path: the 4-term tf32 split carries a per-product error of 2^-22 (vs 2^-24 for fp32), and the
    compensated MUFU sigmoid 4.6 ulp (vs 3.2 ulp for expf): bound 6e-6 scale-relative."""
    for name in ("sim_fbfood_b2", "sim_fbsocial_b1"):
        g = Golden(name)
        coo = orc.batch_coo(g.adjs, g.inst_graph)
        _, traj = orc.forward(g.x, g.params, coo, orc.time_grid(g.maxTime, g.deltaT), return_traj=True)
        batch = make_batch(gn, g)
        W, b = g.params["odefunc.linear.weight"], g.params["odefunc.linear.bias"]
        for k in (0, 38):
            y = traj[k]
            want = torch.stack(orc.rhs(y[0], y[1], y[2], g.x[:, 3].contiguous(), g.x[:, 4].contiguous(), W, b, coo))
            got = gn.rollout.odefunc_eval(y.to(DEV), g.x[:, 3].contiguous().to(DEV), g.x[:, 4].contiguous().to(DEV),
                                          batch, [W.to(DEV), b.to(DEV)] + [b.to(DEV)] * 6).cpu()
            scale = want.abs().max().item()
            err = (got - want).abs().max().item()
            print("variant %d %s k=%d: rhs err %.3e (scale %.3e, rel %.2e)" % (variant, name, k, err, scale, err / scale))
            assert err <= (2e-6 if variant == 0 else 6e-6) * scale, (name, k, err, scale)


@pytest.mark.parametrize("variant", [1, 3], indirect=True, ids=["tcgen05+expf", "tcgen05+mufu"])
def test_variant_large_graph_fp64(gn, variant):
    g = Golden("sim_fbsocial_b1")
    probs = run_cuda(gn, g)[:: g.tstride]
    ref64 = g.probs64.double()
    err_ours = (probs.double() - ref64).abs().max().item()
    err_ref = (g.probs32.double() - ref64).abs().max().item()
    print("variant %d fb-social: err vs fp64 ours %.3e, reference fp32 %.3e" % (variant, err_ours, err_ref))
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)


STEP_KERNELS = {5: "stream", 6: "stream-barrier", 7: "stream-3xtf32", 3: "dual", 0: "generic"}


@pytest.fixture
def step_kernel(gn, request):
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    prev = L.gnode_get_step_kernel()
    _lib.check(L.gnode_set_step_kernel(request.param), "gnode_set_step_kernel")
    yield request.param
    L.gnode_set_step_kernel(prev)


@pytest.mark.parametrize("step_kernel", list(STEP_KERNELS), indirect=True, ids=list(STEP_KERNELS.values()))
@pytest.mark.parametrize("name", STRICT_CASES)
def test_step_kernel_rollout_matches_reference(gn, step_kernel, name):
    """Every structure of the tensor-core step kernel against the reference's own outputs (bar 1e-5)."""
    g = Golden(name)
    probs = run_cuda(gn, g)
    err = (probs[:: g.tstride] - g.probs32).abs().max().item()
    print("step kernel %d %s: max|cuda - reference| = %.3e" % (step_kernel, name, err))
    assert err < 1e-5, err


@pytest.mark.parametrize("step_kernel", [5, 3], indirect=True, ids=["stream", "dual"])
@pytest.mark.parametrize("name", LARGE_CASES)
def test_stream_kernel_large_graphs(gn, step_kernel, name):
    """The TMA-fed step kernels on the power-law goldens (hub relay, CSR-slice overflow, ragged multi-graph tiles)."""
    g = Golden(name)
    probs = run_cuda(gn, g)[:: g.tstride]
    ref64 = g.probs64.double()
    err_ours = (probs.double() - ref64).abs().max().item()
    err_ref = (g.probs32.double() - ref64).abs().max().item()
    print("step kernel %d %s: err vs fp64 ours %.3e, reference fp32 %.3e" % (step_kernel, name, err_ours, err_ref))
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)


def test_step_kernels_agree_on_training_trajectory(gn):
    """Forward with a stored trajectory (training): the pipelined kernels' probabilities of EVERY grid point, including
    the first (encoder) and the last (decode kernel), agree with the generic kernel's."""
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    g = Golden("sim_fbfood_b2")
    prev = L.gnode_get_step_kernel()
    out = {}
    try:
        for k in (3, 5, 6, 0):
            _lib.check(L.gnode_set_step_kernel(k), "gnode_set_step_kernel")
            ps = [p.requires_grad_() for p in dev_params(g.params)]
            dt = gn.rollout.dt_array(orc.time_grid(g.maxTime, g.deltaT))
            out[k] = gn.rollout.rollout(g.x.to(DEV), make_batch(gn, g), dt, ps).detach().cpu()
    finally:
        L.gnode_set_step_kernel(prev)
    for k in (3, 5, 6):
        err = (out[k] - out[0]).abs().max().item()
        print("step kernel %d vs generic (training forward): %.3e" % (k, err))
        assert err < 2e-6, err
    # TMA-fed and LDG-fed pipelines feed the tensor core the same hi / lo operands and run the same update: bitwise equal
    assert torch.equal(out[5], out[3]) and torch.equal(out[6], out[3])


@pytest.mark.parametrize("r_state", [0, 1], ids=["r-plane", "conserved-sum"])
@pytest.mark.parametrize("persistent", [0, 1], ids=["launch-per-step", "cooperative"])
@pytest.mark.parametrize("name", ["sim_fbfood_b2", "ng_mixed_b5"])
def test_pipelined_kernels_are_bitwise_equal_in_inference(gn, name, persistent, r_state):
    """The TMA-fed kernel evaluates row sums, products, accumulator sums and the sigmoid on packed fp32 pairs (FADD2 /
    FMUL2 / FFMA2), the LDG-fed one in scalar form: the same IEEE operations, so the probabilities are bitwise equal in
    every instantiation (with and without the R plane, one launch per step and cooperative). This is the test that catches
    a fused multiply-add the scalar form does not have (ptxas fuses mul.rn.f32x2 + add.rn.f32x2)."""
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    g = Golden(name)
    prev = L.gnode_get_step_kernel(), L.gnode_get_persistent(), L.gnode_get_r_state()
    out = {}
    try:
        L.gnode_set_persistent(persistent); L.gnode_set_r_state(r_state)
        for k in (3, 5):
            _lib.check(L.gnode_set_step_kernel(k), "gnode_set_step_kernel")
            out[k] = run_cuda(gn, g)
    finally:
        L.gnode_set_step_kernel(prev[0]); L.gnode_set_persistent(prev[1]); L.gnode_set_r_state(prev[2])
    assert torch.equal(out[3], out[5])


@pytest.mark.parametrize("persistent", [0, 1], ids=["launch-per-step", "cooperative"])
@pytest.mark.parametrize("step_kernel", [3, 5, 6], indirect=True, ids=["dual", "stream", "stream-barrier"])
@pytest.mark.parametrize("name", ["sim_fbfood_b2", "ng_mixed_b5", "sim_wikivote_b2"])
def test_launch_structures_agree(gn, step_kernel, persistent, name):
    """One launch per Euler step vs the persistent cooperative rollout (per-step operands and TMA maps derived in the
    kernel), inference and training forward: same probabilities as the reference within the case's bar."""
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    g = Golden(name)
    prev = L.gnode_get_persistent()
    try:
        _lib.check(L.gnode_set_persistent(persistent), "gnode_set_persistent")
        inf = run_cuda(gn, g)
        ps = [p.requires_grad_() for p in dev_params(g.params)]
        dt = gn.rollout.dt_array(orc.time_grid(g.maxTime, g.deltaT))
        trn = gn.rollout.rollout(g.x.to(DEV), make_batch(gn, g), dt, ps).detach().cpu()
    finally:
        L.gnode_set_persistent(prev)
    ref64 = g.probs64.double()
    err_ref = (g.probs32.double() - ref64).abs().max().item()
    for what, p in (("inference", inf), ("training forward", trn)):
        err = (p[:: g.tstride].double() - ref64).abs().max().item()
        print("kernel %d persistent %d %s %s: err vs fp64 %.3e (reference fp32 %.3e)" % (step_kernel, persistent, name, what, err, err_ref))
        assert err <= max(1e-5, 2.0 * err_ref), (what, err, err_ref)
