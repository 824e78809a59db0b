"""GPU parity tests: the CUDA path (through the C ABI / drop-in classes) against
  (1) the committed outputs of the reference's own classes (tests/golden/*.npz), and
  (2) the CPU oracle on freshly seeded inputs.
Tolerances (fp32 path; BASELINE.json north_star: 1e-5 max-abs on probabilities):
  * aggregation            bit-exact (same ascending-column sequential fp32 sum as the CPU reference)
  * one rhs evaluation     6e-6 * max|f| scale-relative, teacher-forced on oracle states;
                           the default kernel variant computes S W^T as a 4-term tf32 split on tcgen05
                           (per-product error 2^-22); the fp32 FFMA variant meets 2e-6 (test_variants_gpu.py)
  * probabilities, small   1e-5 max-abs vs the reference fp32 outputs
  * probabilities, large   err(ours, fp64) <= max(1e-5, 2 * err(ref_fp32, fp64))
"""
import numpy as np
import pytest
import torch

from _util import GOLDEN_CASES, LARGE_CASES, STRICT_CASES, Golden
from oracle import gnode_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PARAM_ORDER = ("odefunc.linear.weight", "odefunc.linear.bias", "linearS1.weight", "linearS1.bias",
               "linear3.weight", "linear3.bias", "linearS2.weight", "linearS2.bias")


@pytest.fixture(scope="module")
def gn():
    import gn_ode_sir_b200 as g
    g.build_library()
    return g


def make_batch(gn, g):
    graphs = [gn.DeviceGraph(A) for A in g.adjs]
    return gn.DeviceBatch([graphs[i] for i in g.inst_graph])


def dev_params(params):
    return [params[k].to(DEV) for k in PARAM_ORDER]


def run_cuda(gn, g, batch=None):
    batch = batch or make_batch(gn, g)
    dt = gn.rollout.dt_array(orc.time_grid(g.maxTime, g.deltaT))
    with torch.no_grad():
        probs = gn.rollout.rollout(g.x.to(DEV), batch, dt, dev_params(g.params))
    torch.cuda.synchronize()
    return probs.cpu()


# ---------------------------------------------------------------- a7 aggregation
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_aggregate_bit_exact(gn, name):
    g = Golden(name)
    batch = make_batch(gn, g)
    v = torch.rand(g.M, 64, generator=torch.Generator().manual_seed(7))
    want = orc.neighbour_sum(v, orc.batch_coo(g.adjs, g.inst_graph))
    got = gn.rollout.aggregate(v.to(DEV), batch).cpu()
    assert torch.equal(got, want), (got - want).abs().max()
    got_t = gn.rollout.aggregate(v.to(DEV), batch, transpose=True).cpu()
    assert torch.equal(got_t, want)          # undirected graphs: A^T == A


def test_aggregate_directed_transpose(gn):
    import scipy.sparse
    rng = np.random.RandomState(0)
    A = scipy.sparse.random(300, 300, density=0.03, random_state=rng, format="csr")
    A.data[:] = 1
    graph = gn.DeviceGraph(A)
    assert not graph.symmetric
    batch = gn.DeviceBatch([graph, graph])
    v = torch.rand(600, 64, generator=torch.Generator().manual_seed(3))
    coo = orc.batch_coo([A], [0, 0])
    assert torch.equal(gn.rollout.aggregate(v.to(DEV), batch).cpu(), orc.neighbour_sum(v, coo))
    cooT = orc.batch_coo([A.T.tocsr()], [0, 0])
    assert torch.equal(gn.rollout.aggregate(v.to(DEV), batch, transpose=True).cpu(), orc.neighbour_sum(v, cooT))


# ---------------------------------------------------------------- a5-a8 one rhs evaluation
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("k", [0, 20, 38])
def test_rhs_teacher_forced(gn, name, k):
    g = Golden(name)
    coo = orc.batch_coo(g.adjs, g.inst_graph)
    _, traj = orc.forward(g.x, g.params, coo, orc.time_grid(g.maxTime, g.deltaT), return_traj=True)
    y = traj[k]
    beta, gamma = g.x[:, 3].contiguous(), g.x[:, 4].contiguous()
    W, b = g.params["odefunc.linear.weight"], g.params["odefunc.linear.bias"]
    want = torch.stack(orc.rhs(y[0], y[1], y[2], beta, gamma, W, b, coo))
    batch = make_batch(gn, g)
    got = gn.rollout.odefunc_eval(y.to(DEV), beta.to(DEV), gamma.to(DEV), batch,
                                  [W.to(DEV), b.to(DEV)] + [b.to(DEV)] * 6).cpu()
    scale = want.abs().max().item()
    # large graphs late in the rollout: |S|, |I| ~ 1e3, so the split product's 2^-22 error relative to sum |s_k w_k| is
    # ~1e-3 absolute in the pre-activation while |z| itself is O(10): 5e-5 there (measured 1.4e-5 on wiki-vote, k = 38)
    bar = 5e-5 if name in LARGE_CASES else 6e-6
    assert (got - want).abs().max().item() <= bar * scale + 1e-30, ((got - want).abs().max().item(), scale)
    # conservation: dS + dI + dR == 0 up to rounding of the last subtraction
    assert (got.sum(0)).abs().max().item() <= 1e-6 * scale


# ---------------------------------------------------------------- a1-a9 whole rollout
@pytest.mark.parametrize("name", STRICT_CASES)
def test_rollout_matches_reference_outputs(gn, name):
    g = Golden(name)
    probs = run_cuda(gn, g)
    assert probs.shape == (g.T, g.M, 3)
    err = (probs[:: g.tstride] - g.probs32).abs().max().item()
    assert err < 1e-5, err
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-6


@pytest.mark.parametrize("name", LARGE_CASES)
def test_rollout_large_graph_against_fp64(gn, name):
    """fb-social, openflights, wiki-vote (max degree 1065: hub relay + CSR-slice overflow on real topology) and the
    five-graph training batch of BASELINE configs[2]: the reference's own fp32 run is 4e-5 .. 2e-3 away from its float64
    run there, so the bar is the float64 run with the reference's own fp32 error as the yardstick."""
    g = Golden(name)
    probs = run_cuda(gn, g)[:: g.tstride]
    ref64 = g.probs64.double()
    err_ours = (probs.double() - ref64).abs().max().item()
    err_ref = (g.probs32.double() - ref64).abs().max().item()
    print("%s: err vs fp64 ours %.3e, reference fp32 %.3e" % (name, err_ours, err_ref))
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)
    # and the bulk of the elements agrees with the reference's fp32 run to the 1e-5 bar
    frac_off = ((probs - g.probs32).abs() > max(1e-5, 0.02 * err_ref)).float().mean().item()
    assert frac_off < 0.01, frac_off


def test_rollout_matches_oracle_fresh_seed(gn):
    g = Golden("sim_dolphins_b4")
    params = orc.default_params(64, seed=123)
    N = g.adjs[0].shape[0]
    B = 7                                     # M = 434: not a multiple of the 128-row tile
    x = torch.cat([orc.synthetic_trial(N, 64, 900 + b) for b in range(B)])
    t = orc.time_grid(10, 0.25)               # a different grid: T=40, dt=0.25
    want = orc.forward(x, params, orc.batch_coo(g.adjs, [0] * B), t)
    graph = gn.DeviceGraph(g.adjs[0])
    batch = gn.DeviceBatch([graph] * B)
    with torch.no_grad():
        got = gn.rollout.rollout(x.to(DEV), batch, gn.rollout.dt_array(t), dev_params(params)).cpu()
    assert (got - want).abs().max().item() < 1e-5


def test_single_grid_point_is_decoded_encoder(gn):
    """T=1: no Euler step; output is the decoded encoder state."""
    g = Golden("sim_karate_b1")
    t = orc.time_grid(0.5, 0.5)
    assert len(t) == 1
    want = orc.forward(g.x, g.params, orc.batch_coo(g.adjs, g.inst_graph), t)
    batch = make_batch(gn, g)
    with torch.no_grad():
        got = gn.rollout.rollout(g.x.to(DEV), batch, gn.rollout.dt_array(t), dev_params(g.params)).cpu()
    assert got.shape == (1, g.M, 3)
    assert (got - want).abs().max().item() < 1e-6


# ---------------------------------------------------------------- size-independent properties
def ba_graph(n, m, seed):
    import networkx as nx
    import scipy.sparse
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    A = scipy.sparse.csr_matrix(nx.adjacency_matrix(G))
    A.data[:] = 1
    return A


def test_trial_independence_and_determinism(gn):
    """Rows of different trials never mix: a trial rolled out alone equals, BITWISE, the same
    trial inside a batch (any position), and reruns are bitwise identical (no atomics)."""
    A = ba_graph(5000, 5, 1)                  # power-law: hub degree >> 16 exercises the chunked gather
    N, B = A.shape[0], 6
    params = dev_params(orc.default_params(64, seed=5))
    graph = gn.DeviceGraph(A)
    assert graph.symmetric and graph.max_degree > 64
    xs = [orc.synthetic_trial(N, 64, 40 + b) for b in range(B)]
    dt = gn.rollout.dt_array(orc.time_grid(20, 0.5))
    with torch.no_grad():
        full = gn.rollout.rollout(torch.cat(xs).to(DEV), gn.DeviceBatch([graph] * B), dt, params)
        again = gn.rollout.rollout(torch.cat(xs).to(DEV), gn.DeviceBatch([graph] * B), dt, params)
        assert torch.equal(full, again)
        for b in (0, 3, 5):
            alone = gn.rollout.rollout(xs[b].to(DEV), gn.DeviceBatch([graph]), dt, params)
            assert torch.equal(alone, full[:, b * N:(b + 1) * N])
    p = full.cpu()
    assert torch.isfinite(p).all() and (p.sum(-1) - 1).abs().max() < 1e-6


def test_medium_graph_against_oracle_fp64(gn):
    A = ba_graph(3000, 4, 2)
    N, B = A.shape[0], 2
    params = orc.default_params(64, seed=9)
    x = torch.cat([orc.synthetic_trial(N, 64, 70 + b) for b in range(B)])
    coo = orc.batch_coo([A], [0] * B)
    t = orc.time_grid(20, 0.5)
    ref32 = orc.forward(x, params, coo, t)
    torch.set_default_dtype(torch.float64)
    try:
        ref64 = orc.forward(x.double(), {k: v.double() for k, v in params.items()}, coo, t)
    finally:
        torch.set_default_dtype(torch.float32)
    graph = gn.DeviceGraph(A)
    with torch.no_grad():
        got = gn.rollout.rollout(x.to(DEV), gn.DeviceBatch([graph] * B), gn.rollout.dt_array(t),
                                 dev_params(params)).cpu()
    err_ours = (got.double() - ref64).abs().max().item()
    err_ref = (ref32.double() - ref64).abs().max().item()
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)


# ---------------------------------------------------------------- drop-in classes
@pytest.mark.parametrize("name", ["sim_karate_b8", "ng_mixed_b5"])
def test_dropin_modules_forward(gn, name):
    g = Golden(name)
    if g.variant == "sim":
        of = gn.ode_sim.ODEfunc(g.adjs[0], 0.2, 0.1, g.H, DEV)
        blk = gn.ode_sim.ODEBlock(g.maxTime, g.deltaT, g.adjs[0].shape[0], [0, 1], g.H, of, DEV)
    else:
        of = gn.ode_ngraphs.ODEfunc(g.adjs, g.H, DEV)
        blk = gn.ode_ngraphs.ODEBlock(g.maxTime, g.deltaT, g.H, of, DEV)
    blk.load_state_dict(g.params)
    blk.to(DEV)
    blk.eval()
    with torch.no_grad():
        S, I, R = blk(g.x_as_model_input().to(DEV))
    assert S.shape == (g.T, g.M, 1) and I.shape == S.shape and R.shape == S.shape
    probs = torch.cat((S, I, R), -1).cpu()
    assert (probs - g.probs32).abs().max().item() < 1e-5
    # ODEfunc.forward(t, y) keeps the reference's packed-state signature
    coo = orc.batch_coo(g.adjs, g.inst_graph)
    _, traj = orc.forward(g.x, g.params, coo, orc.time_grid(g.maxTime, g.deltaT), return_traj=True)
    y, bg = traj[5], g.x[:, 3:]
    want = torch.stack(orc.rhs(y[0], y[1], y[2], g.x[:, 3], g.x[:, 4], g.params["odefunc.linear.weight"],
                               g.params["odefunc.linear.bias"], coo))
    if g.variant == "sim":
        packed = torch.cat((y.reshape(3 * g.M, -1), bg)).to(DEV)
        out = of(torch.tensor(0.0), packed).cpu()
        assert out.shape == (4 * g.M, g.H)
        got, tail = out[:3 * g.M].view(3, g.M, -1), out[3 * g.M:]
    else:
        packed = torch.cat((y, bg.unsqueeze(0))).to(DEV)
        out = of(torch.tensor(0.0), packed).cpu()
        assert out.shape == (4, g.M, g.H)
        got, tail = out[:3], out[3]
    assert float(tail.abs().max()) == 0.0
    assert (got - want).abs().max().item() <= 6e-6 * want.abs().max().item()


# ---------------------------------------------------------------- inference R state: hidden pre-activations vs full plane
@pytest.fixture()
def r_state(gn):
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    prev = L.gnode_get_r_state()
    yield L
    L.gnode_set_r_state(prev)


@pytest.mark.parametrize("name", STRICT_CASES)
def test_inference_r_state_modes(gn, r_state, name):
    """Inference carries R either as a 64-float plane (0: the training forward's arithmetic) or not at all (1, default:
    S + I + R is conserved, so hid(R_k) = W3 (S_0 + I_0 + R_0) - hid(S_k) - hid(I_k)). Both meet the 1e-5 bar against the reference's own
    outputs and agree with each other to fp32 rounding; the S and I dynamics are bitwise the same, so only the R
    logit differs."""
    g = Golden(name)
    out = {}
    for mode in (0, 1):
        assert r_state.gnode_set_r_state(mode) == 0
        assert r_state.gnode_get_r_state() == mode
        out[mode] = run_cuda(gn, g)
        assert (out[mode][:: g.tstride] - g.probs32).abs().max().item() < 1e-5
    assert (out[0] - out[1]).abs().max().item() < 2e-6
    assert torch.equal(out[0][0], out[1][0])                 # t = 0 is the decoded encoder state in both
    assert r_state.gnode_set_r_state(2) != 0


def test_inference_r_state_matches_training_forward(gn, r_state):
    """The full-plane inference rollout is bitwise the forward that stores the trajectory (training)."""
    g = Golden("sim_dolphins_b4")
    r_state.gnode_set_r_state(0)
    inf = run_cuda(gn, g)
    batch = make_batch(gn, g)
    dt = gn.rollout.dt_array(orc.time_grid(g.maxTime, g.deltaT))
    ps = [p.requires_grad_() for p in dev_params(g.params)]
    trn = gn.rollout.rollout(g.x.to(DEV), batch, dt, ps).detach().cpu()
    assert torch.equal(inf, trn)
    r_state.gnode_set_r_state(1)
    assert (run_cuda(gn, g) - trn).abs().max().item() < 2e-6
