"""Edge cases of the rollout path the golden fixtures do not reach: hub rows longer than the staged CSR slice,
isolated nodes, self-loops, directed adjacency, and the full-size bench graph (kernel structures against each other)."""
import numpy as np
import pytest
import scipy.sparse
import torch

from oracle import gnode_oracle as orc
from test_parity_gpu import DEV, dev_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gn():
    import gn_ode_sir_b200 as g
    g.build_library()
    return g


def _rollout(gn, A, B, seed, maxTime=20, deltaT=0.5):
    N = A.shape[0]
    params = orc.default_params(64, seed=seed)
    x = torch.cat([orc.synthetic_trial(N, 64, 300 + b) for b in range(B)])
    t = orc.time_grid(maxTime, deltaT)
    want = orc.forward(x, params, orc.batch_coo([A], [0] * B), t)
    graph = gn.DeviceGraph(A)
    with torch.no_grad():
        got = gn.rollout.rollout(x.to(DEV), gn.DeviceBatch([graph] * B), gn.rollout.dt_array(t), dev_params(params)).cpu()
    return got, want, graph


def test_star_hub_isolated_nodes_and_self_loops(gn):
    """One hub with 4000 neighbours (longer than the 1536-entry staged slice and than one 12-row gather round),
    200 isolated nodes (degree 0) and a few self-loops (counted once, like the reference's scatter_add_)."""
    n = 4300
    rows = np.concatenate((np.zeros(4000, dtype=np.int64), np.arange(1, 4001), [7, 9, 4100]))
    cols = np.concatenate((np.arange(1, 4001), np.zeros(4000, dtype=np.int64), [7, 9, 4100]))
    A = scipy.sparse.csr_matrix((np.ones(len(rows), dtype=np.int64), (rows, cols)), shape=(n, n))
    A.sum_duplicates(); A.data[:] = 1; A.sort_indices()
    got, want, graph = _rollout(gn, A, B=3, seed=21, maxTime=6, deltaT=0.5)
    assert graph.max_degree == 4000
    assert torch.isfinite(got).all()
    # the hub's state saturates; compare the probabilities of every node
    assert (got - want).abs().max().item() < 1e-5


def test_directed_graph_rollout(gn):
    """Non-symmetric adjacency: the forward aggregates over the stored (row, col) entries, not their transpose."""
    rng = np.random.RandomState(4)
    A = scipy.sparse.random(700, 700, density=0.01, random_state=rng, format="csr")
    A.data[:] = 1
    A = A.astype(np.int64)
    got, want, graph = _rollout(gn, A, B=2, seed=22, maxTime=10, deltaT=0.5)
    assert not graph.symmetric
    assert (got - want).abs().max().item() < 1e-5


def test_bench_graph_kernel_structures_agree(gn):
    """Full-size epinions stand-in (BA N=75,879: hub tiles, slice overflow, partial last tile), 3 trials: the pipelined
    tensor-core kernels (TMA-fed and LDG-fed operands) against the generic fp32 FFMA kernel that the goldens validate."""
    from gn_ode_sir_b200 import _lib, synth
    L = _lib.lib()
    A = synth.epinions_standin(0)
    N, B = A.shape[0], 3
    params = dev_params(orc.default_params(64, seed=0))
    x = torch.cat([orc.synthetic_trial(N, 64, 500 + b) for b in range(B)]).to(DEV)
    graph = gn.DeviceGraph(A)
    batch = gn.DeviceBatch([graph] * B)
    dt = gn.rollout.dt_array(orc.time_grid(20, 0.5))
    prev_k, prev_v = L.gnode_get_step_kernel(), L.gnode_get_variant()
    out = {}
    try:
        for name, kern, var in (("stream", 5, 3), ("dual", 3, 3), ("ffma", 0, 0)):
            _lib.check(L.gnode_set_step_kernel(kern), "set_step_kernel")
            _lib.check(L.gnode_set_variant(var), "set_variant")
            with torch.no_grad():
                out[name] = gn.rollout.rollout(x, batch, dt, params).cpu()
    finally:
        L.gnode_set_step_kernel(prev_k); L.gnode_set_variant(prev_v)
    for name in ("stream", "dual"):
        err = (out[name] - out["ffma"]).abs().max().item()
        print("%s vs fp32 FFMA kernel on the bench graph: %.3e" % (name, err))
        assert err < 1e-5, (name, err)      # two fp32 arithmetic orders on an ill-conditioned graph (its fp32 run is 1.5e-4 off float64)
    assert (out["dual"].sum(-1) - 1).abs().max().item() < 1e-6


def test_bench_size_batch_against_oracle(gn):
    """BASELINE.json's bench configuration at full graph size (epinions stand-in, BA N=75,879) and a batch large enough
    for the one-launch-per-step regime with the dynamic tile scheduler (16 trials = 1.2M rows): one trial of the batch
    against the CPU oracle (bar 1e-5), the same trial rolled out alone (persistent single-launch regime) bitwise equal
    to its rows inside the batch, probabilities normalised and finite everywhere."""
    from gn_ode_sir_b200 import synth
    A = synth.epinions_standin(0)
    N, B, probe = A.shape[0], 16, 11
    params = orc.default_params(64, seed=0)
    xs = [orc.synthetic_trial(N, 64, b) for b in range(B)]          # the bench recipe (RandomState(1000 + b))
    t = orc.time_grid(20, 0.5)
    graph = gn.DeviceGraph(A)
    dt = gn.rollout.dt_array(t)
    with torch.no_grad():
        full = gn.rollout.rollout(torch.cat(xs).to(DEV), gn.DeviceBatch([graph] * B), dt, dev_params(params))
        alone = gn.rollout.rollout(xs[probe].to(DEV), gn.DeviceBatch([graph]), dt, dev_params(params))
    mine = full[:, probe * N:(probe + 1) * N]
    assert torch.equal(alone, mine)
    assert torch.isfinite(full).all() and (full.sum(-1) - 1).abs().max().item() < 1e-6
    want = orc.forward(xs[probe], params, orc.batch_coo([A], [0]), t)
    err = (mine.cpu() - want).abs().max().item()
    print("bench-size trial vs CPU oracle: %.3e" % err)
    assert err < 1e-5, err


@pytest.mark.parametrize("kernel", [5, 3], ids=["stream", "dual"])
def test_hub_relay_is_bitwise_the_serial_walk(gn, kernel):
    """Rows longer than 512 neighbours are loaded by the whole tile pipeline and added by an in-order relay; the
    result must be BITWISE the serial ascending-column walk (gnode_set_hub_relay(0) switches the relay off), for isolated
    hubs in different tiles, several hubs in one tile, a hub at degree 513 (just above the threshold) and one at 512
    (just below), with batches in the persistent and in the launch-per-step regime."""
    import networkx as nx
    rng = np.random.RandomState(3)
    N = 6000
    A = scipy.sparse.lil_matrix(scipy.sparse.csr_matrix(nx.adjacency_matrix(nx.barabasi_albert_graph(N, 4, seed=2))))
    for hub, deg in ((5, 2500), (6, 1300), (700, 513), (701, 512), (3333, 900), (5999, 777)):
        nb = rng.choice(N, deg, replace=False)
        nb = nb[nb != hub]
        A[hub, nb] = 1
        A[nb, hub] = 1
    A = scipy.sparse.csr_matrix(A)
    A.data[:] = 1
    A.sort_indices()
    deg = np.diff(A.indptr)
    assert (deg > 512).sum() >= 5 and deg.max() > 2400
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    params = dev_params(orc.default_params(64, seed=4))
    graph = gn.DeviceGraph(A)
    dt = gn.rollout.dt_array(orc.time_grid(20, 0.5))
    prev_k = L.gnode_get_step_kernel()
    try:
        _lib.check(L.gnode_set_step_kernel(kernel), "gnode_set_step_kernel")
        for B in (2, 120):                         # 94 tiles (persistent rollout) / 5625 tiles (one launch per step)
            x = torch.cat([orc.synthetic_trial(N, 64, 40 + b) for b in range(B)]).to(DEV)
            batch = gn.DeviceBatch([graph] * B)
            assert L.gnode_set_hub_relay(1) == 0 and L.gnode_get_hub_relay() == 1
            with torch.no_grad():
                relay = gn.rollout.rollout(x, batch, dt, params)
            assert L.gnode_set_hub_relay(0) == 0
            with torch.no_grad():
                serial = gn.rollout.rollout(x, batch, dt, params)
            assert torch.equal(relay, serial), (B, (relay - serial).abs().max().item())
            assert torch.isfinite(relay).all()
    finally:
        L.gnode_set_hub_relay(1)
        L.gnode_set_step_kernel(prev_k)
    # and the relay path itself against the CPU oracle's strictly sequential sum: S' * AI of a hub row after ONE step is
    # exercised by tests/test_parity_gpu.py::test_rollout_large_graph_against_fp64[sim_wikivote_b2] (max degree 1065)


def test_ba2m_trial_against_oracle(gn):
    """BASELINE.json configs[4]: the BA(N=2,000,000, m=10) stress graph (40M stored entries, hubs of degree ~1e4, one
    trial's I' = 512 MB: the gather runs from DRAM, not L2). One trial, maxTime 5 (T = 10), every row against the CPU
    oracle (streaming restatement: chunked neighbour sum, bitwise the oracle's forward, tests/test_oracle_golden.py).
    Hub rows sum ~1e4 neighbours and their hidden state reaches ~1e4 within a few steps, where the reference's own fp32
    run is not reproducible to 1e-5: as for the large golden graphs the criterion is the float64 run with the
    reference's own fp32 error as the yardstick, and all but a handful of rows must agree with the fp32 run to 1e-5."""
    from gn_ode_sir_b200 import synth
    A = synth.ba_stress(0)
    N = A.shape[0]
    assert N == 2_000_000 and A.nnz > 39_000_000
    params = orc.default_params(64, seed=0)
    t = orc.time_grid(5, 0.5)
    x = orc.synthetic_trial(N, 64, 7)
    ref32 = orc.forward_streaming(x, params, A, t)
    torch.set_default_dtype(torch.float64)
    try:
        ref64 = orc.forward_streaming(x.double(), {k: v.double() for k, v in params.items()}, A, t)
    finally:
        torch.set_default_dtype(torch.float32)
    graph = gn.DeviceGraph(A)
    with torch.no_grad():
        got = gn.rollout.rollout(x.to(DEV), gn.DeviceBatch([graph]), gn.rollout.dt_array(t), dev_params(params)).cpu()
    assert got.shape == ref32.shape == (10, N, 3)
    assert torch.isfinite(got).all() and (got.sum(-1) - 1).abs().max().item() < 1e-6
    err_ours = (got.double() - ref64).abs().max().item()
    err_ref = (ref32.double() - ref64).abs().max().item()
    diff = (got - ref32).abs().amax(dim=(0, 2))                       # per node
    n_off = int((diff > 1e-5).sum())
    worst = int(diff.argmax())
    print("BA-2M (T=10): ours vs fp64 %.3e, reference fp32 vs fp64 %.3e, ours vs reference fp32 %.3e (node %d, degree %d); "
          "%d of %d nodes beyond 1e-5" % (err_ours, err_ref, diff.max().item(), worst,
                                         A.indptr[worst + 1] - A.indptr[worst], n_off, N))
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)
    # measured: the reference's fp32 run is 1.5e-2 from float64 here; ours is 1.4e-5 from the reference's fp32 run, with
    # 55 of 2,000,000 nodes beyond 1e-5 -- three orders of magnitude closer to it than it is to exact arithmetic
    assert diff.max().item() <= max(1e-5, 0.01 * err_ref), (diff.max().item(), err_ref)
    assert n_off <= N // 10000, n_off


def test_epinions_standin_maxtime80_against_oracle(gn):
    """BASELINE.json configs[3]'s maxTime sweep at its long end: maxTime 80 (T = 160, 159 Euler steps, where fp32
    drift is largest) on the epinions stand-in, one trial inside a 3-trial batch against the CPU oracle. The reference's
    own fp32 run drifts from its float64 run on this graph (2.9e-4 at T = 40, profiles/r2f_*), so the criterion is the
    one of the large golden graphs: error against float64 <= max(1e-5, 2 x the reference's own fp32 error), and the
    direct fp32 difference is printed."""
    from gn_ode_sir_b200 import synth
    A = synth.epinions_standin(0)
    N, B, probe = A.shape[0], 3, 1
    params = orc.default_params(64, seed=0)
    xs = [orc.synthetic_trial(N, 64, 40 + b) for b in range(B)]
    t = orc.time_grid(80, 0.5)
    assert len(t) == 160
    sel = list(range(0, 160, 8)) + [159]
    ref32 = orc.forward_streaming(xs[probe], params, A, t, out_steps=sel)
    torch.set_default_dtype(torch.float64)
    try:
        ref64 = orc.forward_streaming(xs[probe].double(), {k: v.double() for k, v in params.items()}, A, t, out_steps=sel)
    finally:
        torch.set_default_dtype(torch.float32)
    graph = gn.DeviceGraph(A)
    with torch.no_grad():
        got = gn.rollout.rollout(torch.cat(xs).to(DEV), gn.DeviceBatch([graph] * B), gn.rollout.dt_array(t),
                                 dev_params(params))[:, probe * N:(probe + 1) * N].cpu()[sel]
    err_ours = (got.double() - ref64).abs().max().item()
    err_ref = (ref32.double() - ref64).abs().max().item()
    direct = (got - ref32).abs().max().item()
    print("maxTime 80: ours vs fp64 %.3e, reference fp32 vs fp64 %.3e, ours vs reference fp32 %.3e" % (err_ours, err_ref, direct))
    assert torch.isfinite(got).all()
    assert err_ours <= max(1e-5, 2.0 * err_ref), (err_ours, err_ref)
