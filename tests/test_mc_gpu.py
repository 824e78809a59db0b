"""N3: the Monte-Carlo SIR label kernel (gnode_mc_sir, the process of the reference's sir_torch, ode_nn.py:30-88)
against the SHIPPED karate labels (multi-graph-1/Experiments-seed2-karate, carried in tests/golden/train/) and the
CPU restatement of the process. Random streams differ by construction, so parity is statistical: per (compartment,
time, node) z-scores of the difference of two binomial proportions at 10^4 simulations each."""
import numpy as np
import pytest
import torch

from oracle import gnode_oracle as orc
from test_oracle_golden import load_train_golden

pytestmark = pytest.mark.gpu


def zscores(p1, p2, sims1, sims2):
    return (p1 - p2) / np.sqrt(p1 * (1 - p1) / sims1 + p2 * (1 - p2) / sims2 + 1e-12)


@pytest.fixture(scope="module")
def karate():
    import networkx as nx
    z, A, _, _ = load_train_golden()
    G = nx.from_scipy_sparse_array(A)
    return z, A, G


def test_mc_labels_match_shipped_karate_labels(karate):
    from gn_ode_sir_b200 import harness
    z, A, G = karate
    sims, T = 10000, int(z["maxTime"])
    worst, msq = 0.0, []
    for i in range(4):
        S, I, R = harness.monte_carlo_sir(G, [int(s) for s in z["seeds"][i]], float(z["beta"][i]), float(z["gamma"][i]),
                                          sims=sims, T=T, seed=100 + i)
        ours = np.stack((S[0], I[0], R[0])) / sims                          # [3, T, n]
        ref = np.transpose(z["y"][i], (2, 1, 0))                           # y [n, T, 3] -> [3, T, n]
        zz = zscores(ours[:, 1:], ref[:, 1:], sims, sims)
        worst = max(worst, float(np.abs(zz).max())); msq.append(float((zz ** 2).mean()))
        assert np.abs(ours[:, 1:] - ref[:, 1:]).max() < 0.04
    print("MC kernel vs shipped labels: max|z| %.2f, mean z^2 %s" % (worst, ["%.2f" % m for m in msq]))
    assert worst < 5.0 and max(msq) < 1.5


def test_mc_counts_are_consistent_and_reproducible(karate):
    from gn_ode_sir_b200 import harness
    z, A, G = karate
    sims, T = 2000, 20
    a = harness.monte_carlo_sir(G, [3, 30], 0.3, 0.2, sims=sims, T=T, seed=7)
    b = harness.monte_carlo_sir(G, [3, 30], 0.3, 0.2, sims=sims, T=T, seed=7)
    c = harness.monte_carlo_sir(G, [3, 30], 0.3, 0.2, sims=sims, T=T, seed=8)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)                                         # a pure function of (graph, seeds, rates, key)
    assert not np.array_equal(a[1], c[1])
    S, I, R = (v[0] for v in a)
    assert S.shape == (T, A.shape[0]) and S.dtype == np.float64
    assert np.array_equal((S + I + R)[1:], np.full((T - 1, A.shape[0]), float(sims)))    # every simulation is in one state
    ind = np.zeros(A.shape[0]); ind[[3, 30]] = 1
    assert np.array_equal(I[0], ind) and np.array_equal(S[0], 1 - ind) and not R[0].any()   # t = 0: assigned 0/1 state
    assert (np.diff(S[1:], axis=0) <= 0).all() and (np.diff(R[1:], axis=0) >= 0).all()      # S only falls, R only grows
    assert R[1].sum() > 0 and I[1, [3, 30]].min() > 0.7 * sims
    with pytest.raises(ValueError):
        harness.monte_carlo_sir(G, [999], 0.3, 0.2, sims=10, T=5)


def test_mc_kernel_matches_cpu_process_on_a_larger_graph():
    """fb-food-size power-law graph with self-loops, beta / gamma at the ends of the trial range: kernel vs the numpy
    restatement of the process, 4000 simulations each."""
    import networkx as nx
    from gn_ode_sir_b200 import harness
    G = nx.barabasi_albert_graph(600, 3, seed=2)
    G.add_edge(5, 5); G.add_edge(17, 17)
    A = nx.adjacency_matrix(G)
    sims, T = 4000, 12
    for beta, gamma, seeds in ((0.1, 0.5, [0, 7]), (0.5, 0.1, [100]), (0.25, 0.0, [3, 4, 5])):
        S, I, R = harness.monte_carlo_sir(G, seeds, beta, gamma, sims=sims, T=T, seed=3)
        ours = np.stack((S[0], I[0], R[0])) / sims
        ref = orc.mc_sir_counts(A, seeds, beta, gamma, sims, T, np.random.RandomState(4)) / sims
        zz = zscores(ours[:, 1:], ref[:, 1:], sims, sims)
        print("beta %.2f gamma %.2f: max|z| %.2f mean z^2 %.2f" % (beta, gamma, np.abs(zz).max(), (zz ** 2).mean()))
        assert np.abs(zz).max() < 5.5 and (zz ** 2).mean() < 1.5
