"""Rollout throughput over the BASELINE.json / SURVEY 8d configurations (inference, inputs resident in HBM, CUDA events):
the SHIPPED real graphs (their CSR travels in tests/golden/*.npz and tests/golden/graphs/enron.npz; the pickles do not
exist on the GPU box) next to Barabasi-Albert stand-ins of the same size, trial counts and maxTime sweep on the epinions
stand-in, and a heavier-tailed variant (hub degree ~3k like soc-Epinions1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse, torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth

dev = torch.device("cuda:0")


def heavy_tail(A, n_hubs=8, hub_deg=3000, seed=1):
    """adds n_hubs hubs of degree ~hub_deg to a BA graph (max degree ~3k, like soc-Epinions1)"""
    rng = np.random.RandomState(seed)
    N = A.shape[0]
    rows, cols = [], []
    for h in range(n_hubs):
        nb = rng.choice(N, hub_deg, replace=False)
        nb = nb[nb != h]
        rows += [np.full(len(nb), h), nb]; cols += [nb, np.full(len(nb), h)]
    B = scipy.sparse.csr_matrix((np.ones(sum(len(r) for r in rows), dtype=np.int8), (np.concatenate(rows), np.concatenate(cols))), shape=A.shape)
    C = (A + B).tocsr(); C.sum_duplicates(); C.data[:] = 1; C.sort_indices()
    return C


def run(name, A, B, maxTime, reps=3):
    N = A.shape[0]
    torch.manual_seed(0)
    of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev)
    blk = gn.ode_sim.ODEBlock(maxTime, 0.5, N, [0, 1], 64, of, dev).to(dev).eval()
    x = torch.zeros(B, N, 67)
    for b in range(B):
        rng = np.random.RandomState(1000 + b)
        s = rng.choice(N, 2, replace=False)
        x[b, :, 0] = 1.0; x[b, s, 0] = 0.0; x[b, s, 1] = 1.0
        x[b, :, 3], x[b, :, 4] = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)
    x = x.to(dev)
    T = len(np.arange(0, maxTime, 0.5))
    with torch.no_grad():
        for _ in range(2):
            blk(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            blk(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ns = B * N * (T - 1) / (ms * 1e-3)
    d = np.diff(A.indptr)
    print("%-34s N=%-8d maxdeg=%-6d trials=%-5d T=%-4d %9.3f ms  %.3e node-steps/s  (%.1f %% of 3.18e9)" % (
        name, N, d.max(), B, T, ms, ns, 100 * ns / 3.1826e9), flush=True)
    del blk, of, x
    torch.cuda.empty_cache()


def real_graph(name):
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    fixture = {"karate": ("sim_karate_b1", 0), "dolphins": ("sim_dolphins_b4", 0), "fb-food": ("sim_fbfood_b2", 0),
               "fb-social": ("sim_fbsocial_b1", 0), "openflights": ("sim_openflights_b2", 0), "wiki-vote": ("sim_wikivote_b2", 0)}
    if name == "enron":
        z = np.load(os.path.join(root, "graphs", "enron.npz")); ip, ix = z["indptr"], z["indices"]
    else:
        z = np.load(os.path.join(root, fixture[name][0] + ".npz")); ip, ix = z["g0_indptr"], z["g0_indices"]
    n = len(ip) - 1
    return scipy.sparse.csr_matrix((np.ones(len(ix), dtype=np.int64), ix, ip), shape=(n, n))


if "--small" in sys.argv:       # the latency-bound batches only (at most one tile per pipeline): hub-relay threshold A/B
    for name, Bs in (("karate", (1, 8)), ("fb-food", (8, 32)), ("fb-social", (8, 16)), ("openflights", (8,)), ("wiki-vote", (2, 4))):
        for B in Bs:
            run(name + " (real)", real_graph(name), B, 20, reps=10)
    sys.exit(0)

ep = synth.epinions_standin(0)
print("--- shipped real graphs (reference real_graphs/*.pkl, largest connected component)")
run("karate (real)", real_graph("karate"), 1, 20)
run("karate (real)", real_graph("karate"), 8, 20)
for B in (8, 64, 512):
    run("fb-social (real)", real_graph("fb-social"), B, 20)
run("openflights (real)", real_graph("openflights"), 512, 20)
for B in (64, 512):
    run("wiki-vote (real)", real_graph("wiki-vote"), B, 20)
for B in (32, 128, 256):
    run("enron (real)", real_graph("enron"), B, 20)
print("--- Barabasi-Albert stand-ins of the same sizes")
run("karate-size BA(34,2)", synth.barabasi_albert_csr(34, 2, 0), 1, 20)
for B in (8, 64, 512):
    run("fb-social-size BA(1893,7)", synth.barabasi_albert_csr(1893, 7, 0), B, 20)
run("wiki-vote-size BA(7066,14)", synth.barabasi_albert_csr(7066, 14, 0), 64, 20)
run("enron-size BA(33696,5)", synth.barabasi_albert_csr(33696, 5, 0), 128, 20)
for B in (16, 64, 128):
    run("epinions stand-in BA(75879,5)", ep, B, 20)
for mt in (10, 40, 80):
    run("epinions stand-in, maxTime sweep", ep, 128, mt)
run("epinions stand-in + 8 hubs deg 3k", heavy_tail(ep), 128, 20)
