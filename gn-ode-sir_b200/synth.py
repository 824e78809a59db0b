"""Synthetic scale-free graphs emitted straight to CSR (no networkx object), for the
benchmark workloads of BASELINE.json: the epinions stand-in (real_graphs/epinions.pkl is
missing from the reference checkout) and the BA stress graphs."""
import numpy as np
import scipy.sparse


def barabasi_albert_csr(n, m, seed=0):
    """Preferential attachment: node v >= m attaches to m distinct earlier nodes drawn
    proportionally to degree (+ the initial m-clique-free star used by the classic model).
    Returns a symmetric scipy CSR (int32 indices, sorted, no self-loops, binary)."""
    rng = np.random.RandomState(seed)
    # 'repeated' holds every edge endpoint once: sampling it uniformly == degree-proportional
    repeated = np.empty(2 * m * n, dtype=np.int64)
    fill = 0
    src = np.empty(m * (n - m), dtype=np.int64)
    dst = np.empty(m * (n - m), dtype=np.int64)
    targets = np.arange(m, dtype=np.int64)
    e = 0
    for v in range(m, n):
        src[e:e + m] = v
        dst[e:e + m] = targets
        e += m
        repeated[fill:fill + m] = targets
        repeated[fill + m:fill + 2 * m] = v
        fill += 2 * m
        # m distinct targets for the next node
        chosen = set()
        while len(chosen) < m:
            cand = repeated[rng.randint(0, fill, size=2 * m)]
            for c in cand:
                chosen.add(int(c))
                if len(chosen) == m:
                    break
        targets = np.fromiter(chosen, dtype=np.int64, count=m)
    rows = np.concatenate((src, dst))
    cols = np.concatenate((dst, src))
    A = scipy.sparse.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1
    A.sort_indices()
    return A


EPINIONS_N = 75879        # node count of SNAP soc-Epinions1


def epinions_standin(seed=0):
    """BA(N=75,879, m=5): ~379k undirected edges, mean degree ~10 (epinions: ~10.7)."""
    return barabasi_albert_csr(EPINIONS_N, 5, seed)


def barabasi_albert_csr_fast(n, m, seed=0, chunk=4096):
    """Chunked preferential attachment for the multi-million-node stress graphs (BASELINE.json configs[4]):
    the nodes of one chunk draw their m targets from the endpoint list as it stood at the start of the chunk
    (degree-proportional up to that lag), so the generator is vectorised and runs in seconds for n = 2e6.
    Duplicate targets of a node are merged (a few nodes end with fewer than m new edges). Symmetric CSR."""
    rng = np.random.RandomState(seed)
    n0 = max(m + 1, min(n, 64))
    a0 = barabasi_albert_csr(n0, m, seed).tocoo()
    src = [a0.row[a0.row > a0.col].astype(np.int64)]
    dst = [a0.col[a0.row > a0.col].astype(np.int64)]
    endpoints = np.empty(2 * m * n + 2 * len(src[0]), dtype=np.int64)
    fill = 2 * len(src[0])
    endpoints[:fill] = np.concatenate((src[0], dst[0]))
    v = n0
    while v < n:
        c = min(chunk, n - v, max(64, v // 4))
        nodes = np.arange(v, v + c, dtype=np.int64)
        t = endpoints[rng.randint(0, fill, size=(c, m))]
        s_ = np.repeat(nodes, m)
        t_ = t.reshape(-1)
        src.append(s_); dst.append(t_)
        endpoints[fill:fill + c * m] = s_
        endpoints[fill + c * m:fill + 2 * c * m] = t_
        fill += 2 * c * m
        v += c
    src = np.concatenate(src); dst = np.concatenate(dst)
    rows = np.concatenate((src, dst)); cols = np.concatenate((dst, src))
    A = scipy.sparse.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1
    A.sort_indices()
    return A


def ba_stress(seed=0):
    """BA(N=2,000,000, m=10): ~20M undirected edges, mean degree ~20 (BASELINE.json configs[4])."""
    return barabasi_albert_csr_fast(2_000_000, 10, seed)


def trial_parameters(n_nodes, trial_id, n_seeds=2):
    """(seeds, beta, gamma) of synthetic trial `trial_id`: the compact descriptor of synthetic_trial's dense block."""
    rng = np.random.RandomState(1000 + trial_id)
    seeds = rng.choice(n_nodes, n_seeds, replace=False)
    beta, gamma = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)
    return seeds, beta, gamma


def synthetic_trial(n_nodes, H, trial_id, n_seeds=2):
    """One [N, 3+H] input block in the layout main() of the reference builds (ode_nn_ngraph_sim.py:371-390): columns
    S0 | I0 | R0 | beta gamma 0...; seeds and rates drawn like monitorer-sim.py:116-119 from RandomState(1000 + id)."""
    import torch
    rng = np.random.RandomState(1000 + trial_id)
    seeds = rng.choice(n_nodes, n_seeds, replace=False)
    beta, gamma = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)
    x = torch.zeros(n_nodes, 3 + H, dtype=torch.float32)
    x[seeds, 1] = 1.0
    x[:, 0] = 1.0 - x[:, 1]
    x[:, 3], x[:, 4] = beta, gamma
    return x
