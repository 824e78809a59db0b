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
