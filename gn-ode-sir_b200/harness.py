"""Host-side harness shared by the drop-in experiment scripts at the repo root
(``ode_nn.py`` helpers, ``ode_nn_ngraph_sim.py``, ``ode_nn_ngraphs.py``).

It mirrors the reference's callers of the hot path -- graph loading (ode_nn.py:394-414),
time sub-sampling of the trajectories (ode_nn.py:249-261), the L1 train / eval loops
(ode_nn_ngraph_sim.py:208-295, ode_nn_ngraphs.py:198-264) and the CSV bookkeeping
(ode_nn.py:374-392) -- so that ``monitorer-sim.py`` / ``monitorer-ngraphs.py`` drive this
repo unchanged. Everything that touches the rollout goes through the CUDA drop-in modules.
"""
import csv
import os
import pickle
import time

import numpy as np
import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- graphs (N2)
class GraphHandle:
    """What create_graph hands back as `G` when the adjacency came from the cache: node labels in CSR order plus the CSR
    itself -- enough for everything the experiment scripts do with G (node / edge counts, Monte-Carlo labels). The
    networkx object is rebuilt from the pickle only if something else is asked of it."""

    def __init__(self, path, nodes, A):
        self._path, self._nodes, self.A, self._nx = path, list(nodes), A, None

    def nodes(self):
        return list(self._nodes)

    def number_of_nodes(self):
        return len(self._nodes)

    def number_of_edges(self):
        diag = int(self.A.diagonal().sum())
        return (int(self.A.nnz) - diag) // 2 + diag

    def to_networkx(self):
        if self._nx is None:
            import networkx as nx
            with open(self._path, "rb") as fh:
                G = pickle.load(fh).to_undirected()
            self._nx = G.subgraph(max(nx.connected_components(G), key=len))
        return self._nx

    def __getattr__(self, name):                # anything else: the real networkx graph
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.to_networkx(), name)


def _graph_cache_file(path):
    import hashlib
    st = os.stat(path)
    root = os.environ.get("GNODE_GRAPH_CACHE", os.path.join(os.path.expanduser("~"), ".cache", "gnode_b200", "graphs"))
    key = hashlib.sha1(os.path.abspath(path).encode()).hexdigest()[:16]
    return os.path.join(root, "%s-%d-%d.npz" % (key, st.st_mtime_ns, st.st_size))


def largest_component_csr(G):
    """networkx graph -> (node labels, scipy CSR) of its largest connected component, WITHOUT nx.adjacency_matrix and
    without building a subgraph view: the pattern the reference gets from G.to_undirected() -> G.subgraph(max(cc)) ->
    nx.adjacency_matrix (ode_nn.py:394-414) -- rows in G's node order, binary, symmetric, a self-loop stored once."""
    import scipy.sparse
    import scipy.sparse.csgraph
    nodes = list(G.nodes())
    pos = {v: i for i, v in enumerate(nodes)}
    n = len(nodes)
    e = np.fromiter((pos[w] for uv in G.edges() for w in uv), dtype=np.int64, count=2 * G.number_of_edges()).reshape(-1, 2)
    rows, cols = np.concatenate((e[:, 0], e[:, 1])), np.concatenate((e[:, 1], e[:, 0]))
    A = scipy.sparse.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1
    n_comp, lab = scipy.sparse.csgraph.connected_components(A, directed=False)
    if n_comp > 1:
        sizes = np.bincount(lab)
        # max(nx.connected_components(G), key=len) keeps the FIRST largest component in discovery order
        first_seen = np.full(n_comp, n, dtype=np.int64)
        np.minimum.at(first_seen, lab, np.arange(n))
        best = min(np.flatnonzero(sizes == sizes.max()), key=lambda c: first_seen[c])
        keep = np.flatnonzero(lab == best)
        A = A[keep][:, keep].tocsr()
        nodes = [nodes[i] for i in keep]
    A = A.astype(np.int64)
    A.sort_indices()
    return nodes, A


def load_graph(graph_label, n_nodes=50):
    """pickle -> undirected -> largest connected component -> CSR (reference create_graph, ode_nn.py:394-414), cached on
    disk per (path, mtime, size): a second run of a script on the same dataset neither imports networkx nor touches the
    pickle (enron: 3.3 s of networkx start-up and graph work). Returns (G, A)."""
    import scipy.sparse
    if graph_label == "none":
        import networkx as nx
        G = nx.fast_gnp_random_graph(n_nodes, 0.2)
        return G, nx.adjacency_matrix(G)
    path = graph_label + ".pkl"
    cache = _graph_cache_file(path)
    nodes = A = None
    if os.path.exists(cache):
        try:
            z = np.load(cache, allow_pickle=False)
            n = len(z["indptr"]) - 1
            A = scipy.sparse.csr_matrix((np.ones(len(z["indices"]), dtype=np.int64), z["indices"], z["indptr"]), shape=(n, n))
            nodes = z["nodes"].tolist()
        except Exception:
            nodes = A = None                                  # unreadable cache entry: rebuild it
    if A is None:
        with open(path, "rb") as fh:
            G = pickle.load(fh)
        nodes, A = largest_component_csr(G)
        try:
            node_arr = np.asarray(nodes)
            if node_arr.dtype != object:                      # plain integer / string labels only
                os.makedirs(os.path.dirname(cache), exist_ok=True)
                tmp = cache + ".%d.tmp.npz" % os.getpid()
                np.savez(tmp, indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), nodes=node_arr)
                os.replace(tmp, cache)
        except OSError:
            pass                                              # read-only home: run uncached
    G = GraphHandle(path, nodes, A)
    print("nodes", G.number_of_nodes())
    print("edges", G.number_of_edges())
    return G, A


# ----------------------------------------------------------------------------- trajectories
def sample_unit_times(x, maxTime, deltaT, count=True):
    """Rows int(i/deltaT), i = 0..maxTime-1, of a [T, nodes] trajectory (the reference copies them
    one by one into a CPU tensor; here it is one gather on the tensor's own device, and autograd
    flows through it the same way). count=True returns the per-time node sums instead."""
    idx = torch.tensor([int(i / deltaT) for i in range(int(maxTime))], device=x.device)
    if count:
        return x.sum(dim=1).index_select(0, idx)
    return x.index_select(0, idx)


def prediction_at_unit_times(S, I, R, maxTime, deltaT):
    """[T, M, 1] x3 -> [M, maxTime, 3] as compared with the labels by the loss."""
    cols = [sample_unit_times(torch.squeeze(c, -1), maxTime, deltaT, count=False) for c in (S, I, R)]
    return torch.stack(cols, dim=-1).transpose(0, 1)


def l1_on_rollout(model, criterion, x, y, maxTime, deltaT, instances=None, grad_scale=1.0):
    """Forward + L1 loss on t >= 1 (t = 0 is the given initial condition). Returns (loss, n_items).

    N1: with the drop-in modules, float64 labels and the reference's criterion (nn.L1Loss, mean) the rollout decodes and
    stores ONLY the unit-time grid points int(i/deltaT) (what get_sir_t_nodes_torch would copy out, ode_nn.py:249-261),
    and one kernel computes the float64 loss and the sparse cotangent dL/dprobs that the reverse sweep reads
    (gn_ode_sir_b200.rollout.l1_subsampled). Anything else takes the generic torch path below."""
    from . import rollout as _ro
    steps = _ro.unit_time_steps(maxTime, deltaT)
    fused = (hasattr(model, "rollout_probs") and isinstance(criterion, nn.L1Loss) and criterion.reduction == "mean"
             and y.dtype == torch.float64 and x.is_cuda and len(steps) > 1 and bool(np.all(np.diff(steps) > 0)))
    if fused:
        kw = {"instances": instances} if instances is not None else {}
        probs = model.rollout_probs(x, out_steps=steps, **kw)                    # [maxTime, M, 3]
        target = y.reshape(-1, y.size(-2), y.size(-1))                          # [M, maxTime, 3] float64
        loss = _ro.l1_subsampled(probs, target, skip=1, scale=grad_scale)
        return loss, 3 * (probs.size(0) - 1) * probs.size(1)
    S, I, R = model(x)
    pred = prediction_at_unit_times(S, I, R, maxTime, deltaT)
    target = y.reshape(-1, y.size(-2), y.size(-1))
    loss = criterion(pred[:, 1:, :], target[:, 1:, :])        # float64 labels promote the loss, as in the reference
    if grad_scale != 1.0:                                     # value unchanged, cotangent scaled (data-parallel share)
        loss = loss.detach() + (loss - loss.detach()) * grad_scale
    return loss, 3 * (pred.size(1) - 1) * pred.size(0)


# ----------------------------------------------------------------------------- data parallelism
def init_distributed():
    """One process per GPU under torchrun (WORLD_SIZE > 1): NCCL process group, this rank's GPU made current.
    Returns (rank, world, device). A plain `python script.py` run is rank 0 of 1 on the default device."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if not torch.cuda.is_available():
        return rank, world, torch.device("cpu")
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        # several ranks on one GPU (tests on a single-GPU box) cannot use NCCL: gloo moves the 18 KB gradient instead
        backend = os.environ.get("GNODE_DIST_BACKEND", "nccl" if torch.cuda.device_count() >= world else "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(backend)
    return rank, world, dev


def _dp():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def shared_seed():
    """A random integer all ranks agree on (rank 0's draw): the ranks must shuffle their loaders identically, because
    every global mini-batch is split across them."""
    dist, rank, world = _dp()
    seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)
    if dist is not None:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        seed = seed.to(dev)
        dist.broadcast(seed, src=0)
    return int(seed.item())


def shard_batch(batch, maxTime):
    """This rank's share of one GLOBAL mini-batch and what the loss needs to stay the global mean.

    batch: (x [B, N, 3+H], y [B, N, maxTime, 3]) of the single-graph script -- split by trials -- or
           (list of (x_i, y_i, graph_id_i)) of the multi-graph script -- split by instances, balanced by node count
           (parallel.shard_instances). Returns (x, y, instances or None, share) with share = local items / global items
           (x is None when this rank's share is empty)."""
    from . import parallel
    dist, rank, world = _dp()
    if isinstance(batch, (tuple, list)) and len(batch) == 2 and torch.is_tensor(batch[0]):
        x, y = batch
        if dist is None:
            return x, y, None, 1.0
        lo, hi = parallel.shard_trials(x.size(0), world, rank)
        if hi == lo:
            return None, None, None, 0.0
        return x[lo:hi], y[lo:hi], None, (hi - lo) / float(x.size(0))
    items = list(batch)
    mine = range(len(items)) if dist is None else parallel.shard_instances([it[0].size(0) for it in items], world)[rank]
    if len(mine) == 0:
        return None, None, [], 0.0
    x = torch.cat([items[i][0] for i in mine])
    y = torch.cat([items[i][1] for i in mine])
    total = sum(it[0].size(0) for it in items)
    return x, y, [items[i][2] for i in mine], x.size(0) / float(total)


def run_epoch(model, optimizer, criterion, device, train_batches, val_batches, maxTime, deltaT):
    """One epoch of Adam steps followed by a validation sweep; returns item-weighted mean losses.

    Under torchrun (data parallel) every global mini-batch is split across the ranks (shard_batch); each rank
    back-propagates its share of the GLOBAL mean loss, ONE flat all-reduce sums the 4.6k-float gradient
    (parallel.allreduce_gradients) and every rank applies the identical Adam step -- the update of the single-process
    run on the whole batch."""
    from . import parallel
    dist, rank, world = _dp()
    model.train()
    # the item-weighted loss sums stay on the device: one host read per epoch instead of one per batch (the
    # reference's loss.item() per batch, ode_nn_ngraph_sim.py:247, drains the stream between optimiser steps)
    sums = torch.zeros(4, dtype=torch.float64, device=device)          # train loss x items, items, val loss x items, items
    fwd_time = 0.0
    for batch in train_batches:
        x, y, inst, share = shard_batch(batch, maxTime)
        optimizer.zero_grad()
        if x is not None:
            x, y = x.to(device), y.to(device)
            t0 = time.time()
            loss, n = l1_on_rollout(model, criterion, x, y, maxTime, deltaT, instances=inst, grad_scale=share)
            fwd_time += time.time() - t0
            loss.backward()                                            # cotangent already carries this rank's share
            sums[0] += loss.detach().double() * n
            sums[1] += n
        if dist is not None:
            parallel.allreduce_gradients(model.parameters())
        optimizer.step()
    model.eval()
    with torch.no_grad():
        for batch in val_batches:
            x, y, inst, share = shard_batch(batch, maxTime)
            if x is None:
                continue
            loss, n = l1_on_rollout(model, criterion, x.to(device), y.to(device), maxTime, deltaT, instances=inst)
            sums[2] += loss.double() * n
            sums[3] += n
    if dist is not None:
        buf = sums if dist.get_backend() == "nccl" else sums.cpu()
        dist.all_reduce(buf)
        sums = buf
    tot, items, vtot, vitems = (float(v) for v in sums.tolist())
    if rank == 0:
        print("Time: ", fwd_time)
    return tot / max(items, 1), vtot / max(vitems, 1)


def evaluate(model, criterion, device, batches, maxTime, deltaT):
    dist, rank, world = _dp()
    model.eval()
    sums = torch.zeros(2, dtype=torch.float64, device=device)
    per_batch = []
    with torch.no_grad():
        for batch in batches:
            x, y, inst, share = shard_batch(batch, maxTime)
            part = torch.zeros(2, dtype=torch.float64, device=device)
            if x is not None:
                loss, n = l1_on_rollout(model, criterion, x.to(device), y.to(device), maxTime, deltaT, instances=inst)
                part[0], part[1] = loss.double() * n, n
            if dist is not None:
                buf = part if dist.get_backend() == "nccl" else part.cpu()
                dist.all_reduce(buf)
                part = buf.to(device)
            sums += part
            per_batch.append(float(part[0] / part[1].clamp(min=1)))
    return float(sums[0] / sums[1].clamp(min=1)), per_batch


def fit(model, device, lr, epochs, train_batches, val_batches, test_batches, maxTime, deltaT):
    """Adam + L1 with test evaluation whenever the validation loss improves."""
    from . import parallel
    dist, rank, world = _dp()
    if dist is not None:
        parallel.broadcast_parameters(model, src=0)                  # identical replicas before the first step
    criterion = nn.L1Loss()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    best = dict(val=np.inf, epoch=-1, test=float("nan"), test_all=[], test_time=0.0)
    if rank == 0:
        print("training...")
    for epoch in range(epochs):
        loss, val_loss = run_epoch(model, optimizer, criterion, device, train_batches, val_batches, maxTime, deltaT)
        if rank == 0:
            print("Epoch: {:03d}, Train Loss: {:.10f}, Val Loss: {:.10f}".format(epoch, loss, val_loss))
        if val_loss < best["val"]:
            t0 = time.time()
            test_loss, test_all = evaluate(model, criterion, device, test_batches, maxTime, deltaT)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            best.update(val=val_loss, epoch=epoch, test=test_loss, test_all=test_all, test_time=time.time() - t0)
    return best


# ----------------------------------------------------------------------------- bookkeeping
def append_csv(path_to_csv, columns, values):
    fresh = not os.path.exists(path_to_csv)
    with open(path_to_csv, "w" if fresh else "a+", newline="") as fh:
        w = csv.writer(fh)
        if fresh:
            w.writerow(columns)
        w.writerow(values)
    import pandas as pd
    print(pd.read_csv(path_to_csv))


# ----------------------------------------------------------------------------- Monte-Carlo labels
_mc_graphs = {}


def monte_carlo_sir(G, seed_set, beta, gamma, sims=10000, T=20, device=None, seed=None):
    """Discrete-time Monte-Carlo SIR ground truth: the process of the reference's sir_torch (ode_nn.py:30-88) run by
    the hand-written kernel behind gnode_mc_sir (one CTA per simulation, bit-packed state, Philox random numbers)
    instead of sims x (T-1) iterations of tiny tensor ops with CPU random numbers. Returns per-time-step node COUNTS
    over the simulations, S, I, R each [1, T, n] float64 (t = 0 rows: the 0/1 initial state, like the reference);
    labels = counts / sims. `seed`: RNG key (None = drawn from the OS: runs differ, as with the reference)."""
    import ctypes
    import networkx as nx
    from . import _lib
    from .graph import DeviceGraph
    if not torch.cuda.is_available():
        raise RuntimeError("monte_carlo_sir needs a CUDA device (the reference's sir_torch does too, ode_nn.py:41-50)")
    dev = torch.device(device or "cuda")
    nodes = list(G.nodes())
    pos = {v: i for i, v in enumerate(nodes)}
    adjacency = (lambda: G.A) if isinstance(G, GraphHandle) else (lambda: nx.adjacency_matrix(G, nodelist=nodes))
    missing = [s for s in seed_set if s not in pos]
    if missing:
        raise ValueError("seed nodes %r are not in the graph (after the largest-connected-component cut)" % (missing,))
    key = id(G)
    dg = _mc_graphs.get(key)
    if dg is None or dg[0] is not G:
        with torch.cuda.device(dev):
            dg = (G, DeviceGraph(adjacency()))
        _mc_graphs.clear()
        _mc_graphs[key] = dg
    g = dg[1]
    L = _lib.lib()
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    with torch.cuda.device(dev):
        seeds = torch.tensor([pos[s] for s in seed_set], dtype=torch.int32, device=dev)
        counts = torch.empty((3, int(T), g.n), dtype=torch.float64, device=dev)
        ws_bytes = int(L.gnode_mc_sir_workspace_bytes(g.handle, int(T)))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _lib.check(L.gnode_mc_sir(g.handle, ctypes.c_void_p(seeds.data_ptr()), seeds.numel(), float(beta), float(gamma),
                                  int(sims), int(T), ctypes.c_uint64(seed & (2 ** 64 - 1)), ctypes.c_void_p(counts.data_ptr()),
                                  ctypes.c_void_p(ws.data_ptr()), ws_bytes,
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "gnode_mc_sir")
        out = counts.cpu().numpy()
    return out[0][None], out[1][None], out[2][None]
