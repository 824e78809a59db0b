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


# ----------------------------------------------------------------------------- graphs
def load_graph(graph_label, n_nodes=50):
    """pickle -> undirected -> largest connected component (reference create_graph)."""
    import networkx as nx
    if graph_label != "none":
        with open(graph_label + ".pkl", "rb") as fh:
            G = pickle.load(fh)
        G = G.to_undirected()
        G = G.subgraph(max(nx.connected_components(G), key=len))
        print("nodes", G.number_of_nodes())
        print("edges", G.number_of_edges())
    else:
        G = nx.fast_gnp_random_graph(n_nodes, 0.2)
    return G, nx.adjacency_matrix(G)


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
    return loss, 3 * (pred.size(1) - 1) * pred.size(0)


def run_epoch(model, optimizer, criterion, device, train_batches, val_batches, maxTime, deltaT):
    """One epoch of Adam steps followed by a validation sweep; returns item-weighted mean losses."""
    model.train()
    # the item-weighted loss sums stay on the device: one host read per epoch instead of one per batch (the
    # reference's loss.item() per batch, ode_nn_ngraph_sim.py:247, drains the stream between optimiser steps)
    tot, items, fwd_time = torch.zeros((), dtype=torch.float64, device=device), 0, 0.0
    for x, y in train_batches:
        x, y = x.to(device), y.to(device)
        optimizer.zero_grad()
        t0 = time.time()
        loss, n = l1_on_rollout(model, criterion, x, y, maxTime, deltaT)
        fwd_time += time.time() - t0
        loss.backward()
        optimizer.step()
        tot += loss.detach().double() * n
        items += n
    model.eval()
    vtot, vitems = torch.zeros((), dtype=torch.float64, device=device), 0
    with torch.no_grad():
        for x, y in val_batches:
            loss, n = l1_on_rollout(model, criterion, x.to(device), y.to(device), maxTime, deltaT)
            vtot += loss.double() * n
            vitems += n
    print("Time: ", fwd_time)
    return float(tot.item()) / max(items, 1), float(vtot.item()) / max(vitems, 1)


def evaluate(model, criterion, device, batches, maxTime, deltaT):
    model.eval()
    tot, items, per_batch = 0.0, 0, []
    with torch.no_grad():
        for x, y in batches:
            loss, n = l1_on_rollout(model, criterion, x.to(device), y.to(device), maxTime, deltaT)
            tot += loss.item() * n
            items += n
            per_batch.append(loss.item())
    return tot / max(items, 1), per_batch


def fit(model, device, lr, epochs, train_batches, val_batches, test_batches, maxTime, deltaT):
    """Adam + L1 with test evaluation whenever the validation loss improves."""
    criterion = nn.L1Loss()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    best = dict(val=np.inf, epoch=-1, test=float("nan"), test_all=[], test_time=0.0)
    print("training...")
    for epoch in range(epochs):
        loss, val_loss = run_epoch(model, optimizer, criterion, device, train_batches, val_batches, maxTime, deltaT)
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
def monte_carlo_sir(G, seed_set, beta, gamma, sims=10000, T=20, device=None, chunk=2048, seed=None):
    """Discrete-time Monte-Carlo SIR ground truth (the process of the reference's sir_torch,
    ode_nn.py:30-88), with all simulations of a chunk advanced together instead of one by one:
    per step every (infected u, susceptible v) edge transmits with probability beta (v is infected
    if any of its infected neighbours succeeds) and every infected node recovers with probability
    gamma. Returns per-time-step node COUNTS over the simulations: S, I, R each [1, T, n]."""
    import networkx as nx
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    nodes = list(G.nodes())
    pos = {v: i for i, v in enumerate(nodes)}
    n = len(nodes)
    src = torch.tensor([pos[u] for u, v in G.edges()] + [pos[v] for u, v in G.edges()], device=device)
    dst = torch.tensor([pos[v] for u, v in G.edges()] + [pos[u] for u, v in G.edges()], device=device)
    gen = torch.Generator(device=device)
    if seed is not None:
        gen.manual_seed(seed)
    else:
        gen.seed()
    counts = torch.zeros((3, T, n), dtype=torch.float64, device=device)
    seeds = torch.tensor([pos.get(s, s) for s in seed_set], device=device)
    for c0 in range(0, sims, chunk):
        b = min(chunk, sims - c0)
        I = torch.zeros((b, n), dtype=torch.bool, device=device)
        I[:, seeds] = True
        S, R = ~I, torch.zeros_like(I)
        counts[0, 0] += S.sum(0)
        counts[1, 0] += I.sum(0)
        for t in range(1, T):
            live = I[:, src] & S[:, dst]                              # [b, 2E] edges that can transmit
            hit = live & (torch.rand(live.shape, device=device, generator=gen) < beta)
            newly = torch.zeros((b, n), dtype=torch.int32, device=device).index_add_(1, dst, hit.int()) > 0
            rec = I & (torch.rand(I.shape, device=device, generator=gen) < gamma)
            R |= rec
            I = (I | newly) & ~rec
            S &= ~newly
            counts[0, t] += S.sum(0)
            counts[1, t] += I.sum(0)
            counts[2, t] += R.sum(0)
    out = counts.cpu().numpy()
    return out[0][None], out[1][None], out[2][None]
