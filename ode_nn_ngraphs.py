#!/usr/bin/env python3
"""GN-ODE trained on several graphs and evaluated on an unseen one -- the script
``monitorer-ngraphs.py`` spawns (command line of the reference's ode_nn_ngraphs.py:294-306),
running on the B200-native multi-graph ``ODEfunc`` / ``ODEBlock``."""
import argparse
import os
import pickle

import numpy as np
import torch

import gn_ode_sir_b200  # noqa: F401
from gn_ode_sir_b200 import harness
from gn_ode_sir_b200.ode_ngraphs import ODEBlock, ODEfunc
from ode_nn import csv_trials

INSTANCES_PER_GRAPH = [36, 36, 36, 36, 36, 120]     # five training graphs, one held-out graph (ode_nn_ngraphs.py:311)


def parse_args(argv=None):
    from gn_ode_sir_b200.rollout import check_hidden
    p = argparse.ArgumentParser(description="Neural ODE")
    p.add_argument("--lr", type=float, default=1e-2)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--sim", type=int, default=1000)
    p.add_argument("--deltaT", type=float, default=0.5)
    p.add_argument("--maxTime", type=int, default=20)
    p.add_argument("--hidden", type=int, default=32)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--path_to_save", default="./plots")
    p.add_argument("--trial", type=int, default=32)
    p.add_argument("--dataset", default="none")
    p.add_argument("--train_val_test_ratio", nargs=3, type=float, default=[5e-1, 1e-1, 4e-1])
    p.add_argument("--model", default="ode_nn", type=str)
    args = p.parse_args(argv)
    check_hidden(args.hidden, need_marker=True)          # before any label is loaded or generated
    return args


def create_graphs(graph_label="none"):
    if graph_label == "none":
        return []
    return [harness.load_graph(graph_label[:14] + name)[1] for name in graph_label[14:].split("+")]


def label_dir(graph, path_to_save):
    if graph == "wiki-vote":
        return "./multi-graph-1/Experiments-gpu-seed2-" + graph
    if graph == "enron":
        return "./multi-graph-1/Experiments2-seed2-" + graph
    tag = path_to_save.split("/")[-1].split("-")
    return "./multi-graph-1/" + tag[0] + "-" + tag[1] + "-" + graph


def load_SIR_labels(graph, directory, I_indices, sim):
    key = "-".join(str(i) for i in I_indices)
    out = [pickle.load(open(directory + "/" + graph + "-" + c + "-" + key + ".pkl", "rb")) for c in "SIR"]
    return [v / sim for v in out] if graph == "wiki-vote" else out       # wiki-vote labels are raw counts


def batches(items, batch_size, shuffle, rng=None):
    """Global mini-batches of instances (x_i, y_i, graph_id_i), shuffled once like the reference's loader
    (ode_nn_ngraphs.py:179-196, :360). The harness concatenates each batch -- or, under torchrun, this rank's share of
    it -- along the node axis (ragged batch) and names the instances' graphs, so the marker column is never read back
    from the device."""
    order = (rng or np.random).permutation(len(items)) if shuffle else np.arange(len(items))
    return [[items[j] for j in order[i:i + batch_size]] for i in range(0, len(items), batch_size)]


def main(argv=None):
    args = parse_args(argv)
    A_list = create_graphs(args.dataset)
    print(len(A_list))
    train, val, test = [], [], []
    val_len = INSTANCES_PER_GRAPH[-1] // 2
    for gi, graph in enumerate(args.dataset[14:].split("+")):
        d = label_dir(graph, args.path_to_save)
        take = INSTANCES_PER_GRAPH[min(gi, len(INSTANCES_PER_GRAPH) - 1)]
        seeds = pickle.load(open(d + "/initial-seed.pkl", "rb"))[:take]
        betas = pickle.load(open(d + "/initial-beta.pkl", "rb"))[:take]
        gammas = pickle.load(open(d + "/initial-gamma.pkl", "rb"))[:take]
        n_nodes = A_list[gi].shape[0]
        for i, s in enumerate(seeds):
            x = torch.zeros(n_nodes, 3 + args.hidden, dtype=torch.float)
            x[list(s), 1] = 1.0
            x[:, 0] = 1.0 - x[:, 1]
            x[:, 3], x[:, 4], x[0, 5] = betas[i], gammas[i], gi + 1          # marker: first row names the graph
            y = torch.tensor(np.stack(load_SIR_labels(graph, d, s, args.sim), axis=-1)).transpose(0, 1)
            (train if gi < len(INSTANCES_PER_GRAPH) - 1 else (val if len(val) < val_len else test)).append((x, y, gi))
    # one process per GPU under torchrun: every global mini-batch is split across the ranks by node count, the
    # parameter gradient is summed with one NCCL all-reduce per optimiser step (harness.run_epoch)
    rank, world, device = harness.init_distributed()
    torch.set_default_dtype(torch.float32)
    if rank == 0:
        print(device)
    if device.type != "cuda":
        raise SystemExit("ode_nn_ngraphs.py: the B200 GN-ODE rollout needs a CUDA device (there is no CPU path)")
    odefunc = ODEfunc(A_list, args.hidden, device)
    model = ODEBlock(args.maxTime, args.deltaT, args.hidden, odefunc, device).to(device)
    rng = np.random.RandomState(harness.shared_seed())                    # the same shuffle on every rank
    best = harness.fit(model, device, args.lr, args.epochs, batches(train, args.batch_size, True, rng),
                       batches(val, args.batch_size, False), batches(test, args.batch_size, False),
                       args.maxTime, args.deltaT)
    if rank != 0:
        return
    csv_trials(args.path_to_save + "/Metrics-trials-" + os.path.relpath(args.dataset, "./real_graphs/"),
               ["trial", "model", "lr", "epochs", "deltaT", "maxTime", "hidden", "best_epoch", "val_loss",
                "test_loss", "n_ode_time"],
               [args.trial, args.model, args.lr, args.epochs, args.deltaT, args.maxTime, args.hidden, best["epoch"],
                best["val"], best["test"], best["test_time"]])


if __name__ == "__main__":
    main()
