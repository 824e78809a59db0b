#!/usr/bin/env python3
"""GN-ODE on one graph with many (beta, gamma, seed-set) trials -- the script ``monitorer-sim.py``
spawns (same file name and command line as the reference's ode_nn_ngraph_sim.py:326-343), running
on the B200-native ``ODEfunc`` / ``ODEBlock``."""
import argparse
import os
import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

import gn_ode_sir_b200  # noqa: F401
from gn_ode_sir_b200 import harness
from gn_ode_sir_b200.ode_sim import ODEBlock, ODEfunc
from ode_nn import create_graph, csv_trials, save_trial_to_csv, sir_torch


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Neural ODE")
    p.add_argument("--lr", type=float, default=1e-2)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--sim", type=int, default=1000)
    p.add_argument("--beta", type=float, nargs="+", default=[0.2])
    p.add_argument("--gamma", type=float, nargs="+", default=[0.1])
    p.add_argument("--deltaT", type=float, default=0.5)
    p.add_argument("--maxTime", type=int, default=20)
    p.add_argument("--I_indices", nargs="+", default=[12])
    p.add_argument("--hidden", type=int, default=32)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--path_to_save", default="./plots")
    p.add_argument("--trial", type=int, default=32)
    p.add_argument("--dataset", default="none")
    p.add_argument("--train_val_test_ratio", nargs=3, type=float, default=[5e-1, 1e-1, 4e-1])
    p.add_argument("--model", default="ode_nn", type=str)
    p.add_argument("--out_of_dist", default=False, action="store_true")
    args = p.parse_args(argv)
    from gn_ode_sir_b200.rollout import check_hidden
    check_hidden(args.hidden)                            # before any label is loaded or generated (Monte-Carlo runs are long)
    # seed sets arrive as strings "[a, b]" (monitorer-sim.py:64-65)
    args.I_indices = [[int(v) for v in str(s).strip("[]").split(",") if v.strip()] for s in args.I_indices]
    return args


def load_SIR_labels(dataset, path_to_save, G, I_indices, beta, gamma, sim, maxTime):
    """Label cache keyed by the seed set (file names of the reference, ode_nn_ngraph_sim.py:190-206)."""
    stem = path_to_save + "/" + dataset[14:]
    key = "-".join(str(i) for i in I_indices)
    paths = [stem + "-" + c + "-" + key + ".pkl" for c in "SIR"]
    if all(os.path.exists(p) for p in paths):
        labels = [pickle.load(open(p, "rb")) for p in paths]
        print("ok")
    else:
        S, I, R = sir_torch(G, I_indices, beta, gamma, sim, maxTime)
        labels = [S[0] / sim, I[0] / sim, R[0] / sim]
        for p, v in zip(paths, labels):
            pickle.dump(v, open(p, "wb"))
    return labels


def build_inputs(args, n_nodes, labels):
    """x_i = [S0 | I0 | R0 | beta gamma 0...] of shape [N, 3+H]; y_i = [N, maxTime, 3]."""
    xs, ys = [], []
    for i, seeds in enumerate(args.I_indices):
        x = torch.zeros(n_nodes, 3 + args.hidden, dtype=torch.float)
        x[seeds, 1] = 1.0
        x[:, 0] = 1.0 - x[:, 1]
        x[:, 3], x[:, 4] = args.beta[i], args.gamma[i]
        xs.append(x)
        ys.append(torch.tensor(np.stack(labels[i], axis=-1)).transpose(0, 1))
    return xs, ys


def split_indices(args, n):
    if args.out_of_dist:
        d = pickle.load(open(args.path_to_save + "/out-of-dist-gamma.pkl", "rb"))
        train, val = list(d["train"]), list(d["val"])
        test = [i for i in range(n) if i not in set(train) | set(val)]
        return train, val, test, d["test"]
    a = int(args.train_val_test_ratio[0] * n)
    b = int((args.train_val_test_ratio[0] + args.train_val_test_ratio[1]) * n)
    return list(range(a)), list(range(a, b)), list(range(b, n)), None


def main(argv=None):
    args = parse_args(argv)
    G, A, _ = create_graph(50, args.dataset)
    n_nodes = A.shape[0]
    print(n_nodes)
    if not os.path.exists(args.path_to_save + "/initial-seed.pkl"):
        os.makedirs(args.path_to_save, exist_ok=True)
        pickle.dump(args.I_indices, open(args.path_to_save + "/initial-seed.pkl", "wb"))
        pickle.dump(args.beta, open(args.path_to_save + "/initial-beta.pkl", "wb"))
        pickle.dump(args.gamma, open(args.path_to_save + "/initial-gamma.pkl", "wb"))
    labels = [load_SIR_labels(args.dataset, args.path_to_save, G, s, args.beta[i], args.gamma[i], args.sim, args.maxTime)
              for i, s in enumerate(args.I_indices)]
    xs, ys = build_inputs(args, n_nodes, labels)
    tr, va, te, idx_test = split_indices(args, len(xs))

    # one process per GPU under torchrun: the trials of every global mini-batch are split across the ranks, one NCCL
    # all-reduce of the parameter gradient per optimiser step (harness.run_epoch)
    rank, world, device = harness.init_distributed()
    shuffle_seed = harness.shared_seed()                                  # the same shuffle on every rank

    def loader(idx, batch_size, shuffle):
        ds = TensorDataset(torch.stack([xs[i] for i in idx]), torch.stack([ys[i] for i in idx]))
        return DataLoader(ds, batch_size=batch_size, shuffle=shuffle,
                          generator=torch.Generator().manual_seed(shuffle_seed) if shuffle else None)

    torch.set_default_dtype(torch.float32)
    if rank == 0:
        print(device)
    if device.type != "cuda":
        raise SystemExit("ode_nn_ngraph_sim.py: the B200 GN-ODE rollout needs a CUDA device (there is no CPU path)")
    odefunc = ODEfunc(A, args.beta[0], args.gamma[0], args.hidden, device)
    model = ODEBlock(args.maxTime, args.deltaT, n_nodes, args.I_indices[0], args.hidden, odefunc, device).to(device)
    best = harness.fit(model, device, args.lr, args.epochs, loader(tr, args.batch_size, True),
                       loader(va, args.batch_size, False), loader(te, 1, False), args.maxTime, args.deltaT)
    if rank != 0:
        return
    if not args.out_of_dist:
        save_trial_to_csv(args, best["epoch"], best["val"], best["test"], 0, best["test_time"], 0)
    else:
        rel = os.path.relpath(args.dataset, "./real_graphs/")
        csv_trials(args.path_to_save + "/Out-of-dist-gamma-" + rel, [str(i) for i in idx_test], best["test_all"])
        csv_trials(args.path_to_save + "/Out-of-dist-gamma-trials-" + rel,
                   ["trial", "model", "lr", "epochs", "deltaT", "maxTime", "hidden", "best_epoch", "val_loss",
                    "test_loss", "n_ode_time"],
                   [args.trial, args.model, args.lr, args.epochs, args.deltaT, args.maxTime, args.hidden, best["epoch"],
                    best["val"], best["test"], best["test_time"]])


if __name__ == "__main__":
    main()
