"""Helper module imported by the experiment scripts (same public names as the reference's
``ode_nn.py`` helpers that the GN-ODE scripts use; ode_nn_ngraph_sim.py:23).  The legacy
dense/rk4 model that lived in the reference file is dead code there and is not
reproduced.  Implementations live in gn-ode-sir_b200/harness.py."""
import gn_ode_sir_b200  # noqa: F401  (registers the package)
from gn_ode_sir_b200 import harness as _h


def create_graph(n_nodes, graph_label="none"):
    """Returns (G, A, 0) like the reference (the third slot used to be a dense matrix)."""
    G, A = _h.load_graph(graph_label, n_nodes)
    return G, A, 0


def get_sir_t_nodes_torch(x_rk, maxTime, deltaT, count=True):
    return _h.sample_unit_times(x_rk, maxTime, deltaT, count=count)


def get_sir_t_nodes(x_rk, maxTime, deltaT, count=True):
    import numpy as np
    import torch
    return _h.sample_unit_times(torch.as_tensor(np.asarray(x_rk)), maxTime, deltaT, count=count).numpy()


def csv_trials(path_to_csv, columns, list_to_csv):
    _h.append_csv(path_to_csv, columns, list_to_csv)


def save_trial_to_csv(args, best_epoch, val_loss, test_loss, loss_baseline, n_ode_time, rk_time):
    import os
    row = [args.trial, args.model, args.lr, args.epochs, args.sim, args.train_val_test_ratio, len(args.beta),
           len(args.gamma), args.deltaT, args.maxTime, [len(args.I_indices[0]), len(args.I_indices)], args.hidden,
           best_epoch, val_loss, test_loss, loss_baseline, n_ode_time, rk_time]
    cols = ["trial", "model", "lr", "epochs", "MC sim", "train_val_test_ratio", "beta", "gamma", "deltaT", "maxTime",
            "I_indices", "hidden", "best_epoch", "val_loss", "test_loss", "loss_baseline", "n_ode_time", "rk_time"]
    _h.append_csv(args.path_to_save + "/Metrics-trials-" + os.path.relpath(args.dataset, "./real_graphs/"), cols, row)


def sir_torch(G, seed_set, beta, gamma, sims=10000, T=20):
    return _h.monte_carlo_sir(G, seed_set, beta, gamma, sims=sims, T=T)
