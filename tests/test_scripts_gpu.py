"""The experiment scripts the reference's monitorers spawn, run end to end on the GPU.

monitorer-sim.py builds `python3 ./ode_nn_ngraph_sim.py --lr .. --I_indices "[a, b]" ..` from module
constants (monitorer-sim.py:35-103,209-229) and monitorer-ngraphs.py does the same for
./ode_nn_ngraphs.py; the tests rebuild those command lines (the reference tree does not travel to the
GPU box) on generated fixtures: the karate graph from networkx and Monte-Carlo labels from the
harness simulator."""
import os
import pickle
import subprocess
import sys

import networkx as nx
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def monitorer_sim_argv(script, seeds, betas, gammas, path_to_save, dataset, epochs):
    """createArgs of monitorer-sim.py for many_graph_instances=True."""
    argv = ["python3", script, "--lr", "0.001", "--epochs", str(epochs), "--hidden", "64", "--batch_size", "1"]
    argv += ["--I_indices"] + [str(list(map(int, s))) for s in seeds]
    argv += ["--beta"] + [str(b) for b in betas] + ["--gamma"] + [str(g) for g in gammas]
    argv += ["--deltaT", "0.5", "--maxTime", "20", "--sim", "200", "--dataset", dataset, "--trial", "1",
             "--path_to_save", path_to_save, "--train_val_test_ratio", "0.6", "0.2", "0.2", "--model", "ode_nn"]
    return argv


def test_config1_karate_through_sim_script(tmp_path):
    os.makedirs(tmp_path / "real_graphs")
    pickle.dump(nx.karate_club_graph(), open(tmp_path / "real_graphs" / "karate.pkl", "wb"))
    save = "./multi-graph-1/Experiments-seed2-karate"
    rng = np.random.RandomState(0)
    seeds = [list(rng.choice(34, 2, replace=False)) for _ in range(10)]
    betas, gammas = list(rng.uniform(0.1, 0.5, 10)), list(rng.uniform(0.1, 0.5, 10))
    argv = monitorer_sim_argv(os.path.join(ROOT, "ode_nn_ngraph_sim.py"), seeds, betas, gammas, save,
                              "./real_graphs/karate", epochs=4)
    res = subprocess.run(argv, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    losses = [float(l.split("Train Loss:")[1].split(",")[0]) for l in res.stdout.splitlines() if "Train Loss" in l]
    assert len(losses) == 4 and losses[-1] < losses[0], losses
    d = tmp_path / "multi-graph-1" / "Experiments-seed2-karate"
    assert (d / "initial-seed.pkl").exists() and (d / "Metrics-trials-karate").exists()
    assert pickle.load(open(d / "initial-seed.pkl", "rb")) == [list(map(int, s)) for s in seeds]
    key = "-".join(str(int(i)) for i in seeds[0])
    lab = pickle.load(open(d / ("karate-I-" + key + ".pkl"), "rb"))
    assert lab.shape == (20, 34) and 0.0 <= lab.min() and lab.max() <= 1.0
    # second run hits the label cache ("ok" per trial, ode_nn_ngraph_sim.py:195)
    res2 = subprocess.run(argv[:5] + ["1"] + argv[6:], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert res2.returncode == 0 and res2.stdout.count("ok\n") == 10, res2.stdout[-2000:]


def test_multigraph_script(tmp_path):
    sys.path.insert(0, ROOT)
    import gn_ode_sir_b200  # noqa: F401
    from gn_ode_sir_b200 import harness
    os.makedirs(tmp_path / "real_graphs")
    names = ["ga", "gb", "gc", "gd", "ge", "gf"]
    rng = np.random.RandomState(1)
    for k, name in enumerate(names):
        G = nx.connected_watts_strogatz_graph(20 + 7 * k, 4, 0.3, seed=k)
        pickle.dump(G, open(tmp_path / "real_graphs" / (name + ".pkl"), "wb"))
        d = tmp_path / "multi-graph-1" / ("Experiments-seed2-" + name)
        os.makedirs(d)
        n_inst = 120 if k == 5 else 36
        n = G.number_of_nodes()
        seeds = [[int(v) for v in rng.choice(n, 2, replace=False)] for _ in range(n_inst)]
        betas, gammas = [float(v) for v in rng.uniform(0.1, 0.5, n_inst)], [float(v) for v in rng.uniform(0.1, 0.5, n_inst)]
        for nm, v in (("seed", seeds), ("beta", betas), ("gamma", gammas)):
            pickle.dump(v, open(d / ("initial-%s.pkl" % nm), "wb"))
        for s, b, g in zip(seeds, betas, gammas):
            S, I, R = harness.monte_carlo_sir(G, s, b, g, sims=64, T=20, seed=7)
            for c, arr in zip("SIR", (S, I, R)):
                pickle.dump(arr[0] / 64.0, open(d / ("%s-%s-%s.pkl" % (name, c, "-".join(map(str, s)))), "wb"))
    argv = ["python3", os.path.join(ROOT, "ode_nn_ngraphs.py"), "--lr", "0.001", "--epochs", "2", "--hidden", "64",
            "--batch_size", "8", "--deltaT", "0.5", "--maxTime", "20", "--sim", "64",
            "--dataset", "./real_graphs/" + "+".join(names), "--trial", "1",
            "--path_to_save", "./multi-graph-1/Experiments-seed2-multi", "--train_val_test_ratio", "0.6", "0.2", "0.2",
            "--model", "ode_nn"]
    os.makedirs(tmp_path / "multi-graph-1" / "Experiments-seed2-multi")
    res = subprocess.run(argv, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("Train Loss") == 2
    assert (tmp_path / "multi-graph-1" / "Experiments-seed2-multi" / ("Metrics-trials-" + "+".join(names))).exists()
