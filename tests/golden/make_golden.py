"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CLASSES.

Runs only in the build container (needs /root/reference; see
oracle/ref_harness.py for what is stubbed/restated).  The fixtures pin the
oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_parity_gpu.py)
to outputs of ``ODEfunc``/``ODEBlock`` from
  /root/reference/ode_nn_ngraph_sim.py:37-188   (sim variant)
  /root/reference/ode_nn_ngraphs.py:37-152      (multi-graph variant)
on seeded weights and inputs.

    python tests/golden/make_golden.py          # rewrites every fixture
    python tests/golden/make_golden.py NAME...  # only the named cases

Each npz holds: CSR of every graph used (indptr/indices int32), the instance ->
graph map, x [M,3+H] fp32, the full state_dict, the reference outputs
probs32 [T,M,3] (fp32 run) and probs64 (float64 run of the same weights), the
seed of the probe-loss weight w [T,M,3] and the parameter gradients of L = sum(w*probs) for
grad modes "adjoint" (torchdiffeq semantics) and "discrete" in fp32 and fp64.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402
from oracle import gnode_oracle as orc        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H, MAXTIME, DELTAT = 64, 20, 0.5

# name, variant, graphs, instance->graph list, seed, with_grads, time stride for stored probs
CASES = [
    ("sim_karate_b1", "sim", ["karate"], [0], 0, True, 1),
    ("sim_karate_b8", "sim", ["karate"], [0] * 8, 1, True, 1),
    ("sim_dolphins_b4", "sim", ["dolphins"], [0] * 4, 2, True, 1),
    ("sim_fbfood_b2", "sim", ["fb-food"], [0] * 2, 3, True, 1),
    ("sim_fbsocial_b1", "sim", ["fb-social"], [0], 4, False, 3),
    ("ng_mixed_b5", "ngraphs", ["karate", "dolphins", "fb-food"], [0, 1, 0, 2, 1], 5, True, 1),
    # power-law graphs of BASELINE configs[2]/[3]: wiki-vote (max degree 1065: hub rows, CSR-slice overflow),
    # openflights, and a multi-graph training batch over the five training graphs (ode_nn_ngraphs.py:311-312)
    ("sim_openflights_b2", "sim", ["openflights"], [0] * 2, 7, True, 4),
    ("sim_wikivote_b2", "sim", ["wiki-vote"], [0] * 2, 6, True, 4),
    ("ng_train5_b8", "ngraphs", ["dolphins", "fb-food", "fb-social", "openflights", "wiki-vote"],
     [0, 1, 2, 3, 4, 0, 1, 2], 8, True, 4),
]


def build_models(variant, adjs, inst_graph, seed, grad_mode):
    sim, ng = rh.load_reference(grad_mode)
    torch.set_default_dtype(torch.float32)
    torch.manual_seed(seed)
    if variant == "sim":
        A = adjs[0]
        of = sim.ODEfunc(A, 0.2, 0.1, H, "cpu")
        blk = sim.ODEBlock(MAXTIME, DELTAT, A.shape[0], [0, 1], H, of, "cpu")
    else:
        of = ng.ODEfunc(adjs, H, "cpu")
        blk = ng.ODEBlock(MAXTIME, DELTAT, H, of, "cpu")
    return blk


def make_x(variant, adjs, inst_graph, seed):
    blocks = []
    for i, g in enumerate(inst_graph):
        marker = float(g + 1) if variant == "ngraphs" else 0.0
        blocks.append(orc.synthetic_trial(adjs[g].shape[0], H, 100 * seed + i, graph_marker=marker))
    return blocks


def run(blk, variant, blocks, dtype):
    if variant == "sim":
        x = torch.stack(blocks).to(dtype)          # [B, N, 3+H] as the DataLoader yields it
    else:
        x = torch.cat(blocks).to(dtype)            # [sum N, 3+H] (ode_nn_ngraphs.py:179-196)
    S, I, R = blk(x)
    return torch.cat((S, I, R), -1)                # [T, M, 3]


def grads_of(blk):
    return {k: p.grad.detach().clone() for k, p in blk.named_parameters() if p.grad is not None}


def main():
    only = set(sys.argv[1:])
    for name, variant, gnames, inst_graph, seed, with_grads, tstride in CASES:
        if only and name not in only:
            continue
        adjs = [rh.load_reference_graph(g) for g in gnames]
        blocks = make_x(variant, adjs, inst_graph, seed)
        out = {"variant": variant, "graph_names": np.array(gnames), "inst_graph": np.array(inst_graph, np.int32),
               "maxTime": MAXTIME, "deltaT": DELTAT, "H": H, "tstride": tstride,
               "x": torch.cat(blocks).numpy()}
        for gi, A in enumerate(adjs):
            A = A.tocsr()
            A.sort_indices()
            out["g%d_indptr" % gi] = A.indptr.astype(np.int32)
            out["g%d_indices" % gi] = A.indices.astype(np.int32)
        M = sum(adjs[g].shape[0] for g in inst_graph)
        T = len(np.arange(0, MAXTIME, DELTAT))
        w = torch.randn(T, M, 3, dtype=torch.float32, generator=torch.Generator().manual_seed(1234 + seed))
        for mode in ("adjoint", "discrete"):
            blk = build_models(variant, adjs, inst_graph, seed, mode)
            if mode == "adjoint":
                for k, v in blk.state_dict().items():
                    out["p:" + k] = v.detach().numpy().copy()
            probs = run(blk, variant, blocks, torch.float32)
            if mode == "adjoint":
                out["probs32"] = probs.detach()[::tstride].numpy().copy()
            if with_grads:
                (probs * w).sum().backward()
                for k, g in grads_of(blk).items():
                    out["g32:%s:%s" % (mode, k)] = g.numpy().copy()
            # float64 run of the same weights (default dtype must be switched:
            # torch.zeros(I.size()) at ode_nn_ngraph_sim.py:73 follows it)
            torch.set_default_dtype(torch.float64)
            blk64 = copy.deepcopy(blk).double()
            blk64.zero_grad()
            probs64 = run(blk64, variant, blocks, torch.float64)
            if mode == "adjoint":
                p64 = probs64.detach()[::tstride]
                out["probs64"] = (p64 if with_grads else p64.float()).numpy().copy()
            if with_grads:
                (probs64 * w.double()).sum().backward()
                for k, g in grads_of(blk64).items():
                    out["g64:%s:%s" % (mode, k)] = g.numpy().copy()
            torch.set_default_dtype(torch.float32)
        out["w_seed"] = 1234 + seed        # w = randn(T, M, 3, generator seeded with w_seed)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-18s M=%-6d %8.1f KB" % (name, M, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
