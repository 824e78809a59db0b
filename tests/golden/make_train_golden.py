"""Generates tests/golden/train_karate_b8.npz by running the reference's own train() for ONE mini-batch.

Build container only (needs /root/reference). Pins the N1 path (sub-sampled prediction -> L1 loss -> gradients) to
  /root/reference/ode_nn_ngraph_sim.py:208-250   train(): model(x), get_sir_t_nodes_torch x3, nn.L1Loss on [:,1:,:],
                                                  loss.backward() (torchdiffeq adjoint, restated in oracle/ref_harness.py)
  /root/reference/ode_nn.py:249-261               get_sir_t_nodes_torch
on the SHIPPED karate Monte-Carlo labels (multi-graph-1/Experiments-seed2-karate) and trial parameters
(initial-{seed,beta,gamma}.pkl), with the input / label tensors built as main() builds them (:358-397).
train() is called unmodified with an SGD optimiser of learning rate 0, so the weights stay and the gradients of the
batch are left in .grad; its first return value is the batch's loss.

    python tests/golden/make_train_golden.py
"""
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EXP = os.path.join(rh.REFERENCE_ROOT, "multi-graph-1", "Experiments-seed2-karate")
H, MAXTIME, DELTAT, B, SEED = 64, 20, 0.5, 8, 11


def main():
    sim, _ = rh.load_reference("adjoint")
    A = rh.load_reference_graph("karate")
    n = A.shape[0]
    seeds = pickle.load(open(os.path.join(EXP, "initial-seed.pkl"), "rb"))[:B]
    betas = pickle.load(open(os.path.join(EXP, "initial-beta.pkl"), "rb"))[:B]
    gammas = pickle.load(open(os.path.join(EXP, "initial-gamma.pkl"), "rb"))[:B]
    xs, ys = [], []
    for sd, be, ga in zip(seeds, betas, gammas):
        lab = [pickle.load(open(os.path.join(EXP, "karate-%s-%s.pkl" % (c, "-".join(str(i) for i in sd))), "rb"))
               for c in "SIR"]                                           # each [maxTime, n] float64
        I0 = torch.zeros(n, dtype=torch.float)
        I0[list(sd)] = 1
        bg = torch.zeros(n, H, dtype=torch.float)
        bg[:, 0], bg[:, 1] = be, ga
        xs.append(torch.cat(((1 - I0).unsqueeze(1), I0.unsqueeze(1), torch.zeros(n, 1), bg), -1))
        ys.append(torch.transpose(torch.cat([torch.tensor(np.asarray(l)).unsqueeze(-1) for l in lab], -1), 0, 1))
    x, y = torch.stack(xs), torch.stack(ys)                              # [B, n, 3+H] fp32, [B, n, maxTime, 3] fp64
    torch.set_default_dtype(torch.float32)
    torch.manual_seed(SEED)
    of = sim.ODEfunc(A, 0.2, 0.1, H, "cpu")
    model = sim.ODEBlock(MAXTIME, DELTAT, n, [0, 1], H, of, "cpu")
    params = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    loader = [(x, y)]
    loss, val_loss = sim.train(model, opt, torch.nn.L1Loss(), "cpu", loader, loader, MAXTIME, DELTAT, n)
    out = {"x": x.numpy(), "y": y.numpy(), "loss": np.float64(loss), "val_loss": np.float64(val_loss),
           "maxTime": MAXTIME, "deltaT": DELTAT, "H": H,
           "indptr": A.tocsr().indptr.astype(np.int32), "indices": A.tocsr().indices.astype(np.int32),
           "seeds": np.asarray(seeds, dtype=np.int32), "beta": np.asarray(betas, dtype=np.float64),
           "gamma": np.asarray(gammas, dtype=np.float64)}
    for k, v in params.items():
        out["p:" + k] = v
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["g:" + k] = p.grad.detach().numpy().copy()
    path = os.path.join(OUT, "train", "train_karate_b8.npz")
    np.savez_compressed(path, **out)
    print("loss %.12f val %.12f -> %s (%.1f KB)" % (loss, val_loss, path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
