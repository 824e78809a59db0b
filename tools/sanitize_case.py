"""Small forward + backward case for compute-sanitizer (single-instance tiles, straddling tiles, a partial last tile,
a hub longer than the 12-row gather round, and an isolated 700-degree hub that takes the in-order relay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth
dev = torch.device("cuda:0")
import numpy as np, scipy.sparse
A = scipy.sparse.lil_matrix(synth.barabasi_albert_csr(1500, 3, 0)); N = A.shape[0]; B = 3
nb = np.random.RandomState(0).choice(N, 700, replace=False); nb = nb[nb != 900]
A[900, nb] = 1; A[nb, 900] = 1                       # isolated hub (degree > 512) in tile 7: relay path
A = scipy.sparse.csr_matrix(A); A.data[:] = 1; A.sort_indices()
torch.manual_seed(0)
of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev); blk = gn.ode_sim.ODEBlock(4, 0.5, N, [0, 1], 64, of, dev).to(dev)
x = torch.stack([synth.synthetic_trial(N, 64, b) for b in range(B)]).to(dev)
with torch.no_grad():
    S, I, R = blk(x)
S, I, R = blk(x)
(S.sum() + 2 * I.sum() - R.sum()).backward()
torch.cuda.synchronize()
print("ok", float(S.sum()), float(blk.odefunc.linear.weight.grad.abs().sum()))
