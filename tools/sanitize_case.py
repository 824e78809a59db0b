"""Small forward + backward case for compute-sanitizer (single-instance tiles, straddling tiles, a partial last tile,
a hub longer than the 12-row gather round)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth
from oracle import gnode_oracle as orc
dev = torch.device("cuda:0")
A = synth.barabasi_albert_csr(700, 3, 0); N = A.shape[0]; B = 3
torch.manual_seed(0)
of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev); blk = gn.ode_sim.ODEBlock(4, 0.5, N, [0, 1], 64, of, dev).to(dev)
x = torch.stack([orc.synthetic_trial(N, 64, b) for b in range(B)]).to(dev)
with torch.no_grad():
    S, I, R = blk(x)
S, I, R = blk(x)
(S.sum() + 2 * I.sum() - R.sum()).backward()
torch.cuda.synchronize()
print("ok", float(S.sum()), float(blk.odefunc.linear.weight.grad.abs().sum()))
