"""One training step (forward with trajectory + backward) of the drop-in ODEBlock on the epinions-size graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
A = synth.epinions_standin(0); N = A.shape[0]
torch.manual_seed(0)
of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev); blk = gn.ode_sim.ODEBlock(20, 0.5, N, [0, 1], 64, of, dev).to(dev)
x = torch.stack([synth.synthetic_trial(N, 64, b) for b in range(B)]).to(dev)
w = torch.randn(40, B * N, 3, device=dev)
for _ in range(2):
    blk.zero_grad()
    S, I, R = blk(x)
    (torch.cat((S, I, R), -1) * w).sum().backward()
torch.cuda.synchronize()
print("ok")
