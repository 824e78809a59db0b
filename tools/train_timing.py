"""Forward(+trajectory) and backward timing of the drop-in ODEBlock (CUDA events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth
dev = torch.device("cuda:0")
for name, n, m, B in (("karate-size", 34, 2, 1), ("fb-social-size", 1893, 7, 8), ("fb-social-size", 1893, 7, 64), ("epinions-size", 75879, 5, 4), ("epinions-size", 75879, 5, 8)):
    A = synth.barabasi_albert_csr(n, m, 0); N = A.shape[0]
    torch.manual_seed(0)
    of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev); blk = gn.ode_sim.ODEBlock(20, 0.5, N, [0, 1], 64, of, dev).to(dev)
    x = torch.stack([synth.synthetic_trial(N, 64, b) for b in range(B)]).to(dev)
    w = torch.randn(40, B * N, 3, device=dev)
    def step():
        blk.zero_grad()
        S, I, R = blk(x)
        (torch.cat((S, I, R), -1) * w).sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 5
    tf = tb = 0.0
    for _ in range(reps):
        blk.zero_grad()
        e0.record(); S, I, R = blk(x); e1.record()
        (torch.cat((S, I, R), -1) * w).sum().backward(); e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    with torch.no_grad():
        e0.record(); blk(x); e1.record(); torch.cuda.synchronize()
    ns = B * N * 39
    print("%-15s N=%-6d B=%-3d fwd(train) %8.3f ms  bwd %8.3f ms  fwd(infer) %8.3f ms | fwd+bwd %.3e node-steps/s" % (
        name, N, B, tf / reps, tb / reps, e0.elapsed_time(e1), ns / ((tf + tb) / reps * 1e-3)))
