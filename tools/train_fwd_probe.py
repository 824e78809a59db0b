"""Per-repetition timing of the training forward (trajectory + auxiliary storage) on the epinions-size graph: CUDA events
around the call and host wall clock, to tell device time from host-side (allocator) stalls."""
import os, sys, time
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
for rep in range(7):
    blk.zero_grad()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record(); S, I, R = blk(x); e1.record()
    t1 = time.perf_counter()
    (torch.cat((S, I, R), -1) * w).sum().backward(); e2.record()
    torch.cuda.synchronize()
    print("rep %d  fwd: events %.2f ms, host returned after %.2f ms | bwd events %.2f ms | reserved %.1f GB" % (
        rep, e0.elapsed_time(e1), (t1 - t0) * 1e3, e1.elapsed_time(e2), torch.cuda.memory_reserved() / 2**30), flush=True)
    del S, I, R
