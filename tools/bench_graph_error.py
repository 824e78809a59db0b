"""Accuracy of the step-kernel structures on the bench graph (epinions stand-in, BA N=75,879): 8 trials rolled out by every
kernel given on the command line, against the CPU oracle (the reference's fp32 arithmetic) and against its float64 run.

    python tools/bench_graph_error.py [kernel ...]      (0 = generic fp32 FFMA kernel; a -DGNODE_ABLATIONS build also 10..13)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import _lib, synth
from oracle import gnode_oracle as orc          # checker only
from test_parity_gpu import DEV, dev_params

kernels = [int(k) for k in sys.argv[1:]] or [0, 3, 5]
L = _lib.lib()
L.gnode_set_persistent(0)
A = synth.epinions_standin(0)
N, B = A.shape[0], 8
params = orc.default_params(64, seed=0)
xs = [orc.synthetic_trial(N, 64, b) for b in range(B)]
t = orc.time_grid(20, 0.5)
torch.set_num_threads(os.cpu_count())
coo = orc.batch_coo([A], [0] * B)
want32 = orc.forward(torch.cat(xs), params, coo, t)
want64 = orc.forward(torch.cat(xs).double(), {k: v.double() for k, v in params.items()}, coo, t)
print("reference fp32 vs float64: %.3e" % (want32.double() - want64).abs().max().item())
graph = gn.DeviceGraph(A)
batch = gn.DeviceBatch([graph] * B)
dt = gn.rollout.dt_array(t)
x = torch.cat(xs).to(DEV)
for k in kernels:
    _lib.check(L.gnode_set_variant(0 if k == 0 else 3), "variant")
    _lib.check(L.gnode_set_step_kernel(k), "kernel")
    with torch.no_grad():
        p = gn.rollout.rollout(x, batch, dt, dev_params(params)).cpu()
    e32 = (p - want32).abs().view(p.shape[0], B, N, 3).amax(dim=(0, 2, 3))
    e64 = (p.double() - want64).abs().max().item()
    print("kernel %2d: max|cuda - reference fp32| = %.3e (per trial %s)   vs float64 %.3e"
          % (k, e32.max().item(), " ".join("%.1e" % v for v in e32.tolist()), e64))
