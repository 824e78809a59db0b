"""Where do the step kernels differ on the bench graph? 3 trials of the epinions stand-in (the inputs of
tests/test_edge_cases_gpu.py::test_bench_graph_kernel_structures_agree): stream (5) / dual (3) / generic fp32 FFMA (0)
against each other and, for the worst trial, against the CPU oracle in float32 and float64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import _lib, synth
from oracle import gnode_oracle as orc        # diagnostic tool: the oracle is the checker here

DEV = "cuda:0"
L = _lib.lib()
A = synth.epinions_standin(0)
N, B = A.shape[0], 3
deg = np.diff(A.indptr)
params = orc.default_params(64, seed=0)
ORDER = ("odefunc.linear.weight", "odefunc.linear.bias", "linearS1.weight", "linearS1.bias", "linear3.weight",
         "linear3.bias", "linearS2.weight", "linearS2.bias")
ps = [params[k].to(DEV) for k in ORDER]
xs = [orc.synthetic_trial(N, 64, 500 + b) for b in range(B)]
x = torch.cat(xs).to(DEV)
batch = gn.DeviceBatch([gn.DeviceGraph(A)] * B)
t = orc.time_grid(20, 0.5)
dt = gn.rollout.dt_array(t)
out = {}
for name, kern, var in (("stream", 5, 3), ("dual", 3, 3), ("ffma", 0, 0)):
    L.gnode_set_step_kernel(kern); L.gnode_set_variant(var)
    with torch.no_grad():
        out[name] = gn.rollout.rollout(x, batch, dt, ps).cpu()
L.gnode_set_step_kernel(5); L.gnode_set_variant(3)
for a_, b_ in (("stream", "ffma"), ("dual", "ffma"), ("stream", "dual")):
    d = (out[a_] - out[b_]).abs()                       # [T, M, 3]
    per_t = d.amax(dim=(1, 2))
    row = int(d.amax(dim=(0, 2)).argmax())
    first = int((per_t > 1e-5).nonzero()[0]) if (per_t > 1e-5).any() else -1
    print("%-6s vs %-5s: max %.3e at row %d (trial %d node %d degree %d); first grid point over 1e-5: %d; per-trial max %s; rows over 1e-5: %d" % (
        a_, b_, d.max().item(), row, row // N, row % N, deg[row % N], first,
        ["%.2e" % d[:, k * N:(k + 1) * N].max().item() for k in range(B)], int((d.amax(dim=(0, 2)) > 1e-5).sum())))
    print("        max over rows per grid point:", " ".join("%.1e" % v for v in per_t.tolist()[::3]))
worst = int(torch.stack([(out["stream"] - out["ffma"])[:, k * N:(k + 1) * N].abs().max() for k in range(B)]).argmax())
coo = orc.batch_coo([A], [0])
ref32 = orc.forward(xs[worst], params, coo, t)
torch.set_default_dtype(torch.float64)
ref64 = orc.forward(xs[worst].double(), {k: v.double() for k, v in params.items()}, coo, t)
torch.set_default_dtype(torch.float32)
print("trial %d against the CPU oracle: fp32 oracle vs fp64 %.3e" % (worst, (ref32.double() - ref64).abs().max().item()))
for name in ("stream", "dual", "ffma"):
    mine = out[name][:, worst * N:(worst + 1) * N]
    print("   %-6s vs fp32 oracle %.3e   vs fp64 oracle %.3e" % (name, (mine - ref32).abs().max().item(), (mine.double() - ref64).abs().max().item()))
