"""Per-phase cycle breakdown of step_tc_kernel (debug instrumentation, GNODE_DBG bit 7)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GNODE_DBG"] = str(int(os.environ.get("GNODE_DBG", "0")) | 128)
import numpy as np, torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import _lib, synth
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 32
A = synth.epinions_standin(0); N = A.shape[0]
dev = torch.device("cuda:0")
torch.manual_seed(0)
of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev); blk = gn.ode_sim.ODEBlock(20, 0.5, N, [0, 1], 64, of, dev).to(dev).eval()
x = torch.stack([synth.synthetic_trial(N, 64, b) for b in range(trials)]).to(dev)
L = _lib.lib()
with torch.no_grad():
    blk(x); torch.cuda.synchronize()
    out = (ctypes.c_longlong * 8)(); L.gnode_debug_phase_cycles(out)
    blk(x); torch.cuda.synchronize()
    L.gnode_debug_phase_cycles(out)
v = np.array(list(out), dtype=np.float64)
names_ws = ["WK wait csr", "WK rows (gather+update)", "WK wait S'", "PEA wait free", "PEA load+split+stage", "PEA mma1+E1", "PEB wait csr+a2", "PEB mma2+E2+store"]
names = ["P1 load+split..S1", "P2 mma1+ci+E1..S2", "P3a gather", "P3b update+decode", "S3 wait", "P4 mma2+E2..S4", "P5 store..S5", "-"]
tiles = 39 * ((trials * N + 127) // 128)
print("tiles", tiles, "total cycles/tile %.0f" % (v.sum() / tiles))
names_dual = ["P1 load+split..S1", "P2 mma1+ci+E1..S2", "P3 gather+update", "-", "S3 wait", "P4 mma2+softmax+E2..S4", "P5 store..S5", "-"]
kern = os.environ.get("GNODE_STEP_KERNEL", "3")
if kern == "2": names = names_ws
if kern == "3":
    names = names_dual
    tiles = tiles / 2      # only half 0 of every CTA is instrumented
for n, c in zip(names, v):
    print("%-22s %8.0f cycles/tile  %5.1f%%" % (n, c / tiles, 100 * c / v.sum()))
