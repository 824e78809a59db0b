"""Same-box A/B of library switches on the bench workload (epinions stand-in, inference, inputs resident in HBM, CUDA
events): interleaves the configurations given on the command line and prints node-steps/s for each pass.

    python tools/ab_bench.py [--trials 64] [--rounds 3] r_state=0 r_state=1 dbg=1048576 ...
Each configuration is a comma-separated list of key=value: r_state (gnode_set_r_state), dbg (env GNODE_DBG, read per
rollout), kernel (gnode_set_step_kernel), variant (gnode_set_variant: 0 = FFMA + expf, the reference's arithmetic).
A library built with -DGNODE_ABLATIONS (_build.build_library(force=True, extra_flags=["-DGNODE_ABLATIONS"])) adds the
A/B kernels 7 = 3xTF32 (lo x lo dropped), 10 = round-2h operands (N = 80, rna split packed in place), 11 = N = 160 with
the rna split, 12 = 11 with 3xTF32, and the timing-only 8 = no MMAs, 9 = no lo-operand pass (wrong numerics)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import synth, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=64)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--heavy", action="store_true", help="add 8 hubs of degree ~3000 (max degree like soc-Epinions1)")
ap.add_argument("--spread", action="store_true", help="with --heavy: the 8 hubs sit in 8 different tiles (isolated hubs, as in real graphs)")
ap.add_argument("configs", nargs="+")
args = ap.parse_args()

dev = torch.device("cuda:0")
L = _lib.lib()
A = synth.epinions_standin(0)
if args.heavy:
    import scipy.sparse
    rng = np.random.RandomState(1)
    rows, cols = [], []
    for h in range(8):
        if args.spread:
            h = 77 + 128 * 70 * h
        nb = rng.choice(A.shape[0], 3000, replace=False)
        nb = nb[nb != h]
        rows += [np.full(len(nb), h), nb]; cols += [nb, np.full(len(nb), h)]
    Bm = scipy.sparse.csr_matrix((np.ones(sum(len(r) for r in rows), dtype=np.int8), (np.concatenate(rows), np.concatenate(cols))), shape=A.shape)
    A = (A + Bm).tocsr(); A.sum_duplicates(); A.data[:] = 1; A.sort_indices()
N, B, T = A.shape[0], args.trials, 40
torch.manual_seed(0)
of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, dev)
blk = gn.ode_sim.ODEBlock(20, 0.5, N, [0, 1], 64, of, dev).to(dev).eval()
x = torch.zeros(B, N, 67)
for b in range(B):
    rng = np.random.RandomState(1000 + b)
    s = rng.choice(N, 2, replace=False)
    x[b, :, 0] = 1.0; x[b, s, 0] = 0.0; x[b, s, 1] = 1.0
    x[b, :, 3], x[b, :, 4] = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)
x = x.to(dev)


def apply(cfg):
    os.environ.pop("GNODE_DBG", None)
    L.gnode_set_r_state(1); L.gnode_set_step_kernel(5); L.gnode_set_variant(3)
    for kv in cfg.split(","):
        k, v = kv.split("=")
        if k == "r_state":
            L.gnode_set_r_state(int(v))
        elif k == "dbg":
            os.environ["GNODE_DBG"] = v
        elif k == "kernel":
            L.gnode_set_step_kernel(int(v))
        elif k == "variant":
            L.gnode_set_variant(int(v))
        elif k != "base":
            raise SystemExit("unknown key " + k)


ref = None
with torch.no_grad():
    for r in range(args.rounds):
        for cfg in args.configs:
            apply(cfg)
            S, I, R = blk(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                S, I, R = blk(x)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            p = S._base if S._base is not None else torch.cat((S, I, R), -1)
            if ref is None:
                ref = p.clone()
            print("%-28s %.4e node-steps/s  %.2f ms  max|p - p_first| = %.2e" % (
                cfg, B * N * (T - 1) / (ms * 1e-3), ms, (p - ref).abs().max().item()), flush=True)
            del S, I, R, p
