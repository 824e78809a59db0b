"""Error of every step-kernel structure against the reference's own outputs (tests/golden): max |cuda - reference fp32| on
the well-conditioned cases, and |cuda - float64| beside the reference's own fp32-vs-float64 error on the large ones.

    python tools/kernel_error_table.py [kernel ...]      (default: 0 3 5; a library built with -DGNODE_ABLATIONS also 10..13)
Kernel 0 is run with variant 0 (FFMA + expf: the reference's arithmetic), the others with variant 3 (tcgen05 + MUFU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gn_ode_sir_b200 as gn
from gn_ode_sir_b200 import _lib
from _util import LARGE_CASES, STRICT_CASES, Golden
from test_parity_gpu import run_cuda

kernels = [int(k) for k in sys.argv[1:]] or [0, 3, 5]
L = _lib.lib()
L.gnode_set_persistent(0)            # the A/B kernels exist only as one launch per step
print("%-22s" % "case" + "".join("  k=%-9d" % k for k in kernels) + "  (reference fp32 vs float64)")
for name in STRICT_CASES + LARGE_CASES:
    g = Golden(name)
    row = "%-22s" % name
    large = name in LARGE_CASES
    ref = g.probs64.double() if large else g.probs32
    for k in kernels:
        _lib.check(L.gnode_set_variant(0 if k == 0 else 3), "variant")
        _lib.check(L.gnode_set_step_kernel(k), "kernel")
        p = run_cuda(gn, g)[:: g.tstride]
        row += "  %.3e  " % ((p.double() - ref.double()).abs().max().item())
    if large:
        row += "  %.3e" % (g.probs32.double() - ref).abs().max().item()
    print(row)
