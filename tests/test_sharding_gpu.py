"""Sharded rollout == single-GPU rollout, bitwise per trial: emulated on one device by
running each rank's shard separately (trials never interact in the forward), and gradient all-reduce
equivalence by summing the shards' gradients."""
import pytest
import torch

from _util import Golden
from oracle import gnode_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sharded_forward_and_gradient_sum():
    import gn_ode_sir_b200 as gn
    from gn_ode_sir_b200 import parallel
    g = Golden("sim_dolphins_b4")
    A, N, B, world = g.adjs[0], g.adjs[0].shape[0], 6, 4
    params = orc.default_params(64, seed=21)
    x = torch.stack([orc.synthetic_trial(N, 64, 500 + b) for b in range(B)]).to(DEV)
    w = torch.randn(40, B * N, 3, generator=torch.Generator().manual_seed(2)).to(DEV)

    def block():
        of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, 64, DEV)
        blk = gn.ode_sim.ODEBlock(20, 0.5, N, [0, 1], 64, of, DEV)
        blk.load_state_dict(params)
        return blk.to(DEV)

    full = block()
    S, I, R = full(x)
    probs_full = torch.cat((S, I, R), -1)
    (probs_full * w).sum().backward()
    grads_sum = None
    for rank in range(world):
        lo, hi = parallel.shard_trials(B, world, rank)
        if lo == hi:
            continue
        blk = block()
        S, I, R = blk(x[lo:hi])
        probs = torch.cat((S, I, R), -1)
        assert torch.equal(probs.detach(), probs_full.detach()[:, lo * N:hi * N])      # bitwise per trial
        (probs * w[:, lo * N:hi * N]).sum().backward()
        gs = [p.grad.clone() for p in blk.parameters() if p.grad is not None]
        grads_sum = gs if grads_sum is None else [a + b for a, b in zip(grads_sum, gs)]
    ref = [p.grad for p in full.parameters() if p.grad is not None]
    for a, b in zip(grads_sum, ref):
        assert (a - b).abs().max().item() <= 2e-5 * max(b.abs().max().item(), 1.0)
