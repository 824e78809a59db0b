"""Data-parallel training check, one process per rank under torchrun (NCCL when every rank has its own GPU, gloo when
the ranks share one GPU): the five-graph training batch of tests/golden/ng_train5_b8.npz (BASELINE configs[2]) is split
across the ranks by node count, each rank runs forward + fused L1 + reverse sweep on its share, ONE flat all-reduce sums
the parameter gradient. Checked on every rank against the single-process run of the whole batch:
  * the probabilities of this rank's instances are BITWISE those of the full-batch run (no cross-trial arithmetic),
  * the summed gradient agrees with the full-batch gradient to 2e-5 (scale-relative; different summation order),
  * the global mean loss agrees to 1e-12, and after one Adam step all ranks hold identical weights.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np            # noqa: E402
import torch                  # noqa: E402
import torch.distributed as dist   # noqa: E402
import gn_ode_sir_b200 as gn  # noqa: E402
from gn_ode_sir_b200 import harness, parallel   # noqa: E402
from _util import Golden      # noqa: E402


def main():
    rank, world, dev = harness.init_distributed()
    g = Golden("ng_train5_b8")
    maxTime, deltaT = g.maxTime, g.deltaT
    torch.manual_seed(0)
    of = gn.ode_ngraphs.ODEfunc(g.adjs, g.H, dev)
    blk = gn.ode_ngraphs.ODEBlock(maxTime, deltaT, g.H, of, dev)
    blk.load_state_dict(g.params)
    blk.to(dev)
    if world > 1:
        parallel.broadcast_parameters(blk)
    # instances (x_i, y_i, graph id) with seeded float64 "labels"
    gen = torch.Generator().manual_seed(3)
    items, r0 = [], 0
    for gi in g.inst_graph:
        n = g.adjs[gi].shape[0]
        y = torch.rand(n, maxTime, 3, generator=gen, dtype=torch.float64)
        items.append((g.x[r0:r0 + n], y / y.sum(-1, keepdim=True), gi))
        r0 += n
    steps = gn.rollout.unit_time_steps(maxTime, deltaT)
    crit = torch.nn.L1Loss()
    # ---- single-process run of the whole batch (every rank computes it for comparison)
    xf = torch.cat([it[0] for it in items]).to(dev)
    yf = torch.cat([it[1] for it in items]).to(dev)
    blk.zero_grad()
    probs_full = blk.rollout_probs(xf, out_steps=steps, instances=[it[2] for it in items])
    loss_full = gn.rollout.l1_subsampled(probs_full, yf)
    loss_full.backward()
    grads_full = {k: p.grad.detach().clone() for k, p in blk.named_parameters() if p.grad is not None}
    # ---- data-parallel: this rank's share
    blk.zero_grad()
    x, y, inst, share = harness.shard_batch(items, maxTime)
    mine = parallel.shard_instances([it[0].size(0) for it in items], world)[rank] if world > 1 else list(range(len(items)))
    loss_local = torch.zeros((), dtype=torch.float64, device=dev)
    if x is not None:
        probs = blk.rollout_probs(x.to(dev), out_steps=steps, instances=inst)
        # bitwise per instance against the full-batch run
        starts = np.concatenate(([0], np.cumsum([it[0].size(0) for it in items])))
        off = 0
        for i in mine:
            n = items[i][0].size(0)
            assert torch.equal(probs[:, off:off + n].detach(), probs_full[:, starts[i]:starts[i] + n].detach()), \
                "rank %d: instance %d differs from the full-batch rollout" % (rank, i)
            off += n
        loss = gn.rollout.l1_subsampled(probs, y.to(dev), scale=share)
        loss.backward()
        loss_local = loss.detach() * share
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    parallel.allreduce_gradients(blk.parameters())
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        buf = loss_local if dist.get_backend() == "nccl" else loss_local.cpu()
        dist.all_reduce(buf)
        loss_local = buf.to(dev)
    assert abs(loss_local.item() - loss_full.item()) < 1e-12, (loss_local.item(), loss_full.item())
    worst = 0.0
    for k, p in blk.named_parameters():
        if k in grads_full:
            scale = max(grads_full[k].abs().max().item(), 1e-3)
            worst = max(worst, (p.grad - grads_full[k]).abs().max().item() / scale)
    assert worst < 2e-5, worst
    # one Adam step: replicas stay identical
    opt = torch.optim.Adam(blk.parameters(), lr=1e-3)
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in blk.parameters()])
    if world > 1:
        ref = flat.clone() if dist.get_backend() == "nccl" else flat.cpu().clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref.to(dev), flat), "rank %d: weights diverged after the optimiser step" % rank
    print("DP_CHECK OK rank %d/%d backend %s: %d of %d instances (%d rows), grad err %.2e, loss %.12f, all-reduce %.3f ms" % (
        rank, world, dist.get_backend() if world > 1 else "none", len(mine), len(items), 0 if x is None else x.size(0),
        worst, loss_full.item(), ev0.elapsed_time(ev1)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
