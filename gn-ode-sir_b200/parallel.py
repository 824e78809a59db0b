"""Trial-sharded data parallelism.

Every (beta, gamma, seed-set) trial / graph instance is independent, so the batch is split across
ranks with the graph(s) and the ~19 KB of weights replicated. Inference needs no collective; training
needs ONE all-reduce (sum) of the flattened parameter gradient per optimiser step -- latency-bound at
4.8k floats, so it is issued as a single flat buffer. One process per GPU (torch.distributed, NCCL on
GPUs / gloo in the CPU tests)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index, min_cpus=2):
    """Pin the calling process to the CPUs next to its GPU (NVML's ideal CPU affinity, intersected with the CPUs this
    process may use), BEFORE it allocates pinned host buffers: the pages then come from the GPU's own NUMA node. With one
    process per GPU and every rank streaming inputs and results over PCIe, buffers that land on the other socket cap
    the aggregate host traffic (8 ranks: 120 GB/s measured). Returns the CPU list, or None when nothing was changed
    (no NVML, fewer than `min_cpus` usable CPUs, already bound, or GNODE_NO_NUMA_BIND set)."""
    if os.environ.get("GNODE_NO_NUMA_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = None
        try:
            props = torch.cuda.get_device_properties(device_index)
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            for arg in (bus, bus.encode()):
                try:
                    handle = pynvml.nvmlDeviceGetHandleByPciBusId(arg)
                    break
                except (TypeError, AttributeError):
                    continue
        except Exception:
            handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, 16)            # 1024 CPU bits
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = ideal & allowed
        if len(target) < min_cpus or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return sorted(target)
    except Exception:
        return None


def shard_instances(sizes, world_size):
    """Greedy longest-processing-time split of instances (by node count) over ranks.
    Returns a list of index lists; each list keeps the original relative order."""
    sizes = [int(s) for s in sizes]
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += sizes[i]
    return [sorted(p) for p in parts]


def shard_trials(n_trials, world_size, rank):
    """Contiguous split of equally sized trials (sim variant): rank r owns [lo, hi)."""
    base, extra = divmod(n_trials, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_gradients(parameters, group=None, average=False):
    """Sum (or average) the gradients of `parameters` across ranks with ONE all-reduce on a flat
    buffer. Parameters without a gradient on this rank contribute zeros (ranks may hold empty shards)."""
    params = [p for p in parameters if p.requires_grad]
    if not params or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights (replicas must be identical)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class ShardedRollout:
    """Runs a drop-in ODEBlock on this rank's share of a sim-variant batch x [B, N, 3+H].
    loss_fn(S, I, R, lo, hi) must return this shard's SUM loss; `step` all-reduces the gradients so
    that every rank applies the identical optimiser update."""

    def __init__(self, block, group=None):
        self.block = block
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def local_slice(self, n_trials):
        return shard_trials(n_trials, self.world, self.rank)

    def forward_local(self, x):
        lo, hi = self.local_slice(x.size(0))
        if hi == lo:
            return None, lo, hi
        return self.block(x[lo:hi]), lo, hi

    def step(self, optimizer, x, loss_fn):
        optimizer.zero_grad()
        out, lo, hi = self.forward_local(x)
        loss = None
        if out is not None:
            loss = loss_fn(out[0], out[1], out[2], lo, hi)
            loss.backward()
        allreduce_gradients(self.block.parameters(), self.group)
        optimizer.step()
        return loss
