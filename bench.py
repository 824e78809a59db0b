#!/usr/bin/env python
"""Benchmark of the GN-ODE rollout hot path (BASELINE.json metric: rollout node-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (config.workload): epinions-scale rollout inference -- BASELINE.json configs[3].
The epinions pickle is missing from the reference checkout, so the graph is the synthetic
stand-in BA(N=75,879, m=5, seed=0); trials are the synthetic (beta, gamma, I0)
draws of monitorer-sim.py:116-119; weights are the default nn.Linear init under seed 0.
One "step" = one full rollout (T-1 = 39 Euler steps + encoder + decoder) of the rank's trials.
Trials are independent -> sharded across ranks with the graph replicated, no collective on the
data path; per-rank trial count is fixed, so scaling is "weak".

Prints ONE JSON line (rank 0). value = node-steps/s with inputs resident in HBM; e2e = same
metric through ODEBlock.forward with pinned-host input and output copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np    # noqa: E402
import torch          # noqa: E402

H = 64
MAXTIME, DELTAT = 20, 0.5
ALGO_BYTES_PER_NODE_STEP = 2060.0      # SURVEY 8d / BASELINE.md section 3 (fixed denominator)
R_STATE_NOTE = {
    1: "inference carries R as hid(R) = W3 R (4 floats per row, exact by linearity of R' = gamma I'); the 64-float R plane is "
       "neither read nor written, i.e. 512 of the 2060 algorithmic bytes per node-step are not moved -- the roofline denominator "
       "stays 2060 B (SURVEY 8d); GNODE_R_STATE=full keeps the plane",
    0: "full 64-float R plane (GNODE_R_STATE=full)"}
JSON_OUT = sys.stdout
UNIT = "node-steps/s"
METRIC = "GN-ODE rollout node-steps/s (epinions)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(rows):
    """DRAM bytes per step-kernel launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed
    `ncu --set full` capture (profiles/step_kernel_traffic.json holds bytes per row of that capture)."""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["dram_bytes_per_row"]) * rows
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


WORKLOADS = {
    "epinions": ("epinions stand-in BA(N=75879,m=5,seed=0)", "BASELINE.json configs[3]"),
    "ba2m": ("BA(N=2000000,m=10,seed=0) stress graph (chunked preferential attachment, ~20M edges)", "BASELINE.json configs[4]"),
}


def build_workload(trials, trial_offset, workload="epinions"):
    from gn_ode_sir_b200 import synth
    A = synth.epinions_standin(seed=0) if workload == "epinions" else synth.ba_stress(seed=0)
    N = A.shape[0]
    x = torch.zeros(trials, N, 3 + H, dtype=torch.float32)
    for b in range(trials):
        rng = np.random.RandomState(1000 + trial_offset + b)
        seeds = rng.choice(N, 2, replace=False)
        beta, gamma = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)
        x[b, :, 0] = 1.0
        x[b, seeds, 0] = 0.0
        x[b, seeds, 1] = 1.0
        x[b, :, 3], x[b, :, 4] = beta, gamma
    return A, x


def cpu_reference_sample(A, trials, n_points, repeats=1):
    """Times the oracle port of the reference's CPU path (torch ops, per-step host-side
    block_diag rebuild as at ode_nn_ngraph_sim.py:68-71) on a bounded sample. The only place
    where bench.py touches oracle/ (cpu_baseline leg and the --impl reference arm)."""
    import scipy.sparse
    from oracle import gnode_oracle as orc
    N = A.shape[0]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = orc.default_params(H, seed=0)
    x = torch.cat([orc.synthetic_trial(N, H, b) for b in range(trials)])
    t = orc.time_grid(MAXTIME, DELTAT)[:n_points]

    def rebuild():
        bd = scipy.sparse.block_diag([A for _ in range(trials)])
        return torch.LongTensor(np.vstack((bd.row, bd.col)))

    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.forward(x, params, None, t, rebuild_index=rebuild)
        best = min(best, time.perf_counter() - t0)
    units = trials * N * (n_points - 1)
    return units / best, cores, "%d trial(s) x %d Euler steps of the bench workload (%d node-steps, %.1f s)" % (
        trials, n_points - 1, units, best)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    A, _ = build_workload(0, 0, args.workload)
    for _ in range(args.warmup):
        cpu_reference_sample(A, 1, 3)
    vals, times = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, cores, sample = cpu_reference_sample(A, args.ref_trials, args.ref_points)
        times.append(time.perf_counter() - t0)
        vals.append(v)
    value = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, world),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), file=JSON_OUT, flush=True)


def workload_config(args, world, r_state=None):
    name, cfg = WORKLOADS[args.workload]
    return {"workload": "%s rollout inference, H=64, T=40 (maxTime=20, deltaT=0.5), %d trials per GPU (%s)" % (name, args.trials, cfg),
            "trials_per_gpu": args.trials, "global_trials": args.trials * world,
            "nodes": 75879 if args.workload == "epinions" else 2000000,
            "euler_steps": int(len(np.arange(0, MAXTIME, DELTAT)) - 1), "parallelism": "trial-sharded dp%d, graph replicated" % world,
            "l2_policy": "no flush: per-step working set (state+I' of all trials, >3 GB) exceeds the 126 MB L2",
            "r_state": R_STATE_NOTE[r_state] if r_state is not None else None}


def main():
    # stdout carries exactly ONE JSON line: everything else that libraries write to file descriptor 1 (NCCL prints its
    # version banner there under torchrun) is sent to stderr, and the JSON line goes to the saved descriptor.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trials", type=int, default=128, help="trials per GPU")
    ap.add_argument("--workload", default="epinions", choices=sorted(WORKLOADS),
                    help="epinions = the metric's configuration (default); ba2m = the 2M-node stress graph (use --trials 8)")
    ap.add_argument("--e2e-chunk", type=int, default=32, help="trials per pipelined chunk of the e2e loop")
    ap.add_argument("--ref-trials", type=int, default=2)
    ap.add_argument("--ref-points", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GN-ODE rollout has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1:
        # one process per GPU: host buffers next to the GPU (8 ranks x 28 GB/s of PCIe traffic must not cross sockets)
        from gn_ode_sir_b200.parallel import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)
        print("bench.py: rank %d bound to %s" % (rank, ("%d CPUs near GPU %d" % (len(numa_cpus), local_rank)) if numa_cpus else "no CPU set (unchanged)"),
              file=sys.stderr, flush=True)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    import gn_ode_sir_b200 as gn
    from gn_ode_sir_b200 import _lib
    gn.build_library()
    L = _lib.lib()

    A, x_host = build_workload(args.trials, rank * args.trials, args.workload)
    N = A.shape[0]
    T = len(np.arange(0, MAXTIME, DELTAT))
    torch.manual_seed(0)
    of = gn.ode_sim.ODEfunc(A, 0.2, 0.1, H, dev)
    blk = gn.ode_sim.ODEBlock(MAXTIME, DELTAT, N, [0, 1], H, of, dev).to(dev).eval()
    units_per_step = args.trials * N * (T - 1)
    rows = args.trials * N

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident measurement (value, roofline)
    x_dev = x_host.to(dev)
    with torch.no_grad():
        for _ in range(args.warmup):
            blk(x_dev)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = int(L.gnode_launch_count())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        ev[0].record()
        for _ in range(args.steps):
            S, I, R = blk(x_dev)
        ev[1].record()
        barrier()
        clocks = sampler.stop()
        launches = int(L.gnode_launch_count()) - launches0
        ms_total = max_over_ranks(ev[0].elapsed_time(ev[1]))
    ms_per_step = ms_total / args.steps
    value = world * units_per_step / (ms_per_step * 1e-3)
    # dominant kernel = the fused Euler-step kernel: (T-1) launches per rollout, each over all rows;
    # the encoder launch (1 of T) is timed in the same stream window and is charged to the step kernel
    # (conservative: makes the per-launch time slightly larger).
    step_launches = args.steps * (T - 1)
    step_ms = ev[0].elapsed_time(ev[1]) / step_launches
    peak, peak_src = measured_peak()
    achieved = rows * ALGO_BYTES_PER_NODE_STEP / (step_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch(rows)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": "gnode::step_dual_kernel (fused Euler step, tcgen05; encoder + final decode launches charged to it)", "launch_ms": step_ms, "rows_per_launch": rows,
                "algorithmic_bytes_per_node_step": ALGO_BYTES_PER_NODE_STEP, "peak_source": peak_src,
                "traffic_source": "profiles/step_kernel_traffic.json (ncu --set full dram bytes per row of the committed capture) x rows"}
    del S, I, R

    # ---------------- end-to-end through the public API with host buffers
    # The caller's loop: pinned host x -> device, ODEBlock.forward, probabilities -> pinned host. The trials
    # are fed in chunks over three streams so that the PCIe copies of chunk c-1 / c+1 overlap the rollout
    # of chunk c (plain PyTorch stream code around the drop-in module; every byte is copied every step; the
    # K timed steps form one continuous stream of chunks, timed from the first H2D to the last D2H).
    bc = max(1, min(args.e2e_chunk, args.trials))
    n_chunks = (args.trials + bc - 1) // bc
    x_pin = x_host.pin_memory()
    del x_host                                   # one host copy per rank (8 ranks share the box's RAM)
    out_pin = torch.empty((n_chunks, T, bc * N, 3), dtype=torch.float32).pin_memory()   # chunk-major host result
    h2d_bytes = x_pin.numel() * 4
    d2h_bytes = T * rows * 3 * 4
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    xd = [torch.empty((bc, N, 3 + H), dtype=torch.float32, device=dev) for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n_steps):
        """n_steps passes over this rank's trials as ONE continuous stream of chunks (a long job does not drain the
        pipeline between batches): H2D of chunk i+1 and D2H of chunk i-1 overlap the rollout of chunk i."""
        main = torch.cuda.current_stream()
        for st in (s_in, s_cmp, s_out):
            st.wait_stream(main)
        for i in range(n_steps * n_chunks):
            c = i % n_chunks
            b0, b1, buf = c * bc, min(args.trials, (c + 1) * bc), i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[buf])                       # rollout of chunk i-2 has consumed xd[buf]
                xd[buf][:b1 - b0].copy_(x_pin[b0:b1], non_blocking=True)
                ev_in[buf].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[buf])
                S, I, R = blk(xd[buf][:b1 - b0])                        # views of one [T, M_c, 3] buffer
                probs_c = S._base if S._base is not None else torch.cat((S, I, R), -1)
                ev_free[buf].record(s_cmp)
                ev_cmp[buf].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[buf])
                out_pin[c, :, :(b1 - b0) * N].copy_(probs_c, non_blocking=True)
                probs_c.record_stream(s_out)
        for st in (s_in, s_cmp, s_out):
            main.wait_stream(st)

    with torch.no_grad():
        e2e_steps(max(1, min(args.warmup, 2)))
        barrier()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record()
        e2e_steps(args.steps)
        ev2[1].record()
        barrier()
        e2e_ms = max_over_ranks(ev2[0].elapsed_time(ev2[1])) / args.steps
    e2e = {"value": world * units_per_step / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms, "chunks": n_chunks}

    # ---------------- CPU baseline (oracle port of the reference's CPU path), rank 0, N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_reference_sample(A, args.ref_trials, args.ref_points)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, int(L.gnode_get_r_state())),
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
