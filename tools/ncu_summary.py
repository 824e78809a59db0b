"""`ncu -i X.ncu-rep --page raw --csv` -> the per-kernel summary committed under profiles/ (and, with --traffic, the
DRAM bytes per row that bench.py scales into roofline.traffic).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/<name>_ncu_summary.csv [--rows ROWS --traffic profiles/step_kernel_traffic.json --note "..."]
The first profiled launch whose kernel name contains --kernel (default step_dual_kernel) is summarised."""
import argparse, csv, json

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__t_sector_op_read_hit_rate.pct", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__shared_mem_config_size", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]
ap = argparse.ArgumentParser()
ap.add_argument("raw"); ap.add_argument("out")
ap.add_argument("--kernel", default="step_dual_kernel")
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--traffic", default=None)
ap.add_argument("--note", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
rec = next(r for r in rows[2:] if a.kernel in r[ix["Kernel Name"]])


def find(name):
    hits = [h for h in hdr if h == name or h.endswith("." + name)]
    for h in hits:
        if rec[ix[h]] != "":
            return h
    return hits[0] if hits else None


out = [("kernel", "", rec[ix["Kernel Name"]])]
for m in METRICS:
    h = find(m)
    if h:
        out.append((m, units[ix[h]], rec[ix[h]]))
stalls = [(h, units[ix[h]], rec[ix[h]]) for h in hdr if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h and rec[ix[h]] not in ("", "0")]
stalls.sort(key=lambda s: -float(s[2].replace(",", "")))
out += [(h.split(".")[-1] if "." in h else h, u, v) for h, u, v in stalls[:12]]
with open(a.out, "w", newline="") as fh:
    w = csv.writer(fh); w.writerow(["metric", "unit", "value"]); w.writerows(out)
d = {k: (u, v) for k, u, v in out}
if a.traffic and a.rows:
    def gb(k):
        u, v = d[k]; v = float(v.replace(",", ""))
        return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}[u]
    r, wv = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    json.dump({"kernel": "gnode::" + a.kernel, "capture": a.note, "rows": a.rows, "dram_bytes_read": r, "dram_bytes_write": wv,
               "units": "Gbyte", "dram_bytes_per_row": (r + wv) * 1e9 / a.rows}, open(a.traffic, "w"), indent=1)
for k, u, v in out:
    print(k, u, v)
