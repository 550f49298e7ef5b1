"""Summaries of ncu captures for profiles/ (run in the build container on files brought back in gpurun_out/):
    python tools/ncu_summary.py full  <capture.ncu-rep> <out.json> [<traffic.json> <kernel label> <config label> <algorithmic bytes>]
    python tools/ncu_summary.py list  <launches.csv>   <out.json>
"""
import collections, csv, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value) * mult


def full(rep, out, traffic=None, kernel=None, config=None, alg_bytes=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ti = hdr.index("gpu__time_duration.sum")
    vals = max(rows[2:], key=lambda r: float(r[ti]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ti], 1.0))   # longest launch of the capture
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h):
            d[h] = {"unit": u, "value": v}
    d["kernel_name"] = vals[hdr.index("Kernel Name")]
    json.dump(d, open(out, "w"), indent=1)
    if traffic:
        rd = to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"])
        wr = to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
        json.dump({"kernel": kernel, "config": config, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "traffic_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": int(alg_bytes),
                   "source": "ncu --set full, " + out}, open(traffic, "w"), indent=1)


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["Metric Unit"]]
                agg[d["Kernel Name"].split("(")[0][:80]].append(v)
    total = sum(sum(v) for v in agg.values())
    top = sorted(({"kernel": k, "count": len(v), "ms": sum(v), "share": sum(v) / total, "mean_us": 1e3 * sum(v) / len(v)} for k, v in agg.items()),
                 key=lambda e: -e["ms"])
    json.dump({"launches": sum(len(v) for v in agg.values()), "total_ms": total, "top": top[:10]}, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(*sys.argv[2:])
    else:
        launch_list(*sys.argv[2:])
