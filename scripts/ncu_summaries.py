"""Turn the scratch ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/ncu_summaries.py launches gpurun_out/launches_r01c.csv "<header note>"  > profiles/r01c_launches_summary.csv
    python scripts/ncu_summaries.py raw gpurun_out/hot_r01c.ncu-rep "<header note>"        > profiles/r01c_ncu_full_summary.txt
"""
import collections
import csv
import re
import subprocess
import sys

RAW_METRICS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("tome::", "")


def launches(path, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        us = v / 1000 if unit.startswith("n") else v if unit.startswith("u") else v * 1000
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    print(f"# {note}")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's `kernels` shares, not absolutes")
    print(f"# total {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches")
    print("kernel,launches,total_ms,share,avg_us")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{n},{us / 1000:.3f},{us / tot:.4f},{us / n:.1f}')


def raw(path, note):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {note}")
    print(f"# report: {path} (scratch); selected raw metrics per captured launch")
    for r in rows[2:]:
        print("---")
        for m in RAW_METRICS:
            if m in ix:
                print(f"{m} [{units[ix[m]]}] = {r[ix[m]]}")
        stalls = {h: float(r[ix[h]] or 0) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
        print("top stalls (warps per issue): " + ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for k, v in top))


def traffic(path, note):
    """profiles/roofline_traffic.json: mean dram read + write bytes per launch of every kernel in an ncu --set full report."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = short(r[ix["Kernel Name"]]).split("<")[0]
        b = sum(float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += b
    print(json.dumps({k: {"bytes_per_launch": v[1] / v[0], "launches": v[0], "source": f"{note} ({path}, ncu --set full)"} for k, v in agg.items()}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
