"""Turn the ncu artefacts of one round (gpurun_out/*.ncu-rep, launches CSV) into the small text/JSON
summaries committed under profiles/.  Run in the build container (ncu reads reports without a GPU):

    python profiles/summarize.py r1
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return [dict(zip(rows[0], zip(rows[1], r))) for r in rows[2:]]


def to_bytes(unit, value):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * scale.get(unit, 1)


def summarize_report(rep, label):
    out = []
    traffic = None
    for k in raw_page(rep):
        name = k.get("Kernel Name", ("", "?"))[1]
        out.append(f"# {label}: {name}")
        for key in KEYS:
            if key in k:
                out.append(f"{key:70s} {k[key][1]:>16s} {k[key][0]}")
        for key, (unit, val) in k.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", key)
            if m and float(val) >= 0.05:
                out.append(f"stall/{m.group(1):64s} {float(val):16.3f} warps per issue")
        if "dram__bytes_read.sum" in k:
            traffic = to_bytes(*k["dram__bytes_read.sum"]) + to_bytes(*k["dram__bytes_write.sum"])
            out.append(f"{'dram traffic per launch (read + write)':70s} {traffic / 1e6:16.3f} MB")
        out.append("")
    return "\n".join(out), traffic


def summarize_launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
    unit_i = hdr.index("Metric Unit")
    agg = defaultdict(list)
    order = []
    for r in rows[1:]:
        v = float(r[val_i].replace(",", ""))
        v = v / 1e3 if r[unit_i] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*", "", r[name_i]).replace("aat::<unnamed>::", "")
        agg[name].append(v)
        order.append((name, v))
    total = sum(sum(v) for v in agg.values())
    lines = [f"# {tag}: kernel launches of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e` under",
             "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
             f"{'kernel':60s} {'launches':>8s} {'avg us':>10s} {'share':>7s}"]
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{name[:60]:60s} {len(v):8d} {sum(v) / len(v):10.2f} {100 * sum(v) / total:6.1f}%")
    return "\n".join(lines) + "\n"


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    traffic = {}
    for rep, label, wl in ((f"prof_logmel_{tag}.ncu-rep", "logmel, config 2", None),
                           (f"prof_pool_{tag}.ncu-rep", "pool, config 2", "c2"),
                           (f"prof_pool_c3_{tag}.ncu-rep", "pool, config 3", "c3"),
                           (f"prof_pool_c4_{tag}.ncu-rep", "pool, config 4", "c4"),
                           (f"prof_boundaries_{tag}.ncu-rep", "boundaries, config 2", None)):
        path = os.path.join(SRC, rep)
        if not os.path.exists(path):
            continue
        text, t = summarize_report(path, label)
        with open(os.path.join(OUT, f"{tag}_{os.path.splitext(rep)[0].replace('_' + tag, '')}.txt"), "w") as f:
            f.write(text)
        if wl and t:
            traffic[wl] = t
    lp = os.path.join(SRC, f"launches_{tag}.csv")
    if os.path.exists(lp):
        with open(os.path.join(OUT, f"{tag}_launch_shares.txt"), "w") as f:
            f.write(summarize_launches(lp, tag))
    tp = os.path.join(OUT, "pool_traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update(traffic)
    json.dump(old, open(tp, "w"), indent=1, sort_keys=True)
    print("wrote summaries for", tag, "traffic", traffic)


if __name__ == "__main__":
    main()
