"""Turn the ncu outputs of tools/gpu_round.sh into the tracked summaries under profiles/.

  python tools/ncu_summarise.py <tag> [<out-prefix>]
reads  gpurun_out/<tag>_launches.csv  (ncu --metrics gpu__time_duration.sum --csv launch list)
       gpurun_out/<tag>_full*_raw.csv (ncu -i <rep> --page raw --csv of the --set full captures)
writes profiles/<out>_launch_summary.csv, profiles/<out>_ncu_full_summary.csv, profiles/<out>_traffic.json
"""
import collections
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def short(name):
    name = name.replace("void ", "").replace("mode::", "")
    return name.split("(")[0]


def launches(tag, out):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        agg.setdefault(short(r[ik]), []).append(us)
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{out}_launch_summary.csv"), "w") as f:
        f.write(f"# {out} ncu launch list summary: the steps of bench.py only (REPMODE_BENCH_FAST=2: no e2e / CPU / per-kernel"
                " legs; REPMODE_BENCH_GRAPH=0 REPMODE_OVERLAP=0: eager, one stream); cold-cache serialised times: compare"
                " SHARES, not absolutes\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 --warmup 3"
                "   (tools/gpu_round.sh)\n")
        f.write("kernel,launches,avg_us,total_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k},{len(v)},{sum(v) / len(v):.2f},{sum(v) / 1e3:.3f},{100 * sum(v) / total:.1f}\n")
    print("launch summary:", len(agg), "kernels,", f"{total / 1e3:.2f} ms")


def full(tag, out):
    paths = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_full*_raw.csv")))
    lines, traffic = [], collections.OrderedDict()
    units = None
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        idx = [hdr.index(k) if k in hdr else None for k in KEEP]
        if units is None:
            units = [rows[1][i] if i is not None else "" for i in idx]
        for r in rows[2:]:
            vals = [r[i] if i is not None else "" for i in idx]
            vals[0] = short(vals[0])
            lines.append(vals)
            try:
                us = float(vals[3].replace(",", ""))
                rd, wr = float(vals[5].replace(",", "")), float(vals[6].replace(",", ""))
                mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                ur, uw = mul.get(rows[1][idx[5]], 1e6), mul.get(rows[1][idx[6]], 1e6)
                b = rd * ur + wr * uw
                traffic.setdefault(vals[0], []).append({"dram_bytes": b, "duration_us": us, "dram_GBps": b / us / 1e3})
            except ValueError:
                pass
    with open(os.path.join(ROOT, "profiles", f"{out}_ncu_full_summary.csv"), "w") as f:
        f.write(f"# {out}: ncu --set full --clock-control none [--import-source on] -k regex:<kernel> -s <skip> -c <n> "
                "(REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 REPMODE_OVERLAP=0 python bench.py --steps 3 --warmup 3; "
                "tools/gpu_round.sh); one row per captured launch\n")
        w = csv.writer(f)
        w.writerow(KEEP)
        w.writerow(units or [])
        w.writerows(lines)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{out}_traffic.json"), "w"), indent=1)
    print("full summary:", len(lines), "launches;", {k: round(sum(t["dram_bytes"] for t in v) / len(v) / 1e6, 1) for k, v in traffic.items()}, "MB/launch")


if __name__ == "__main__":
    tag = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else tag
    if os.path.exists(os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")):
        launches(tag, out)
    full(tag, out)
