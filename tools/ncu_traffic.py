#!/usr/bin/env python
"""Turns the raw pages of the round's ncu --set full captures (tools/r2_call5.sh: `ncu -i ... --page raw --csv --metrics ...`) into
profiles/traffic.json, the DRAM bytes per launch that bench.py reports as roofline.traffic.

usage: tools/ncu_traffic.py <tag> <workload>=<raw.csv> ...     e.g.  tools/ncu_traffic.py r2e atrium1m=gpurun_out/r2e_full_atrium1m_raw.csv

A "launch" is what bench.py's roofline divides by: one stage of one bounce depth, i.e. ALL material-class launches of k_shade of that
depth together. The captures hold two consecutive depths (0 and 1); the entry is their mean."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TSCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def parse(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(head)}
    out = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
        base = re.match(r"[A-Za-z_0-9]+", name).group(0)
        get = lambda m, table: float(r[ix[m]]) * table[units[ix[m]]]
        out.append(dict(kernel=base, name=name.split("(")[0], read=get("dram__bytes_read.sum", SCALE), write=get("dram__bytes_write.sum", SCALE),
                        us=get("gpu__time_duration.sum", TSCALE), regs=int(r[ix["launch__registers_per_thread"]]),
                        warps_active=float(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                        l1_hit=float(r[ix["l1tex__t_sector_hit_rate.pct"]]), l2_hit=float(r[ix["lts__t_sector_hit_rate.pct"]]),
                        threads_per_inst=float(r[ix["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                        issue=float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in ix else float("nan"),
                        alu=float(r[ix["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]) if "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active" in ix else float("nan"),
                        fma=float(r[ix["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]]) if "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active" in ix else float("nan"),
                        inst=float(r[ix["smsp__inst_executed.sum"]])))
    return out


def main():
    tag = sys.argv[1]
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    lines = []
    for arg in sys.argv[2:]:
        workload, path = arg.split("=")
        launches = parse(path)
        # split into depths: a new depth starts at every k_trace_closest
        depths = []
        for l in launches:
            if l["kernel"] == "k_trace_closest" or not depths:
                depths.append([])
            depths[-1].append(l)
        lines.append(f"## {workload} ({os.path.basename(path)}; {len(depths)} consecutive bounce depths after the warm-up pass)")
        lines.append("| depth | kernel | us | DRAM read MB | DRAM write MB | regs | warps active % | L1 hit % | L2 hit % | threads/inst | issue active % | ALU pipe % | FMA pipe % | warp inst (M) |")
        lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
        for d, ls in enumerate(depths):
            for l in ls:
                lines.append(f"| {d} | {l['name']} | {l['us']:.1f} | {l['read'] / 1e6:.1f} | {l['write'] / 1e6:.1f} | {l['regs']} | {l['warps_active']:.1f} | "
                             f"{l['l1_hit']:.1f} | {l['l2_hit']:.1f} | {l['threads_per_inst']:.1f} | {l['issue']:.1f} | {l['alu']:.1f} | {l['fma']:.1f} | {l['inst'] / 1e6:.1f} |")
        entry = {}
        for kernel in ("k_trace_closest", "k_shade", "k_trace_shadow", "k_trace_enum"):
            per_depth = [sum(l["read"] + l["write"] for l in ls if l["kernel"] == kernel) for ls in depths]
            detail = ", ".join("depth %d (%.1f + %.1f MB)" % (d, sum(l["read"] for l in ls if l["kernel"] == kernel) / 1e6,
                                                               sum(l["write"] for l in ls if l["kernel"] == kernel) / 1e6) for d, ls in enumerate(depths))
            entry[kernel] = {"dram_bytes_per_launch": sum(per_depth) / len(per_depth),
                             "source": f"ncu --set full (profiles/{tag}_ncu_full_{workload}_raw.csv): dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                                       f"stage's launches of one depth, mean of {detail}"}
        traffic[workload] = entry
        lines.append("")
    json.dump(traffic, open(traffic_path, "w"), indent=2)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
