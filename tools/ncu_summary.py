#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into profiles/<name>.md (+ traffic json).

    python tools/ncu_summary.py gpurun_out/env_step_r01.ncu-rep profiles/r01_env_step [--traffic-json profiles/env_step_traffic.json --envs 65536]

Reads the report here (no GPU needed) with `ncu -i ... --page raw --csv` and `--page source --csv`.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    traffic_json = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    raw = ncu_csv(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    lines = [f"# ncu summary of `{rep}`", "", f"kernels captured: {len(rows)}", ""]
    kn = hdr.index("Kernel Name")
    lines.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " |")
    lines.append("|---|---|" + "---|" * len(rows))
    lines.append("| kernel | | " + " | ".join(r[kn][:60] for r in rows) + " |")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows) + " |")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    traf = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows]
    lines += ["", f"DRAM traffic per launch (read+write): {[round(t / 1e6, 1) for t in traf]} MB", ""]

    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 10:
            cur["rows"].append(r)
    if blocks:
        b = blocks[0]
        h = b["hdr"]
        ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(int(r[ie]) for r in b["rows"])
        tots = max(1, sum(int(r[isamp]) for r in b["rows"]))
        lines += [f"## SASS profile of launch 0 (`{b['name'][:70]}`)", "",
                  f"warp-instructions executed: {tot}  ({len(b['rows'])} SASS instructions)", ""]
        st = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        totst = collections.Counter()
        for r in b["rows"]:
            for i, x in st:
                totst[x] += int(r[i] or 0)
        s = max(1, sum(totst.values()))
        lines += ["| stall reason (all samples) | share |", "|---|---|"]
        lines += [f"| {x} | {100 * c / s:.1f}% |" for x, c in totst.most_common(8)]
        op, ops = collections.Counter(), collections.Counter()
        for r in b["rows"]:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
            o = m.group(2).split(".")[0] if m else "?"
            op[o] += int(r[ie])
            ops[o] += int(r[isamp])
        lines += ["", "| opcode | executed | share | stall samples |", "|---|---|---|---|"]
        lines += [f"| {o} | {c} | {100 * c / tot:.1f}% | {100 * ops[o] / tots:.1f}% |" for o, c in op.most_common(16)]
        lines += ["", "hottest SASS lines by stall samples:", "", "```"]
        for r in sorted(b["rows"], key=lambda r: -int(r[isamp]))[:14]:
            rs = {x: int(r[i] or 0) for i, x in st if int(r[i] or 0) > 0}
            top = ", ".join(f"{k}={v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:2])
            lines.append(f"{r[isamp]:>5} samples  exec {r[ie]:>8}  {r[ia].strip()[:64]:64s} {top}")
        lines.append("```")
    open(outbase + ".md", "w").write("\n".join(lines) + "\n")
    if traffic_json:
        json.dump({"dram_bytes_per_launch": sum(traf) / len(traf),
                   "source": f"{outbase}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(traf)} launches)"},
                  open(traffic_json, "w"))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
