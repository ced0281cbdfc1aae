#!/usr/bin/env python
"""Summarises an `ncu --set full` report (.ncu-rep) into the text/JSON kept under profiles/.

usage: python tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/r1_xxx [warps_per_launch]
writes <out>.txt (key metrics, pipe utilisation, stall reasons, executed-opcode mix per warp)
and prints the dram bytes per launch that bench.py reports as roofline.traffic.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = raw[0], raw[1]
    lines = []
    summary = {}
    for li, row in enumerate(raw[2:]):
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        lines.append("== launch %d: %s  grid %s block %s" % (li, d.get("Kernel Name", "?"), d.get("Grid Size", "?"), d.get("Block Size", "?")))
        for k in KEYS:
            if k in d:
                lines.append("  %-78s %s %s" % (k, d[k], u.get(k, "")))
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.05:
                        lines.append("  stall %-72s %s" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[k]))
                except ValueError:
                    pass
        if li == 0:
            def val(k):
                v = float(d[k].replace(",", ""))
                unit = u.get(k, "")
                return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(unit, 1)
            try:
                summary["dram_bytes_per_launch"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                summary["kernel"] = d.get("Kernel Name")
                summary["inst_executed"] = float(d["smsp__inst_executed.sum"].replace(",", ""))
            except Exception as e:  # noqa
                summary["error"] = str(e)
    # opcode mix of the first kernel from the source page
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    try:
        h = src[1]
        iE, iS = h.index("Instructions Executed"), h.index("Source")
        ops = collections.Counter()
        tot = 0
        for r in src[2:]:
            if not r or r[0].startswith("Kernel Name") or r[0] == "Address":
                break
            try:
                n = int(r[iE])
            except ValueError:
                continue
            m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[iS].strip())
            ops[m.group(2) if m else "?"] += n
            tot += n
        warps = float(sys.argv[3]) if len(sys.argv) > 3 else None
        lines.append("== executed SASS opcode mix of launch 0 (%d warp-instructions%s)" % (tot, (", %.1f per warp" % (tot / warps)) if warps else ""))
        for op, n in ops.most_common(40):
            lines.append("  %-10s %12d  %5.1f%%%s" % (op, n, 100.0 * n / tot, ("  %7.1f/warp" % (n / warps)) if warps else ""))
        summary["inst_per_warp"] = tot / warps if warps else None
    except Exception as e:  # noqa
        lines.append("(source page unavailable: %s)" % e)
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
