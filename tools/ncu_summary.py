#!/usr/bin/env python
"""Summarise an ncu report of bench.py's dominant kernel into profiles/ (run in the authoring
container: `python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <tag> [workload]`).

Writes profiles/<tag>_costvol_metrics.csv (selected raw metrics, one column per captured launch),
profiles/<tag>_costvol_opcodes.txt (executed warp instructions per 32 samples by opcode, from the
SASS source page) and updates profiles/traffic.json (DRAM bytes per launch, read by bench.py)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "sintel_1024x436_L51_F2_cubic"
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 2 * 446464 * 51 / 32.0      # one launch = both neighbour frames
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit = rows[0], rows[1]
    with open(os.path.join(out, "%s_costvol_metrics.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        name_i = hdr.index("Kernel Name")
        w.writerow(["kernel", ""] + [r[name_i] for r in rows[2:]])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, unit[i]] + [r[i] for r in rows[2:]])
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = sum(float(r[rd]) * scale[unit[rd]] + float(r[wr]) * scale[unit[wr]] for r in rows[2:]) / (len(rows) - 2)
    tj = os.path.join(out, "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d[workload] = traffic
    d["_source"] = d.get("_source", {})
    d["_source"][workload] = "%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches)" % (tag, len(rows) - 2)
    d["_capture"] = "profiles/%s_costvol_metrics.csv (ncu --set full --clock-control none, round 2 final kernel: one launch = both neighbour frames)" % tag
    d["_captures"] = d.get("_captures", {})
    d["_captures"][workload] = d["_capture"]
    ie = hdr.index("smsp__inst_executed.sum")
    d["_inst_executed"] = d.get("_inst_executed", {})
    d["_inst_executed"][workload] = sum(float(r[ie].replace(",", "")) for r in rows[2:]) / (len(rows) - 2)
    json.dump(d, open(tj, "w"), indent=1, sort_keys=True)

    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sec = None
    for r in srows:
        if r and r[0] == "Kernel Name":
            if sec is not None:
                break
            sec = {"name": r[1], "rows": []}
        elif r and r[0] == "Address":
            sec["hdr"] = r
        elif sec is not None and r:
            sec["rows"].append(r)
    ia, isrc = sec["hdr"].index("Instructions Executed"), sec["hdr"].index("Source")
    hist = collections.Counter()
    for r in sec["rows"]:
        op = r[isrc].split()
        o = op[1] if op[0].startswith("@") else op[0]
        hist[o.split(".")[0]] += int(r[ia])
    tot = sum(hist.values())
    with open(os.path.join(out, "%s_costvol_opcodes.txt" % tag), "w") as f:
        f.write("%s\nexecuted warp instructions per warp-sample (32 pixel*labels): %.1f\n" % (sec["name"], tot / units))
        for k, v in hist.most_common():
            f.write("%-12s %8.2f %5.1f%%\n" % (k, v / units, 100.0 * v / tot))
    # the hot loop: the SASS lines executed at least half as often as the most executed one, in address order
    mx = max(int(r[ia]) for r in sec["rows"])
    with open(os.path.join(out, "%s_costvol_hotloop.sass" % tag), "w") as f:
        f.write("// %s\n// SASS of the label loop (lines executed >= 50%% as often as the hottest line), from the ncu source page;\n"
                "// columns: address, warp-level executions, instruction\n" % sec["name"])
        iaddr = sec["hdr"].index("Address")
        for r in sec["rows"]:
            if int(r[ia]) * 2 >= mx:
                f.write("%s %10d  %s\n" % (r[iaddr], int(r[ia]), r[isrc]))
    print("traffic per launch: %.1f MB" % (traffic / 1e6))


if __name__ == "__main__":
    main()
