"""tools/ncu_summary.py -- text summary of an .ncu-rep (run where ncu is installed; no GPU needed).

Writes the headline metrics of every profiled launch plus the stall breakdown and the hottest SASS
lines of the first one.  Usage: python tools/ncu_summary.py report.ncu-rep > profiles/NAME.txt
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of %s" % rep.split("/")[-1])
    for n, row in enumerate(raw[2:]):
        print("\n## launch %d: %s" % (n, row[idx["Kernel Name"]][:100]))
        for k in KEYS:
            if k in idx:
                print("%-70s %s %s" % (k, row[idx[k]], units[idx[k]]))
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(row[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (b, a) for a, b in sorted(stalls, reverse=True)[:7]))
    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        i2 = {h: i for i, h in enumerate(h2)}
        rows = [r for r in src[2:] if len(r) >= len(h2)]

        def f(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        if "# Samples" in i2:
            tot = sum(f(r[i2["# Samples"]]) for r in rows) or 1.0
            print("\n## hottest SASS lines of the first kernel (share of stall samples; long-scoreboard / barrier / short-scoreboard)")
            for r in sorted(rows, key=lambda r: -f(r[i2["# Samples"]]))[:14]:
                print("%5.1f%%  lsb=%-6.0f bar=%-6.0f ssb=%-6.0f %s" % (100 * f(r[i2["# Samples"]]) / tot, f(r[i2["stall_long_sb"]]),
                      f(r[i2["stall_barrier"]]), f(r[i2["stall_short_sb"]]), r[i2["Source"]].strip()[:90]))
            ops = {}
            for r in rows:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i2["Source"]])
                if m:
                    ops[m.group(2)] = ops.get(m.group(2), 0) + f(r[i2["Instructions Executed"]])
            print("\n## executed warp instructions by opcode (millions): " +
                  ", ".join("%s %.0f" % (k, v / 1e6) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))


if __name__ == "__main__":
    main()
