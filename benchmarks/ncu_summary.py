"""Markdown summary of an `ncu --set full` report: per kernel launch the duration, occupancy, issue / pipe
utilisation, the warp-stall breakdown (average warps stalled per issued instruction), DRAM bytes, and the top
stalled SASS opcodes of the source page.  Usage: python benchmarks/ncu_summary.py report.ncu-rep > profiles/x.md"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU / MUFU pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (%)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "no_instruction", "wait", "not_selected", "math_pipe_throttle",
          "mio_throttle", "lg_throttle", "barrier", "dispatch_stall", "branch_resolving", "membar", "sleeping"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none: `{rep.split('/')[-1]}`\n")
    for r in data:
        print(f"## `{r[col['Kernel Name']]}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in col and r[col[key]] not in ("", "n/a"):
                print(f"| {label} | {r[col[key]]} {units[col[key]]} |")
        print("\nWarp stalls (average warps stalled per issued instruction):\n")
        print("| reason | warps |\n|---|---|")
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in col:
                try:
                    v = float(r[col[key]])
                except ValueError:
                    continue
                if v >= 0.01:
                    print(f"| {s} | {v:.2f} |")
        print()
    src = page(rep, "source")
    # the source page lists one kernel after another: 'Kernel Name' row, header row, instruction rows
    kernel, agg, ex = None, None, None

    def flush():
        if kernel is None or not agg:
            return
        tot, totex = sum(agg.values()), sum(ex.values())
        print(f"### stall samples by SASS opcode: `{kernel[:90]}`\n")
        print("| opcode | stall samples | share | executed (warp instructions) | share of executed |\n|---|---|---|---|---|")
        for m, v in agg.most_common(14):
            print(f"| {m} | {v} | {100 * v / max(tot, 1):.1f} % | {ex[m]} | {100 * ex[m] / max(totex, 1):.1f} % |")
        print()

    for r in src:
        if r and r[0] == "Kernel Name":
            flush()
            kernel, agg, ex = r[1], collections.Counter(), collections.Counter()
            continue
        if not r or r[0] == "Address" or kernel is None or len(r) < 6:
            continue
        m = re.sub(r"^\s*(@!?U?P\d+\s+)?", "", r[1]).split()
        if not m:
            continue
        try:
            agg[m[0]] += int(r[4])
            ex[m[0]] += int(r[5])
        except ValueError:
            pass
    flush()


if __name__ == "__main__":
    main()
