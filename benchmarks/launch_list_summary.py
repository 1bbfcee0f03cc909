"""Markdown summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file list.csv <command>`): launches and total time per kernel, share of the listed time.
Usage: python benchmarks/launch_list_summary.py list.csv [title] > profiles/x.md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        rows.append(r)
    if hdr is None:
        print("no launches found")
        return
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) <= v:
            continue
        try:
            t = float(r[v].replace(",", ""))
        except ValueError:
            continue
        unit = r[u]
        t_us = t / 1000.0 if unit in ("ns", "nsecond") else (t * 1000.0 if unit in ("ms", "msecond") else t)
        name = r[k].split("(")[0]
        agg[name] += t_us
        cnt[name] += 1
    tot = sum(agg.values())
    ours = sum(t for n, t in agg.items() if "ax::" in n or n.startswith("k_") or "k_tc" in n or "k_narrow" in n)
    print(f"# ncu launch list: {title}\n")
    print(f"{sum(cnt.values())} launches, {tot / 1000.0:.3f} ms listed (cold-cache, serialised); kernels of "
          f"libai4xde_b200.so: {100 * ours / max(tot, 1e-9):.1f} % of the listed time\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, t in agg.most_common(40):
        print(f"| `{n[:110]}` | {cnt[n]} | {t:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
