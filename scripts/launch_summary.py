"""Per-kernel share table (profiles/rNN_launches_summary.md) from the ncu launch list of the bench command:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu
usage: python scripts/launch_summary.py gpurun_out/launches.csv [bench.json] > profiles/r01_launches_summary.md"""
import collections
import csv
import json
import sys


def main(path, bench=None):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, gi, vi = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value")
    tot = collections.OrderedDict()
    cnt = collections.Counter()
    fine = 0.0
    nrow = 0
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        nrow += 1
        name = r[ki].replace("void ", "").split("(")[0]
        ms = float(r[vi].replace(",", "")) / 1e6
        tot[name] = tot.get(name, 0.0) + ms
        cnt[name] += 1
        if name.startswith("k_relax_ws") and r[gi].startswith("(129,"):
            fine += ms
    total = sum(tot.values())
    print("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (first %d launches), gpu__time_duration.sum per kernel" % nrow)
    print("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file launches.csv "
          "python bench.py --steps 2 --warmup 3 --no-cpu")
    print("# (times under ncu are cold-cache and serialised: compare SHARES with bench.py's kernel_ms_per_step, not absolutes)")
    print()
    print("| kernel | launches | total ms | share |")
    print("|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("| %s | %d | %.2f | %.1f %% |" % (k, cnt[k], v, 100 * v / total))
    print()
    line = "finest-level k_relax_ws launches (grid 129): %.2f ms = %.1f %% of the listed GPU time" % (fine, 100 * fine / total)
    if bench:
        d = json.load(open(bench))
        line += " (bench.py roofline.share_of_step: %.1f %%)" % (100 * d["roofline"]["share_of_step"])
    print(line)


if __name__ == "__main__":
    main(*sys.argv[1:3])
