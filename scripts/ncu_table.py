"""Per-kernel summary table (profiles/rNN_ncu_kernels.md) from the raw pages written by scripts/ncu_kernels.sh.
usage: python scripts/ncu_table.py profiles/ncu_r01 > profiles/r01_ncu_kernels.md"""
import csv
import json
import os
import sys

ALG = {"lap": 16, "rhs": 46, "residual": 30, "correct": 24, "restrict": 10, "prolong": 10, "relax": 30}  # SURVEY.md 8(d), nl = 4
UNITS = {"lap": 4096 * 4096 * 4, "rhs": 4096 * 4096 * 4, "residual": 4096 * 4096 * 4, "correct": 4096 * 4096 * 4,
         "restrict": 4096 * 4096 * 4, "prolong": 4096 * 4096 * 4, "relax": 4096 * 4096 * 4}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def col(h, u, v, name, scale=False):
    if name not in h:
        c = [x for x in h if x.endswith(name)]
        if not c:
            return None
        name = c[0]
    i = h.index(name)
    if v[i] == "":
        return None
    x = float(v[i].replace(",", ""))
    return x * SCALE.get(u[i], 1.0) if scale else x


def main(d):
    peak = 6547.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print("# ncu --set full, finest-level launch of every hot-path kernel, 4096^2 x nl=4 (one timestep, scripts/one_step.py)")
    print("# command per kernel: ncu --set full --clock-control none -k regex:<name> --launch-skip <s> --launch-count <c> "
          "python scripts/one_step.py  (scripts/ncu_kernels.sh); table: scripts/ncu_table.py")
    print("# alg B/unit = SURVEY.md 8(d) algorithmic bytes per cell-layer; HBM peak = MEASURED_PEAKS.json %.1f GB/s; "
          "times under ncu are cold-cache, serialised" % peak)
    print()
    print("| kernel | time us | grid | block | regs | dram rd MB | dram wr MB | dram %peak | sm %peak | L1 hit % | L2 hit % | "
          "warps active % | issue active % | fp64 pipe % | alg B/unit | alg GB/s | dram GB/s | dram/alg |")
    print("|" + "---|" * 18)
    for name in ("lap", "rhs", "residual", "correct", "restrict", "prolong", "relax"):
        p = os.path.join(d, name + ".raw.csv")
        if not os.path.exists(p):
            continue
        rows = list(csv.reader(open(p)))
        h, u = rows[0], rows[1]
        for v in rows[2:]:
            t = col(h, u, v, "gpu__time_duration.sum", True)
            rd = col(h, u, v, "dram__bytes_read.sum", True)
            wr = col(h, u, v, "dram__bytes_write.sum", True)
            alg = ALG[name] * UNITS[name]
            sweeps = ""
            if name == "relax":  # algorithmic bytes are per sweep; the launch fuses the cycle's nrelax sweeps
                sweeps = " x nsweeps"
            print("| %s | %.1f | %d | %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d%s | %.0f | %.0f | %.2f |" % (
                v[h.index("Kernel Name")].replace("void ", "").split("(")[0], t * 1e6, col(h, u, v, "launch__grid_size"),
                col(h, u, v, "launch__block_size"), col(h, u, v, "launch__registers_per_thread"), rd / 1e6, wr / 1e6,
                (rd + wr) / t / 1e9 / peak * 100,
                col(h, u, v, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0.,
                col(h, u, v, "l1tex__t_sector_hit_rate.pct") or 0., col(h, u, v, "lts__t_sector_hit_rate.pct") or 0.,
                col(h, u, v, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0.,
                col(h, u, v, "sm__inst_issued.avg.pct_of_peak_sustained_active") or 0.,
                col(h, u, v, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") or 0.,
                ALG[name], sweeps, alg / t / 1e9, (rd + wr) / t / 1e9, (rd + wr) / alg))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/ncu_r01")
