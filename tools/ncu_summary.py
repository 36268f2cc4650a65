"""Summarise an .ncu-rep (ncu --set full) as one line per kernel launch: duration, DRAM bytes, DRAM %, tensor %, L2 %, occupancy.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x.txt]
"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__cycles_active.avg", "cyc_act"), ("sm__cycles_elapsed.avg", "cyc"),
        ("smsp__inst_executed.sum", "inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes")]


def to_unit(val, unit, want):
    v = float(val.replace(",", ""))
    if want == "us":
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    if want.endswith("MB"):
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    return v


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# id kernel " + " ".join(n for _, n in WANT))
    for r in rows[2:]:
        vals = []
        for m, n in WANT:
            if m in col and r[col[m]] != "":
                v = to_unit(r[col[m]], units[col[m]], n)
                vals.append(f"{n}={v:.1f}" if n not in ("regs", "grid", "inst", "cyc", "cyc_act") else f"{n}={v:.0f}")
        print(r[col["ID"]], r[col["Kernel Name"]][:44], " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
