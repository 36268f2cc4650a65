"""Per-kernel aggregate of an ncu launch list (gpu__time_duration.sum, optionally dram bytes): launches, mean us, total ms,
mean DRAM MB read / written per launch.   python tools/launch_summary.py gpurun_out/X_launches.csv [--all]
Launches are serialised and cold-cache under ncu: compare shares, not absolutes."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    only_ours = "--all" not in sys.argv
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(k, {}).setdefault(row["Metric Name"], []).append(v)
    total = sum(sum(d.get("gpu__time_duration.sum", [])) for k, d in agg.items() if ("ap::" in k or not only_ours))
    print(f"# {path}: kernels of libapnetg.so{'' if only_ours else ' and others'}, {total / 1e6:.2f} ms serialised")
    print(f"{'kernel':66s} {'n':>5s} {'mean us':>9s} {'total ms':>9s} {'share':>6s} {'rd MB':>8s} {'wr MB':>8s}")
    for k, d in agg.items():
        if only_ours and "ap::" not in k:
            continue
        t = d.get("gpu__time_duration.sum", [0.0])
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        name = k.replace("void ", "").replace("ap::<unnamed>::", "").replace("ap::", "")[:66]
        print(f"{name:66s} {len(t):5d} {sum(t) / len(t) / 1e3:9.1f} {sum(t) / 1e6:9.2f} {sum(t) / total * 100:5.1f}% "
              f"{(sum(rd) / len(rd) / 1e6 if rd else float('nan')):8.2f} {(sum(wr) / len(wr) / 1e6 if wr else float('nan')):8.2f}")


if __name__ == "__main__":
    main()
