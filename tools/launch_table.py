"""One forward of the ncu launch list (gpu__time_duration per launch) as a table, optionally next to a second list.
    python tools/launch_table.py gpurun_out/A_launches.csv [gpurun_out/B_launches.csv] [--first N]
Kernel launches are serialised and cold-cache under ncu: compare shares, not absolutes."""
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi = hdr.index("Grid Size")
    out = []
    for r in rd:
        try:
            out.append((r[ki], float(r[vi].replace(",", "")) / 1000.0, r[gi]))
        except ValueError:
            pass
    return out


def one_forward(L, which=3):
    """71 consecutive launches between two occurrences of the stem kernel (one forward, cyclically shifted)."""
    stems = [i for i, (k, _, _) in enumerate(L) if "stem_umma" in k]
    which = min(which, len(stems) - 2)
    return L[stems[which]:stems[which + 1]]


def short(k):
    k = k.replace("ap::", "").replace("void ", "")
    return k[:46]


if __name__ == "__main__":
    paths = [a for a in sys.argv[1:] if not a.startswith("--")]
    fw = [one_forward(load(p)) for p in paths]
    for j, F in enumerate(fw):
        print(f"== {paths[j]}: {len(F)} launches, {sum(t for _, t, _ in F):.1f} us serialised")
        agg = {}
        for k, t, g in F:
            a = agg.setdefault(short(k), [0, 0.0])
            a[0] += 1
            a[1] += t
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"   {t:9.1f} us  {n:3d}x  {k}")
    if "--all" in sys.argv:
        for j, F in enumerate(fw):
            print(f"-- {paths[j]}")
            for i, (k, t, g) in enumerate(F):
                print(f"{i:3d} {t:8.1f} {g:>14s} {short(k)}")
