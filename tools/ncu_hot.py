"""Top stall lines of one kernel from an .ncu-rep source page (SASS view).
    python tools/ncu_hot.py rep kernel_regex [topN]
"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
# several kernels may be concatenated; take the first block
blocks, cur = [], []
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [ln]
    else:
        cur.append(ln)
if cur: blocks.append(cur)
blk = blocks[int(sys.argv[4]) if len(sys.argv) > 4 else 0]
print(blk[0][:120])
rows = list(csv.reader(blk[1:]))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[1:]:
    try:
        s = int(r[ci["# Samples"]])
    except Exception:
        continue
    tot += s
    data.append((s, r))
data.sort(key=lambda x: -x[0])
print("total samples", tot)
for s, r in data[:top]:
    st = sorted(((int(r[ci[h]]), h[6:]) for h in stalls if r[ci[h]] not in ("", "0")), reverse=True)[:3]
    print(f"{s:7d} {100*s/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {st}")
