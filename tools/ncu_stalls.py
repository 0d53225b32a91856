"""Summarise the source page of an .ncu-rep: total stall samples by reason and the hottest SASS instructions.
usage: python tools/ncu_stalls.py file.ncu-rep [kernel-index] [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the csv holds one block per kernel: a "Kernel Name" line, a header line, then instructions
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = blocks[k]
print(blk[0][:160])
rd = csv.DictReader(io.StringIO("\n".join(blk[1:])))
rows = list(rd)
stall_cols = [c for c in rd.fieldnames if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
alls = sum(tot.values())
print("total samples", alls, " instructions executed (warp)", sum(int(r["Instructions Executed"] or 0) for r in rows))
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {c:28s} {v:8d} {100.0 * v / max(alls, 1):5.1f}%")
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:top]:
    main = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f"{int(r['# Samples']):7d} {100.0 * int(r['# Samples']) / max(alls, 1):5.1f}% {main[6:]:14s} {r['Source'].strip()[:110]}")
