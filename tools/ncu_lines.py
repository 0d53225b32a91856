"""Per CUDA source line: executed warp instructions and stall samples, from the cuda,sass source page of an .ncu-rep.
usage: python tools/ncu_lines.py file.ncu-rep kernel-substring [top-n]"""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# blocks start with "File Path" / "Function Name"
blocks, cur = [], None
for i, l in enumerate(lines):
    if l.startswith('"Function Name"'):
        cur = {"name": l, "rows": []}
        blocks.append(cur)
    elif cur is not None and not l.startswith('"File Path"'):
        cur["rows"].append(l)
# a kernel has one block per source FILE it inlines code from: report the one with the most instructions
best = None
for b in blocks:
    if want not in b["name"]:
        continue
    rd0 = csv.reader(io.StringIO("\n".join(b["rows"])))
    h0 = next(rd0)
    k0 = h0.index("Instructions Executed")
    n0 = 0
    for r in rd0:
        try:
            n0 += int(r[k0])
        except (ValueError, IndexError):
            pass
    if best is None or n0 > best[0]:
        best = (n0, b)
for b in ([best[1]] if best else []):
    rd = csv.reader(io.StringIO("\n".join(b["rows"])))
    hdr = next(rd)
    iline, isrc = 0, 1
    iinst = hdr.index("Instructions Executed")
    ithr = hdr.index("Thread Instructions Executed")
    isamp = hdr.index("# Samples")
    agg = {}
    for r in rd:
        if len(r) <= iinst or not r[iline]:
            continue
        try:
            key = (int(r[iline]), r[isrc].strip())
            agg.setdefault(key, [0, 0, 0])
            agg[key][0] += int(r[iinst]); agg[key][1] += int(r[ithr]); agg[key][2] += int(r[isamp])
        except ValueError:
            pass
    tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
    print(b["name"][:150]); print(f"total warp instr {tot}, samples {tots}")
    for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{ln:5d} {100*v[0]/max(tot,1):5.1f}% inst  {100*v[2]/max(tots,1):5.1f}% stall  thr/inst {v[1]/max(v[0],1):4.1f}  {src[:90]}")
    # optional region summary: extra args "name:lo-hi"
    regs = [a for a in sys.argv[4:] if ":" in a]
    if regs:
        print("regions:")
        for r in regs:
            name, rng = r.split(":"); lo, hi = (int(v) for v in rng.split("-"))
            vi = sum(v[0] for (ln, _), v in agg.items() if lo <= ln <= hi); vs = sum(v[2] for (ln, _), v in agg.items() if lo <= ln <= hi)
            vt = sum(v[1] for (ln, _), v in agg.items() if lo <= ln <= hi)
            print(f"  {name:18s} {100*vi/max(tot,1):5.1f}% inst ({vi/1e6:6.2f}M) {100*vs/max(tots,1):5.1f}% stall  thr/inst {vt/max(vi,1):4.1f}")
    break
