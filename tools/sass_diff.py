"""No GPU needed: which kernels of two builds of libsmesh_b200.so differ in their SASS (addresses and encodings stripped)?
usage: python tools/sass_diff.py old.so new.so   - used to show that an opt-in variant leaves the shipped kernels untouched."""
import sys, re, subprocess, hashlib
def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, name, buf = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name: d[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            buf.append(re.sub(r"\s+", " ", re.sub(r"/\*.*?\*/", "", line)))
    if name: d[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return d
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
norm = lambda n: n.replace("ELb0EEEvNS0_11ScatterArgsE", "EEEvNS0_11ScatterArgsE")
b2 = {norm(k): v for k, v in b.items()}
print(len(a), "functions before,", len(b), "after")
for k in sorted(set(a) | set(b2)):
    if a.get(k) != b2.get(k):
        print("DIFF" if k in a and k in b2 else ("NEW " if k in b2 else "GONE"), k)
