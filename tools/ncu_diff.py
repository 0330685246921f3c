"""Per-source-line executed-instruction diff of two .ncu-rep captures of the same kernel: python tools/ncu_diff.py A B [top]"""
import collections
import csv
import io
import subprocess
import sys


def per_line(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, H, k = None, None, 0
    d = collections.Counter()
    ops = collections.Counter()
    curline = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Kernel Name":
            k += 1
            if k > 1:
                break
            continue
        if r[0] == "Line No":
            H = r
            continue
        if H is None or len(r) < len(H):
            continue
        if r[0].isdigit() and r[2] == "-":
            curline = (cur, r[1].strip()[:80])
            continue
        if r[2].startswith("0x"):
            try:
                n = int(r[H.index("Instructions Executed")])
            except ValueError:
                continue
            d[curline] += n
            t = r[3].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += n
    return d, ops


a, oa = per_line(sys.argv[1])
b, ob = per_line(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print("total A %d  B %d  diff %d" % (sum(a.values()), sum(b.values()), sum(a.values()) - sum(b.values())))
keys = set(a) | set(b)
for key in sorted(keys, key=lambda k: -abs(a[k] - b[k]))[:top]:
    print("%+9d  A %8d B %8d  %s: %s" % (a[key] - b[key], a[key], b[key], key[0], key[1]))
print("opcodes:", ", ".join("%s %+d" % (o, oa[o] - ob[o]) for o in sorted(set(oa) | set(ob), key=lambda o: -abs(oa[o] - ob[o]))[:25]))
