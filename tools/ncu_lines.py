"""Per CUDA-C source line view of an .ncu-rep (needs -lineinfo and --import-source on): samples, executed warp
instructions and the dominant stall reasons, hottest lines first.   python tools/ncu_lines.py rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []
cur_file = None
H = None
seen_kernel = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break
        continue
    if r[0] == "Line No":
        H = r
        continue
    if H is None or len(r) < len(H) or not r[0].isdigit() or r[2] != "-":
        continue
    d = dict(zip(H[4:], r[4:]))
    try:
        smp = int(d["# Samples"])
        ins = int(d["Instructions Executed"])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    lines.append((smp, ins, cur_file, int(r[0]), r[1].strip(), stalls))
tot_s = sum(x[0] for x in lines)
tot_i = sum(x[1] for x in lines)
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
agg = {}
for smp, ins, f, ln, src, st in lines:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100. * v / max(sum(agg.values()), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
for smp, ins, f, ln, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
    s3 = " ".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %-70s | %s" % (100. * smp / tot_s, 100. * ins / tot_i, f, ln, src[:70], s3))
