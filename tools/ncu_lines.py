"""Per-source-line stall samples of one launch in an .ncu-rep (cuda,sass correlated view).
usage: ncu_lines.py REPORT LAUNCH_INDEX [TOP]"""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
agg = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] == "":
        continue
    try:
        line = int(r[0]); samples = int(r[4]); inst = int(r[7])
    except ValueError:
        continue
    k = (cur_file, line)
    a = agg.setdefault(k, [0, 0, r[1].strip()[:90]])
    a[0] += samples; a[1] += inst
tot = sum(a[0] for a in agg.values()) or 1
print(f"total samples {tot}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}%  {a[1]:>10}  {f}:{l}  {a[2]}")
