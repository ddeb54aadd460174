"""Warp-stall samples of one kernel aggregated per CUDA source line (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      "regex:" + pat, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, out, stalls = None, None, [], {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[2] == "-":
        try:
            n = int(r[6])
        except ValueError:
            continue
        best = ("", 0)
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                v = int(r[i] or 0)
                stalls[h] = stalls.get(h, 0) + v
                if v > best[1]:
                    best = (h, v)
        out.append((n, cur.split("/")[-1], int(r[0]), r[1].strip()[:95], int(r[7]), best[0][6:]))
tot = sum(o[0] for o in out)
print("samples", tot, "warp-inst", sum(o[4] for o in out))
for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:7]:
    print(f"  {k:26s} {100 * v / tot:5.1f}%")
for n, f, l, s, ie, st in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100 * n / tot:5.1f}% {f}:{l:<4d} inst={ie:9d} {st:14s} {s}")
