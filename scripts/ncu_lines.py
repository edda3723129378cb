"""Per-source-line executed warp instructions and stall samples of one kernel from an .ncu-rep
(captured with --import-source on; compiled with -lineinfo).  Development aid."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur_file = ""
lines = {}
total = 0
samples = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iex = hdr.index("Instructions Executed")
        ismp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0]:   # a source line row (aggregated)
        try:
            key = (cur_file, int(r[0]))
        except ValueError:
            continue
        try:   # (source text may contain unescaped quotes: index from the end of the row)
            ex, sm = int(r[iex - len(hdr)] or 0), int(r[ismp - len(hdr)] or 0)
        except ValueError:
            continue
        e = lines.setdefault(key, [0, 0, r[1].strip()[:90]])
        e[0] += ex
        e[1] += sm
        total += ex
        samples += sm
print(f"total warp instructions {total:,}  samples {samples:,}")
for key, (ex, sm, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]:12s}{key[1]:5d} {100.0 * ex / max(total, 1):6.2f}% inst {100.0 * sm / max(samples, 1):6.2f}% smp  {src}")
