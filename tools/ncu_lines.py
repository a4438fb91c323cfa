"""Per-source-line view of an `ncu --set full --import-source on` capture: joins the SASS page of the report (samples and
executed instructions per instruction) with `nvdisasm -g` line information of the same cubin, in instruction order.
Usage: python tools/ncu_lines.py report.ncu-rep <kernel regex> <cubin> <mangled function substring> [top N]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, cubin, fun = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        break   # next kernel of the report
    body.append(dict(zip(hdr, r)))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith("\t.section") or l.startswith(".section"):
        on = False
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        on = fun in m.group(1)
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
if len(lines) != len(body):
    print(f"warning: {len(lines)} instructions in the cubin vs {len(body)} in the report", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
for ln, b in zip(lines, body):
    agg[ln][0] += int(b["# Samples"])
    agg[ln][1] += int(b["Instructions Executed"])
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"samples {ts}  warp instructions {ti}")
src = {}
for (ln, v) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if ln:
        f = "xsarsea_b200/csrc/" + ln[0]
        if f not in src:
            try:
                src[f] = open(f).read().splitlines()
            except OSError:
                src[f] = []
        if 0 < ln[1] <= len(src[f]):
            text = src[f][ln[1] - 1].strip()[:100]
    print(f"{100 * v[0] / max(ts, 1):5.1f}% samples {100 * v[1] / max(ti, 1):5.1f}% instr  {ln}  {text}")
