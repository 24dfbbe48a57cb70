#!/usr/bin/env python
"""scripts/hot_lines.py -- warp-stall samples of one kernel of an .ncu-rep, aggregated by CUDA source line.
Joins `ncu --page source --csv` (SASS rows with sample counts) with `nvdisasm -g` line annotations of the same cubin
by instruction order.  Usage: python scripts/hot_lines.py REP.ncu-rep KERNEL_REGEX MANGLED_SUBSTR [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = Path(__file__).resolve().parents[1] / "msckf_mono_b200" / "libmsckf_b200.so"
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(so)], cwd=tmp, check=True, capture_output=True)
cub = max(tmp.glob("*.cubin"), key=lambda p: p.stat().st_size)
sass = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout.splitlines()
# instruction -> (file, line) for the wanted function(s): the kernel itself plus any non-inlined callee named in argv
lines_of = []
cur_file, cur_line, active = None, None, False
for ln in sass:
    if ln.startswith(".text."):
        active = mangled in ln
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = Path(m.group(1)).name, int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append((cur_file, cur_line))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
si, ie = H.index("# Samples"), H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[hdr + 1:]
nxt = next((i for i, r in enumerate(body) if r and r[0] in ("Address", "Kernel Name")), len(body))
body = body[:nxt]  # (a second launch of the same kernel follows)
print(f"SASS rows: ncu {len(body)}  nvdisasm {len(lines_of)}")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for k, r in enumerate(body):
    if len(r) <= max(si, ie):
        continue
    key = lines_of[k] if k < len(lines_of) else ("?", 0)
    a = agg[key]
    a[0] += int(r[si] or 0)
    a[1] += int(r[ie] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[2][H[c][6:]] += v
tot = sum(a[0] for a in agg.values())
src_cache = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src_cache:
        cand = list((Path(__file__).resolve().parents[1] / "msckf_mono_b200" / "csrc").glob(f or "?"))
        src_cache[f] = cand[0].read_text().splitlines() if cand else []
    text = src_cache[f][l - 1].strip()[:110] if src_cache[f] and 0 < l <= len(src_cache[f]) else ""
    st = ", ".join(f"{k} {v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * a[0] / max(tot, 1):5.1f}%  inst {a[1]:8d}  {f}:{l:<5d} [{st}]  {text}")
