"""gpurun_out/r2_parity.jsonl (appended to by the GPU tests, tests/util_neus.py:log_parity) -> profiles/r2_parity.json:
the numbers actually measured on the B200, one record per (test, case, engine, ...) -- the last run wins."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "r2_parity.jsonl")
dst = os.path.join(ROOT, "profiles", "r2_parity.json")
recs = {}
KEYS = ("test", "case", "golden", "engine", "R", "H", "P", "bg", "mode", "add_no_texture", "texture_cast_light")
if os.path.exists(dst):          # every gpurun call starts with an empty gpurun_out/: accumulate across calls
    for r in json.load(open(dst)):
        recs[tuple((k, r[k]) for k in KEYS if k in r)] = r
for line in open(src):
    line = line.strip()
    if not line:
        continue
    r = json.loads(line)
    key = tuple((k, r[k]) for k in KEYS if k in r)
    recs[key] = r
out = sorted(recs.values(), key=lambda r: json.dumps([r.get(k) for k in ("test", "case", "golden", "engine")], default=str))
json.dump(out, open(dst, "w"), indent=1)
print(f"{len(out)} records -> {dst}")
