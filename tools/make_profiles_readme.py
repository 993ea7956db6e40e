"""Rebuild the table at the top of profiles/README.md from the committed bench JSON lines."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*.json"))):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    g = d.get("roofline_gather") or {}
    ev = d.get("eval") or {}
    fc = d.get("eval_full_catalogue") or {}
    cb = d.get("cpu_baseline") or {}
    rows.append((os.path.basename(f), d["dtype"], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"],
                 g.get("achieved"), ev.get("value"), fc.get("value"), cb.get("value"), d["config"].get("table_update", "")))
print("| file | dtype | GPUs | train seq/s | ms/step | e2e seq/s | emb-gather GB/s | eval users/s | full-catalogue users/s | CPU port seq/s | table |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
fmt = lambda v, nd=0: "" if v is None else (f"{v:,.{nd}f}")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {fmt(r[3])} | {fmt(r[4], 3)} | {fmt(r[5])} | {fmt(r[6])} | {fmt(r[7])} | {fmt(r[8])} | {fmt(r[9], 1)} | {r[10]} |")
