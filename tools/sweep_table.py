"""profiles/r2_sweep.md from profiles/r2_sweep_n*.jsonl (tests/sweep.py output)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
data = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", "r2_sweep_n%d.jsonl" % n)
    if os.path.exists(p):
        for line in open(p):
            line = line.strip()
            if line.startswith("{"):
                d = json.loads(line)
                data.setdefault((d["P"], d["F"]), {})[n] = d
print("| Gaussians | F | reference ms (1 GPU) | ours ms (1 GPU) | speed-up | ours views/s, N=1 | N=2 | N=4 | N=8 | N=8 / N=1 |")
print("|---|---|---|---|---|---|---|---|---|---|")
for (P, F), row in sorted(data.items()):
    d1 = row.get(1)
    v = lambda n: ("%.1f" % row[n]["ours_views_per_s"]) if n in row else "-"
    ref = d1 and d1.get("ref_ms")
    sc = ("%.2fx" % (row[8]["ours_views_per_s"] / d1["ours_views_per_s"])) if (8 in row and d1) else "-"
    print("| %d | %d | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        P, F, ("%.1f" % ref) if ref else "- (not run)", ("%.3f" % d1["ours_ms_per_step"]) if d1 else "-",
        ("%.1fx" % d1["speedup_vs_ref_1gpu"]) if (d1 and d1.get("speedup_vs_ref_1gpu")) else "-", v(1), v(2), v(4), v(8), sc))
