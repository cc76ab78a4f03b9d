#!/usr/bin/env python
"""Assemble profiles/r02_fused_pair_final.txt, profiles/r02_launches.csv and profiles/traffic.json from what
`bash profiles/tools/final_pass.sh` (run on the GPU box through gpurun) left under gpurun_out/.  Runs in the build container (needs ncu
for reading the report, no GPU)."""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "gpurun_out")
rep = os.path.join(G, "prof_r2f.ncu-rep")
out = []
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out.append("# Round 2 final: fused_pair_kernel<true,0> (two-source tile kernel), BASELINE configs[1] (640x480, batch 16, 2 sources, 4 scales)")
out.append(f"# built from commit {commit} (+ working tree); gpurun 'bash profiles/tools/final_pass.sh' on one B200; only the bench line was timed outside a profiler")
out.append("\n## bench.py (default flags, same build, not under ncu)")
out.append(open(os.path.join(G, "r2f_bench.log")).read().strip().splitlines()[-1])
out.append("\n## pytest -m gpu (same call)")
out.append(open(os.path.join(G, "r2f_pytest.log")).read().strip().splitlines()[-1])

# launch list
shutil.copy(os.path.join(G, "r2f_launches.csv"), os.path.join(ROOT, "profiles", "r02_launches.csv"))
rows = list(csv.reader(open(os.path.join(G, "r2f_launches.csv"))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
ci = {k: i for i, k in enumerate(h)}
d = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    n = r[ci["Kernel Name"]][:70]
    d[n][0] += 1
    d[n][1] += v
tot = sum(v[1] for v in d.values())
out.append("\n## ncu --metrics gpu__time_duration.sum launch list of 'bench.py --steps 2 --warmup 1 --no-cpu --no-eager --no-train --no-train-big'"
           " (profiles/r02_launches.csv), summed per kernel")
for n, (k, v) in sorted(d.items(), key=lambda x: -x[1][1])[:8]:
    out.append(f"{n:72s} {k:4d} launches {v / 1e3:10.1f} us {100 * v / tot:5.1f}%")
mine = {n: v for n, v in d.items() if "dvs::" in n}
tile = sum(v[1] for n, v in mine.items() if "fused_pair" in n)
out.append(f"(the list covers the device-resident leg, the e2e pipeline legs and their warm-ups; the tile kernel is "
           f"{100 * tile / sum(v[1] for v in mine.values()):.1f} % of the product kernels' time; live, bench.py has kernel_ms / ms_per_step "
           f"with the L2 flush outside and the launch gaps inside the step)")

# key metrics
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hh, uu, vv = r[0], r[1], r[2]
keys = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "tools", "ncu_keys.py"), rep], capture_output=True, text=True).stdout
out.append("\n## ncu --set full --clock-control none --import-source on, 4th launch of the kernel (gpurun_out/prof_r2f.ncu-rep)")
out.append(keys.rstrip())
val = {k: (x, u) for k, u, x in zip(hh, uu, vv)}
def to_bytes(k):
    x, u = val[k]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(x.replace(",", "")) * mult
traffic = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
json.dump({"fused_tile_kernel_dram_bytes": traffic,
           "source": "profiles/r02_fused_pair_final.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch of fused_pair_kernel<true,0>, "
                     "ncu --set full, config 2"}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))

# per-phase split + opcode mix
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tmp = "/tmp/_sass_final.csv"
open(tmp, "w").write(sass)
seg = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "tools", "segment_profile.py"), tmp], capture_output=True, text=True).stdout
out.append("\n## per-phase split by block barriers (profiles/tools/segment_profile.py)")
out.append("# seg 0 tile load + constants, 1 identity terms + edge weights, 3 warp (gather), 4 statistics, 5 gradient, 6-8 up-sample adjoint + block reductions")
out.append(seg.rstrip())
rows = list(csv.reader(sass.splitlines()))
hi = next(i for i, r_ in enumerate(rows) if r_ and r_[0] == "Address")
hdr = rows[hi]
ci = {k: i for i, k in enumerate(hdr)}
c = Counter()
total = 0
for r_ in rows[hi + 1:]:
    if len(r_) != len(hdr):
        continue
    src = r_[ci["Source"]]
    ie = int(r_[ci["Instructions Executed"]] or 0)
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    c[op.split(".")[0]] += ie
    total += ie
out.append(f"\nopcode mix (warp-instructions executed, {total} total):")
out.append(", ".join(f"{k} {100 * v / total:.1f}%" for k, v in c.most_common(24)))
open(os.path.join(ROOT, "profiles", "r02_fused_pair_final.txt"), "w").write("\n".join(out) + "\n")
print("traffic", traffic, "instructions", total)
