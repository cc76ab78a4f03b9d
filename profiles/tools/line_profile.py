#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line / enclosing function.

    ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all libdvsloss.so ; nvdisasm -g -c dvs_fused.sm_100a.cubin > all.sass
    python profiles/tools/line_profile.py sass.csv all.sass '<mangled kernel name>' [core.cuh]

The ncu CLI does not print per-line metrics for the CUDA view, so instructions are matched by order with
nvdisasm's line annotations (same cubin).
"""
import csv
import re
import sys
from collections import defaultdict


def parse_sass(path, kernel):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
    cur = ("?", 0)
    out = []
    for l in lines[start + 1:]:
        if l.startswith("//---") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", l):
            op = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip().rstrip(";").strip()
            out.append((cur, op))
    return out


def functions_of(path):
    """line -> enclosing function name (crude: lines matching 'DVS_HD ... name(' or '__global__')."""
    names = {}
    cur = "?"
    for n, l in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:DVS_HD|__global__|static|inline)\b.*?(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            cur = m.group(1)
        names[n] = cur
    return names


def main():
    sass_csv, all_sass, kernel = sys.argv[1:4]
    core = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ann = parse_sass(all_sass, kernel)
    ci = {h: i for i, h in enumerate(hdr)}
    n = min(len(body), len(ann))
    if len(body) != len(ann):
        print(f"# warning: {len(body)} profiled instructions vs {len(ann)} disassembled")
    fn = functions_of(core) if core else {}
    by_line, by_fn, by_op = defaultdict(lambda: [0, 0]), defaultdict(lambda: [0, 0]), defaultdict(int)
    tot_i = tot_s = 0
    for r, ((f, ln), op) in zip(body[:n], ann[:n]):
        ie = int(r[ci["Instructions Executed"]] or 0)
        ss = int(r[ci["# Samples"]] or 0)
        by_line[(f, ln)][0] += ie
        by_line[(f, ln)][1] += ss
        key = fn.get(ln, "?") if f.endswith("core.cuh") else f
        by_fn[key][0] += ie
        by_fn[key][1] += ss
        by_op[op.split()[0].split(".")[0] if not op.startswith("@") else op.split()[1].split(".")[0]] += ie
        tot_i += ie
        tot_s += ss
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    print("\n== by function (inst %, stall-sample %)")
    for k, (a, b) in sorted(by_fn.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:28s} {100 * a / tot_i:6.2f}% {100 * b / max(tot_s, 1):6.2f}%")
    print("\n== top lines")
    for (f, ln), (a, b) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{f}:{ln:<5d} {100 * a / tot_i:6.2f}% {100 * b / max(tot_s, 1):6.2f}%  {fn.get(ln, '') if f.endswith('core.cuh') else ''}")
    print("\n== top lines by stall samples (line, samples %, dominant stall reasons)")
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_line_r = defaultdict(lambda: defaultdict(int))
    for r, ((f, ln), op) in zip(body[:n], ann[:n]):
        for h in reasons:
            v = int(r[ci[h]] or 0)
            if v:
                by_line_r[(f, ln)][h] += v
    for (f, ln), (a, b) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:30]:
        top = sorted(by_line_r[(f, ln)].items(), key=lambda kv: -kv[1])[:3]
        print(f"{f}:{ln:<5d} {100 * b / max(tot_s, 1):6.2f}%  {fn.get(ln, '') if f.endswith('core.cuh') else '':18s} " +
              " ".join(f"{k[6:]}={v}" for k, v in top))
    print("\n== by opcode")
    for k, a in sorted(by_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{k:12s} {100 * a / tot_i:6.2f}%")


if __name__ == "__main__":
    main()
