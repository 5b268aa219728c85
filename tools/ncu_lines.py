#!/usr/bin/env python
"""Correlates an ncu SASS-level source page with CUDA source lines using nvdisasm line info.

usage: tools/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top N]
Needs: ncu, cuobjdump, nvdisasm on PATH; muax_b200/libmzsearch.so built with -lineinfo from the same sources.
Prints, per (file:line) with inlining context collapsed to the innermost frame, the share of executed
warp-instructions and of stall samples.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kname = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "muax_b200", "libmzsearch.so")], cwd=tmp,
                   check=True, capture_output=True)
    dis = ""
    for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):  # one per translation unit
        out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout
        if kname in out:
            dis = out
            break
    # split per function
    lines_by_off = {}
    cur_fn, cur_line, in_fn = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            in_fn = kname in cur_fn
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines_by_off[int(m.group(1), 16)] = (cur_line, m.group(2))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # one section per profiled kernel: a "Kernel Name" row, a header row starting with "Address", then the body
    want = sys.argv[4] if len(sys.argv) > 4 else kname.split("ILi")[0].split("kernel")[0]
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next((i for i in starts if want in rows[i][1]), starts[0])
    end = next((i for i in starts if i > sec), len(rows))
    hdr_i = next(i for i in range(sec, end) if rows[i] and rows[i][0] == "Address")
    hdr = rows[hdr_i]
    body = [dict(zip(hdr, r)) for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
    base = int(body[0]["Address"], 16)
    agg, tot_i, tot_s = {}, 0.0, 0.0
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    for r in body:
        off = int(r["Address"], 16) - base
        line = lines_by_off.get(off, (None, ""))[0] or ("?", 0)
        inst = float(r["Instructions Executed"] or 0)
        samp = float(r["# Samples"] or 0)
        a = agg.setdefault(line, {"inst": 0.0, "samp": 0.0, "stalls": {}})
        a["inst"] += inst
        a["samp"] += samp
        for c in stall_cols:
            v = float(r[c] or 0)
            if v:
                a["stalls"][c] = a["stalls"].get(c, 0) + v
        tot_i += inst
        tot_s += samp
    print(f"total warp-instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
        if f not in srcs:
            for cand in (os.path.join(ROOT, "muax_b200", "csrc", f), os.path.join(ROOT, "include", f)):
                if os.path.exists(cand):
                    srcs[f] = open(cand).read().splitlines()
        text = srcs.get(f, [""] * (l + 1))[l - 1].strip() if f in srcs and 0 < l <= len(srcs[f]) else ""
        st = ",".join(f"{k[6:]}:{100 * v / max(a['samp'], 1):.0f}" for k, v in
                      sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"{f}:{l:<4d} inst {100 * a['inst'] / tot_i:5.1f}%  samples {100 * a['samp'] / tot_s:5.1f}%  [{st}]  {text[:90]}")


if __name__ == "__main__":
    main()
