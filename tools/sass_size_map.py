#!/usr/bin/env python
"""SASS bytes per source line of one kernel: where the code of a kernel comes from.

usage: tools/sass_size_map.py <kernel mangled-name substring> [library.so | file.cubin] [top N]
Needs cuobjdump and nvdisasm on PATH and a library built with -lineinfo.  Every instruction is attributed to the
OUTERMOST frame of its inline chain (the line of the kernel body that caused it) and, in a second table, to the
innermost frame.  This is the map that showed the tree-warp kernel's instruction-fetch bound: loops
with run-time trip counts unrolled x4 by nvcc accounted for most of the 75 KB one simulation walked through.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_listing(lib, kname):
    if lib.endswith(".cubin"):
        cubins = [lib]
    else:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
        cubins = sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin"))
    for cubin in cubins:
        out = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
        lines = out.splitlines()
        starts = [i for i, ln in enumerate(lines) if re.match(r"\s*\.section\s+\.text\.", ln)]
        for n, i in enumerate(starts):
            if kname in lines[i]:
                end = starts[n + 1] if n + 1 < len(starts) else len(lines)
                return lines[i].split(",")[0].split(".text.")[-1], lines[i:end]
    raise SystemExit(f"no kernel matching {kname!r} in {lib}")


def main():
    kname = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "muax_b200", "libmzsearch.so")
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    name, lines = kernel_listing(lib, kname)
    outer, inner = collections.Counter(), collections.Counter()
    chain, in_annot, total = [], False, 0
    for ln in lines:
        if "//## File" in ln:
            frames = [(os.path.basename(f), int(n)) for f, n in re.findall(r'"([^"]+)", line (\d+)', ln)]
            chain = (chain + frames) if in_annot else frames
            in_annot = True
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            in_annot = False
            total += 16
            if chain:
                outer[chain[-1]] += 16
                inner[chain[0]] += 16
    print(f"{name}: {total} bytes of SASS ({total // 16} instructions)")
    for title, table in (("outermost frame (line of the kernel / function body)", outer), ("innermost frame", inner)):
        print(f"-- by {title}")
        for (f, l), b in table.most_common(top):
            text = ""
            for cand in (os.path.join(ROOT, "muax_b200", "csrc", f), os.path.join(ROOT, "include", f)):
                if os.path.exists(cand):
                    src = open(cand).read().splitlines()
                    text = src[l - 1].strip()[:80] if 0 < l <= len(src) else ""
            print(f"  {f}:{l:<5d} {b:7d} B  {text}")


if __name__ == "__main__":
    main()
