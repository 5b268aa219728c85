set -x
O=gpurun_out/r2i; mkdir -p $O
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python tools/bench_recurrent.py 2>&1 | grep -E "tc clk|us_per_call" > $O/tc_clk_all.txt
python - <<PY
import re
seen=set()
for ln in open("$O/tc_clk_all.txt"):
    if ln.startswith("tc clk"):
        sig=re.sub(r"-?\d+", "", ln)[:200]
        key=tuple(re.findall(r"k16=\d+ n=\d+", ln))
        if key in seen: continue
        seen.add(key); print(ln.strip())
    else: print(ln.strip()[:300])
PY
