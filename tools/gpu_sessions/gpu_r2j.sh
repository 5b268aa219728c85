set -x
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python tools/bench_recurrent.py 2>&1 | grep -E "tc clk|us_per_call" > $O/tc_clk_all.txt
python - <<PY
import re
seen=set()
for ln in open("$O/tc_clk_all.txt"):
    if ln.startswith("tc clk"):
        key=tuple(re.findall(r"k16=\d+ n=\d+", ln))
        if key in seen: continue
        seen.add(key); print(ln.strip())
PY
timeout 300 python tools/bench_recurrent.py > $O/recurrent_micro.txt 2>&1; cat $O/recurrent_micro.txt
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_$w.json
done
for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32_$w.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)), d.get("details",{}).get("mean_path_depth"))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
