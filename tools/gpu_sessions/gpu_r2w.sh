set -x
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_$w.json
done
python bench.py 2>&1 | tail -1 > $O/bench_n1.json
B=4096 NS=50 python tools/act_overhead.py 2>&1 | head -1 > $O/act_overhead.txt
MZ_TC_DUMP=1 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep -E "tc clk|bs clk|tc program|^  s[0-9]" | tail -24 > $O/clk.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
cat $O/act_overhead.txt
