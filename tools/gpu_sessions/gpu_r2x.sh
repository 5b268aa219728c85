set -x
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_cached.json
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_nocache.so python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_nocache.json
python bench.py --steps 20 --warmup 5 --workload cartpole_mlp_e8_b1024_sim50 2>&1 | tail -1 > $O/bench_cached_b1024.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.4f kernel_ms %.4f value %.1fM e2e %.1fM"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
