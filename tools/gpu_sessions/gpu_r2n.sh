set -x
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/smoke.txt; cat $O/smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_atari.json
timeout 300 python bench.py --workload atari_conv_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_atari_conv.json
timeout 300 python bench.py --workload atari_conv_e256_b1024_sim50 --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32_atari_conv.json
python - <<PY
import json,glob,re
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
