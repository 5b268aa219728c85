set -x
O=gpurun_out/r2m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
for v in "" _cold _u2; do
  for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32; do
    MZ_LIB_PATH=$PWD/muax_b200/libmzsearch$v.so timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32${v}_$w.json
  done
done
for S in 16 32; do
  MZ_TC_STAGE_KB=$S timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_stage${S}_atari.json
done
timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32_atari.json
python - <<PY
import json,glob,re
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
