set -x
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stochastic.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest_tc.txt
cat $O/pytest_tc.txt
for S in 1 0; do
  for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200; do
    MZ_TC_SPLIT_BACKUP=$S timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_split${S}_$w.json
  done
done
timeout 300 python bench.py --workload atari_conv_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_atari_conv.json
timeout 300 python bench.py --workload atari_conv_e256_b1024_sim50 --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32_atari_conv.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
python - <<PY
import json,glob,re
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
python tools/launch_summary.py $O/launches_bf16_atari.csv | head -8
