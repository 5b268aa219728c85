set -x
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
for T in 1 0; do
  for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200; do
    MZ_TW_SMEM_TREE=$T timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_smemtree${T}_$w.json
  done
done
timeout 300 python bench.py --workload atari_conv_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_atari_conv.json
MZ_TC_DUMP=1 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep -E "tc clk|tc program|^  s[0-9]" | tail -30 > $O/tc_clk_search_atari.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
python tools/launch_summary.py $O/launches_bf16_atari.csv | head -6
