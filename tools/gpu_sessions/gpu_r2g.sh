set -x
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/smoke.txt; cat $O/smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32 lunarlander_notebook_e64_b4096_sim200 atari_mlp_e256_b1024_sim50 atari_conv_e256_b1024_sim50 cartpole_mlp_e8_b1024_sim50; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/wl_$w.json
done
for w in lunarlander_notebook_e64_b4096_sim200 atari_mlp_e256_b1024_sim50 atari_conv_e256_b1024_sim50 lunarlander_mlp_e64_b4096_sim200; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_$w.json
done
timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --engine stepwise 2>&1 | tail -1 > $O/stepwise_fp32_atari.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recurrent_tc_kernel -c 1 -s 60 -o $O/recurrent_tc_atari -f python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/n1.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)), d.get("details",{}).get("mean_path_depth"))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
