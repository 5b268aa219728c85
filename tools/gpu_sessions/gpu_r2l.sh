set -x
O=gpurun_out/r2l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
for H in 0 1; do
for P in 0 1; do
  MZ_TC_TMEM_H=$H MZ_TW_SELECT_PREFETCH=$P timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_h${H}_p${P}_atari.json
done
done
for w in lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200; do
  for P in 0 1; do
  MZ_TC_TMEM_H=1 MZ_TW_SELECT_PREFETCH=$P timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_h1_p${P}_$w.json
  done
done
for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/fp32_$w.json
done
MZ_TC_TMEM_H=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:treewarp_search_kernel -c 1 -s 3 -o $O/treewarp_lunar -f python bench.py --steps 1 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 > $O/n1.log 2>&1
python - <<PY
import json,glob,re
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
