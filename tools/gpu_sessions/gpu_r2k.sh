set -x
O=gpurun_out/r2k; mkdir -p $O
for H in 0 1; do
  MZ_TC_TMEM_H=$H timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest_tc_h$H.txt
  cat $O/pytest_tc_h$H.txt
  MZ_TC_TMEM_H=$H MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python tools/bench_recurrent.py 2>&1 | grep -E "tc clk|us_per_call" > $O/tc_clk_h$H.txt
  MZ_TC_TMEM_H=$H timeout 300 python tools/bench_recurrent.py > $O/recurrent_micro_h$H.txt 2>&1; cat $O/recurrent_micro_h$H.txt
  for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200 lunarlander_mlp_e64_b4096_sim200; do
    MZ_TC_TMEM_H=$H timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_h${H}_$w.json
  done
done
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep "tc clk" | tail -3 > $O/tc_clk_search_atari.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bf16_atari.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l1.log 2>&1
python - <<PY
import json,glob,re
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
for f in sorted(glob.glob("$O/tc_clk_*.txt")):
    seen=set(); print(f)
    for ln in open(f):
        if ln.startswith("tc clk"):
            key=tuple(re.findall(r"k16=\d+ n=\d+", ln))
            if key in seen: continue
            seen.add(key); print(ln.strip())
PY
