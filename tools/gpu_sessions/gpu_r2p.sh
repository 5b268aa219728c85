set -x
O=gpurun_out/r2p; mkdir -p $O
MZ_TC_DUMP=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -40 > $O/pytest_tc.txt
tail -5 $O/pytest_tc.txt
for P in 1 0; do
  for w in atari_mlp_e256_b1024_sim50 lunarlander_notebook_e64_b4096_sim200; do
    MZ_TC_PIPE=$P timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_pipe${P}_$w.json
  done
done
MZ_TC_DUMP=1 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep -E "tc clk|tc program|^  s[0-9]" | tail -30 > $O/tc_clk_search_atari.txt
timeout 300 python tools/bench_recurrent.py > $O/recurrent_micro.txt 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
cat $O/recurrent_micro.txt | tail -5
