set -x
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_conv.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest_tc.txt
tail -3 $O/pytest_tc.txt
for C in 1 0; do
  MZ_TW_CACHED_SCORES=$C timeout 120 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_cached${C}.json
done
for kb in 48 72; do
  MZ_TC_STAGE_KB=$kb timeout 120 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_stage${kb}.json
done
MZ_TC_STAGE_KB=72 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep -E "tc clk|bs clk" | tail -4 > $O/clk.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
