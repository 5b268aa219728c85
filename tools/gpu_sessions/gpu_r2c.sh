# round-2: tree-warp v2 (warp-uniform walks, parallel backup) + first light of the tcgen05 recurrent kernel
set -x
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "treewarp or seeded or full_sizes or golden or root_supplied" 2>&1 | tail -15 > $O/pytest_treewarp.txt
cat $O/pytest_treewarp.txt
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s 2>&1 | tail -40 > $O/pytest_tc.txt
cat $O/pytest_tc.txt
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32 lunarlander_notebook_e64_b4096_sim200; do
  for lg in 8 16; do
    MZ_TREEWARP_LANES=$lg timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg${lg}_$w.json
  done
done
MZ_TREEWARP_K=16 timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_k16_lunar.json
MZ_TREEWARP_K=48 timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_k48_lunar.json
for lg in 8 16; do
  MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so MZ_TREEWARP_LANES=$lg timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 1 --warmup 3 2>&1 | grep "cta 1 warp" | tail -3 > $O/clk_lg$lg.txt
done
cat $O/clk_lg*.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM launches %d depth %.2f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["gpu_launches"], d["config"]["mean_path_depth"]))
    except Exception as e: print(f, "ERR", open(f).read()[-400:])
PY
