set -x
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "treewarp or edge_cases or seeded" 2>&1 | tail -8 > $O/pytest.txt
cat $O/pytest.txt
W=lunarlander_mlp_e64_b4096_sim200
for lg in 8 16; do MZ_TREEWARP_LANES=$lg timeout 300 python bench.py --workload $W --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg$lg.json; done
timeout 300 python bench.py --workload lunarlander_gumbel_e64_b4096_sim32 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_gumbel.json
timeout 300 python bench.py --workload lunarlander_notebook_e64_b4096_sim200 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_notebook.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM launches %d depth %.2f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["gpu_launches"], d["config"]["mean_path_depth"]))
    except Exception as e: print(f, "ERR", open(f).read()[-400:])
PY
