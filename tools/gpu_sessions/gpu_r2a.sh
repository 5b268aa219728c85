# round-2 first light: parity of the restructured library + the tree-warp engine, then its speed at C3 / C4
set -x
O=gpurun_out/r2a; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32 lunarlander_notebook_e64_b4096_sim200; do
  for lg in 8 16 32; do
    MZ_TREEWARP_LANES=$lg timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg${lg}_$w.json
  done
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --engine resident 2>&1 | tail -1 > $O/res_$w.json
done
MZ_TREEWARP_K=0 timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_k0_lunar.json
MZ_TREEWARP_K=48 timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_k48_lunar.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d["config"]["workload"], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d depth %.2f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d["config"]["mean_path_depth"]))
    except Exception as e: print(f, "ERR", open(f).read()[-400:])
PY
