# tree-warp engine with cached selection scores: parity (whole GPU suite) + the C3 / C3-notebook / C4 lines
set -x
O=gpurun_out/r2ab; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/wl_$w.json
  python tools/bench_line.py $O/wl_$w.json
done
