# Final check of round 2 on the compact tree-warp kernel: GPU suite, smoke, both bench arms, the tree-warp workloads,
# launch list of the C3 workload.  Everything lands in gpurun_out/final3/
O=gpurun_out/final3; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q < /dev/null 2>&1 | tail -4 > $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" < /dev/null 2>&1 | tail -2 > $O/smoke.txt
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 < /dev/null 2>&1 | tail -1 > $O/bench_reference.json
timeout 200 python bench.py < /dev/null 2>&1 | tail -1 > $O/bench_n1.json
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32; do
  timeout 120 python bench.py --workload $w --steps 5 --warmup 3 < /dev/null 2>&1 | tail -1 > $O/wl_$w.json
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_lunar.csv python bench.py --steps 2 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 < /dev/null > $O/l2.log 2>&1
cat $O/pytest_gpu.txt $O/smoke.txt
for f in $O/*.json; do echo "$(basename $f) $(python tools/bench_line.py < $f)"; done
