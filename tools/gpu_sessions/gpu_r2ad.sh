# A/B of MZ_TW_COMPACT (rolled run-time loops in the tree-warp kernel) on one box + the GPU suite on the new default
O=gpurun_out/r2ad; mkdir -p $O
run() {  # run <tag> <workload> [env...]
  local tag=$1 w=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $w --steps 5 --warmup 3 < /dev/null 2>&1 | tail -1 > $O/${tag}_$w.json
  python tools/bench_line.py "$tag $w" < $O/${tag}_$w.json
}
C3=lunarlander_mlp_e64_b4096_sim200; NB=lunarlander_notebook_e64_b4096_sim200; C4=lunarlander_gumbel_e64_b4096_sim32
AB=MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_ab.so
run compact $C3
run unrolled $C3 $AB
run compact $NB
run unrolled $NB $AB
run compact $C4
run unrolled $C4 $AB
run compact $C3
run unrolled $C3 $AB
timeout 300 python -m pytest tests -m gpu -q -x < /dev/null 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio -k regex:treewarp_search -c 1 -s 3 --clock-control none python bench.py --workload $C3 --steps 1 --warmup 3 < /dev/null 2>&1 | grep -E "treewarp_search|duration|inst_executed|issue_active|no_instruction" > $O/ncu_compact.txt
cat $O/ncu_compact.txt
