# A/B of the tree-warp engine's cached selection scores (MZ_TW_CACHED) and of its lanes / warps knobs, one box
O=gpurun_out/r2ac; mkdir -p $O
run() {  # run <tag> <workload> [env...]
  local tag=$1 w=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $w --steps 5 --warmup 3 < /dev/null 2>&1 | tail -1 > $O/${tag}_$w.json
  python tools/bench_line.py "$tag $w" < $O/${tag}_$w.json
}
C3=lunarlander_mlp_e64_b4096_sim200; NB=lunarlander_notebook_e64_b4096_sim200
run cached $C3
run nocache $C3 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_nocache.so
run cached $NB
run nocache $NB MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_nocache.so
run cached_w7 $C3 MZ_TREEWARP_WARPS=7
run nocache_w7 $C3 MZ_TREEWARP_WARPS=7 MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_nocache.so
run cached_l8 $C3 MZ_TREEWARP_LANES=8
run cached_l32 $C3 MZ_TREEWARP_LANES=32
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:treewarp_search -c 1 -s 3 --clock-control none python bench.py --workload $C3 --steps 1 --warmup 3 < /dev/null 2>&1 | grep -E "treewarp_search|duration|inst_executed|issue_active" > $O/ncu_cached.txt
cat $O/ncu_cached.txt
