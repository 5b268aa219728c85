# Last A/B of the round: split walk (MZ_TW_SPLITWALK) and out-of-line threefry continuation (MZ_TW_COLD_NOINLINE) on
# the compact tree-warp kernel; GPU suite on the default library and on the split-walk one.
O=gpurun_out/r2ag; mkdir -p $O
run() {  # run <tag> <workload> [env...]
  local tag=$1 w=$2; shift 2
  env "$@" timeout 100 python bench.py --workload $w --steps 5 --warmup 3 < /dev/null 2>&1 | tail -1 > $O/${tag}_$w.json
  python tools/bench_line.py "$tag $w" < $O/${tag}_$w.json
}
C3=lunarlander_mlp_e64_b4096_sim200; NB=lunarlander_notebook_e64_b4096_sim200
SP=MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_split.so; CO=MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_cold.so
timeout 200 python -m pytest tests -m gpu -q < /dev/null 2>&1 | tail -2 > $O/pytest_default.txt; cat $O/pytest_default.txt
run default $C3
run split $C3 $SP
run cold $C3 $CO
run default $NB
run split $NB $SP
run cold $NB $CO
env $SP timeout 200 python -m pytest tests -m gpu -q < /dev/null 2>&1 | tail -2 > $O/pytest_split.txt; cat $O/pytest_split.txt
