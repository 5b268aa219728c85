# compact tree-warp kernel: A/B of the cached selection scores and of MZ_TW_UNROLL = 2 on top of it, warps / lanes knobs,
# and a full ncu capture of the new default (summarised on the box)
O=gpurun_out/r2ae; mkdir -p $O
run() {  # run <tag> <workload> [env...]
  local tag=$1 w=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $w --steps 5 --warmup 3 < /dev/null 2>&1 | tail -1 > $O/${tag}_$w.json
  python tools/bench_line.py "$tag $w" < $O/${tag}_$w.json
}
C3=lunarlander_mlp_e64_b4096_sim200; NB=lunarlander_notebook_e64_b4096_sim200
CA=MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_cached.so; U2=MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_u2.so
run compact $C3
run cached $C3 $CA
run u2 $C3 $U2
run compact $NB
run cached $NB $CA
run u2 $NB $U2
run compact_w7 $C3 MZ_TREEWARP_WARPS=7
run compact_l8 $C3 MZ_TREEWARP_LANES=8
run compact_k16 $C3 MZ_TREEWARP_K=16
run compact_k64 $C3 MZ_TREEWARP_K=64
name=treewarp_compact_full
timeout 200 ncu --set full --clock-control none --import-source on -k regex:treewarp_search -c 1 -s 3 -o $O/$name -f python bench.py --workload $C3 --steps 1 --warmup 3 < /dev/null > $O/$name.log 2>&1
python tools/ncu_summary.py $O/$name.ncu-rep > $O/${name}_summary.txt 2>&1
python tools/ncu_lines.py $O/$name.ncu-rep treewarp_search_kernel 30 > $O/${name}_lines.txt 2>&1
rm -f $O/$name.ncu-rep
cat $O/${name}_summary.txt; head -16 $O/${name}_lines.txt
