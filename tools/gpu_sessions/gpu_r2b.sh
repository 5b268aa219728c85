# phase clocks + ncu of the tree-warp engine at the C3 shapes
set -x
O=gpurun_out/r2b; mkdir -p $O
for lg in 8 16 32; do
  MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so MZ_TREEWARP_LANES=$lg timeout 300 python bench.py --workload lunarlander_mlp_e64_b4096_sim200 --steps 1 --warmup 3 > $O/clk_lg$lg.txt 2>&1
done
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so MZ_TREEWARP_LANES=16 timeout 300 python bench.py --workload lunarlander_notebook_e64_b4096_sim200 --steps 1 --warmup 3 > $O/clk_notebook_lg16.txt 2>&1
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so MZ_TREEWARP_LANES=16 timeout 300 python bench.py --workload lunarlander_gumbel_e64_b4096_sim32 --steps 1 --warmup 3 > $O/clk_gumbel_lg16.txt 2>&1
MZ_TREEWARP_LANES=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:treewarp_search -c 1 -s 3 -o $O/treewarp_lunar_full -f python bench.py --steps 2 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 > $O/ncu.log 2>&1
grep -h "cta 1 warp" $O/clk_*.txt | sort | uniq -c | head -40
