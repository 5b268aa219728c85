set -x
O=gpurun_out/r2u; mkdir -p $O
for C in 1 0; do
MZ_TW_CACHED_SCORES=$C MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep -E "bs clk" | tail -4 > $O/clk_cached$C.txt
done
