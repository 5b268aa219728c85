O=gpurun_out/r2v; mkdir -p $O
python tools/act_overhead.py > $O/act_overhead.txt 2>&1
B=4096 NS=50 python tools/act_overhead.py 2>&1 | head -3 >> $O/act_overhead.txt
MZ_TREEWARP_K=50 timeout 120 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_K50.json
MZ_TC_STAGE_KB=64 timeout 120 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/bf16_stage64.json
