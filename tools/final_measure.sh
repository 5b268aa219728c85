# Round-end measurement bundle (run under gpurun on one B200); everything lands in gpurun_out/final/
# RESIDENT=1 also re-captures the CTA-resident engine's ncu reports (unchanged kernels otherwise keep their profiles).
set -x
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/smoke.txt
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json
python bench.py 2>&1 | tail -1 > $O/bench_n1.json
python bench.py --engine fused_lane2 2>&1 | tail -1 > $O/bench_n1_lane2.json
MZ_WARP_LANES=8 MZ_WARP_PRODUCERS=1 python bench.py 2>&1 | tail -1 > $O/bench_n1_warp_l8p1.json
MZ_WARP_PRODUCERS=0 python bench.py 2>&1 | tail -1 > $O/bench_n1_warp_prepass.json
ENGINE=auto TAG=final/wl bash tools/bench_all.sh > $O/workloads_auto.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --steps 2 --warmup 3 > $O/l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_search -c 1 -s 3 -o $O/warp_full -f python bench.py --steps 2 --warmup 3 > $O/n1.log 2>&1
if [ -n "$RESIDENT" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_lunar.csv python bench.py --steps 2 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 > $O/l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident_search -c 1 -s 3 -o $O/resident_lunar_full -f python bench.py --steps 2 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 > $O/n2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident_search -c 1 -s 3 -o $O/resident_atari_full -f python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 > $O/n3.log 2>&1
fi
cat $O/pytest_gpu.txt $O/smoke.txt $O/workloads_auto.txt
