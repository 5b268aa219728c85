# DRAM bytes per launch of the dominant kernel of the C2 (cartpole 1024 x 50) and C4 (Gumbel 4096 x 32) workloads, for
# profiles/traffic.json (roofline.traffic of those bench lines)
O=gpurun_out/r2ah; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 25 ncu --metrics $M -k regex:warp_search -c 1 -s 3 --clock-control none python bench.py --workload cartpole_mlp_e8_b1024_sim50 --steps 1 --warmup 3 < /dev/null 2>&1 | grep -E "warp_search|dram__|duration" > $O/traffic_c2.txt &
timeout 25 ncu --metrics $M -k regex:treewarp_search -c 1 -s 3 --clock-control none python bench.py --workload lunarlander_gumbel_e64_b4096_sim32 --steps 1 --warmup 3 < /dev/null 2>&1 | grep -E "treewarp_search|dram__|duration" > $O/traffic_c4.txt &
wait
cat $O/traffic_c2.txt $O/traffic_c4.txt
