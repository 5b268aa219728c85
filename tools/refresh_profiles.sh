# Copies the artefacts of tools/final_measure.sh (gpurun_out/final/) into profiles/ as round-1 evidence.
O=gpurun_out/final
python tools/ncu_summary.py $O/warp_full.ncu-rep > profiles/r01_warp_final_summary.txt
python tools/ncu_lines.py $O/warp_full.ncu-rep warp_search_kernelILi2ELi8ELi16ELi10ELi16 30 > profiles/r01_warp_final_lines.txt 2>&1
python tools/launch_summary.py $O/launches_headline.csv "bench.py --steps 2 --warmup 3 (headline workload, engine auto = warp)" > profiles/r01_launches_headline_final_summary.txt
cp $O/launches_headline.csv profiles/r01_launches_headline_final.csv
if [ -f $O/resident_lunar_full.ncu-rep ]; then
python tools/ncu_summary.py $O/resident_lunar_full.ncu-rep > profiles/r01_resident_lunar_summary.txt
python tools/ncu_lines.py $O/resident_lunar_full.ncu-rep resident_search_kernelILi4ELb1 25 resident > profiles/r01_resident_lunar_lines.txt 2>&1
python tools/ncu_summary.py $O/resident_atari_full.ncu-rep > profiles/r01_resident_atari_summary.txt
python tools/ncu_lines.py $O/resident_atari_full.ncu-rep resident_search_kernelILi32ELb0 25 resident > profiles/r01_resident_atari_lines.txt 2>&1
python tools/launch_summary.py $O/launches_lunar.csv "bench.py --workload lunarlander_mlp_e64_b4096_sim200 (engine auto = resident)" > profiles/r01_launches_lunar_resident_summary.txt
cp $O/launches_lunar.csv profiles/r01_launches_lunar_resident.csv
fi
cp $O/bench_n1.json profiles/r01_bench_n1_final.json
cp $O/bench_n1_lane2.json profiles/r01_bench_n1_lane2_same_run.json
cp $O/bench_n1_warp_l8p1.json profiles/r01_bench_n1_warp_8lanes_1producer.json
cp $O/bench_n1_warp_prepass.json profiles/r01_bench_n1_warp_prepass_kernel.json
cp $O/bench_reference.json profiles/r01_bench_reference_arm.json
cp $O/workloads_auto.txt profiles/r01_bench_workloads_auto.txt
cp $O/pytest_gpu.txt profiles/r01_pytest_gpu.txt
for w in $O/wl_*.json; do cp $w profiles/r01_bench_$(basename $w | sed 's/^wl_//'); done
[ -f gpurun_out/bench_n2_warp.json ] && cp gpurun_out/bench_n2_warp.json profiles/r01_bench_n2_warp.json
python - <<'PY'
import json, re
t = json.load(open('profiles/traffic.json'))
def dram(f):
    s = open(f).read()
    def val(k):
        m = re.search(k + r' = ([0-9.]+) (\w+)', s)
        return float(m.group(1)) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1}[m.group(2)]
    return int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
a = dram('profiles/r01_warp_final_summary.txt')
l2 = dram('profiles/r01_lane2_v3_final_summary.txt')
t['cartpole_mlp_e8_b4096_sim50'] = {'auto': a, 'fused_warp': a, 'fused': a, 'fused_lane2': l2}
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
print('traffic', a, l2)
PY
