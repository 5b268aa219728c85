set -x
O=gpurun_out/r2d; mkdir -p $O
W=lunarlander_mlp_e64_b4096_sim200
MZ_TREEWARP_LANES=32 timeout 300 python bench.py --workload $W --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg32.json
MZ_TREEWARP_PREFETCH=1 timeout 300 python bench.py --workload $W --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg16_prefetch.json
MZ_TREEWARP_PREFETCH=1 MZ_TREEWARP_LANES=32 timeout 300 python bench.py --workload $W --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg32_prefetch.json
MZ_TREEWARP_PREFETCH=1 MZ_TREEWARP_LANES=8 timeout 300 python bench.py --workload $W --steps 5 --warmup 3 2>&1 | tail -1 > $O/tw_lg8_prefetch.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:treewarp_search -c 1 -s 3 -o $O/treewarp_v2_lunar -f python bench.py --steps 2 --warmup 3 --workload $W > $O/ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:treewarp_search -c 1 -s 3 -o $O/treewarp_v2_notebook -f python bench.py --steps 2 --warmup 3 --workload lunarlander_notebook_e64_b4096_sim200 > $O/ncu2.log 2>&1
# tcgen05 throughput mode: C5 heads and notebook nets, fp32 stepwise for reference
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM launches %d depth %.2f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["gpu_launches"], d["config"]["mean_path_depth"]))
    except Exception as e: print(f, "ERR", open(f).read()[-400:])
PY
