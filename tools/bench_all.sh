for w in cartpole_mlp_e8_b4096_sim50 cartpole_mlp_e8_b1024_sim50 lunarlander_gumbel_e64_b4096_sim32 lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 atari_mlp_e256_b1024_sim50; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --engine ${ENGINE:-resident} 2>&1 | tail -1 > gpurun_out/${TAG:-res}_$w.json; done; python - <<PY
import json,glob,os
for f in sorted(glob.glob("gpurun_out/${TAG:-res}_*.json")):
    try:
        d=json.load(open(f)); print(d["config"]["workload"], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d depth %.2f cpu %.1fM"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d["config"]["mean_path_depth"], d["cpu_baseline"]["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-300:])
PY
