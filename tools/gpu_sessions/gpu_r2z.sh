O=gpurun_out/r2z; mkdir -p $O
for k in 8 10 12 13 16; do
  MZ_GROUP_K=$k python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_k$k.json
done
MZ_WARP_PRODUCERS=4 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_p4.json
MZ_WARP_PRODUCERS=4 MZ_GROUP_K=12 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_p4_k12.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.4f kernel_ms %.4f value %.1fM e2e %.1fM"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
