set -x
O=gpurun_out/r2y; mkdir -p $O
for cfg in "16 2" "8 2" "8 1" "16 1" "16 3"; do
  set -- $cfg
  MZ_WARP_LANES=$1 MZ_WARP_PRODUCERS=$2 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_l$1_p$2.json
done
ncu --set full --clock-control none --import-source on -k regex:warp_search -c 1 -s 3 -o $O/warp_full -f python bench.py --steps 2 --warmup 3 > $O/n1.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.4f kernel_ms %.4f value %.1fM e2e %.1fM"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
