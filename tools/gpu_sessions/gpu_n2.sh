set -x
N=${N:-2}
O=gpurun_out/n$N; mkdir -p $O
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -5 > $O/pytest_sharded.txt; cat $O/pytest_sharded.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" 2>&1 | tail -1; }
timeout 300 python bench.py 2>&1 | tail -1 > $O/bench_n1.json
run --steps 20 --warmup 3 > $O/bench_weak.json
run --steps 20 --warmup 3 --scaling strong > $O/bench_strong.json
run --steps 5 --warmup 3 --workload atari_conv_e256_b1024_sim50 > $O/bench_conv_weak.json
run --steps 5 --warmup 3 --workload atari_conv_e256_b1024_sim50 --precision bf16 > $O/bench_conv_weak_bf16.json
run --steps 5 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/bench_atari_weak_bf16.json
run --impl reference --steps 3 --warmup 1 > $O/bench_reference.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "n %d ms %.3f value %.1fM e2e %.1fM"%(d["n_gpus"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), d.get("details",{}).get("exchange"), d["config"])
    except Exception as e: print(f, "ERR", open(f).read()[-600:])
PY
