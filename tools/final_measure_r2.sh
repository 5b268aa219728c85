# Round-2 measurement bundle (run under gpurun on one B200); everything lands in gpurun_out/final2/
set -x
O=gpurun_out/final2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/smoke.txt
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json
python bench.py 2>&1 | tail -1 > $O/bench_n1.json
for w in cartpole_mlp_e8_b1024_sim50 lunarlander_gumbel_e64_b4096_sim32 lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 atari_mlp_e256_b1024_sim50 atari_conv_e256_b1024_sim50; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > $O/wl_$w.json
done
for w in lunarlander_mlp_e64_b4096_sim200 lunarlander_notebook_e64_b4096_sim200 atari_mlp_e256_b1024_sim50 atari_conv_e256_b1024_sim50; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --precision bf16 2>&1 | tail -1 > $O/wl_bf16_$w.json
done
# launch lists (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --steps 2 --warmup 3 > $O/l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_lunar.csv python bench.py --steps 2 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200 > $O/l2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_atari_bf16.csv python bench.py --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16 > $O/l3.log 2>&1
# full captures of the top kernels; summarised on the box (the .ncu-rep files together exceed what gpurun brings back)
cap() {  # cap <name> <kernel regex> <mangled-name substring for the line attribution> <skip> <top> <bench args...>
  local name=$1 rx=$2 sub=$3 skip=$4 top=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -s $skip -o $O/$name -f python bench.py "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/${name}_summary.txt 2>&1
  python tools/ncu_lines.py $O/$name.ncu-rep $sub $top > $O/${name}_lines.txt 2>&1
  rm -f $O/$name.ncu-rep
}
cap warp_full warp_search warp_search_kernelILi2ELi8ELi16ELi10ELi16 3 32 --steps 2 --warmup 3
cap treewarp_lunar_full treewarp_search_kernel treewarp_search_kernel 3 30 --steps 1 --warmup 3 --workload lunarlander_mlp_e64_b4096_sim200
cap recurrent_tc_atari_full recurrent_tc_kernel recurrent_tc_kernel 80 30 --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16
cap backup_select_atari_full tw_backup_select_kernel tw_backup_select_kernel 80 25 --steps 2 --warmup 3 --workload atari_mlp_e256_b1024_sim50 --precision bf16
# tcgen05 kernel timeline (instrumented build)
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python bench.py --workload atari_mlp_e256_b1024_sim50 --steps 1 --warmup 3 --precision bf16 2>&1 | grep "tc clk" | tail -2 > $O/tc_clk_search_atari.txt
MZ_LIB_PATH=$PWD/muax_b200/libmzsearch_clk.so timeout 300 python tools/bench_recurrent.py 2>&1 | grep -E "tc clk|us_per_call" > $O/tc_clk_micro.txt
timeout 300 python tools/bench_recurrent.py > $O/recurrent_micro.txt 2>&1
# memory checker on the throughput mode and the stochastic path (small cases)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stochastic.py -m gpu -x -q -k "batched or stochastic_muzero or stays_close" > $O/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/memcheck.txt
tail -5 $O/memcheck.txt
cat $O/pytest_gpu.txt $O/smoke.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "ms %.3f kernel_ms %.3f value %.1fM e2e %.1fM launches %d frac %.4f"%(d["ms_per_step"], d.get("roofline",{}).get("kernel_ms",0), d["value"]/1e6, d["e2e"]["value"]/1e6, d["gpu_launches"], d.get("roofline",{}).get("frac",0)))
    except Exception as e: print(f, "ERR", open(f).read()[-300:])
PY
