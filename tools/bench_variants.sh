# A/B harness: runs a few workloads against every library under muax_b200/variants/ (MZ_LIB_PATH override)
WL=${WL:-"cartpole_mlp_e8_b4096_sim50 lunarlander_mlp_e64_b4096_sim200 lunarlander_gumbel_e64_b4096_sim32"}
for lib in muax_b200/variants/libmz_*.so; do
  v=$(basename $lib .so)
  for w in $WL; do
    MZ_LIB_PATH=$PWD/$lib timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --engine ${ENGINE:-resident} 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v', d['config']['workload'], 'ms %.3f'%d['ms_per_step'], 'value %.1fM'%(d['value']/1e6))
except Exception as e: print('$v $w ERR', e)
"
  done
done
