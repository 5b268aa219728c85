# Copies the artefacts of tools/final_measure_r2.sh (gpurun_out/final2/) into profiles/ as round-2 evidence.
O=gpurun_out/final2
for pair in warp_full:warp treewarp_lunar_full:treewarp_lunar recurrent_tc_atari_full:recurrent_tc_atari backup_select_atari_full:backup_select_atari; do
  src=${pair%%:*}; dst=${pair##*:}
  cp $O/${src}_summary.txt profiles/r02_${dst}_summary.txt
  cp $O/${src}_lines.txt profiles/r02_${dst}_lines.txt
done
python tools/launch_summary.py $O/launches_headline.csv "bench.py --steps 2 --warmup 3 (headline workload, engine auto = warp)" > profiles/r02_launches_headline_summary.txt
python tools/launch_summary.py $O/launches_lunar.csv "bench.py --workload lunarlander_mlp_e64_b4096_sim200 (engine auto = tree-warp)" > profiles/r02_launches_lunar_treewarp_summary.txt
python tools/launch_summary.py $O/launches_atari_bf16.csv "bench.py --workload atari_mlp_e256_b1024_sim50 --precision bf16 (throughput mode)" > profiles/r02_launches_atari_bf16_summary.txt
cp $O/launches_headline.csv profiles/r02_launches_headline.csv
cp $O/launches_lunar.csv profiles/r02_launches_lunar_treewarp.csv
cp $O/launches_atari_bf16.csv profiles/r02_launches_atari_bf16.csv
cp $O/bench_n1.json profiles/r02_bench_n1_final.json
cp $O/bench_reference.json profiles/r02_bench_reference_arm.json
cp $O/pytest_gpu.txt profiles/r02_pytest_gpu.txt
cp $O/memcheck.txt profiles/r02_memcheck.txt
cp $O/recurrent_micro.txt profiles/r02_recurrent_micro.txt
for w in $O/wl_*.json; do cp $w profiles/r02_bench_$(basename $w | sed 's/^wl_//'); done
python - <<'PY'
import re
for src, dst, head in (("gpurun_out/final2/tc_clk_micro.txt", "profiles/r02_recurrent_tc_timeline_final.txt",
                        "# recurrent_tc_kernel timeline, final round-2 kernel (MZ_TC_CLOCKS build, CTA 0, SM cycles since kernel start; mz_recurrent standalone: fp32 gather / fp32 next-state store)"),
                       ("gpurun_out/final2/tc_clk_search_atari.txt", "profiles/r02_recurrent_tc_timeline_final_search.txt",
                        "# same, inside the search at the C5 shapes (bf16 tree embeddings; `A` includes the programmatic-dependent-launch wait for the kernel before)")):
    out, seen = [head, ""], set()
    try:
        for ln in open(src):
            if ln.startswith("tc clk"):
                key = tuple(re.findall(r"k16=\d+ n=\d+", ln))
                if key in seen:
                    continue
                seen.add(key)
                parts = ln.split("|")
                out.append(parts[0].strip() + " | " + parts[1].strip())
                out += ["    " + x.strip() for x in parts[2:]]
            else:
                out.append(ln.strip()[:200])
        open(dst, "w").write("\n".join(out) + "\n")
    except FileNotFoundError:
        pass
PY
python - <<'PY'
import json, re
t = json.load(open('profiles/traffic.json'))
def dram(f, kernel=None):
    s = open(f).read()
    def val(k):
        m = re.search(k + r' = ([0-9.]+) (\w+)', s)
        return float(m.group(1)) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1}[m.group(2)]
    return int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
try:
    a = dram('profiles/r02_warp_summary.txt')
    t['cartpole_mlp_e8_b4096_sim50'].update({'auto': a, 'fused_warp': a, 'fused': a})
    tw = dram('profiles/r02_treewarp_lunar_summary.txt')
    t.setdefault('lunarlander_mlp_e64_b4096_sim200', {}).update({'auto': tw, 'treewarp': tw, 'fused': tw})
    rt = dram('profiles/r02_recurrent_tc_atari_summary.txt')
    t.setdefault('atari_mlp_e256_b1024_sim50', {}).update({'bf16': rt})
    json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
    print('traffic', a, tw, rt)
except Exception as e:
    print('traffic not updated:', e)
PY
