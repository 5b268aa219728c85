# compute-sanitizer racecheck + memcheck over the warp engine's variant tests (run under gpurun); log in gpurun_out/san/
mkdir -p gpurun_out/san
(
echo "# compute-sanitizer --tool racecheck on tests/test_gpu_parity.py::test_warp_engine_variants (A=2 S=5 B=64 NS=24, table depth 2; 8/16 lanes, 0/2/3 producers) + golden vector c1_muzero_seed0 on engine 8"
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -k "(warp_engine_variants and 2-5-64) or (test_golden_vectors and c1_muzero_seed0 and 8)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -5
echo "# compute-sanitizer --tool memcheck on all warp-engine variant tests"
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -k "warp_engine_variants" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -5
) > gpurun_out/san/warp_sanitizer.txt 2>&1
cat gpurun_out/san/warp_sanitizer.txt
