"""Condenses bench.py's JSON line: engine, ms / act (events), search-kernel ms, device and end-to-end sims/s."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(sys.argv[1] if len(sys.argv) > 1 else "", d["config"].get("engine"), "ms %.4f kernel %.4f value %.1fM e2e %.1fM" % (
        d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"] / 1e6, d["e2e"]["value"] / 1e6))
