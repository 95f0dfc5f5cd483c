"""Reads a bench.py JSON line on stdin and prints the few numbers used while tuning."""
import json, sys
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
tag = sys.argv[1] if len(sys.argv) > 1 else ""
k = j.get("roofline", {}).get("kernels", {})
print(tag, "score %.2f M paths/s (%.3f ms; proj %.3f recur %.3f) e2e %.2f M" % (
    j["value"] / 1e6, j["ms_per_step"], k.get("proj", {}).get("launch_ms", 0), k.get("recur", {}).get("launch_ms", 0), j["e2e"]["value"] / 1e6))
t = j.get("train")
if t and "ms_per_step" in t:
    print(tag, "train %.3f ms/step = %.2f M paths/s, e2e %.2f M, launches/step %.0f, coll bytes/step %.0f" % (
        t["ms_per_step"], t["value"] / 1e6, t["e2e"]["value"] / 1e6, t["gpu_launches_per_step"], t["collectives"]["bytes_per_step_per_rank"]))
if "parity" in j:
    print(tag, "parity", j["parity"])
