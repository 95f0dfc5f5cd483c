"""Times the gate phase of the tensor-core forward for several KPRN_TC_FLAGS settings (debug / tuning)."""
import os, subprocess, sys, json
for fl in sys.argv[1:] or ["0", "1", "2", "4", "7"]:
    env = dict(os.environ, KPRN_TC_FLAGS=fl)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-cpu-baseline"], env=env,
                         capture_output=True, text=True).stdout.strip().splitlines()
    try:
        j = json.loads(out[-1])
        print("flags", fl, "value %.2fM" % (j["value"] / 1e6), "ms/step %.3f" % j["ms_per_step"], "gate ms %.3f" % j["roofline"]["launch_ms"], flush=True)
    except Exception as e:
        print("flags", fl, "failed", out[-3:], flush=True)
