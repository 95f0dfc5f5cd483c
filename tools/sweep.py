"""BASELINE configs[4]: embedding/hidden sweep d in {64,128,256,512}, batch = 2^16 paths, 1 GPU.
Writes profiles/<tag>_sweep.md (paths/s, gather GB/s, achieved TFLOP/s and fraction of the measured bf16 peak)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = []
for d, conv in ((64, "A"), (128, "A"), (256, "A"), (512, "A"), (256, "B")):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--d", str(d), "--conv", conv, "--steps", "10", "--warmup", "3",
                          "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=280)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print("d=%d failed: %s" % (d, out.stderr[-400:]))
        continue
    r = j["roofline"]
    tr = j.get("train") or {}
    rows.append(("%d (%s)" % (d, conv), j["value"], j["e2e"]["value"], j["hbm"]["gather_gbs"], r.get("gate_phase_frac"), r["mode"], r["kernels"],
                 tr.get("ms_per_step"), tr.get("value")))
    print(d, conv, j["value"], r["mode"], flush=True)
with open(os.path.join(ROOT, "profiles", "%s_sweep.md" % tag), "w") as f:
    f.write("# %s -- dimension sweep (BASELINE configs[4]): convention A (d_t=d_e=d_r=H=d) and, last row, convention B (d_e=d/2, d_t=d_r=d/4, H=d: the reference ratio), 65 536 paths/step, T=6, 1xB200\n\n" % tag)
    f.write("`frac` = algorithmic flops of the gate phase / CUDA-event time / measured sustained bf16 peak (MEASURED_PEAKS.json); "
            "gather GB/s = algorithmic gather bytes (T*(12+4D) per path) x paths/s.\n\n")
    f.write("| d (conv) | paths/s (device) | paths/s (e2e) | gather GB/s | gate-phase frac | engine | proj ms | recur ms | train ms/step | train paths/s |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for d, v, e, g, fr, mode, k, tms, tv in rows:
        f.write("| %s | %.2f M | %.2f M | %.0f | %s | %s | %s | %s | %s | %s |\n" % (
            d, v / 1e6, e / 1e6, g, ("%.3f" % fr) if fr else "-", mode,
            ("%.3f" % k["proj"]["launch_ms"]) if "proj" in k else "-", ("%.3f" % k["recur"]["launch_ms"]) if "recur" in k else "-",
            ("%.2f" % tms) if tms else "-", ("%.2f M" % (tv / 1e6)) if tv else "-"))
    f.write("\nTraining column: `kprn_train_step_device`, precision 3; the fused tensor-core backward covers H <= 256 with d_e in {128, 256} "
            "(d = 64 and d = 512 run the generic engine's backward).\n")
print(open(os.path.join(ROOT, "profiles", "%s_sweep.md" % tag)).read())
