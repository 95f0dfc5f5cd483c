#!/bin/bash
# Runs correctness (tools/tc_debug.py mid) + bench for every kprn_b200/variants/libkprn_*.so; one summary line per variant.
mkdir -p gpurun_out
for lib in kprn_b200/variants/libkprn_*.so; do
  n=$(basename $lib .so); n=${n#libkprn_}
  export KPRN_B200_LIB=$PWD/$lib
  timeout 120 python tools/tc_debug.py mid > gpurun_out/var_${n}_dbg.log 2>&1 || echo "$n: tc_debug FAILED/timeout"
  timeout 180 python bench.py --no-cpu-baseline --steps 60 --warmup 5 > gpurun_out/var_${n}_bench.log 2>&1 || echo "$n: bench FAILED/timeout"
  python - "$n" <<'PY'
import json,sys,re
n=sys.argv[1]
try:
    errs=[float(x) for x in re.findall(r"rel-to-rms ([0-9.e+-]+)", open("gpurun_out/var_%s_dbg.log"%n).read())]
    line=[l for l in open("gpurun_out/var_%s_bench.log"%n) if l.startswith("{")][-1]
    j=json.loads(line); k=j["roofline"]["kernels"]
    print("%-10s value %.2fM e2e %.2fM proj %.4f ms recur %.4f ms  max rel-to-rms %.2e" % (n, j["value"]/1e6, j["e2e"]["value"]/1e6, k["proj"]["launch_ms"], k["recur"]["launch_ms"], max(errs)))
except Exception as ex:
    print(n, "parse failed", ex)
PY
done
unset KPRN_B200_LIB
