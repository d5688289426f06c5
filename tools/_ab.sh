python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/s4_tests.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/s4_w16.json 2>gpurun_out/s4.err
BPMF_B200_LIB=$PWD/seismic_bpmf_b200/lib/libbpmf_b200_w8.so $B > gpurun_out/s4_w8.json 2>>gpurun_out/s4.err
BPMF_TC_DEBUG=8 $B > gpurun_out/s4_w16dbg8.json 2>>gpurun_out/s4.err
BPMF_TC_DEBUG=4 $B > gpurun_out/s4_w16dbg4.json 2>>gpurun_out/s4.err
tail -3 gpurun_out/s4_tests.log
for f in w16 w8 w16dbg8 w16dbg4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"], d["clocks"]["sm_mhz"], d.get("parity")["max_abs_err_vs_oracle"])
except Exception as e: print("$f", "ERR", e)
PY
done
