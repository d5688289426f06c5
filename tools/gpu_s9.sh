#!/bin/bash
# pair kernel (cta_group::2) bring-up: small parity first (under a short timeout), then tests, then A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/r2s9_smoke.txt 2>&1 <<'PY'
import numpy as np
from oracle import cpu
from seismic_bpmf_b200 import fast_matched_filter as fmf
from tests import _synth
for (T,S,C,L,N) in [(20,4,3,400,60000),(33,3,2,800,40000),(5,2,1,100,9000)]:
    tp, mv, w, d, planted = _synth.mf_case(T,S,C,L,N, seed=T)
    want = cpu.matched_filter(tp, mv, w, d, 1)
    for prec in ("fp16","fp16x3"):
        got = fmf.matched_filter(tp, mv, w, d, 1, arch="gpu", precision=prec, check_zeros=False)
        print(T,S,C,L,N,prec,"max err", float(np.abs(got-want).max()), flush=True)
PY
echo "smoke rc=$?" >> gpurun_out/r2s9_smoke.txt
cat gpurun_out/r2s9_smoke.txt
if grep -q "smoke rc=0" gpurun_out/r2s9_smoke.txt; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2s9_tests.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/r2s9_tests.log
  Q="--no-e2e --no-cpu-baseline --no-beam --no-cfg4 --no-libc --no-variants"
  for pair in 1 0 1 0; do
    BPMF_TC_PAIR=$pair timeout 300 python bench.py --steps 6 --warmup 3 $Q >> gpurun_out/r2s9_pair$pair.json 2>> gpurun_out/r2s9_ab.err
  done
  tail -c 300 gpurun_out/r2s9_tests.log
fi
