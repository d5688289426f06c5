#!/bin/bash
# Round-end sequence on one B200 (what the driver runs, plus the beamforming figures and one ncu capture):
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh <tag>'
cd "${GRAFT_REPO_ROOT:-.}"
tag=${1:-validate}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $out/tests.log 2>&1; echo "tests: $(tail -1 $out/tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 400 $out/bench_ref.json; echo
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err; python - <<P
import json
d=json.loads([l for l in open("$out/bench.json") if l.startswith("{")][-1])
print("bench: value %.1f e2e %.1f frac %.3f; beam frac %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("beamform", {}).get("roofline", {}).get("frac")))
P
for args in "" "--keep 10" "--sources 125000"; do
  n=$(echo beam$args | tr -d ' -')
  timeout 200 python tools/bench_beam.py $args > $out/$n.json 2> $out/$n.err
  python -c "import json; d=json.loads([l for l in open('$out/$n.json') if l.startswith('{')][-1]); print('$n', 'ms %.1f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:beam_group_kernel -c 1 -o $out/beam_group_full python tools/bench_beam.py --sources 8000 --samples 216000 --steps 1 --warmup 0 > $out/ncu.log 2>&1; tail -2 $out/ncu.log
