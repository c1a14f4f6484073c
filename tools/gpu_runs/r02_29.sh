#!/bin/bash
# round-2 evidence: default bench line (parity legs, cpu baseline, extra.configs), reference arm, racecheck of smoke()
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
/usr/bin/time -v -o gpurun_out/r02_bench_default.time timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "default bench rc=$?"; grep "Elapsed" gpurun_out/r02_bench_default.time
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/r02_bench_reference.json
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck_smoke.log
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'cpu', d['cpu_baseline'])
print(json.dumps(d['config'].get('parity'))[:1500])
print(json.dumps(d.get('extra'))[:3000])
PY
