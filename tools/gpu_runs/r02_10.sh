#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 900 python tools/config_sweep.py > gpurun_out/r02_config_sweep_10.json 2> gpurun_out/r02_config_sweep_10.err ) 2>&1 | tail -3
echo rc=$?; tail -5 gpurun_out/r02_config_sweep_10.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_config_sweep_10.json').read().strip().splitlines()[-1])
for k,v in d.items():
    if k=='c3':
        print('c3', v.get('error') or (len(v['entries']), v['wall_s'], v['hbm_bound_entries_below_0.70_at_streaming_sizes']))
        for e in v.get('entries', []): print('   ', e['op'], e['size'], e.get('dtype',''), e['ms'], e['achieved'], e['unit'], e['frac'])
    else:
        print(k, {a:b for a,b in v.items() if a not in ('clocks',)})
PY
