#!/bin/bash
# C2/C4 per-class times after freeing per-step graphs; compute-sanitizer racecheck + memcheck of smoke()
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/config_prof.py 2>&1 | tail -8
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -8 gpurun_out/r02_racecheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r02_memcheck_smoke.log
