#!/bin/bash
# round 2, state check: the whole -m gpu suite (no -x: list every failure), smoke(), the default bench line
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r02_gputests_full.log 2>&1; tail -6 gpurun_out/r02_gputests_full.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | tail -4; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print(d['ms_per_step'], d['stages_ms'], d['e2e']['ms_per_step'], d['gpu_launches'])
print({k:d['roofline'][k] for k in ('frac','frac_of_fp4_measured','kernel_ms','traffic')})
print(d['roofline_nj']['kernel'], d['roofline_nj']['us_per_iteration'])
PY
