#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite, the default bench line, the VCF-text workloads, NJ phase accounting
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/r02_gputests.log 2>&1; echo "gpu tests rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/r02_gputests.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ); echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
for w in c1-vcf c2-vcf; do timeout 600 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "bench $w rc=$?"; tail -3 gpurun_out/r02_bench_$w.err; done
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | tail -4
timeout 300 python tools/nj_profile.py 500 1000 2000 3000 > gpurun_out/r02_nj_profile.txt 2>&1; tail -12 gpurun_out/r02_nj_profile.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print(d['ms_per_step'], d['stages_ms'], d['e2e']['ms_per_step'], d['gpu_launches'])
print({k:d['roofline'][k] for k in ('frac','peak','peak_fp4_measured','frac_of_fp4_measured','kernel_ms','fp4_probe')})
print(d['roofline_nj']['us_per_iteration'], d['roofline_nj']['frac'])
print(d['extra'].get('c5',{}).get('ms_per_step'), d['extra'].get('c5',{}).get('stages_ms'))
PY
