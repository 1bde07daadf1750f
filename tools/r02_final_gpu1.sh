#!/bin/bash
# round 2, final state on one GPU: the -m gpu suite, smoke(), the default bench line, the VCF-text workloads, the
# reference arm, bootstrap timing, converter throughput, entry-point profile; then the ncu launch list of the bench
# command and one full capture of every kernel of the timed step at the bench's own size
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r02_gputests_final.log 2>&1; tail -6 gpurun_out/r02_gputests_final.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err ) 2>&1 | tail -4; tail -3 gpurun_out/r02_bench_final.err
for w in c1-vcf c2-vcf; do timeout 600 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "bench $w rc=$?"; tail -2 gpurun_out/r02_bench_$w.err; done
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | tail -4
timeout 300 python tools/bench_bootstrap.py 1000,1000000,1000,256 3000,10000000,100,256 2>&1 | tail -3; cp gpurun_out/bootstrap_timing.json gpurun_out/r02_bootstrap_timing.json
timeout 300 python tools/bench_vcf.py > gpurun_out/r02_vcf_ingest_throughput.json 2> gpurun_out/r02_vcf_ingest.err; grep "text_GBps" gpurun_out/r02_vcf_ingest_throughput.json
timeout 600 python tools/r02_entry_profile.py 1000000 > gpurun_out/r02_entry_profile_c3.json 2> gpurun_out/r02_entry_profile_c3.err; tail -12 gpurun_out/r02_entry_profile_c3.json
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || { tail -5 gpurun_out/r02_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_c4.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
# skip the warm-up step's launches of each kernel (-s 1 per kernel name match is not available: capture both steps' launches, 2 per kernel)
ncu --set full --import-source on --clock-control none -k regex:'pack_kernel|planes_to_e2m1_kernel|pair_counts_tc2_kernel|normalise_kernel|nj_lazy_kernel' -c 12 -f -o gpurun_out/r02_step_kernels $CMD > gpurun_out/r02_ncu_step.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r02_ncu_step.log; ls -la gpurun_out/r02_step_kernels.ncu-rep
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_final.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print({k:d['roofline'][k] for k in ('frac','frac_of_fp4_measured','kernel_ms','traffic')}, d['roofline_pack']['frac'], d['roofline_pack']['traffic'])
print(d['roofline_nj']['kernel'], d['roofline_nj']['us_per_iteration'], d['extra']['c5']['ms_per_step'], d['extra']['c5']['stages_ms'])
PY
