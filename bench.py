#!/usr/bin/env python
"""Benchmark of the population-phylogeny hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], "3,000 samples x 10M SNPs, sites sharded over 2/4/8
B200", the shape the north-star target is quoted on. It fits one B200 (30 GB of site-major byte genotypes +
7.5 GB of bit planes). Default "scaling": "strong" — the named configuration's 10M sites are split over the N
ranks, as BASELINE.json words it, so every N measures the SAME job; `--scaling weak` gives every GPU the full 10M
sites instead. The default run also prints two bounded sub-records under "extra": "weak" (N > 1) and "c5"
(BASELINE configs[4], 10,000 x 5M with 10 % missing genotypes, which BASELINE names for 8 GPUs).

A step = one pass of the hot path over the whole batch:
    pack (byte genotypes -> bit planes) -> all-pairs counts -> [NCCL allreduce of the
    int32 count matrices] -> fp64 distances -> neighbor-joining tree.
  value   whole-job sample-pair x sites per second, inputs resident in HBM, CUDA-event timed,
          max over ranks.
  e2e     the same metric through Engine.tree_from_host_codes: genotypes start in pinned host
          memory (2-bit site-major rows by default, --e2e-format u8 for one byte per genotype),
          host->device copies (chunked, overlapped with pack + counts) and the device->host read
          of the result (distance matrix, join list -> Newick) are timed.
  roofline      the pair-count kernel: tensor-core kernel against 4 x MEASURED_PEAKS bf16 and against the fp4
                rate measured in this run by the probes library; popcount kernel against the integer-pipe
                peak measured in this run.
  roofline_nj   the NJ kernel against measured HBM (4 n^2 bytes of live upper triangle per iteration).
  cpu_baseline  the CPU oracle (oracle/pf_oracle.c, all host threads) on a bounded sample,
                rank 0 at N=1 only.
`--impl reference` times that CPU port alone (the reference's own engine for distance + NJ is
the absent external `treebest`; oracle/__init__.py).

`--workload c1-vcf | c2-vcf`: BASELINE configs[0] / configs[1] as real VCF TEXT through the reference-facing
entry point (`phyloforge -s/-S cfg`): file -> character matrix -> packed genotypes -> distances -> NJ -> tree
files; the CPU leg beside it is the reference's own converter (the unmodified one when /root/reference is
present, else its restatement oracle/ref_convert.py), which is all of this path the reference runs itself.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (samples, sites, miss_ppm, description)
    "c1": (50, 100_000, 0, "50 samples x 100k SNPs (BASELINE configs[0])"),
    "c3": (1000, 1_000_000, 0, "1,000 samples x 1M SNPs (BASELINE configs[2])"),
    "c4": (3000, 10_000_000, 0, "3,000 samples x 10M SNPs (BASELINE configs[3])"),
    "c5": (10000, 5_000_000, 100_000, "10,000 samples x 5M SNPs, 10% missing (BASELINE configs[4])"),
}
VCF_WORKLOADS = {
    # name: (kind, samples, records, miss_ppm, description)
    "c1-vcf": ("snp", 50, 100_000, 0, "VCF text, 50 samples x 100k biallelic SNPs through `phyloforge -s` (BASELINE configs[0])"),
    "c2-vcf": ("sv", 500, 100_000, 50_000, "SV VCF text, 500 samples x 100k SVs, 5% missing, through `phyloforge -S` (BASELINE configs[1])"),
}
METRIC = "sample_pair_sites_per_s"
UNIT = "pair-sites/s"
SEED = 42
DTYPE = ("e2m1 x e2m1 -> f32 tensor-core indicator products (exact: integers < 2^24 per CTA) -> int32 counts; "
         "f64 distances + NJ; u32 bit planes")
# dram__bytes_read.sum + dram__bytes_write.sum per pair-count launch, from the committed
# ncu --set full capture of this command (profiles/); None until measured
NCU_TRAFFIC_BYTES_TC = 70_360_000_000  # r02, this build, the bench's own launch (3,000 x 10M, N=1): 66.3 / 72.4 / 71.5 GB read + 0.29 GB written in three captured launches of pair_counts_tc2_kernel (profiles/r02_ncu_step_kernels_raw_subset.csv; r01: 54.05 GB)
# pack_kernel over 10,000,000 sites x 3,000 samples (the bench's launch): 33.60 GB read + 7.50 GB written against 30.0 + 7.5 GB
# algorithmic (same capture); planes_to_e2m1_x4_kernel over the same sites: 7.52 GB read + 16.27 GB written against 7.5 + 15.4 GB
# (4.00 ms: 5.95 TB/s, 0.91 of the measured HBM peak)
NCU_TRAFFIC_BYTES_PACK = 41_105_000_000
NCU_TRAFFIC_SITES_PACK = 10_000_000
NCU_TRAFFIC_BYTES_CONV = 23_787_000_000
NCU_TRAFFIC_BYTES = 47_233_049_744  # r01: 46.45 GB read + 0.78 GB written (profiles/r01_ncu_pair_counts_csa_raw_subset.csv), N=1


def pair_sites(n, m):
    return n * (n - 1) // 2 * m


def workload_config(name: str, world: int, scaling: str) -> dict:
    """The `config` object of the JSON line. Both arms (ours / --impl reference) print exactly this dict."""
    if name in VCF_WORKLOADS:
        kind, n, m, miss, desc = VCF_WORKLOADS[name]
        return {"workload": desc, "samples": n, "sites": m, "miss_ppm": miss, "gpus": world, "scaling": scaling,
                "parallelism": "one GPU (the VCF text workloads run through the single-process entry point)",
                "l2": "every step re-reads the VCF file (page cache) and re-parses it; inputs exceed L2 except c1-vcf (23 MB), "
                      "whose step writes > 126 MB of intermediate buffers between passes",
                "seed": SEED}
    n, m, miss, desc = WORKLOADS[name]
    m_total = m * world if scaling == "weak" else m
    ld = (n + 3) // 4 * 4
    per_gpu = m if scaling == "weak" else -(-m // world)
    return {"workload": desc + ("" if world == 1 or scaling == "strong" else
                                f" per GPU: {m_total} sites over {world} GPUs (weak scaling)"),
            "samples": n, "sites": m_total, "miss_ppm": miss, "gpus": world, "scaling": scaling,
            "parallelism": f"sites sharded over {world} GPU(s), int32 counts allreduce (NCCL), NJ replicated",
            "l2": ("inputs (%.1f GB/GPU of byte genotypes) exceed L2; no flush needed" % (per_gpu * ld / 1e9)
                   if per_gpu * ld > 256e6 else
                   "inputs (%.0f MB/GPU) are smaller than L2 and not flushed: a parity-size workload, not a bench line"
                   % (per_gpu * ld / 1e6)),
            "seed": SEED}


# ---------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------- CPU arm
def cpu_threads_all() -> int:
    """Use every host core for the CPU arm, whatever OMP_NUM_THREADS says: torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which made the round-1 reference arm 16x slower under torchrun."""
    from oracle import cpu as oracle
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return oracle.set_num_threads(cores)


def cpu_port_run(n, m, miss_ppm, sample_sites, with_nj=True):
    """The CPU oracle on `sample_sites` sites of the workload (all host threads) + NJ at the
    full sample count. Returns timings and the extrapolated whole-path throughput."""
    import numpy as np
    from oracle import cpu as oracle
    sample_sites = max(64, min(sample_sites, m))
    t0 = time.perf_counter()
    codes = oracle.synth_codes(n, 0, sample_sites, seed=SEED, miss_ppm=miss_ppm)      # site-major (untimed input)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    sm = np.ascontiguousarray(codes.T)                       # site-major -> sample-major + bit planes = "pack"
    lo, hi = oracle.planes64(sm)
    t_pack = time.perf_counter() - t0
    t0 = time.perf_counter()
    counts = oracle.pair_counts_bits(lo, hi)
    t_counts = time.perf_counter() - t0
    t0 = time.perf_counter()
    dist, _ = oracle.normalise(counts, "ibs")
    t_norm = time.perf_counter() - t0
    t_nj = 0.0
    if with_nj:
        t0 = time.perf_counter()
        oracle.nj(dist)
        t_nj = time.perf_counter() - t0
    per_site = (t_pack + t_counts) / sample_sites
    full_s = per_site * m + t_norm + t_nj
    return {"threads": oracle.num_threads(), "sample_sites": sample_sites, "gen_s": t_gen, "pack_s": t_pack,
            "counts_s": t_counts, "normalise_s": t_norm, "nj_s": t_nj, "extrapolated_full_s": full_s,
            "value": pair_sites(n, m) / full_s,
            "counts_only_value": pair_sites(n, sample_sites) / t_counts}


CPU_PORT_NOTE = ("The reference's own engine for counts + NJ (treebest) is absent: this is the restated oracle "
                 "(oracle/pf_oracle.c, gcc -O3 -fopenmp), whose counts loop is an unblocked row-pair popcount - a soft "
                 "baseline, stated, never the target")


def run_reference_arm(args):
    """`--impl reference`: the CPU port of the path, all host threads, bounded sample per step.
    The sample per step is sized from a short calibration pass so that the whole run (W + K steps + one full NJ)
    stays near --cpu-arm-seconds whatever K and W the driver passes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload in VCF_WORKLOADS:
        return run_reference_arm_vcf(args)
    threads = cpu_threads_all()
    n, m, miss, desc = WORKLOADS[args.workload]
    if args.scaling == "weak":
        m = m * max(1, args.gpus)          # same global configuration as our arm at this N
    t_start = time.perf_counter()
    sample = args.cpu_sample_sites
    if sample <= 0:
        probe = cpu_port_run(n, m, miss, 1 << 14, with_nj=False)
        per_site = (probe["gen_s"] + probe["pack_s"] + probe["counts_s"]) / probe["sample_sites"]
        budget = max(5.0, args.cpu_arm_seconds * 0.8) / max(1, args.warmup + args.steps)
        sample = int(min(1 << 18, max(1 << 12, budget / per_site)) // 64 * 64)
    runs = []
    for i in range(args.warmup + args.steps):
        r = cpu_port_run(n, m, miss, sample, with_nj=(i == args.warmup))   # NJ (independent of M) once
        if i >= args.warmup:
            runs.append(r)
    nj_s = runs[0]["nj_s"]
    per_site = statistics.mean((r["pack_s"] + r["counts_s"]) / r["sample_sites"] for r in runs)
    norm_s = statistics.mean(r["normalise_s"] for r in runs)
    full_s = per_site * m + norm_s + nj_s
    value = pair_sites(n, m) / full_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": statistics.mean(r["pack_s"] + r["counts_s"] + r["normalise_s"] for r in runs) * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u64 bit planes + popcount -> int64 counts; f64 distances + NJ",
        "data": "synthetic", "config": workload_config(args.workload, max(1, args.gpus), args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": (f"{runs[0]['sample_sites']} of {m} sites x all {n} samples per step "
                                    f"(pack + counts, scaled linearly in sites) + normalise + one full NJ at N={n} "
                                    f"({nj_s:.2f} s); whole-path time extrapolated to {full_s:.1f} s. " + CPU_PORT_NOTE),
                         "nj_seconds": nj_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


# ---------------------------------------------------------------------- GPU arm
def _measure(eng, args, n, m, miss, scaling, steps, warmup, world, rank, with_e2e, with_split):
    """Time `steps` passes of the hot path (after `warmup` untimed ones) on this rank's site shard of an
    n x m workload; returns a dict of measurements (times are max over ranks where noted)."""
    import torch
    import torch.distributed as dist
    from phyloforge_b200 import sharding

    dev = eng.device
    names = [f"S{i:05d}" for i in range(n)]
    if scaling == "weak":
        # every rank holds one full copy's worth of sites of the named configuration: the global
        # matrix has world x m sites (rank r owns sites [r m, (r + 1) m)); N = 1 is the named config
        m_total = m * world
        site0, site1 = rank * m, (rank + 1) * m
    else:
        # the named configuration is split: contiguous word-aligned site ranges (phyloforge_b200/sharding.py)
        m_total = m
        wb, we = sharding.shard_words((m + 31) // 32, world, rank)
        site0, site1 = wb * 32, min(we * 32, m)
    my_sites = site1 - site0
    my_words = (my_sites + 31) // 32
    ld = (n + 3) // 4 * 4

    # ---- inputs resident in HBM (untimed): this rank's site shard, site-major byte codes
    codes = eng.synth_codes(n, site0, my_sites, seed=SEED, miss_ppm=miss)
    planes = eng.alloc_planes(n, max(my_words, 1))
    counts = eng.alloc_counts(n)
    torch.cuda.synchronize()

    ran_variant = [args.variant]
    stage_ev = {k: [] for k in ("pack", "counts", "allreduce", "normalise", "nj")}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(record: bool):
        counts.zero_()
        e0 = ev()
        eng.pack_codes(codes, n, planes, my_words, n_sites=my_sites)
        e1 = ev()
        ran_variant[0] = eng.pair_counts(planes, n, my_words, counts, variant=args.variant)
        e2 = ev()
        eng.allreduce_counts(counts)
        e3 = ev()
        dist_m, nz = eng.normalise(counts, "ibs")
        e4 = ev()
        merge, length = eng.nj(dist_m, destroy=True)
        e5 = ev()
        if record:
            for k, a, b in (("pack", e0, e1), ("counts", e1, e2), ("allreduce", e2, e3), ("normalise", e3, e4),
                            ("nj", e4, e5)):
                stage_ev[k].append((a, b))
        return merge, length

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(False)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    time.sleep(0.25)
    launches0 = int(eng.lib.pf_launch_count())
    _nj_counters(eng)   # (reset)
    t_wall0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        step(True)
    stop.record()
    barrier()
    launches = int(eng.lib.pf_launch_count()) - launches0
    nj_counters = _nj_counters(eng)
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    elapsed_ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed_ms.item()) / steps
    stages_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in stage_ev.items()}
    del codes
    out = {"n": n, "m_total": m_total, "my_sites": my_sites, "my_words": my_words, "ld": ld, "ms_per_step": ms_per_step,
           "value": pair_sites(n, m_total) / (ms_per_step * 1e-3), "stages_ms": stages_ms, "clocks": clocks,
           "variant": ran_variant[0], "gpu_launches": launches, "tc_split": None, "e2e": None,
           "tc_launches": getattr(eng, "tc_launches", 1) if ran_variant[0] == 3 else 0,
           "nj_counters": nj_counters, "nj_calls": steps}
    # split of the counts stage into its two kernels (untimed extra passes, CUDA events on the launching stream
    # recorded inside pf_pair_counts_tc): operand-stream conversion and the tensor-core GEMM launch(es)
    if with_split and ran_variant[0] == 3:
        import ctypes as C
        from phyloforge_b200._lib import check
        check(eng.lib.pf_pair_counts_tc_timing(1, None))
        conv, gemm = [], []
        for _ in range(3):
            counts.zero_()
            eng.pair_counts(planes, n, my_words, counts, variant=args.variant)
            ms2 = (C.c_float * 2)()
            check(eng.lib.pf_pair_counts_tc_timing(1, ms2))
            conv.append(float(ms2[0]))
            gemm.append(float(ms2[1]))
        check(eng.lib.pf_pair_counts_tc_timing(0, None))
        out["tc_split"] = {"conversion_ms": statistics.mean(conv), "gemm_ms": statistics.mean(gemm)}
    del planes, counts

    # ---- end to end from pinned host memory through the public host-buffer API
    if with_e2e:
        chunk_sites = args.e2e_chunk_sites
        packed2 = args.e2e_format == "2bit"
        ld_host = ((n + 15) // 16 * 4) if packed2 else ld
        host = eng.pinned_host_buffer((my_sites, ld_host))
        for c0 in range(0, my_sites, chunk_sites):                 # fill the host buffer (untimed)
            c1 = min(my_sites, c0 + chunk_sites)
            tmp = eng.synth_codes(n, site0 + c0, c1 - c0, seed=SEED, miss_ppm=miss)
            host[c0:c1].copy_(eng.to_2bit_rows(tmp, n) if packed2 else tmp)
            del tmp
        torch.cuda.synchronize()
        e2e_steps = max(1, min(steps, args.e2e_steps))
        res = eng.tree_from_host_codes(host, n, names, metric="ibs", chunk_sites=chunk_sites, variant=args.variant,
                                       packed2=packed2)
        barrier()
        t0 = time.perf_counter()
        h2d_ms = []
        for _ in range(e2e_steps):
            res = eng.tree_from_host_codes(host, n, names, metric="ibs", chunk_sites=chunk_sites, variant=args.variant,
                                           packed2=packed2)
            h2d_ms.append(res.timings_ms.get("h2d", 0.0))
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        h2d_bytes = int(my_sites) * ld_host
        my_rate = torch.tensor([h2d_bytes / max(1e-9, statistics.mean(h2d_ms) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        rates = [my_rate.clone() for _ in range(world)]
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            dist.all_gather(rates, my_rate)
        out["e2e"] = {"value": pair_sites(n, m_total) / float(t_e2e.item()), "unit": UNIT,
                      "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(res.d2h_bytes),
                      "host_format": ("site-major 2-bit genotypes, 4 samples per byte" if packed2 else "site-major u8 genotype codes"),
                      "ms_per_step": float(t_e2e.item()) * 1e3, "steps": e2e_steps,
                      "h2d_gbps_per_rank": [round(float(r.item()), 2) for r in rates],
                      "h2d_note": "bytes of this rank's rows / time from its first to its last host->device copy "
                                  "(copy-stream events); the copies overlap pack + counts of the previous chunk",
                      "api": "Engine.tree_from_host_codes (pinned host genotype rows -> Newick + distance matrix)",
                      "timing": "host wall clock between barriers + cuda synchronize (includes H2D, D2H, host Newick assembly)"}
        del host
    eng.release_scratch()
    torch.cuda.empty_cache()
    return out


def _nj_counters(eng):
    """Counters of the bounded-scan NJ kernel since the last reading (pf_nj_profile [4..7]): row rescans, list
    evaluations, iterations, matrices it joined."""
    import ctypes as C
    from phyloforge_b200._lib import check
    out = (C.c_ulonglong * 8)()
    check(eng.lib.pf_nj_profile(0, out))
    return [int(x) for x in list(out)[4:8]]


def _roofline_nj(n, nj_ms, hbm_peak, peak_src, counters=None, calls=1):
    """NJ against the bytes of a full scan: 4 na^2 bytes of live upper triangle per iteration. The
    full-scan kernel (below 3,500 nodes) reads them from L2 and is bound by its per-iteration latency chain; the
    bounded-scan kernel (from 3,500 nodes) reads a small fraction of them (`rows_rescanned_per_iteration` of na rows)
    - `achieved` is then an EFFECTIVE rate, which is how it can exceed what a full scan could reach."""
    scan_bytes = sum(4 * na * na for na in range(4, n + 1))
    ach = scan_bytes / (nj_ms * 1e-3) / 1e9
    bounded = bool(counters and counters[3] > 0)
    out = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
           "kernel": ("nj_lazy_kernel (bounded scan, one 16-CTA cluster)" if bounded else
                      ("nj_fast_kernel" if n <= 11000 else "nj_kernel")), "kernel_ms": nj_ms,
           "us_per_iteration": nj_ms * 1e3 / max(1, n - 3), "iterations": n - 3, "algorithmic_bytes": scan_bytes,
           "peak_source": peak_src,
           "note": "algorithmic bytes = sum over iterations of 4 na^2 (fp64 upper triangle of the live matrix), ~4 n^3 / 3, "
                   "i.e. what a full scan reads. Below ~3,900 nodes the matrix (8 n^2 bytes) is L2-resident and the "
                   "full-scan kernel is bound by its per-iteration latency chain (one grid-wide exchange + update), not by "
                   "HBM; the bounded-scan kernel reads only the rows whose lower bound does not exclude the minimum, so "
                   "its rate against these bytes is an effective one: us_per_iteration is the number to read"}
    if bounded:
        it = max(1, counters[2])
        out["bounded_scan"] = {"rows_rescanned_per_iteration": counters[0] / it, "list_evaluations_per_iteration": counters[1] / it,
                               "calls_counted": calls}
    return out


def _fp4_peak(eng):
    """Sustained rate of tcgen05.mma.cta_group::2 kind::mxf4 on every CTA pair of the device at once, timed with
    CUDA events around a ~0.4 s launch (probes library; operands resident in shared memory)."""
    import ctypes as C
    import torch
    from phyloforge_b200 import _lib
    try:
        probes = _lib.load_probes()
        tops, ms = C.c_double(), C.c_double()
        out = {}
        for n_cols in (208, 240):
            _lib.check_probe(probes.pf_probe_umma_fp4_2cta_peak(n_cols, 200_000, 4, C.byref(tops), C.byref(ms),
                                                               torch.cuda.current_stream().cuda_stream))
            reps = int(max(200_000, 200_000 * 400.0 / max(1e-3, ms.value)))
            _lib.check_probe(probes.pf_probe_umma_fp4_2cta_peak(n_cols, reps, 4, C.byref(tops), C.byref(ms),
                                                               torch.cuda.current_stream().cuda_stream))
            out[n_cols] = {"tops": tops.value, "ms": ms.value}
        return out
    except Exception as e:  # noqa: BLE001 - a missing probe must not lose the bench line
        return {"error": str(e)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from phyloforge_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload in VCF_WORKLOADS:
        rc = run_ours_vcf(args, world, rank, local_rank)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc
    eng = Engine(local_rank)

    n, m, miss, desc = WORKLOADS[args.workload]
    main = _measure(eng, args, n, m, miss, args.scaling, args.steps, args.warmup, world, rank,
                    with_e2e=not args.no_e2e, with_split=True)
    extra = {}
    if not args.no_extra:
        xs, xw = max(1, min(args.steps, args.extra_steps)), 2
        if world > 1 and args.scaling == "strong":
            w = _measure(eng, args, n, m, miss, "weak", xs, xw, world, rank, with_e2e=not args.no_e2e, with_split=False)
            extra["weak"] = {"config": workload_config(args.workload, world, "weak"), "value": w["value"], "unit": UNIT,
                             "ms_per_step": w["ms_per_step"], "steps": xs, "warmup": xw, "stages_ms": w["stages_ms"],
                             "e2e": w["e2e"], "scaling": "weak"}
        if args.workload == "c4":
            n5, m5, miss5, _ = WORKLOADS["c5"]
            c5 = _measure(eng, args, n5, m5, miss5, "strong", xs, 1, world, rank, with_e2e=False, with_split=False)
            extra["c5"] = {"config": workload_config("c5", world, "strong"), "value": c5["value"], "unit": UNIT,
                           "ms_per_step": c5["ms_per_step"], "steps": xs, "warmup": 1, "stages_ms": c5["stages_ms"],
                           "nj_seconds": c5["stages_ms"]["nj"] * 1e-3, "gemm_launches": c5["tc_launches"],
                           "scaling": "strong", "pair_counts_variant": c5["variant"],
                           "nj_counters": c5["nj_counters"], "nj_calls": c5["nj_calls"]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    my_sites, my_words, ld = main["my_sites"], main["my_words"], main["ld"]
    stages_ms, counts_ms, tc_split = main["stages_ms"], main["stages_ms"]["counts"], main["tc_split"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"
    pair_words = n * (n - 1) // 2 * my_words       # algorithmic: unordered pairs x words of this rank
    achieved_pw = pair_words / (counts_ms * 1e-3)
    variant_ran = main["variant"]
    if variant_ran == 3:
        # tensor-core indicator GEMMs on 4-bit (E2M1) operands: 2 multiply-accumulates per unordered pair
        # per site (4 with sample-specific missingness)
        bf16 = peaks.get("bf16_tflops", 1590.0)
        peak_tops = 4.0 * bf16                      # dense fp4 = 4 x dense bf16 on the same tensor cores (9 vs 2.25 PFLOP/s nominal)
        launches = main["tc_launches"]              # 2 = the missing-data launch (TT, VV) ran too
        macs_per = 4.0 if launches == 2 else 2.0
        macs = macs_per * pair_sites(n, my_sites)
        gemm_ms = tc_split["gemm_ms"] if tc_split else counts_ms
        achieved_tops = 2.0 * macs / (gemm_ms * 1e-3) / 1e12
        stage_tops = 2.0 * macs / (counts_ms * 1e-3) / 1e12
        fp4 = _fp4_peak(eng)
        fp4_tops = fp4.get(208, {}).get("tops") if isinstance(fp4, dict) else None
        roofline = {"bound": "tensor", "achieved": achieved_tops, "peak": peak_tops, "unit": "TOP/s (fp4)",
                    "frac": achieved_tops / peak_tops, "traffic": NCU_TRAFFIC_BYTES_TC,
                    "kernel": "pair_counts_tc2_kernel (tcgen05.mma.cta_group::2 kind::mxf4 on CTA pairs, E2M1 operands)",
                    "kernel_ms": gemm_ms, "stream_conversion_ms": tc_split["conversion_ms"] if tc_split else None,
                    "counts_stage_ms": counts_ms, "frac_of_counts_stage": stage_tops / peak_tops,
                    "peak_fp4_measured": fp4_tops,
                    "frac_of_fp4_measured": (achieved_tops / fp4_tops) if fp4_tops else None,
                    "frac_of_counts_stage_vs_fp4_measured": (stage_tops / fp4_tops) if fp4_tops else None,
                    "fp4_probe": fp4,
                    "ops_per_pair_site": {"fp4_mac": macs_per}, "gemm_launches": launches,
                    "frac_of_int8_tensor_peak": achieved_tops / (2.0 * bf16),
                    "peak_source": "peak = 4 x MEASURED_PEAKS.json bf16_tflops (%s; the file has no fp4 entry; nominal dense "
                                   "fp4 is 9000 TOP/s, 4 x the 2250 of bf16). peak_fp4_measured = the same instruction "
                                   "(256 x 208 x 64 kind::mxf4 on CTA pairs) issued back to back on resident operands by every "
                                   "pair of the device for ~0.4 s, CUDA events around the launch, in this run (probes "
                                   "library): the harsher denominator" % ("of measured" if peaks else "of fallback 1590"),
                    "note": "algorithmic work = unordered sample pairs x sites x 2 MAC (x.x' and h.h'; 4 MAC with "
                            "sample-specific missingness: + t.t', v.v'); the kernel also computes ~16% redundant "
                            "pairs in the 256 x 208 tiles that straddle the diagonal. frac = the GEMM kernel alone "
                            "(kernel_ms, CUDA events around its launch inside pf_pair_counts_tc, untimed extra passes); "
                            "frac_of_counts_stage also charges the planes -> 4-bit operand stream conversion "
                            "(planes_to_e2m1_kernel, stream_conversion_ms) to it, as the timed step does"}
    else:
        import ctypes as C
        from phyloforge_b200 import _lib
        probes = _lib.load_probes()

        def int_pipe(kind):
            a, b = C.c_double(), C.c_double()
            _lib.check_probe(probes.pf_microbench_int_pipe(kind, 1500, C.byref(a), C.byref(b),
                                                           torch.cuda.current_stream().cuda_stream))
            return a.value
        popc_ops, lop3_ops = int_pipe(0), int_pipe(1)
        plain_popc_peak_pw = popc_ops / (2.0 if miss == 0 else 3.0)
        # Instructions the kernel needs per unordered pair per 32-site word (DESIGN.md "Kernels"; counted in the SASS of
        # pair_counts_kernel): plain popcount (variant 1) 2 LOP3 + 2 POPC without missing data, 4.5 LOP3 + 3 POPC with;
        # carry-save (variant 0/2) 4 LOP3 + 1 POPC (+1 IMAD) without missing data; chunks with missing genotypes run the plain body
        csa = variant_ran != 1
        if miss == 0:
            lop_per_pw, popc_per_pw = (4.0, 1.0) if csa else (2.0, 2.0)
        else:
            lop_per_pw, popc_per_pw = 4.5, 3.0
        peak_pw = min(lop3_ops / lop_per_pw, popc_ops / popc_per_pw)
        binding = "LOP3" if lop3_ops / lop_per_pw < popc_ops / popc_per_pw else "POPC"
        roofline = {"bound": "int-pipe (%s binding)" % binding, "achieved": achieved_pw * 32 / 1e12,
                    "peak": peak_pw * 32 / 1e12, "unit": "Tpair-sites/s", "frac": achieved_pw / peak_pw,
                    "traffic": NCU_TRAFFIC_BYTES,
                    "kernel": "pair_counts_kernel<%s>" % ("carry-save" if csa else "plain"), "kernel_ms": counts_ms,
                    "ops_per_pair_word": {"lop3": lop_per_pw, "popc": popc_per_pw},
                    "frac_of_plain_popcount_roofline": achieved_pw / plain_popc_peak_pw,
                    "peak_source": "pf_microbench_int_pipe measured in this run: POPC %.3g ops/s, LOP3 %.3g ops/s; "
                                   "peak = min(LOP3 rate / %.1f, POPC rate / %.1f) pair-words/s x 32 sites "
                                   "(MEASURED_PEAKS.json has no integer-pipe peak)" % (popc_ops, lop3_ops, lop_per_pw,
                                                                                      popc_per_pw),
                    "note": "algorithmic pair-words = unordered sample pairs x 32-site words of this rank; "
                            "frac_of_plain_popcount_roofline is against SURVEY 8(d)'s POPC-only bound "
                            "(POPC rate / 2 or 3 per pair-word), which the carry-save kernel may exceed"}
    pack_bytes = my_sites * ld + my_words * ((n + 63) // 64) * 64 * 8
    roofline_pack = {"bound": "hbm", "achieved": pack_bytes / (stages_ms["pack"] * 1e-3) / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": pack_bytes / (stages_ms["pack"] * 1e-3) / 1e9 / hbm_peak,
                     "traffic": int(NCU_TRAFFIC_BYTES_PACK * my_sites / NCU_TRAFFIC_SITES_PACK) if n == 3000 else None,
                     "traffic_note": "ncu dram bytes (read + write) of the pack_kernel launch over %d sites x 3,000 samples "
                                     "(41.1 GB against 37.5 GB algorithmic, x 1.10; profiles/r02_ncu_step_kernels_raw_"
                                     "subset.csv), scaled linearly to the %d sites this rank packs; the operand-stream "
                                     "conversion of the same sites moved %.2f GB against 22.9 algorithmic in 4.00 ms "
                                     "(0.91 of the HBM peak; same capture)"
                                     % (NCU_TRAFFIC_SITES_PACK, my_sites, NCU_TRAFFIC_BYTES_CONV / 1e9),
                     "kernel": "pack_kernel", "kernel_ms": stages_ms["pack"], "peak_source": hbm_src}
    roofline_nj = _roofline_nj(n, stages_ms["nj"], hbm_peak, hbm_src, main.get("nj_counters"), main.get("nj_calls", 1))
    if "c5" in extra:
        extra["c5"]["roofline_nj"] = _roofline_nj(WORKLOADS["c5"][0], extra["c5"]["stages_ms"]["nj"], hbm_peak, hbm_src,
                                                  extra["c5"].pop("nj_counters", None), extra["c5"].pop("nj_calls", 1))

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads_all()
        sample = args.cpu_sample_sites if args.cpu_sample_sites > 0 else 1 << 18
        r = cpu_port_run(n, m, miss, sample, with_nj=True)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": (f"{r['sample_sites']} of {m} sites x all {n} samples (pack {r['pack_s']:.2f} s + "
                                   f"counts {r['counts_s']:.2f} s, scaled linearly in sites) + normalise + full NJ at "
                                   f"N={n} ({r['nj_s']:.2f} s): whole path extrapolated to {r['extrapolated_full_s']:.1f} s. "
                                   + CPU_PORT_NOTE),
                        "counts_only_value": r["counts_only_value"], "nj_seconds": r["nj_s"]}

    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": workload_config(args.workload, world, args.scaling),
        "detail": {"sites_per_gpu": my_sites, "pair_counts_variant": variant_ran,
                   "pair_counts_variant_requested": args.variant},
        "nj_seconds": stages_ms["nj"] * 1e-3, "stages_ms": stages_ms,
        "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
        "gpu_launches_note": "pf_launch_count() after - before the timed region on rank 0: every kernel launch of "
                             "libphyloforge_b200.so (pack, operand-stream conversions, tensor-core GEMM launches, normalise, NJ)",
        "roofline": roofline, "roofline_pack": roofline_pack, "roofline_nj": roofline_nj, "cpu_baseline": cpu_baseline,
        "extra": extra,
    }
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------- VCF text workloads
def _make_vcf(eng, name, path):
    """Synthetic VCF text of a VCF workload (genotype codes from the device generator, rendered by
    phyloforge_b200/synth.py), written to `path`. Returns (bytes, n, m)."""
    from phyloforge_b200 import synth
    kind, n, m, miss, _ = VCF_WORKLOADS[name]
    names = [f"S{i:04d}" for i in range(n)]
    codes = eng.synth_codes(n, 0, m, seed=SEED, miss_ppm=miss)[:, :n].cpu().numpy()
    text = synth.snp_vcf_bytes(codes, names)[0] if kind == "snp" else synth.sv_vcf_bytes(codes, names)[0]
    with open(path, "wb") as f:
        f.write(text)
    return len(text), n, m


def _vcf_paths(name):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else ROOT + "/gpurun_out"
    d = os.path.join(base, f"pf_bench_{name}_{os.getpid()}")
    os.makedirs(d, exist_ok=True)
    return d, os.path.join(d, "in.vcf"), os.path.join(d, "run.cfg"), os.path.join(d, "out")


def run_ours_vcf(args, world, rank, local_rank):
    import contextlib
    import io
    import shutil
    import torch
    from phyloforge_b200 import run as pf_run
    from phyloforge_b200 import vcf as pf_vcf
    from phyloforge_b200.engine import Engine
    if rank != 0:
        return 0                      # the text entry point is one process; other ranks have nothing to do
    eng = Engine(local_rank)
    kind, n, m, miss, desc = VCF_WORKLOADS[args.workload]
    d, vcf_path, cfg_path, out_dir = _vcf_paths(args.workload)
    try:
        nbytes, _, _ = _make_vcf(eng, args.workload, vcf_path)
        section, flag, tool = ("snp_opt", "-s", "treebest") if kind == "snp" else ("sv_opt", "-S", "nj")
        with open(cfg_path, "w") as f:
            f.write(f"[{section}]\nvcf_file = {vcf_path}\nout_path = {out_dir}\ntree_software = {tool}\nthread = 8\n"
                    "distance = ibs\n")
        load = pf_vcf.load_snp_vcf if kind == "snp" else pf_vcf.load_sv_vcf

        def dev_step():
            # device-timed leg: text (page cache -> pinned -> HBM) -> parse -> pack -> counts -> distances -> NJ
            p = load(eng, vcf_path, keep_chars=False)
            return p, eng.tree_from_planes(p.planes, p.n_samples, p.n_words, [str(i) for i in range(p.n_samples)],
                                           metric="ibs", n_sites=p.n_sites)

        def e2e_step():
            with contextlib.redirect_stdout(io.StringIO()):
                pf_run.main([flag, cfg_path])

        for _ in range(args.warmup):
            dev_step()
        torch.cuda.synchronize()
        launches0 = int(eng.lib.pf_launch_count())
        sampler = ClockSampler(local_rank)
        sampler.start()
        t_wall0 = time.perf_counter()
        t_parse = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            p, res = dev_step()
            torch.cuda.synchronize()
            t_parse.append(time.perf_counter() - t0)
        t_wall1 = time.perf_counter()
        launches = int(eng.lib.pf_launch_count()) - launches0
        clocks = sampler.stop(t_wall0, t_wall1)
        columns = p.n_sites
        # conversion alone (file -> packed planes), for the parser's roofline
        t_conv = []
        for _ in range(max(3, args.steps)):
            t0 = time.perf_counter()
            load(eng, vcf_path, keep_chars=False)
            torch.cuda.synchronize()
            t_conv.append(time.perf_counter() - t0)
        for _ in range(2):
            e2e_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        out_bytes = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)
                        if os.path.isfile(os.path.join(out_dir, f)))
        ms_per_step = statistics.mean(t_parse) * 1e3
        work = pair_sites(n, columns)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        conv_s = min(t_conv)
        cpu = None
        if not args.no_cpu_baseline:
            cpu = _cpu_convert(args.workload, vcf_path, d)
            cpu["gpu_convert_seconds"] = conv_s
            cpu["speedup_conversion"] = cpu["seconds_full"] / conv_s
        line = {
            "metric": METRIC, "value": work / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "text bytes -> u8 codes -> 2-bit planes; " + DTYPE, "data": "synthetic",
            "config": workload_config(args.workload, world, args.scaling),
            "detail": {"vcf_bytes": nbytes, "matrix_columns": columns, "records": m,
                       "value_is": "VCF file in the page cache -> GPU parse -> packed planes -> counts -> distances -> NJ join list "
                                   "(vcf.load_*_vcf + Engine.tree_from_planes), host wall clock with device synchronise"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": work / t_e2e, "unit": UNIT, "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
                    "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": out_bytes,
                    "api": f"phyloforge {flag} run.cfg (phyloforge_b200.run.main): VCF file -> all_snp.fa / SV.matrix, "
                           "distance matrix file, tree file on disk",
                    "d2h_note": "bytes of the artefacts written (character matrix, distance matrix, tree)"},
            "roofline_parse": {"bound": "hbm", "achieved": nbytes / conv_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": nbytes / conv_s / 1e9 / hbm_peak, "traffic": None,
                               "kernel": "pf_line_index + %s_parse_kernel + select + pack" % kind, "kernel_ms": conv_s * 1e3,
                               "genotypes_per_s": n * m / conv_s,
                               "note": "whole conversion from the file (page cache -> pinned buffer -> H2D -> kernels), best of "
                                       "%d; bounded by the host read + PCIe copy, not by the kernels" % len(t_conv)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), file=args.out, flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return 0


def _cpu_convert(name, vcf_path, work_dir):
    """The reference's CPU half of the VCF workloads: its converter. The UNMODIFIED reference when /root/reference
    is present (kind "reference", oracle/run_reference.py in its own process), else the restatement
    oracle/ref_convert.py (kind "port", single Python process). Bounded: the SV converter is quadratic in the
    reference (766 s at full size) and the port takes minutes too, so both run on a prefix of the records and the
    full-size figure is extrapolated linearly (a lower bound for the reference's super-linear SV converter)."""
    from oracle import run_reference
    kind, n, m, miss, _ = VCF_WORKLOADS[name]
    cores = os.cpu_count() or 1
    use_ref = run_reference.available()
    sample = m if (kind == "snp" and use_ref) else (20_000 if kind == "snp" else 4_000)
    sample = min(sample, m)
    sub = os.path.join(work_dir, "cpu_sample.vcf")
    with open(vcf_path, "rb") as f, open(sub, "wb") as g:
        k = 0
        for line in f:
            g.write(line)
            if not line.startswith(b"#"):
                k += 1
                if k >= sample:
                    break
    t0 = time.perf_counter()
    if use_ref:
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), kind, sub, os.path.join(work_dir, "ref_out")]
        if kind == "snp":
            cmd += [str(cores), "treebest"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        info = json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {"error": r.stderr[-300:]}
        seconds = float(info.get("total_s", time.perf_counter() - t0))
        used = cores if kind == "snp" else 1
        kind_s = "reference"
    else:
        from oracle import ref_convert
        text = open(sub, encoding="utf-8").read()
        t0 = time.perf_counter()
        (ref_convert.snp_convert if kind == "snp" else ref_convert.sv_convert)(text)
        seconds = time.perf_counter() - t0
        used, kind_s = 1, "port"
    full = seconds * m / sample
    return {"value": n * m / full, "unit": "genotypes/s (conversion only: the reference runs nothing else of this path itself)",
            "cores": used, "kind": kind_s, "seconds_sample": seconds, "seconds_full": full,
            "sample": f"first {sample} of {m} records x all {n} samples through "
                      + ("the unmodified reference converter (oracle/run_reference.py)" if use_ref else
                         "oracle/ref_convert.py (restated converter, one Python process)")
                      + (", scaled linearly to the full file" if sample < m else "")}


def run_reference_arm_vcf(args):
    import shutil
    kind, n, m, miss, desc = VCF_WORKLOADS[args.workload]
    d, vcf_path, cfg_path, out_dir = _vcf_paths(args.workload)
    try:
        # input text from the CPU generator (same counter-based codes as the device generator)
        from oracle import cpu as oracle
        from phyloforge_b200 import synth
        names = [f"S{i:04d}" for i in range(n)]
        codes = oracle.synth_codes(n, 0, m, seed=SEED, miss_ppm=miss)
        text = synth.snp_vcf_bytes(codes, names)[0] if kind == "snp" else synth.sv_vcf_bytes(codes, names)[0]
        with open(vcf_path, "wb") as f:
            f.write(text)
        cpu = _cpu_convert(args.workload, vcf_path, d)
        line = {"impl": "reference", "metric": "genotypes_per_s", "value": cpu["value"], "unit": "genotypes/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": cpu["seconds_sample"] * 1e3,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "text",
                "data": "synthetic", "config": workload_config(args.workload, max(1, args.gpus), args.scaling),
                "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": "genotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "conversion only (VCF -> character matrix): the reference delegates distance + NJ to the absent treebest / iqtree"}
        print(json.dumps(line), file=args.out, flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return 0


def _claim_stdout():
    """Keep the process's stdout for the single JSON line: libraries that print to fd 1 (NCCL's version
    banner) are sent to stderr for the rest of the run. Returns a file object bound to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + sorted(VCF_WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the named configuration's sites are split over the GPUs, as BASELINE.json "
                         "words configs[3]; weak: every GPU gets the named configuration's site count")
    ap.add_argument("--variant", type=int, default=0, help="pair-count kernel: 0 auto, 1 plain popcount, 2 carry-save popcount, 3 tcgen05 4-bit (E2M1) GEMMs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the bounded sub-records (weak scaling at N > 1, BASELINE configs[4])")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-format", default="2bit", choices=["2bit", "u8"],
                    help="host genotype format of the end-to-end path: 2 bits (4 samples per byte) or one byte per genotype")
    ap.add_argument("--e2e-chunk-sites", type=int, default=1 << 19)
    ap.add_argument("--cpu-sample-sites", type=int, default=0,
                    help="sites per CPU-arm step; 0 = sized from a calibration pass so that the arm takes ~--cpu-arm-seconds")
    ap.add_argument("--cpu-arm-seconds", type=float, default=60.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    args.out = _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
