#!/usr/bin/env python
"""Benchmark of the population-phylogeny hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], "3,000 samples x 10M SNPs", the shape the
north-star target is quoted on; it fits one B200 (30 GB of site-major byte genotypes + 7.5 GB of
bit planes), so the same total work is used at every N and the site axis is sharded over the
ranks. Default "scaling": "weak" — every GPU holds the named configuration's 10M sites (the global
matrix has N x 10M sites; N = 1 is exactly the named configuration); `--scaling strong` splits the 10M
sites over the GPUs instead (round-1 measurements of both are in profiles/).

A step = one pass of the hot path over the whole batch:
    pack (byte genotypes -> bit planes) -> all-pairs counts -> [NCCL allreduce of the
    int32 count matrices] -> fp64 distances -> neighbor-joining tree.
  value   whole-job sample-pair x sites per second, inputs resident in HBM, CUDA-event timed,
          max over ranks.
  e2e     the same metric through Engine.tree_from_host_codes: genotypes start in pinned host
          memory (2-bit site-major rows by default, --e2e-format u8 for one byte per genotype),
          host->device copies (chunked, overlapped with pack + counts) and the device->host read
          of the result (distance matrix, join list -> Newick) are timed.
  roofline      the pair-count kernel against the integer-pipe peak measured live by
                pf_microbench_int_pipe (MEASURED_PEAKS.json has no integer peaks).
  cpu_baseline  the CPU oracle (oracle/pf_oracle.c, all host threads) on a bounded sample,
                rank 0 at N=1 only.
`--impl reference` times that CPU port alone (the reference's own engine for distance + NJ is
the absent external `treebest`; oracle/__init__.py).
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
METRIC = "sample_pair_sites_per_s"
UNIT = "pair-sites/s"
SEED = 42
# dram__bytes_read.sum + dram__bytes_write.sum per pair-count launch, from the committed
# ncu --set full capture of this command (profiles/); None until measured
NCU_TRAFFIC_BYTES_TC = 54_051_894_000  # r01: 52.15 GB read + 1.90 GB written per launch of pair_counts_tc2_kernel (profiles/r01_ncu_pair_counts_tc2_raw_subset.csv), N=1
NCU_TRAFFIC_BYTES = 47_233_049_744  # r01: 46.45 GB read + 0.78 GB written (profiles/r01_ncu_pair_counts_csa_raw_subset.csv), N=1


def pair_sites(n, m):
    return n * (n - 1) // 2 * m


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
def cpu_port_run(n, m, miss_ppm, sample_sites, with_nj=True):
    """The CPU oracle on `sample_sites` sites of the workload (all host threads) + NJ at the
    full sample count. Returns timings and the extrapolated whole-path throughput."""
    import numpy as np
    from oracle import cpu as oracle
    sample_sites = min(sample_sites, m)
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


def run_reference_arm(args):
    """`--impl reference`: the CPU port of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, m, miss, desc = WORKLOADS[args.workload]
    if args.scaling == "weak":
        m = m * max(1, args.gpus)          # same global configuration as our arm at this N
    sample = args.cpu_sample_sites
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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32 bit-planes / int64 counts / f64",
        "data": "synthetic", "config": {"workload": desc, "samples": n, "sites": m, "miss_ppm": miss},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": runs[0]["threads"], "kind": "port",
                         "sample": (f"{runs[0]['sample_sites']} of {m} sites x all {n} samples per step "
                                    f"(pack + counts, scaled linearly in sites) + normalise + one full NJ at N={n} "
                                    f"({nj_s:.2f} s); whole-path time extrapolated to {full_s:.1f} s. The reference's "
                                    "own engine for this step (treebest) is absent: this is the restated oracle"),
                         "nj_seconds": nj_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


# ---------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from phyloforge_b200.engine import Engine
    from phyloforge_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    dev = eng.device

    n, m, miss, desc = WORKLOADS[args.workload]
    names = [f"S{i:05d}" for i in range(n)]
    if args.scaling == "weak":
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

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    t_wall0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        merge, length = step(True)
    stop.record()
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    elapsed_ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = pair_sites(n, m_total) / (ms_per_step * 1e-3)
    stages_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in stage_ev.items()}
    counts_ms = stages_ms["counts"]
    del codes
    # split of the counts stage into its two kernels (untimed extra passes, CUDA events on the launching stream
    # recorded inside pf_pair_counts_tc): operand-stream conversion and the tensor-core GEMM launch(es)
    tc_split = None
    if ran_variant[0] == 3:
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
        tc_split = {"conversion_ms": statistics.mean(conv), "gemm_ms": statistics.mean(gemm)}

    # ---- end to end from pinned host memory through the public host-buffer API
    e2e = None
    if not args.no_e2e:
        chunk_sites = args.e2e_chunk_sites
        packed2 = args.e2e_format == "2bit"
        ld_host = ((n + 15) // 16 * 4) if packed2 else ld
        host = torch.empty((my_sites, ld_host), dtype=torch.uint8).pin_memory()
        for c0 in range(0, my_sites, chunk_sites):                 # fill the host buffer (untimed)
            c1 = min(my_sites, c0 + chunk_sites)
            tmp = eng.synth_codes(n, site0 + c0, c1 - c0, seed=SEED, miss_ppm=miss)
            host[c0:c1].copy_(eng.to_2bit_rows(tmp, n) if packed2 else tmp)
            del tmp
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        res = None
        for _ in range(1):
            res = eng.tree_from_host_codes(host, n, names, metric="ibs", chunk_sites=chunk_sites, variant=args.variant,
                                           packed2=packed2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = eng.tree_from_host_codes(host, n, names, metric="ibs", chunk_sites=chunk_sites, variant=args.variant,
                                           packed2=packed2)
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": pair_sites(n, m_total) / float(t_e2e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(my_sites) * ld_host, "d2h_bytes_per_step": int(res.d2h_bytes),
               "host_format": ("site-major 2-bit genotypes, 4 samples per byte" if packed2 else "site-major u8 genotype codes"),
               "ms_per_step": float(t_e2e.item()) * 1e3, "steps": e2e_steps,
               "api": "Engine.tree_from_host_codes (pinned host genotype rows -> Newick + distance matrix)",
               "timing": "host wall clock between barriers + cuda synchronize (includes H2D, D2H, host Newick assembly)"}
        del host

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pair counts), against the live integer-pipe peak
    popc_ops, _ = eng.microbench_int_pipe(0, 1500)
    lop3_ops, _ = eng.microbench_int_pipe(1, 1500)
    pair_words = n * (n - 1) // 2 * my_words       # algorithmic: unordered pairs x words of this rank
    # Instructions the kernel needs per unordered pair per 32-site word (DESIGN.md "Kernels";
    # counted in the SASS of pair_counts_kernel):
    #   plain popcount (variant 1): 2 LOP3 + 2 POPC without missing data, 4.5 LOP3 + 3 POPC with
    #   carry-save (variant 0/2)  : 4 LOP3 + 1 POPC (+1 IMAD) without missing data; chunks that hold
    #                               missing genotypes run the plain body
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    achieved_pw = pair_words / (counts_ms * 1e-3)
    plain_popc_peak_pw = popc_ops / (2.0 if miss == 0 else 3.0)
    variant_ran = ran_variant[0]
    if variant_ran == 3:
        # tensor-core indicator GEMMs on 4-bit (E2M1) operands: 2 multiply-accumulates per unordered pair
        # per site (4 with sample-specific missingness)
        bf16 = peaks.get("bf16_tflops", 1590.0)
        peak_tops = 4.0 * bf16                      # dense fp4 = 4 x dense bf16 on the same tensor cores (9 vs 2.25 PFLOP/s nominal)
        launches = getattr(eng, "tc_launches", 1)    # 2 = the missing-data launch (TT, VV) ran too
        macs_per = 4.0 if launches == 2 else 2.0
        macs = macs_per * pair_sites(n, my_sites)
        gemm_ms = tc_split["gemm_ms"] if tc_split else counts_ms
        achieved_tops = 2.0 * macs / (gemm_ms * 1e-3) / 1e12
        stage_tops = 2.0 * macs / (counts_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved_tops, "peak": peak_tops, "unit": "TOP/s (fp4)",
                    "frac": achieved_tops / peak_tops, "traffic": NCU_TRAFFIC_BYTES_TC,
                    "kernel": "pair_counts_tc2_kernel (tcgen05.mma.cta_group::2 kind::mxf4 on CTA pairs, E2M1 operands)",
                    "kernel_ms": gemm_ms, "stream_conversion_ms": tc_split["conversion_ms"] if tc_split else None,
                    "counts_stage_ms": counts_ms, "frac_of_counts_stage": stage_tops / peak_tops,
                    "ops_per_pair_site": {"fp4_mac": macs_per}, "gemm_launches": launches,
                    "frac_of_popcount_kernel_roofline": achieved_pw / (min(lop3_ops / 4.0, popc_ops / 1.0) if miss == 0
                                                                       else plain_popc_peak_pw),
                    "frac_of_plain_popcount_roofline": achieved_pw / plain_popc_peak_pw,
                    "frac_of_int8_tensor_peak": achieved_tops / (2.0 * bf16),
                    "peak_source": "4 x MEASURED_PEAKS.json bf16_tflops (%s; the file has no fp4 entry; nominal dense "
                                   "fp4 is 9000 TOP/s, 4 x the 2250 of bf16)" % ("of measured" if peaks else "of fallback 1590"),
                    "note": "algorithmic work = unordered sample pairs x sites x 2 MAC (x.x' and h.h'; 4 MAC with "
                            "sample-specific missingness: + t.t', v.v'); the kernel also computes ~16% redundant "
                            "pairs in the 256 x 208 tiles that straddle the diagonal. frac = the GEMM kernel alone "
                            "(kernel_ms, CUDA events around its launch inside pf_pair_counts_tc, untimed extra passes); "
                            "frac_of_counts_stage also charges the planes -> 4-bit operand stream conversion "
                            "(planes_to_e2m1_kernel, stream_conversion_ms) to it, as the timed step does"}
    else:
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
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    pack_bytes = my_sites * ld + my_words * ((n + 63) // 64) * 64 * 8
    roofline_pack = {"bound": "hbm", "achieved": pack_bytes / (stages_ms["pack"] * 1e-3) / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": pack_bytes / (stages_ms["pack"] * 1e-3) / 1e9 / hbm_peak,
                     "traffic": None, "kernel": "pack_kernel", "kernel_ms": stages_ms["pack"],
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(n, m, miss, args.cpu_sample_sites, with_nj=True)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                        "sample": (f"{r['sample_sites']} of {m} sites x all {n} samples (pack {r['pack_s']:.2f} s + "
                                   f"counts {r['counts_s']:.2f} s, scaled linearly in sites) + normalise + full NJ at "
                                   f"N={n} ({r['nj_s']:.2f} s): whole path extrapolated to {r['extrapolated_full_s']:.1f} s"),
                        "counts_only_value": r["counts_only_value"], "nj_seconds": r["nj_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u32 bit-planes / int32 counts / f64 distances+NJ", "data": "synthetic",
        "config": {"workload": desc + ("" if world == 1 or args.scaling == "strong" else
                                       f" per GPU: {m_total} sites over {world} GPUs (weak scaling)"),
                   "samples": n, "sites": m_total, "miss_ppm": miss, "sites_per_gpu": my_sites,
                   "parallelism": f"sites sharded over {world} GPU(s), int32 counts allreduce (NCCL), NJ replicated",
                   "l2": "inputs (%.1f GB/GPU of byte genotypes) exceed L2; no flush needed" % (my_sites * ld / 1e9),
                   "pair_counts_variant": ran_variant[0], "pair_counts_variant_requested": args.variant, "seed": SEED},
        "nj_seconds": stages_ms["nj"] * 1e-3, "stages_ms": stages_ms,
        "clocks": clocks, "e2e": e2e, "gpu_launches": (7 if ran_variant[0] == 3 else 4) * args.steps,   # pack, (2 stream conversions, 2 GEMM launches | popcount), normalise, NJ
        "roofline": roofline, "roofline_pack": roofline_pack, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets the named configuration's site count (default); "
                         "strong: the named configuration is split over the GPUs")
    ap.add_argument("--variant", type=int, default=0, help="pair-count kernel: 0 auto, 1 plain popcount, 2 carry-save popcount, 3 tcgen05 4-bit (E2M1) GEMMs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-format", default="2bit", choices=["2bit", "u8"],
                    help="host genotype format of the end-to-end path: 2 bits (4 samples per byte) or one byte per genotype")
    ap.add_argument("--e2e-chunk-sites", type=int, default=1 << 19)
    ap.add_argument("--cpu-sample-sites", type=int, default=1 << 18)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    args.out = _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
