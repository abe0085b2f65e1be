#!/usr/bin/env python
"""bench.py -- target loci called / s on the synthetic N0030 panel (BASELINE.json config 2), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's algorithm on the host CPU (oracle port)

A "step" is one pass of the hot path (K1 prep -> K2 sorts -> K3 tile pileup -> K4 statistics) over one batch:
``--intervals`` seeded intervals of the N0030 panel BED per GPU, ~3 000 barcodes per locus, ~4 read pairs per
barcode, 2 x 150 bp.  ``value`` is measured with the batch resident in HBM, ``e2e`` through
``GpuCaller.call()`` with pinned host buffers (H2D + kernels + D2H inside the timed region).  The panel is sharded
across GPUs by BED interval (weak scaling: every rank gets its own ``--intervals`` intervals); loci are independent,
so there is no data-path collective -- torch.distributed is used for the barrier and the max-over-ranks only.
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

PANEL_BED = os.path.join(ROOT, "tests", "golden", "n0030_panel.bed")
METRIC = "target_loci_called_per_sec"
UNIT = "loci/s"
ALGO_BYTES_PER_EVENT = 35.0     # reads in once (2.7 B/event) + every 16-byte event written once and read once
UMIS_PER_LOCUS, RPB = 3000, 4.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--intervals", type=int, default=96, help="panel intervals per GPU in one batch")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-loci", type=int, default=0, help="loci in the CPU-baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--whole-reads", action="store_true", help="upload whole reads instead of the target windows (store_lo/store_len)")
    ap.add_argument("--pipeline-intervals", type=int, default=12,
                    help="intervals of the rank-0 batch run once through the whole CLI path (BAM decode -> files); 0 = skip")
    return ap.parse_args()


def make_batch(args, rank, world):
    import pickle
    from smcounter_b200.synth import SynthSpec, make_panel_mp, panel_intervals_from_bed
    from smcounter_b200.targets import build_loci
    cache = os.environ.get("SMC_BENCH_CACHE")           # tuning sessions: reuse the generated batch between runs
    if cache:
        cache = "%s.%d_%d_%d_%d_%d" % (cache, args.intervals, args.seed, rank, world, int(args.whole_reads))
        if os.path.exists(cache):
            with open(cache, "rb") as fh:
                return pickle.load(fh)
    ivs = panel_intervals_from_bed(PANEL_BED, limit=args.intervals * world, seed=args.seed)
    if world > 1:
        # the product's own multi-GPU plan (shard.assign_intervals): BED intervals to ranks, balanced by estimated work
        # (uniform depth here, so ~ interval length + a margin for the reads hanging over both ends), not by interval count.
        # Margin measured at N=4 on B200: 150 -> 5.14 ms max step, 300 -> 4.88, 600 -> 4.84 (ideal 4.54)
        from smcounter_b200.shard import assign_intervals
        margin = float(os.environ.get("SMC_BENCH_MARGIN", "500"))
        shard, _ = assign_intervals([float(e - s) + margin for (_, s, e) in ivs], world)
        mine = [iv for iv, g in zip(ivs, shard) if g == rank]
    else:
        mine = ivs
    spec = SynthSpec(umis_per_locus=UMIS_PER_LOCUS, rpb=RPB, snv_every=1000, snv_vaf=0.01, indel_every=12000, indel_vaf=0.01)
    soa, refs, truth = make_panel_mp(mine, spec, seed=args.seed + 17 * rank)
    full = soa                                          # whole reads: the CPU legs (oracle, BAM writer) need them
    if not args.whole_reads:
        # what the BAM decoder hands over: only the bases a pileup over the targets can see (store_lo / store_len), packed
        soa = soa.trim_to_targets(mine)
    loci, bed_order = build_loci(mine, soa.chroms, refs)
    if cache:
        with open(cache, "wb") as fh:
            pickle.dump((mine, soa, refs, loci, bed_order, full), fh, protocol=4)
    return mine, soa, refs, loci, bed_order, full


def vc_params():
    from smcounter_b200.caller import VcParams
    # --mtDepth 3000 --rpb 4.0 --mtDrop 0 --minBQ 20 --minMQ 30 --hpLen 10 (ds = 6000: no down-sampling)
    return VcParams(mtDepth=3000, rpb=4.0, minBQ=20, minMQ=30, hpLen=10, mismatchThr=6.0, mtDrop=0, maxMT=0, primerDist=2)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line): NVML every 5 ms when
    nvidia_ml_py is importable (the timed region is only a few hundred ms), else nvidia-smi every ~0.1 s."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.source = "nvidia-smi"

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.device
        if vis:
            try:
                idx = int(vis.split(",")[self.device])
            except Exception:
                pass
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        self.source = "nvml"
        while not self.stop_flag:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in names:
                if r & bit:
                    self.reasons.add(name)
            time.sleep(0.005)

    def _run_smi(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                s = [x.strip() for x in out.split(",")] if out else []
                if len(s) >= 8:
                    if s[1].replace(".", "").isdigit():
                        self.sm.append(float(s[1]))
                    if s[2].replace(".", "").isdigit():
                        self.mx.append(float(s[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


# --------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's vc() on the host cores (bounded sample of the same workload)
# --------------------------------------------------------------------------------------------------------------
_G = {}
CPU_WHAT = {"reference": "the reference's own vc_wrapper() (oracle/_ref = /root/reference/smCounter.py via oracle/ref_build.py; pysam stand-in over "
                         "in-memory reads, plain dict/set containers)",
            "port": "oracle/smcounter_oracle.py (Python-3 port; oracle/_ref not built)"}


def cpu_kind():
    """'reference': oracle/_ref (the reference's own smCounter.py made runnable by oracle/ref_build.py, plain containers,
    pysam stand-in) is present; 'port': only the oracle restatement is."""
    from oracle import ref_build
    return "reference" if ref_build.available() else "port"


def _cpu_worker(job):
    from oracle import smcounter_oracle as orc
    chrom, pos = job
    p = _G["prm"]
    if _G.get("ref") is not None:          # the reference's vc_wrapper(), exactly as its Pool calls it (smCounter.py:684)
        return _G["ref"].vc_wrapper("bench.bam", chrom, pos, p.minBQ, p.minMQ, p.mtDepth, p.rpb, p.hpLen, p.mismatchThr, p.mtDrop,
                                    p.maxMT, p.primerDist, "bench.fa")
    return orc.vc(_G["index"], chrom, pos, p.minBQ, p.minMQ, p.mtDepth, p.rpb, p.hpLen, p.mismatchThr, p.mtDrop, p.maxMT,
                  p.primerDist, _G["refs"])


def cpu_sample_setup(soa, refs, loci, n_loci_sample):
    """Pick consecutive loci from the middle of the batch and the reads that overlap them (host objects for the oracle)."""
    import numpy as np
    from oracle import smcounter_oracle as orc
    from smcounter_b200.soa import soa_to_records
    n_loci_sample = max(1, min(int(n_loci_sample), loci.n))
    a = max(0, loci.n // 2 - n_loci_sample // 2)
    sel = np.arange(a, min(loci.n, a + n_loci_sample))
    ends = soa.ref_end()
    mask = np.zeros(soa.n, dtype=bool)
    for rid in np.unique(loci.ref_id[sel]):
        p = loci.pos0[sel[loci.ref_id[sel] == rid]]
        # one window per run of nearby loci (an interval), so that reads between distant intervals are not dragged in
        cuts = np.flatnonzero(np.diff(p) > 1000)
        for lo_i, hi_i in zip(np.concatenate(([0], cuts + 1)), np.concatenate((cuts, [len(p) - 1]))):
            lo, hi = int(p[lo_i]), int(p[hi_i]) + 1
            mask |= (soa.ref_id == rid) & (soa.pos < hi) & (ends > lo)
    recs = soa_to_records(soa.select(np.flatnonzero(mask)), orc.Read)
    _G["index"] = orc.ReadIndex(recs)
    _G["refs"] = refs
    _G["prm"] = vc_params()
    _G["ref"] = None
    if cpu_kind() == "reference":
        from oracle import ref_build, ref_shims
        _G["ref"] = ref_build.load("native", inline_pool=False)
        ref_shims.register_bam("bench.bam", _G["index"])
        ref_shims.register_fasta("bench.fa", refs)
    jobs = [(soa.chroms[int(loci.ref_id[i])], str(int(loci.pos0[i]) + 1)) for i in sel]
    return jobs, len(recs)


def cpu_run(jobs, cores, pool=None):
    """loci/s of the oracle port over ``jobs`` on ``cores`` worker processes (the reference's Pool(nCPU) fan-out,
    smCounter.py:683-685); the pool is created and warmed outside the timed region."""
    import multiprocessing as mp
    own = None
    if cores > 1 and pool is None:
        own = pool = mp.get_context("fork").Pool(cores)
        pool.map(_cpu_worker, jobs[:cores], chunksize=1)
    t0 = time.perf_counter()
    if cores > 1:
        rows = pool.map(_cpu_worker, jobs, chunksize=1)
    else:
        rows = [_cpu_worker(j) for j in jobs]
    dt = time.perf_counter() - t0
    if own is not None:
        own.close(); own.join()
    events = sum(int(r.split("\t")[5]) for r in rows if r.split("\t")[5])
    return len(jobs) / dt, dt, events


def pipeline_leg(args, mine, soa, refs, device):
    """The whole reference-facing path once, stage by stage, on the first ``--pipeline-intervals`` intervals of the rank-0
    batch: BAM file -> libsmc_bamio decode -> SoA -> smc_call_batch -> 45-column rows -> repeat filters -> the three output
    files (what smCounter.main() does, smCounter.py:645-909).  north_star asks for the end-to-end figure including the
    BAM decode beside the kernel-only one; the BAM itself is written outside the timed stages."""
    import shutil
    import tempfile
    import numpy as np
    from smcounter_b200 import bam, repeats, writers
    from smcounter_b200.shard import reads_for_intervals
    from smcounter_b200.smCounter import call_loci
    ivs = mine[:args.pipeline_intervals]
    sub = soa.select(reads_for_intervals(soa, ivs, soa.chroms))
    tmp = tempfile.mkdtemp(prefix="smc_pipe_")
    try:
        path = os.path.join(tmp, "reads.bam")
        bam.write_bam(path, sub, refs.lengths)
        t = [time.perf_counter()]
        reads = bam.read_bam(path, ivs, threads=os.cpu_count() or 1, trim=True)
        t.append(time.perf_counter())
        rows = call_loci(reads, ivs, refs, vc_params(), gpus=1, devices=[device], stage_times=(st := {}))
        t.append(time.perf_counter())
        target_rows = [(c, str(s), str(e)) for (c, s, e) in ivs]
        trf, rm = repeats.build_repeat_regions(target_rows, [], [])
        rows = repeats.apply_repeat_filters(rows, trf, rm)
        writers.write_outputs(rows, os.path.join(tmp, "out"), 3000, 0)
        t.append(time.perf_counter())
        n_loci = sum(e - s for (_, s, e) in ivs)
        return {"value": n_loci / (t[3] - t[0]), "unit": UNIT, "loci": n_loci, "reads": int(reads.n), "intervals": len(ivs),
                "bam_mb": os.path.getsize(path) / 1e6, "ms_bam_decode": 1e3 * (t[1] - t[0]), "ms_call_loci": 1e3 * (t[2] - t[1]),
                "ms_gpu_call": st.get("ms_gpu_call"), "ms_format_rows": st.get("ms_format_rows"), "gpu_call_detail": {k: v for k, v in st.items() if k not in ("ms_gpu_call", "ms_format_rows")},
                "ms_filters_and_writers": 1e3 * (t[3] - t[2]),
                "what": "BAM decode (libsmc_bamio, all host threads) + smc_call_batch + row formatting + repeat filters + the three "
                        "output files, once, on a sub-batch of the rank-0 panel batch"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        # The Python-2 reference cannot run (no python2 / pysam); its algorithm is timed through the oracle port, on
        # rank 0 only, with all host cores.  Each step is a bounded sample of the same workload.
        if rank != 0:
            return 0
        args.intervals = min(args.intervals, 16)     # the sample only needs the reads around its loci
        mine, _, refs, loci, bed_order, soa = make_batch(args, 0, 1)
        n_sample = args.cpu_loci or 48 * cores
        jobs, nrec = cpu_sample_setup(soa, refs, loci, n_sample)
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        for _ in range(min(args.warmup, 1)):
            cpu_run(jobs[:max(1, 2 * cores)], cores, pool)
        vals, times, ev = [], [], 0
        for _ in range(args.steps):
            v, dt, ev = cpu_run(jobs, cores, pool)
            vals.append(v); times.append(dt)
        if pool is not None:
            pool.close(); pool.join()
        value = len(jobs) * len(times) / sum(times)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
                "config": {"workload": "cfg2 synthetic N0030 194-gene panel, 3000 UMIs/locus, rpb 4, 2x150bp",
                           "sample": "%d consecutive loci, %d reads, %d pileup events per step" % (len(jobs), nrec, ev)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                                 "sample": "%d loci per step x %d steps (%s, multiprocessing.Pool(%d))" % (len(jobs), args.steps, CPU_WHAT[cpu_kind()], cores)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from smcounter_b200.caller import GpuCaller, LocusResults

    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = None
    if world > 1 and all_cpus is not None:
        # one process per GPU: run on the cores next to this GPU so that the pinned host buffers (first touch) live on its
        # NUMA node and the uploads of the 8 ranks do not all cross the socket interconnect
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(local_rank)
            words = nv.nvmlDeviceGetCpuAffinity(h, (max(all_cpus) + 64) // 64)
            local = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1} & set(all_cpus)
            if local:
                os.sched_setaffinity(0, local)
                numa = len(local)
        except Exception:
            numa = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mine, soa, refs, loci, bed_order, soa_full = make_batch(args, rank, world)

    # pinned host copies of every buffer that crosses the ABI
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    for f in ("ref_id", "pos", "flag", "mapq", "nm", "l_seq", "seq_off", "qual_off", "cigar_off", "n_cigar", "umi", "frag_id", "seq",
              "qual", "cigar"):
        setattr(soa, f, pin(getattr(soa, f)))
    for f in ("store_lo", "store_len"):
        if getattr(soa, f) is not None:
            setattr(soa, f, pin(getattr(soa, f)))
    for f in ("ref_id", "pos0", "ref_base"):
        setattr(loci, f, pin(getattr(loci, f)))

    caller = GpuCaller(vc_params(), device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only: batch resident in HBM
    caller.upload(soa, loci)
    for _ in range(args.warmup):
        caller.run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, kp_ms, launches, tms = 0.0, 0.0, 0, None
    for _ in range(args.steps):
        caller.run()
        tms = caller.timings()
        dev_ms += tms["ms_total_device"]; kp_ms += tms["ms_k_pileup"]; launches += tms["kernel_launches"]
    barrier()
    wall_s = time.perf_counter() - t0
    out = caller.download(LocusResults(loci.n, max(int(tms["n_dyn"]), 16), pinned=True))
    n_dyn = out.n_dyn

    # ---------------- end to end: host buffers in, host buffers out, every step
    for _ in range(min(args.warmup, 2)):
        caller.call(soa, loci, out=out)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        caller.call(soa, loci, out=out)
        e2e_tm = caller.timings()
    barrier()
    e2e_s = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_s_max = allmax(dev_ms / 1000.0)
    wall_s_max = allmax(wall_s)
    e2e_s_max = allmax(e2e_s)
    loci_total = allsum(float(loci.n))
    events_total = allsum(float(tms["n_pileup_events"]))
    reads_total = allsum(float(soa.n))
    # device time is what the CUDA events on the library's launch stream bracket for each step (it contains the
    # few host round-trips a step needs); wall clock is reported beside it.
    value = loci_total * args.steps / dev_s_max
    e2e_value = loci_total * args.steps / e2e_s_max

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        algo_bytes = ALGO_BYTES_PER_EVENT * float(tms["n_pileup_events"])
        kp_avg_s = (kp_ms / args.steps) / 1000.0
        traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one k_pileup launch (ncu --set full, same batch)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k_pileup_traffic.json")))
            if int(tr["pileup_events"]) == int(tms["n_pileup_events"]):
                traffic = int(tr["dram_bytes_read"]) + int(tr["dram_bytes_write"])
        except Exception:
            pass
        achieved = algo_bytes / kp_avg_s / 1e9 if kp_avg_s > 0 else 0.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * dev_s_max / args.steps, "wall_ms_per_step": 1000.0 * wall_s_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
                "config": {"workload": "cfg2 synthetic N0030 194-gene panel (primers...coding.bed geometry), 3000 UMIs/locus, rpb 4, "
                                       "2x150bp; one batch = %d seeded panel intervals per GPU" % args.intervals,
                           "intervals_per_gpu": args.intervals, "loci_total": int(loci_total), "reads_total": int(reads_total),
                           "pileup_events_per_step": int(events_total), "tile_events_per_step_rank0": int(tms["n_tile_events"]),
                           "params": "mtDepth 3000 rpb 4.0 mtDrop 0 minBQ 20 minMQ 30",
                           "l2": "inputs larger than L2 (%.0f MB resident per GPU), no flush needed" % (soa.nbytes() / 1e6),
                           "reads": "whole reads" if soa.store_lo is None else "bases / qualities trimmed to each read's target window (store_lo / store_len), "
                                    "%.0f of %d bases per read stored" % (float(soa.store_len.mean()), int(soa.l_seq.max())),
                           "parallelism": "panel sharded by BED interval (balanced by estimated events), no collective",
                           "host_affinity": ("GPU-local cores (%d) while pinning and uploading" % numa) if numa else "unbound"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_tm["bytes_h2d"]),
                        "d2h_bytes_per_step": int(e2e_tm["bytes_d2h"]), "ms_per_step": 1000.0 * e2e_s_max / args.steps,
                        "ms_h2d": e2e_tm["ms_h2d"], "ms_device": e2e_tm["ms_total_device"], "ms_d2h": e2e_tm["ms_d2h"],
                        "h2d_gbs": e2e_tm["bytes_h2d"] / e2e_tm["ms_h2d"] / 1e6 if e2e_tm["ms_h2d"] > 0 else None,
                        "upload": "pipelined: scalars first, bases/qualities in %d chunks on a copy stream, %d pileup launch pairs as they "
                                  "arrive (ms_device overlaps ms_h2d)" % (e2e_tm["pipe_chunks"], e2e_tm["pipe_launches"])},
                "gpu_launches": int(launches),
                "stage_ms_rank0": {k: tms[k] for k in ("ms_prep", "ms_sort", "ms_pileup", "ms_stats", "ms_k_pileup", "ms_k_gather", "ms_k_merge")},
                "roofline": {"bound": "hbm", "kernel": "k_gather + k_merge (K3: tile pileup = event expansion + fragment merge, then calProb / PI / consensus)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                             "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_event": ALGO_BYTES_PER_EVENT,
                             "ms_per_launch": 1000.0 * kp_avg_s, "traffic": traffic,
                             "whole_step": {"what": "the same 35 B x pileup events over the WHOLE device step (read sort, prep, tile events, "
                                                    "both pileup kernels, ALT / filters / Fisher) of rank 0",
                                            "achieved": algo_bytes / ((dev_ms / args.steps) / 1000.0) / 1e9 if dev_ms > 0 else 0.0,
                                            "frac": algo_bytes / ((dev_ms / args.steps) / 1000.0) / 1e9 / peak if dev_ms > 0 and peak else None},
                             "events_per_s": float(tms["n_pileup_events"]) / kp_avg_s if kp_avg_s > 0 else 0.0},
                "clocks": sampler.summary(), "n_dyn_alleles_rank0": int(n_dyn), "n_fisher_rank0": int(tms["n_fisher"])}
    caller.close()

    if all_cpus is not None and numa:
        os.sched_setaffinity(0, all_cpus)          # the host-side legs below may use every core again
    if rank == 0 and args.pipeline_intervals > 0:
        try:
            line["pipeline"] = pipeline_leg(args, mine, soa_full, refs, local_rank)
        except Exception as e:          # the extra leg must never cost the bench line
            line["pipeline"] = {"error": repr(e)}

    if rank == 0 and not args.no_cpu_baseline:
        n_sample = args.cpu_loci or 160 * cores          # ~10-15 s of host work
        jobs, nrec = cpu_sample_setup(soa_full, refs, loci, n_sample)
        v, dt, ev = cpu_run(jobs, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                                "sample": "%d consecutive loci of the rank-0 batch (%d reads, %d pileup events), %s "
                                          "with multiprocessing.Pool(%d), %.1f s" % (len(jobs), nrec, ev, CPU_WHAT[cpu_kind()], cores, dt),
                                "archived_reference": "4.15 loci/s on 10 processes at DP~58k (example run log, 2017 hardware)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
