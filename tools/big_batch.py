"""Manual large-batch check (not part of the suite: ~2-3 min of generation): a panel batch near the per-batch limits
(>= 3 GiB of bases + qualities untrimmed) through one smc_call_batch, whole reads and trimmed, checked with the size-independent
properties of tests/test_gpu_scale.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_scale as T
from smcounter_b200.caller import GpuCaller, VcParams
from smcounter_b200.synth import SynthSpec, make_panel_mp, panel_intervals_from_bed
from smcounter_b200.targets import build_loci

n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 520
ivs = panel_intervals_from_bed(os.path.join(ROOT, "tests", "golden", "n0030_panel.bed"), limit=n_iv, seed=3)
t = time.time()
soa, refs, truth = make_panel_mp(ivs, SynthSpec(umis_per_locus=3000, rpb=4.0, snv_every=1000, snv_vaf=0.01, indel_every=12000, indel_vaf=0.01), seed=3)
print("generated %d reads, %.2f GiB payload in %.0f s" % (soa.n, (soa.seq.nbytes + soa.qual.nbytes) / 2**30, time.time() - t), flush=True)
loci, bed_order = build_loci(ivs, soa.chroms, refs)
prm = VcParams(mtDepth=3000, rpb=4.0)
c = GpuCaller(prm, 0)
t = time.time(); res = c.call(soa, loci); tm = c.timings()
print("whole reads: %.0f ms call, device %.1f ms, h2d %.1f ms (%.1f GB/s), %d loci, %d events, chunks %d" % (
    1e3 * (time.time() - t), tm["ms_total_device"], tm["ms_h2d"], tm["bytes_h2d"] / tm["ms_h2d"] / 1e6, loci.n, tm["n_pileup_events"], tm["pipe_chunks"]), flush=True)
T._check_properties(res, soa, loci)
print("properties ok", flush=True)
c.run(); again = c.download(None)
for f in ("loc", "cnt", "pi", "alt_allele", "fl1"):
    assert np.array_equal(getattr(res, f), getattr(again, f)), f
t = time.time(); tr = soa.trim_to_targets(ivs); print("trim %.0f s -> %.2f GiB" % (time.time() - t, (tr.seq.nbytes + tr.qual.nbytes) / 2**30), flush=True)
t = time.time(); res_t = c.call(tr, loci); tm = c.timings()
print("trimmed: %.0f ms call, device %.1f ms, h2d %.1f ms" % (1e3 * (time.time() - t), tm["ms_total_device"], tm["ms_h2d"]), flush=True)
for f in ("loc", "cnt", "pi", "alt_allele", "alt_pi", "fl1", "fl2"):
    assert np.array_equal(getattr(res, f), getattr(res_t, f)), f
print("trimmed == whole, resident rerun == pipelined call; loci/s resident: %.0f" % (loci.n / (tm["ms_prep"] + tm["ms_sort"] + tm["ms_pileup"] + tm["ms_stats"]) * 1e3))
c.close()
