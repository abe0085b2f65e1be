"""The differential-fuzz corpus shared by the GPU parity tests (CUDA path vs oracle) and the reference-code parity tests
(oracle vs the reference's own vc() under oracle/ref_shims): panel shape, read-error knobs and every vc() parameter are
drawn from a generator seeded by the case number."""
from __future__ import annotations

import numpy as np

from smcounter_b200.caller import VcParams
from smcounter_b200.synth import SynthSpec


def fuzz_case(seed):
    """-> (intervals, SynthSpec, VcParams) of fuzz case ``seed``."""
    rng = np.random.default_rng(seed)
    ivs = []
    for _ in range(int(rng.integers(1, 5))):
        c = "chr%d" % int(rng.integers(1, 4))
        s = int(rng.integers(200, 3000))
        L = int(rng.choice([1, 2, 7, 31, 33, 64, 90, 150]))
        if any(c == cc and s < ee + 400 and ss < s + L + 400 for (cc, ss, ee) in ivs):
            continue
        ivs.append((c, s, s + L))
    rpb = float(rng.choice([1.1, 2.0, 3.5, 6.0]))
    umis = int(rng.choice([8, 25, 60, 150]))
    spec = SynthSpec(umis_per_locus=umis, rpb=rpb, snv_every=int(rng.choice([0, 20, 60])), snv_vaf=float(rng.choice([0.02, 0.2, 0.6])),
                     indel_every=int(rng.choice([0, 45, 120])), indel_vaf=float(rng.choice([0.05, 0.3])),
                     softclip_frac=float(rng.choice([0.0, 0.1, 0.5])), lowmapq_frac=float(rng.choice([0.0, 0.1, 0.4])),
                     n_frac=float(rng.choice([0.0, 0.002])), pcr_err_per_frag=float(rng.choice([0.0, 0.01, 0.05])),
                     q_values=(37, 30, 12) if rng.random() < 0.7 else (40, 22, 19), q_probs=(0.85, 0.10, 0.05) if rng.random() < 0.7 else (0.4, 0.3, 0.3))
    prm = VcParams(mtDepth=int(rng.choice([umis, max(2, umis // 3)])), rpb=rpb, minBQ=int(rng.choice([0, 13, 20, 25, 31])),
                   minMQ=int(rng.choice([0, 20, 30, 60])), hpLen=int(rng.choice([4, 8, 10])), mismatchThr=float(rng.choice([0.5, 2.0, 6.0, 100.0])),
                   mtDrop=int(rng.choice([0, 0, 1, 2])), maxMT=int(rng.choice([0, 0, 12])), primerDist=int(rng.choice([0, 2, 10])))
    return ivs, spec, prm


# hand-written cases (also the GPU parity cases of tests/test_gpu_parity.py)
CASES = {
    "snv_basic": (dict(umis_per_locus=60, rpb=3.0, snv_every=50, snv_vaf=0.1), dict(mtDepth=60, rpb=3.0), [("chr1", 1000, 1200), ("chr2", 500, 560)], 7),
    "indel_heavy": (dict(umis_per_locus=50, rpb=3.0, snv_every=40, snv_vaf=0.08, indel_every=30, indel_vaf=0.08), dict(mtDepth=50, rpb=3.0), [("chr1", 2000, 2180)], 11),
    "mtdrop1_rpb8": (dict(umis_per_locus=40, rpb=8.6, snv_every=60, snv_vaf=0.05, indel_every=70, indel_vaf=0.03), dict(mtDepth=40, rpb=8.6, mtDrop=1, hpLen=8), [("chr17", 41243700, 41243860)], 20170410),
    "low_rpb": (dict(umis_per_locus=120, rpb=1.2, snv_every=30, snv_vaf=0.2, n_frac=0.01), dict(mtDepth=120, rpb=1.2), [("chrX", 100, 260)], 3),
    "strict_bq_mq": (dict(umis_per_locus=60, rpb=4.0, snv_every=45, snv_vaf=0.5, lowmapq_frac=0.2, softclip_frac=0.3), dict(mtDepth=60, rpb=2.0, minBQ=31, minMQ=50, mismatchThr=3.0, primerDist=5), [("chr3", 700, 900)], 5),
    "ragged_intervals": (dict(umis_per_locus=30, rpb=3.0, snv_every=25, snv_vaf=0.9, indel_every=45, indel_vaf=0.4), dict(mtDepth=30, rpb=3.0), [("chr1", 100, 101), ("chr1", 140, 173), ("chr1", 173, 175), ("chr2", 5, 70), ("chr1", 150, 160)], 13),
    # down-sampling fires on most loci (ds = 2 * mtDepth = 36 < ~60 barcodes): exercises random.seed(pos)/random.sample (:497-498)
    "downsample": (dict(umis_per_locus=60, rpb=3.0, snv_every=35, snv_vaf=0.15), dict(mtDepth=18, rpb=3.0), [("chr1", 1000, 1080)], 17),
}


# the depths of BASELINE.json's configurations, a few loci each: part of the committed reference-rows fixture
# (tests/golden/ref_rows.json), i.e. the GPU path is held against rows the reference's own code printed at these depths
DEEP_CASES = {
    "deep_cfg2": (dict(umis_per_locus=3000, rpb=4.0, snv_every=5, snv_vaf=0.01), dict(mtDepth=3000, rpb=4.0), [("chr1", 5000, 5010)], 31),
    "deep_cfg1": (dict(umis_per_locus=4000, rpb=9.8, snv_every=4, snv_vaf=0.02, indel_every=7, indel_vaf=0.02), dict(mtDepth=3612, rpb=8.6, mtDrop=1, hpLen=8),
                  [("chr17", 41243700, 41243706)], 32),
    "deep_cfg3": (dict(umis_per_locus=20000, rpb=4.0, snv_every=2, snv_vaf=0.005), dict(mtDepth=20000, rpb=4.0), [("chr2", 7000, 7003)], 33),
}


def case_inputs(name):
    spec_kw, prm_kw, intervals, seed = CASES[name] if name in CASES else DEEP_CASES[name]
    return intervals, SynthSpec(**spec_kw), VcParams(**prm_kw), seed
