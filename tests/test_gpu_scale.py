"""-m gpu: the CUDA path at the sizes and shapes BASELINE.json names (configs 2, 3 and 5), where the oracle can no
longer be run over every locus.  Each test checks

  * size-independent properties of the outputs (coverage recomputed independently from the read spans, the counter
    identities of smCounter.py:368-460 / :482-532, idempotence of a resident re-run, independence of the result from how
    the BED intervals are sharded over contexts), and
  * a seeded random sample of loci (always including the spiked variant sites) against the CPU oracle, row for row.
"""
import numpy as np
import pytest

from smcounter_b200 import _ffi
from smcounter_b200.caller import GpuCaller, VcParams
from smcounter_b200.rows import device_hp_flags, format_rows
from smcounter_b200.synth import SynthSpec, make_panel_mp, panel_intervals_from_bed
from smcounter_b200.targets import build_loci, loc_list

pytestmark = pytest.mark.gpu


def _coverage_from_spans(soa, loci):
    """cvg per locus (smCounter.py:368) recomputed on the host: reads with pos <= p < reference_end (htslib column
    membership; deletions count, there are no N ops in the synthetic reads)."""
    ends = soa.ref_end()
    out = np.zeros(loci.n, dtype=np.int64)
    key = (loci.ref_id.astype(np.int64) << 32) | loci.pos0.astype(np.int64)
    lo = np.searchsorted(key, (soa.ref_id.astype(np.int64) << 32) | soa.pos.astype(np.int64), side="left")
    hi = np.searchsorted(key, (soa.ref_id.astype(np.int64) << 32) | ends, side="left")
    diff = np.zeros(loci.n + 1, dtype=np.int64)
    np.add.at(diff, lo, 1)
    np.add.at(diff, hi, -1)
    out[:] = np.cumsum(diff)[:-1]
    return out


def _check_properties(res, soa, loci):
    L, Cn = res.loc, res.cnt
    cvg = L[_ffi.L_CVG].astype(np.int64)
    assert np.array_equal(cvg, _coverage_from_spans(soa, loci)), "cvg differs from the read spans"
    n = loci.n
    dyn_allele = np.zeros(n, dtype=np.int64)
    if res.n_dyn:
        np.add.at(dyn_allele, res.dyn_locus[:res.n_dyn], res.dyn_cnt[:res.n_dyn, _ffi.C_ALLELE])
    # every pileup read is counted under exactly one allele (:379, :401, :459)
    assert np.array_equal(Cn[:, _ffi.C_ALLELE, :].sum(axis=0) + dyn_allele, cvg)
    for a in range(_ffi.SMC_NFIXED):
        if a == _ffi.A_DEL:       # in-deletion reads have no strand tally (:454-457 sits in the regular-base branch)
            assert not Cn[a, _ffi.C_FWD].any() and not Cn[a, _ffi.C_REV].any()
            continue
        assert np.array_equal(Cn[a, _ffi.C_FWD] + Cn[a, _ffi.C_REV], Cn[a, _ffi.C_ALLELE])
        assert (Cn[a, _ffi.C_LOWQ] <= Cn[a, _ffi.C_ALLELE]).all()
        assert (Cn[a, _ffi.C_R1LE] <= Cn[a, _ffi.C_R1TOT]).all() and (Cn[a, _ffi.C_R2LE] <= Cn[a, _ffi.C_R2TOT]).all()
        assert (Cn[a, _ffi.C_R2PLE] <= Cn[a, _ffi.C_R2TOT]).all()
        assert (Cn[a, _ffi.C_STRONG] <= Cn[a, _ffi.C_MT]).all()
    assert (L[_ffi.L_USEDMT] <= L[_ffi.L_NBC]).all() and (L[_ffi.L_NBC] <= L[_ffi.L_ALLMT]).all()
    assert (L[_ffi.L_USEDFRAG] <= L[_ffi.L_ALLFRAG]).all() and (L[_ffi.L_ALLFRAG] <= cvg).all()
    assert (L[_ffi.L_ALLMT] <= L[_ffi.L_ALLFRAG]).all()
    assert (L[_ffi.L_MT10] <= L[_ffi.L_MT7]).all() and (L[_ffi.L_MT7] <= L[_ffi.L_MT5]).all()
    assert (L[_ffi.L_MT5] <= L[_ffi.L_MT3]).all() and (L[_ffi.L_MT3] <= L[_ffi.L_USEDMT]).all()
    mt = Cn[:, _ffi.C_MT, :].sum(axis=0).astype(np.int64)
    if res.n_dyn:
        np.add.at(mt, res.dyn_locus[:res.n_dyn], res.dyn_cnt[:res.n_dyn, _ffi.C_MT])
    assert (mt <= L[_ffi.L_USEDMT]).all()                       # at most one consensus allele per used barcode (:514-523)
    assert not (L[_ffi.L_STATUS] & ~1).any(), "unexpected status bits (down-sampling / overflow)"
    assert np.isfinite(res.pi).all() and (res.pi >= 0).all()


_ORC = {}


def _oracle_one(cp):
    from oracle import smcounter_oracle as orc
    prm = _ORC["prm"]
    c, p = cp
    return orc.vc(_ORC["index"], c, str(p + 1), prm.minBQ, prm.minMQ, prm.mtDepth, prm.rpb, prm.hpLen, prm.mismatchThr, prm.mtDrop,
                  prm.maxMT, prm.primerDist, _ORC["refs"])


def _oracle_rows_for(soa, refs, prm, positions):
    """Oracle rows for [(chrom, pos0)] using only the reads that overlap them; one forked worker per host core (the oracle
    is the reference's per-locus Python: ~0.1 - 1 s per locus at these depths)."""
    import multiprocessing
    import os
    from oracle import smcounter_oracle as orc
    from smcounter_b200.soa import soa_to_records
    cidx = {c: i for i, c in enumerate(soa.chroms)}
    ends = soa.ref_end()
    mask = np.zeros(soa.n, dtype=bool)
    by_chrom = {}
    for (c, p) in positions:
        by_chrom.setdefault(c, []).append(p)
    for c, ps in by_chrom.items():
        ps = np.sort(np.asarray(ps))
        on = soa.ref_id == cidx[c]
        # a read is needed iff some requested position lies in [pos, end)
        nxt = np.searchsorted(ps, soa.pos[on], side="left")
        hit = np.zeros(int(on.sum()), dtype=bool)
        ok = nxt < len(ps)
        hit[ok] = ps[nxt[ok]] < ends[on][ok]
        mask[np.flatnonzero(on)[hit]] = True
    _ORC.update(index=orc.ReadIndex(soa_to_records(soa.select(np.flatnonzero(mask)), orc.Read)), refs=refs, prm=prm)
    workers = min(os.cpu_count() or 1, 32, max(1, len(positions)))
    if workers == 1:
        return [_oracle_one(cp) for cp in positions]
    with multiprocessing.get_context("fork").Pool(workers) as pool:
        return pool.map(_oracle_one, positions, chunksize=1)


def _sample_positions(intervals, truth, rng, n_random, n_truth):
    allpos = [(c, p) for (c, s, e) in intervals for p in range(s, e)]
    inside = set(allpos)
    picks = [allpos[i] for i in rng.choice(len(allpos), size=min(n_random, len(allpos)), replace=False)]
    spiked = [(c, p) for kind in ("snv", "ins", "del") for (c, p, *_) in truth[kind] if (c, p) in inside]
    picks += spiked[:n_truth]
    return list(dict.fromkeys(picks))


def _run(soa, intervals, refs, prm):
    loci, bed_order = build_loci(intervals, soa.chroms, refs)
    caller = GpuCaller(prm, 0)
    res = caller.call(soa, loci)
    tm = caller.timings()
    return caller, res, loci, bed_order, tm


def _rows_by_position(rows, intervals):
    return {cp: r for cp, r in zip([(c, int(p) - 1) for (c, p) in loc_list(intervals)], rows)}


def _compare_with_oracle(g_rows, intervals, soa, refs, prm, truth, seed, n_random, n_truth):
    from helpers import diff_rows
    picks = _sample_positions(intervals, truth, np.random.default_rng(seed), n_random, n_truth)
    want = _oracle_rows_for(soa, refs, prm, picks)
    by_pos = _rows_by_position(g_rows, intervals)
    problems = diff_rows([by_pos[cp] for cp in picks], want)
    assert not problems, "\n".join(problems)
    return len(picks)


def test_cfg2_panel_batch_properties_sharding_and_oracle_sample():
    """BASELINE config 2 at bench depth: N0030 panel intervals, ~3 000 barcodes per locus, rpb 4, 2 x 150 bp (a 24-interval
    batch, ~0.7 M reads, large enough for the library to pipeline the upload on its own)."""
    import os
    from smcounter_b200.smCounter import call_loci
    bed = os.path.join(os.path.dirname(__file__), "golden", "n0030_panel.bed")
    ivs = panel_intervals_from_bed(bed, limit=24, seed=7)
    spec = SynthSpec(umis_per_locus=3000, rpb=4.0, snv_every=400, snv_vaf=0.01, indel_every=1500, indel_vaf=0.01)
    prm = VcParams(mtDepth=3000, rpb=4.0)
    soa, refs, truth = make_panel_mp(ivs, spec, seed=11)
    caller, res, loci, bed_order, tm = _run(soa, ivs, refs, prm)
    try:
        assert tm["pipe_chunks"] >= 2 and tm["pipe_launches"] >= 2, tm
        assert tm["n_pileup_events"] > 3e7
        _check_properties(res, soa, loci)
        hp = device_hp_flags(caller, res, soa, loci, soa.chroms, refs, prm.hpLen)
        g_rows = format_rows(res, soa, loci, soa.chroms, refs, prm.hpLen, bed_order, hp_flags=hp)
        # the same reads trimmed to their target windows (what the BAM decoder delivers): identical bits from fewer bytes
        trimmed = soa.trim_to_targets(ivs)
        assert trimmed.seq.nbytes + trimmed.qual.nbytes < 0.7 * (soa.seq.nbytes + soa.qual.nbytes)
        res_t = caller.call(trimmed, loci)
        for f in ("loc", "cnt", "pi", "alt_allele", "alt_pi", "fl1", "fl2"):
            assert np.array_equal(getattr(res, f), getattr(res_t, f)), f
        res = caller.call(soa, loci)
        # idempotence: the resident batch run again, one launch instead of per-chunk launches -> identical bits
        caller.run()
        again = caller.download(None)
        for f in ("loc", "cnt", "pi", "alt_allele", "alt_pi", "fl1", "fl2"):
            assert np.array_equal(getattr(res, f), getattr(again, f)), f
    finally:
        caller.close()
    n = _compare_with_oracle(g_rows, ivs, soa, refs, prm, truth, seed=3, n_random=200, n_truth=24)
    assert n >= 200
    # the same panel as three shards (three contexts, depth-balanced interval groups): rows identical to the single batch
    assert call_loci(soa, ivs, refs, prm, gpus=3, devices=[0, 0, 0]) == g_rows


@pytest.mark.parametrize("umis", [5500, 10400])
def test_cfg1_example_shaped_every_locus_against_the_oracle(umis):
    """BASELINE config 1 in its synthetic form (example.bam is not distributed): the parameters of
    /root/reference/example/run.example.sh:4-24 (mtDepth 3612, rpb 8.6, mtDrop 1, hpLen 8, minBQ 20, minMQ 30, mismatchThr 6.0,
    primerDist 2) on an example-shaped amplicon -- ~4 100 barcodes per locus at 9.8 fragments per barcode (the example: 4 162),
    and a deep variant at ~7 800 barcodes where ds = 2 * 3612 = 7 224 fires and the product draws the Python-2
    random.sample() mask itself.  EVERY locus is compared with the oracle, row for row."""
    from helpers import diff_rows
    from smcounter_b200.smCounter import call_loci
    ivs = [("chr17", 41243700, 41243800)]
    spec = SynthSpec(umis_per_locus=umis, rpb=9.8, snv_every=25, snv_vaf=0.01, indel_every=60, indel_vaf=0.02)
    prm = VcParams(mtDepth=3612, rpb=8.6, minBQ=20, minMQ=30, hpLen=8, mismatchThr=6.0, mtDrop=1, maxMT=0, primerDist=2)
    soa, refs, truth = make_panel_mp(ivs, spec, seed=20170410)
    g_rows = call_loci(soa, ivs, refs, prm, gpus=1)
    picks = [(c, p) for (c, s, e) in ivs for p in range(s, e)]
    want = _oracle_rows_for(soa, refs, prm, picks)
    umt = [int(r.split("\t")[9]) for r in want]
    mt = [int(r.split("\t")[7]) for r in want]
    print("cfg1-shaped: %d loci, MT %d..%d, UMT %d..%d, DP %s..%s" % (len(want), min(mt), max(mt), min(umt), max(umt), want[0].split("\t")[5], want[-1].split("\t")[5]))
    if umis > 7224:
        assert max(umt) == 7224 and sum(1 for u in umt if u == 7224) > 50        # down-sampling fired (smCounter.py:486-500)
    else:
        assert max(umt) < 7224 and min(mt) > 2000
    problems = diff_rows(g_rows, want)
    assert not problems, "\n".join(problems)


def test_cfg3_deep_low_vaf_shape():
    """BASELINE config 3 shape: 20 000 barcodes per locus (E ~ 1e5 reads per locus, hundreds of units per tile), 0.5 % VAF
    spike-ins, mtDepth 20000 (ds = 40 000: no down-sampling)."""
    ivs = [("chr7", 55000, 55200)]
    spec = SynthSpec(umis_per_locus=20000, rpb=4.0, snv_every=12, snv_vaf=0.005)
    prm = VcParams(mtDepth=20000, rpb=4.0)
    soa, refs, truth = make_panel_mp(ivs, spec, seed=2)
    caller, res, loci, bed_order, tm = _run(soa, ivs, refs, prm)
    try:
        _check_properties(res, soa, loci)
        assert int(res.loc[_ffi.L_USEDMT].max()) > 10000
        hp = device_hp_flags(caller, res, soa, loci, soa.chroms, refs, prm.hpLen)
        g_rows = format_rows(res, soa, loci, soa.chroms, refs, prm.hpLen, bed_order, hp_flags=hp)
    finally:
        caller.close()
    n = _compare_with_oracle(g_rows, ivs, soa, refs, prm, truth, seed=5, n_random=200, n_truth=16)      # every locus of the batch
    assert n >= 200


def test_cfg5_many_loci_many_contigs_shape():
    """BASELINE config 5 shape, scaled: 1 200 intervals x 200 bp over 24 contigs (240 000 loci, > 7 000 tiles), shallow and
    with log-normal depth per interval; also run as 8 depth-balanced shards (the 8-GPU plan, here on one device)."""
    from smcounter_b200.smCounter import call_loci
    rng = np.random.default_rng(4)
    ivs = []
    for k in range(1200):
        c = "chr%d" % (1 + k % 24)
        s = 1000 + 700 * (k // 24) + int(rng.integers(0, 100))
        ivs.append((c, s, s + 200))
    spec = SynthSpec(umis_per_locus=10, rpb=2.0, snv_every=500, snv_vaf=0.3, indel_every=3000, indel_vaf=0.3, depth_sigma=0.5)
    prm = VcParams(mtDepth=10, rpb=2.0, maxMT=400)
    soa, refs, truth = make_panel_mp(ivs, spec, seed=4)
    caller, res, loci, bed_order, tm = _run(soa, ivs, refs, prm)
    try:
        assert loci.n == 240000
        _check_properties(res, soa, loci)
        hp = device_hp_flags(caller, res, soa, loci, soa.chroms, refs, prm.hpLen)
        g_rows = format_rows(res, soa, loci, soa.chroms, refs, prm.hpLen, bed_order, hp_flags=hp)          # forked formatting workers
        assert format_rows(res, soa, loci, soa.chroms, refs, prm.hpLen, bed_order[:20000], workers=1, hp_flags=hp) == g_rows[:20000]
    finally:
        caller.close()
    n = _compare_with_oracle(g_rows, ivs, soa, refs, prm, truth, seed=9, n_random=200, n_truth=40)
    assert n >= 200
    assert call_loci(soa, ivs, refs, prm, gpus=8, devices=[0] * 8) == g_rows
    # one GPU, the target streamed through it in batches of at most 20 000 loci (the per-batch limits, scaled down)
    st = {}
    assert call_loci(soa, ivs, refs, prm, gpus=1, stage_times=st, batch_limits={"max_loci": 20000}) == g_rows
    assert st["batches"] >= 12


def test_call_loci_across_all_visible_gpus():
    """The CLI's multi-GPU path: one host thread and one context per visible GPU (up to 4), BED intervals sharded by
    estimated depth, rows interleaved back into BED order -- identical to the single-GPU rows.  (On a one-GPU box this runs
    two shards on the same device.)"""
    import torch
    from smcounter_b200.smCounter import call_loci
    ndev = min(4, torch.cuda.device_count())
    ivs = [("chr%d" % (1 + k % 3), 1000 + 400 * k, 1000 + 400 * k + 40 + 13 * (k % 5)) for k in range(14)]
    spec = SynthSpec(umis_per_locus=80, rpb=3.0, snv_every=60, snv_vaf=0.1, indel_every=90, indel_vaf=0.1)
    prm = VcParams(mtDepth=80, rpb=3.0)
    soa, refs, _ = make_panel_mp(ivs, spec, seed=6, workers=2)
    one = call_loci(soa, ivs, refs, prm, gpus=1)
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    many = call_loci(soa.trim_to_targets(ivs), ivs, refs, prm, gpus=len(devices), devices=devices)
    assert many == one and len(one) == sum(e - s for (_, s, e) in ivs)
    # intervals larger than the per-batch limits (here: 25 loci; in production a whole-chromosome BED line or a very deep
    # amplicon) are cut into consecutive sub-intervals and streamed; rows unchanged
    st = {}
    assert call_loci(soa, ivs, refs, prm, gpus=1, batch_limits={"max_loci": 25}, stage_times=st) == one
    assert st["batches"] > len(ivs)


def test_cfg4_indel_and_repeat_heavy_panel_through_the_cli(tmp_path):
    """BASELINE config 4 shape, scaled: panel intervals on three contigs, 5 % of the molecules carry a 1-4 bp indel every 40 bp,
    synthetic simpleRepeat (3-column) and RepeatMasker (4-column: Simple_repeat / Low_complexity / Satellite / other) tracks
    over ~10 % of the target, all filters active.  smCounter.main() from BAM + BED + FASTA + the two repeat BEDs to the three
    output files, byte-identical to the oracle's restatement of main() (smCounter.py:645-909)."""
    import os
    from oracle import smcounter_oracle as orc
    from smcounter_b200 import bam, smCounter
    from smcounter_b200.soa import soa_to_records
    bed_path = os.path.join(os.path.dirname(__file__), "golden", "n0030_panel.bed")
    ivs = panel_intervals_from_bed(bed_path, limit=9, seed=21)
    spec = SynthSpec(umis_per_locus=120, rpb=4.0, snv_every=150, snv_vaf=0.05, indel_every=40, indel_vaf=0.05, softclip_frac=0.1)
    soa, refs, _ = make_panel_mp(ivs, spec, seed=21, workers=4)
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as fh:
        for c in soa.chroms:
            s = refs.fetch(c, 0, refs.get_reference_length(c))
            fh.write(">%s\n" % c)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
    bed = tmp_path / "target.bed"
    bed_lines = ["%s\t%d\t%d\n" % iv for iv in ivs]
    bed.write_text("".join(bed_lines))
    # repeat tracks: every third interval gets a tandem repeat over its first 30 bp and a RepeatMasker element over the next 25
    kinds = ("Simple_repeat", "Low_complexity", "Satellite", "L1")
    trf_rows, rm_rows = [], []
    for k, (c, s, e) in enumerate(sorted(ivs)):
        if k % 3 == 0:
            trf_rows.append((c, str(s), str(min(e, s + 30))))
            rm_rows.append((c, str(s + 20), str(min(e, s + 55)), kinds[(k // 3) % 4]))
    trf = tmp_path / "simpleRepeat.bed"; trf.write_text("".join("\t".join(r) + "\n" for r in trf_rows))
    rm = tmp_path / "SR_LC_SL.bed"; rm.write_text("".join("\t".join(r) + "\n" for r in rm_rows))
    bam_path = tmp_path / "reads.bam"
    bam.write_bam(str(bam_path), soa, refs.lengths)
    prefix = str(tmp_path / "out")
    smCounter.argParseInit()
    thr = smCounter.main({"outPrefix": prefix, "bamFile": str(bam_path), "bedTarget": str(bed), "mtDepth": 120, "rpb": 4.0, "nCPU": 4,
                          "refGenome": str(fa), "bedTandemRepeats": str(trf), "bedRepeatMaskerSubset": str(rm)})
    want_thr, all_txt, cut_txt, cut_vcf = orc.run(soa_to_records(soa, orc.Read), bed_lines, refs, mtDepth=120, rpb=4.0, threshold=0,
                                                   outPrefix=prefix, trf_rows=trf_rows, rm_rows=rm_rows)
    assert thr == want_thr
    got_all = open(prefix + ".smCounter.all.txt").read()
    assert got_all == all_txt
    assert open(prefix + ".smCounter.cut.txt").read() == cut_txt
    assert open(prefix + ".smCounter.cut.vcf").read() == cut_vcf
    assert "RepT" in all_txt and any(t in all_txt for t in ("RepS", "LowC", "SL", "Other_Repeat")) and "INDEL" in all_txt


def test_example_bam_against_the_committed_golden_output(tmp_path):
    """BASELINE config 1 proper: /root/reference/example/run.example.sh on the real example.bam, diffed against the committed
    example.smCounter.all.txt / .cut.txt / .cut.vcf.  The BAM (76 MB) and hg19 are not distributed with the reference and there
    is no network here, so this runs only where they are supplied:

        SMC_EXAMPLE_BAM=/path/example.bam  SMC_HG19=/path/ucsc.hg19.fasta  [SMC_TRF_BED=simpleRepeat.bed  SMC_RM_BED=SR_LC_SL.nochr.bed]

    Expected: every integer column identical; PI columns equal as printed except for last-digit effects of the summation order
    (DESIGN.md section 5); the 8 rows at UMT = 7224 depend on the Python-2 random.sample emulation."""
    import os
    bam_path, fa = os.environ.get("SMC_EXAMPLE_BAM"), os.environ.get("SMC_HG19")
    if not bam_path or not fa:
        pytest.skip("set SMC_EXAMPLE_BAM and SMC_HG19 to run the reference's example (files not distributed with the reference)")
    from smcounter_b200 import smCounter
    gold = os.path.join(os.path.dirname(__file__), "golden")
    prefix = str(tmp_path / "example")
    smCounter.argParseInit()
    thr = smCounter.main({"outPrefix": prefix, "bamFile": bam_path, "bedTarget": os.path.join(gold, "example.bed"), "mtDepth": 3612, "rpb": 8.6,
                          "nCPU": os.cpu_count() or 1, "minBQ": 20, "minMQ": 30, "hpLen": 8, "mismatchThr": 6.0, "mtDrop": 1, "maxMT": 0,
                          "primerDist": 2, "threshold": 0, "refGenome": fa, "bedTandemRepeats": os.environ.get("SMC_TRF_BED", "none"),
                          "bedRepeatMaskerSubset": os.environ.get("SMC_RM_BED", "none")})
    assert thr == 58                                                       # ceil(14 + 0.012 * 3612), example.smCounter.cut.txt
    got = open(prefix + ".smCounter.all.txt").read().split("\n")
    want = open(os.path.join(gold, "example.smCounter.all.txt")).read().split("\n")
    assert got[0] == want[0] and len(got) == len(want)
    int_cols = [5, 6, 7, 8, 9, 11, 13, 15] + list(range(16, 20)) + list(range(24, 32)) + list(range(36, 40))
    bad_int, bad_any = [], []
    for g, w in zip(got[1:], want[1:]):
        if g == w:
            continue
        gf, wf = g.split("\t"), w.split("\t")
        bad_any.append((wf[1], [(i, gf[i], wf[i]) for i in range(min(len(gf), len(wf))) if gf[i] != wf[i]][:6]))
        if wf[9] != "7224" and any(gf[i] != wf[i] for i in int_cols):
            bad_int.append(bad_any[-1])
    print("example.bam: %d rows, %d differ in any column, %d differ in an integer column outside the down-sampled rows" % (len(want) - 2, len(bad_any), len(bad_int)))
    for b in bad_any[:40]:
        print("  ", b)
    if "SMC_TRF_BED" not in os.environ:
        bad_any = [b for b in bad_any if not all(i == 44 for (i, _, _) in b[1])]       # FILTER column carries RepT in the golden run
    assert not bad_int, bad_int[:10]
    assert len(bad_any) <= 40
