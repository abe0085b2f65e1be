"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar: every integer tally bit-exact, PI / Fisher p within 1e-9 relative, rows byte-identical."""
import pytest

from smcounter_b200.caller import VcParams
from smcounter_b200.synth import SynthSpec

pytestmark = pytest.mark.gpu

from fuzz import CASES, fuzz_case


@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_case(name):
    from helpers import run_case
    spec_kw, prm_kw, intervals, seed = CASES[name]
    problems, stats, _ = run_case(intervals, SynthSpec(**spec_kw), VcParams(**prm_kw), seed)
    print(name, stats)
    assert not problems, "\n".join(problems)
    assert stats["events"] > 0


def _ref_rows_fixture():
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_rows.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("name", sorted(_ref_rows_fixture()))
def test_product_rows_equal_the_reference_codes_rows(name):
    """tests/golden/ref_rows.json holds the rows the REFERENCE'S OWN vc() printed (smCounter.py run through oracle/ref_build.py with
    Python-2 container order; made by tests/golden/make_ref_rows.py).  The product path -- call_loci(): C-ABI, CUDA kernels,
    its own Py2 down-sampling emulation, host row formatting -- must print the same bytes."""
    from fuzz import case_inputs
    from smcounter_b200.smCounter import call_loci
    from smcounter_b200.synth import make_panel
    want = _ref_rows_fixture()[name]
    if name.startswith("fuzz"):
        seed = int(name[4:])
        ivs, spec, prm = fuzz_case(seed)
    else:
        ivs, spec, prm, seed = case_inputs(name)
    soa, refs, _ = make_panel(ivs, spec, seed=seed)
    got = call_loci(soa, ivs, refs, prm, gpus=1)
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, "%d rows differ, first: %r vs %r" % (len(bad), bad[0][0], bad[0][1])


def test_cli_files_equal_the_files_of_the_references_main(tmp_path):
    """The reference's main() (oracle/_ref: BED -> vc per locus -> bedtools repeat filters -> three files) against the product
    CLI on the same BAM / BED / FASTA / repeat tracks: all three output files byte-identical."""
    from oracle import ref_build, ref_shims
    from oracle import smcounter_oracle as orc
    if not ref_build.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    from smcounter_b200 import bam, smCounter
    from smcounter_b200.soa import soa_to_records
    from smcounter_b200.synth import make_panel
    ivs = [("chr1", 1000, 1150), ("chr2", 600, 640), ("chr1", 1100, 1120)]
    spec = SynthSpec(umis_per_locus=80, rpb=3.0, snv_every=40, snv_vaf=0.2, indel_every=60, indel_vaf=0.15)
    soa, refs, _ = make_panel(ivs, spec, seed=31)
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as fh:
        for c in soa.chroms:
            s = refs.fetch(c, 0, refs.get_reference_length(c))
            fh.write(">%s\n" % c)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
    bed = tmp_path / "target.bed"
    bed.write_text("track name=t\n" + "".join("%s\t%d\t%d\n" % iv for iv in ivs))
    trf = tmp_path / "trf.bed"; trf.write_text("chr1\t1010\t1040\nchr2\t0\t700\n")
    rm = tmp_path / "rm.bed"
    rm.write_text("chr1\t1030\t1060\tSimple_repeat\nchr1\t1055\t1100\tLow_complexity\nchr1\t1101\t1105\tSatellite\nchr2\t610\t620\tL1\n")
    bam_path = tmp_path / "reads.bam"
    bam.write_bam(str(bam_path), soa, refs.lengths)
    common = {"bedTarget": str(bed), "mtDepth": 80, "rpb": 3.0, "bedTandemRepeats": str(trf), "bedRepeatMaskerSubset": str(rm), "threshold": 20}
    smCounter.argParseInit()
    thr = smCounter.main(dict(common, outPrefix=str(tmp_path / "gpu"), bamFile=str(bam_path), refGenome=str(fa)))
    ref = ref_build.load("py2")
    ref.parser = None
    rbam = ref_shims.register_bam("mem:cli.bam", soa_to_records(soa, orc.Read))
    rfa = ref_shims.register_fasta("mem:cli.fa", refs)
    thr_ref = ref.main(dict(common, outPrefix=str(tmp_path / "ref"), bamFile=rbam, refGenome=rfa, bedtoolsPath="/nowhere/"))
    ref_shims.clear_registries()
    assert thr == thr_ref == 20
    for ext in (".smCounter.all.txt", ".smCounter.cut.txt", ".smCounter.cut.vcf"):
        a = open(str(tmp_path / "gpu") + ext).read(); b = open(str(tmp_path / "ref") + ext).read()
        if ext.endswith(".vcf"):        # the sample column is named after outPrefix (smCounter.py:817)
            a = a.replace(str(tmp_path / "gpu"), "X"); b = b.replace(str(tmp_path / "ref"), "X")
        assert a == b, ext
    assert open(str(tmp_path / "gpu") + ".smCounter.cut.txt").read().count("\n") > 3


def test_downsampling_mask_from_oracle():
    """mtDepth far below the real depth: ds = 2*mtDepth fires on most loci; the CUDA path applies the read-selection mask
    the oracle produced (north_star) and must agree on everything downstream of it."""
    from helpers import run_case
    spec = SynthSpec(umis_per_locus=60, rpb=3.0, snv_every=35, snv_vaf=0.15)
    problems, stats, _ = run_case([("chr1", 1000, 1080)], spec, VcParams(mtDepth=18, rpb=3.0), seed=17)
    assert stats["n_downsampled"] > 20
    assert not problems, "\n".join(problems)


def test_downsampling_drawn_by_the_product_matches_oracle_py2_sampler():
    """No external mask: the product lists the barcodes on the device (smc_list_barcodes), emulates CPython-2
    random.seed(pos)/random.sample over the Py2 dict order on the host, and re-runs with its own mask.  The oracle does
    the same with its independent restatement (sampler='py2'); rows must be byte-identical."""
    from helpers import oracle_run
    from smcounter_b200.smCounter import call_loci
    from smcounter_b200.synth import make_panel
    ivs = [("chr1", 2000, 2050), ("chr2", 400, 420)]
    prm = VcParams(mtDepth=15, rpb=3.0)
    soa, refs, _ = make_panel(ivs, SynthSpec(umis_per_locus=50, rpb=3.0, snv_every=30, snv_vaf=0.2, indel_every=45, indel_vaf=0.1), seed=23)
    o_rows, details = oracle_run(soa, ivs, refs, prm)
    assert sum(1 for d in details if d.get("nBC", 0) > d.get("ds", 1 << 30)) > 10
    g_rows = call_loci(soa, ivs, refs, prm, gpus=1)
    assert g_rows == o_rows
    # the same through the trimmed + packed encoding (listing kernels, mask re-run) and as two shards
    assert call_loci(soa.trim_to_targets(ivs), ivs, refs, prm, gpus=2, devices=[0, 0]) == o_rows


def test_cli_end_to_end_files(tmp_path):
    """smCounter.main(): BAM + BED + FASTA + repeat tracks on disk -> .all.txt / .cut.txt / .cut.vcf, byte-identical to
    the oracle's restatement of main() (smCounter.py:645-909) on the same inputs."""
    import argparse
    from oracle import smcounter_oracle as orc
    from smcounter_b200 import bam, smCounter
    from smcounter_b200.soa import soa_to_records
    from smcounter_b200.synth import make_panel
    ivs = [("chr1", 1000, 1150), ("chr2", 600, 640), ("chr1", 1100, 1120)]
    spec = SynthSpec(umis_per_locus=80, rpb=3.0, snv_every=40, snv_vaf=0.2, indel_every=60, indel_vaf=0.15)
    soa, refs, _ = make_panel(ivs, spec, seed=31)
    # files
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as fh:
        for c in soa.chroms:
            s = refs.fetch(c, 0, refs.get_reference_length(c))
            fh.write(">%s\n" % c)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
    bed = tmp_path / "target.bed"
    bed_lines = ["track name=t\n"] + ["%s\t%d\t%d\n" % iv for iv in ivs]
    bed.write_text("".join(bed_lines))
    trf_rows = [("chr1", "1010", "1040"), ("chr2", "0", "700")]
    rm_rows = [("chr1", "1030", "1060", "Simple_repeat"), ("chr1", "1055", "1100", "Low_complexity"), ("chr1", "1101", "1105", "Satellite"),
               ("chr2", "610", "620", "L1")]
    trf = tmp_path / "trf.bed"; trf.write_text("".join("\t".join(r) + "\n" for r in trf_rows))
    rm = tmp_path / "rm.bed"; rm.write_text("".join("\t".join(r) + "\n" for r in rm_rows))
    bam_path = tmp_path / "reads.bam"
    bam.write_bam(str(bam_path), soa, refs.lengths)
    prefix = str(tmp_path / "out")
    smCounter.argParseInit()
    thr = smCounter.main({"outPrefix": prefix, "bamFile": str(bam_path), "bedTarget": str(bed), "mtDepth": 80, "rpb": 3.0,
                          "refGenome": str(fa), "bedTandemRepeats": str(trf), "bedRepeatMaskerSubset": str(rm), "threshold": 20})
    assert thr == 20
    # oracle on the same inputs (FASTA contents identical to SparseRef: windows + N elsewhere)
    want_thr, all_txt, cut_txt, cut_vcf = orc.run(soa_to_records(soa, orc.Read), bed_lines, refs, mtDepth=80, rpb=3.0, threshold=20,
                                                   outPrefix=prefix, trf_rows=trf_rows, rm_rows=rm_rows)
    assert open(prefix + ".smCounter.all.txt").read() == all_txt
    assert open(prefix + ".smCounter.cut.txt").read() == cut_txt
    assert open(prefix + ".smCounter.cut.vcf").read() == cut_vcf
    assert cut_txt.count("\n") > 3 and "RepT" in all_txt and ("RepS" in all_txt or "LowC" in all_txt)


def test_fragment_code_storage_overflow_retries_with_worst_case_layout():
    """Single-read fragments (R1 only) with dynamic alleles nearly everywhere: a unit needs more extension rows than the
    compact fragment-code layout holds, k_gather raises GF_CODE_FULL and the batch is re-run with the 3x layout.
    Results must still be bit-exact."""
    from helpers import run_case
    import numpy as np
    spec = SynthSpec(umis_per_locus=90, rpb=2.0, snv_every=40, snv_vaf=0.1, indel_every=9, indel_vaf=0.9, n_frac=0.02)

    def r1_only(soa):
        return soa.select(np.flatnonzero((soa.flag & 0x40) != 0))

    problems, stats, _ = run_case([("chr1", 3000, 3090)], spec, VcParams(mtDepth=90, rpb=2.0), seed=41, mutate=r1_only)
    print(stats)
    assert stats["code_mult"] == 3, "the test input no longer overflows the compact layout"
    assert not problems, "\n".join(problems)


def test_barcodes_deeper_than_the_prior_table():
    """> 192 fragments in one barcode: the PCR prior comes from device pow() instead of the host-built table."""
    from helpers import run_case
    spec = SynthSpec(umis_per_locus=3, rpb=260.0, snv_every=30, snv_vaf=0.3)
    problems, stats, _ = run_case([("chr2", 900, 960)], spec, VcParams(mtDepth=3, rpb=8.0, maxMT=50), seed=43)
    print(stats)
    assert not problems, "\n".join(problems)


def test_many_dynamic_alleles_grow_the_table():
    """More distinct indel / N alleles than the initial dynamic-allele table holds: GF_DYN_FULL -> x4 -> re-run."""
    import os
    from helpers import run_case
    spec = SynthSpec(umis_per_locus=40, rpb=2.0, snv_every=50, snv_vaf=0.1, indel_every=7, indel_vaf=0.3, n_frac=0.05)
    os.environ["SMC_DYN_CAP0"] = "64"
    try:
        problems, stats, _ = run_case([("chr1", 500, 700)], spec, VcParams(mtDepth=40, rpb=2.0), seed=47)
    finally:
        del os.environ["SMC_DYN_CAP0"]
    print(stats)
    assert stats["dyn_capacity"] > 64 and stats["n_dyn"] > 32
    assert not problems, "\n".join(problems)


def test_empty_and_uncovered_batches():
    """No reads, no loci, and reads that miss every locus: zero-coverage rows, no kernel faults."""
    import numpy as np
    from helpers import gpu_run, oracle_run
    from smcounter_b200.synth import make_panel
    ivs = [("chr1", 1000, 1040)]
    prm = VcParams(mtDepth=20, rpb=3.0)
    soa, refs, _ = make_panel(ivs, SynthSpec(umis_per_locus=20, rpb=3.0), seed=3)
    # (a) no reads at all
    empty = soa.select(np.zeros(0, dtype=np.int64))
    rows, res, loci, _, tm = gpu_run(empty, ivs, refs, prm)
    o_rows, _ = oracle_run(empty, ivs, refs, prm)
    assert rows == o_rows and tm["n_pileup_events"] == 0 and all(r.endswith("Zero_Coverage") for r in rows)
    # (b) reads that do not overlap the targets
    far = [("chr1", 1500, 1520)]        # inside the contig, beyond every read
    rows, res, loci, _, tm = gpu_run(soa, far, refs, prm)
    o_rows, _ = oracle_run(soa, far, refs, prm)
    assert rows == o_rows and tm["n_pileup_events"] == 0
    # (c) no loci
    rows, res, loci, _, tm = gpu_run(soa, [], refs, prm)
    assert rows == [] and loci.n == 0


def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


PIPE_IVS = [("chr1", 1000, 1150), ("chr1", 4000, 4100), ("chr2", 300, 420), ("chr3", 50, 130)]
PIPE_SPEC = dict(umis_per_locus=50, rpb=3.0, snv_every=40, snv_vaf=0.1, indel_every=35, indel_vaf=0.1, n_frac=0.01)


def test_pipelined_upload_chunks_match_oracle():
    """smc_call_batch with the bases / qualities uploaded in 5 chunks on the copy stream (forced: the test batch is far
    below the size where the library pipelines on its own): the pileup kernels are launched per chunk for the units
    whose reads have arrived; every tally, PI and row must still match the oracle."""
    from helpers import run_case
    problems, stats, _ = _with_env("SMC_PIPE_CHUNKS", "5", lambda: run_case(PIPE_IVS, SynthSpec(**PIPE_SPEC), VcParams(mtDepth=50, rpb=3.0), seed=53))
    print(stats)
    assert stats["pipe_chunks"] == 5 and stats["pipe_launches"] >= 3
    assert not problems, "\n".join(problems)


def test_pipelined_upload_with_payload_out_of_read_order():
    """Bases / qualities stored in REVERSE read order: a read's bytes are not in its chunk, the layout check fires and
    every unit waits for the last chunk (one launch pair).  Results unchanged."""
    import numpy as np
    from helpers import run_case

    def reverse_payload(soa):
        sb = (soa.l_seq.astype(np.int64) + 1) // 2
        lq = soa.l_seq.astype(np.int64)
        order = np.arange(soa.n)[::-1]
        new_seq_off = np.zeros(soa.n, np.int64); new_qual_off = np.zeros(soa.n, np.int64)
        new_seq_off[order] = np.concatenate(([0], np.cumsum(sb[order])))[:-1]
        new_qual_off[order] = np.concatenate(([0], np.cumsum(lq[order])))[:-1]
        seq = np.zeros_like(soa.seq); qual = np.zeros_like(soa.qual)
        for r in range(soa.n):
            seq[new_seq_off[r]:new_seq_off[r] + sb[r]] = soa.seq[soa.seq_off[r]:soa.seq_off[r] + sb[r]]
            qual[new_qual_off[r]:new_qual_off[r] + lq[r]] = soa.qual[soa.qual_off[r]:soa.qual_off[r] + lq[r]]
        soa.seq, soa.qual, soa.seq_off, soa.qual_off = seq, qual, new_seq_off, new_qual_off
        return soa

    problems, stats, _ = _with_env("SMC_PIPE_CHUNKS", "4", lambda: run_case(PIPE_IVS[:2], SynthSpec(**PIPE_SPEC), VcParams(mtDepth=50, rpb=3.0),
                                                                          seed=59, mutate=reverse_payload))
    print(stats)
    assert stats["pipe_chunks"] == 4 and stats["pipe_launches"] == 1
    assert not problems, "\n".join(problems)


def test_pipelined_call_equals_resident_rerun():
    """smc_call_batch (pipelined) and a later smc_run_resident + smc_download of the same, now resident, batch give
    bit-identical outputs -- including the FP64 ones (the PI sums are order independent by construction)."""
    import numpy as np
    from smcounter_b200.caller import GpuCaller
    from smcounter_b200.synth import make_panel
    from smcounter_b200.targets import build_loci
    prm = VcParams(mtDepth=50, rpb=3.0)
    soa, refs, _ = make_panel(PIPE_IVS, SynthSpec(**PIPE_SPEC), seed=61)
    loci, _ = build_loci(PIPE_IVS, soa.chroms, refs)

    def go():
        c = GpuCaller(prm, 0)
        a = c.call(soa, loci)
        assert c.timings()["pipe_chunks"] == 6
        c.run()
        b = c.download(None)
        c.close()
        return a, b

    a, b = _with_env("SMC_PIPE_CHUNKS", "6", go)
    assert a.n_dyn == b.n_dyn and a.n_dyn > 0
    for f in ("loc", "cnt", "pi", "max_allele", "second_allele", "alt_allele", "alt_pi", "second_pi", "fl1", "fl2", "biallelic"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("fisher_p", "fisher_or"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f
    for f in ("dyn_locus", "dyn_kind", "dyn_site", "dyn_len", "dyn_iskey", "dyn_cnt", "dyn_pi"):
        assert np.array_equal(getattr(a, f)[:a.n_dyn], getattr(b, f)[:a.n_dyn]), f


def test_device_hp_lowcomp_matches_oracle():
    """smc_hp_lowcomp (k_hp_lowcomp, one warp per candidate) against the oracle's isHPorLowComp() restatement
    (smCounter.py:122-177) on every position of a sequence with planted homopolymers and dinucleotide repeats, for SNP,
    insertion and deletion alleles, three hpLen values, and positions at both contig ends."""
    import random
    from oracle import smcounter_oracle as orc
    from smcounter_b200 import rows
    from smcounter_b200.caller import GpuCaller
    rng = random.Random(3)
    rnd = lambda k: "".join(rng.choice("ACGT") for _ in range(k))
    seq = "GGGGGGGGGGGG" + rnd(150) + "A" * 12 + rnd(60) + "AT" * 15 + rnd(40) + "CCCCCCCCCG" + rnd(90) + "TGTGTGTGTGTGTGTGTGTGTG" + "n" + rnd(7)
    ref = orc.DictFasta({"c": seq})
    caller = GpuCaller(VcParams(mtDepth=10, rpb=2.0), 0)
    try:
        seen = set()
        for hp in (8, 10, 3):
            cands, want = [], []
            for pos0 in range(len(seq)):
                for (r, a) in ((seq[pos0].upper(), "C"), (seq[pos0].upper(), seq[pos0].upper() + "TT"), (seq[pos0:pos0 + 3].upper(), seq[pos0].upper()),
                               ("A", "AAAAAAAAAAAAA")):
                    win, wpos = rows.hp_window("c", pos0, hp, r, a, ref)
                    cands.append((win, wpos, r, a))
                    want.append(orc.is_hp_or_low_comp("c", str(pos0 + 1), hp, r, a, ref))
            flags = caller.hp_lowcomp(hp, cands)
            got = [(bool(f & 1), bool(f & 2)) for f in flags.tolist()]
            bad = [(k, cands[k], got[k], want[k]) for k in range(len(want)) if got[k] != tuple(want[k])]
            assert not bad, bad[:5]
            seen.update(got)
        assert seen == {(False, False), (True, False), (False, True), (True, True)}
        assert len(caller.hp_lowcomp(8, [])) == 0
    finally:
        caller.close()


def test_packed_payload_without_offsets():
    """smc_reads_soa with seq_off / qual_off / cigar_off = NULL (payload packed in read order): the offsets are derived on
    the device; results must match the oracle, with and without the chunked upload; a payload whose size contradicts the
    per-read lengths is refused."""
    import numpy as np
    from helpers import run_case
    from smcounter_b200.caller import GpuCaller
    from smcounter_b200.synth import make_panel
    from smcounter_b200.targets import build_loci
    spec = SynthSpec(**PIPE_SPEC)
    prm = VcParams(mtDepth=50, rpb=3.0)

    def packed(soa):
        out = soa.repack()
        assert out.packed and out.is_packed()
        return out

    for chunks in ("1", "5"):
        problems, stats, _ = _with_env("SMC_PIPE_CHUNKS", chunks, lambda: run_case(PIPE_IVS, spec, prm, seed=67, mutate=packed))
        print(stats)
        assert not problems, "\n".join(problems)
    soa, refs, _ = make_panel(PIPE_IVS[:1], spec, seed=67)
    soa = soa.repack()
    soa.seq = np.ascontiguousarray(soa.seq[:-3])           # three bytes short
    loci, _ = build_loci(PIPE_IVS[:1], soa.chroms, refs)
    c = GpuCaller(prm, 0)
    try:
        with pytest.raises(RuntimeError, match="packed payload"):
            c.call(soa, loci)
    finally:
        c.close()


@pytest.mark.parametrize("nq,seq_bits", [(3, 2), (7, 2), (20, 2), (3, 4)])
def test_compact_upload_encodings_give_identical_bits(nq, seq_bits):
    """ABI v3 wire encodings (16-bit scalars, 2- / 4-bit quality codes behind a LUT, 2-bit bases with the non-ACGT ones in a
    side list) are expanded on the device into the same arrays the plain upload makes: every count, PI and row must equal
    the oracle's on the plain reads, through the single-shot and the chunked upload.  The synthetic reads carry 'N's and
    insertion sites, so allele names are read back from the compact bases too."""
    import numpy as np
    from helpers import run_case
    qs = tuple(int(q) for q in np.linspace(40, 8, nq).round())
    spec = SynthSpec(**dict(PIPE_SPEC, q_values=qs, q_probs=tuple([1.0 / nq] * nq)))
    prm = VcParams(mtDepth=50, rpb=3.0)
    seen = {}

    def enc(soa):
        out = soa.trim_to_targets(PIPE_IVS).compact(seq_bits_wanted=seq_bits, scalar_bits_min=16 if nq == 7 else 8)
        seen.update(qual_bits=out.qual_bits, seq_bits=out.seq_bits, n_exc=0 if out.seq_exc is None else len(out.seq_exc[0]), scalar_bits=out.scalar_bits,
                    umi=str(out.umi.dtype), ref_id=str(out.ref_id.dtype))
        return out

    def twelve_nt(soa):                 # the synthetic barcodes are 16-mers (33-bit codes); QIAseq's are 12-mers: 25-bit codes, sent as uint32
        soa.umi = (np.uint64(1) << np.uint64(24)) | (soa.umi & np.uint64(0xFFFFFF))
        return soa

    for chunks, short in (("1", False), ("5", False), ("5", True)):
        problems, stats, _ = _with_env("SMC_PIPE_CHUNKS", chunks, lambda: run_case(PIPE_IVS, spec, prm, seed=71, mutate=twelve_nt if short else None,
                                                                                   gpu_mutate=enc))
        print(stats, seen)
        assert not problems, "\n".join(problems)
        assert seen["umi"] == ("uint32" if short else "uint64") and seen["ref_id"] == "uint8"
    assert seen["qual_bits"] == (2 if nq + 1 <= 4 else 4 if nq + 1 <= 16 else 8) and seen["scalar_bits"] == (16 if nq == 7 else 8)   # + quality 2 of the 'N's
    assert seen["seq_bits"] == seq_bits and (seq_bits == 4 or seen["n_exc"] > 0)


def test_fragment_ids_must_be_dense():
    """frag_id is part of the read sort key ((barcode slot, frag_id) in one sort), sized from n_reads: ids that are not
    dense first-appearance numbers (>= 2^ceil(log2 n_reads)) are refused instead of being sorted wrongly."""
    import numpy as np
    from smcounter_b200.caller import GpuCaller
    from smcounter_b200.synth import make_panel
    from smcounter_b200.targets import build_loci
    ivs = [("chr1", 1000, 1040)]
    soa, refs, _ = make_panel(ivs, SynthSpec(umis_per_locus=20, rpb=3.0), seed=71)
    loci, _ = build_loci(ivs, soa.chroms, refs)
    soa.frag_id = (soa.frag_id.astype(np.uint64) + np.uint64(4 * soa.n)).astype(np.uint32)
    c = GpuCaller(VcParams(mtDepth=20, rpb=3.0), 0)
    try:
        with pytest.raises(RuntimeError, match="dense id"):
            c.call(soa, loci)
    finally:
        c.close()


def test_reads_trimmed_to_their_targets():
    """smc_reads_soa with a stored window per read (store_lo / store_len): only the query bases between a read's first and
    last target position are in seq / qual.  The oracle sees the whole reads, the CUDA path the trimmed ones; every tally, PI,
    allele string and row must match -- with soft clips, indels (kept whole), N bases, the chunked upload and the packed offsets."""
    from helpers import run_case
    ivs = [("chr1", 1000, 1060), ("chr1", 1200, 1203), ("chr2", 300, 420), ("chr3", 50, 51)]
    spec = SynthSpec(umis_per_locus=60, rpb=3.0, snv_every=30, snv_vaf=0.1, indel_every=40, indel_vaf=0.1, n_frac=0.005, softclip_frac=0.4)
    prm = VcParams(mtDepth=60, rpb=3.0)
    for chunks in ("1", "4"):
        problems, stats, (soa, *_rest) = _with_env("SMC_PIPE_CHUNKS", chunks, lambda: run_case(
            ivs, spec, prm, seed=73, gpu_mutate=lambda s: s.trim_to_targets(ivs)))
        print(stats)
        assert stats["payload_bytes"] < 0.75 * (soa.seq.nbytes + soa.qual.nbytes)
        assert not problems, "\n".join(problems)


def test_malformed_stored_windows_are_refused():
    import numpy as np
    from smcounter_b200.caller import GpuCaller
    from smcounter_b200.synth import make_panel
    from smcounter_b200.targets import build_loci
    ivs = [("chr1", 1000, 1040)]
    prm = VcParams(mtDepth=20, rpb=3.0)
    soa, refs, _ = make_panel(ivs, SynthSpec(umis_per_locus=20, rpb=3.0), seed=79)
    loci, _ = build_loci(ivs, soa.chroms, refs)
    good = soa.trim_to_targets(ivs)
    k = int(np.flatnonzero((good.store_len > 8) & (good.store_len < good.l_seq))[0])

    def broken(edit):
        t = soa.trim_to_targets(ivs)
        t.packed = False                      # keep the explicit offsets: only the window itself is wrong
        edit(t)
        return t

    cases = (broken(lambda t: t.store_lo.__setitem__(k, t.store_lo[k] + 1)),                 # odd start
             broken(lambda t: t.store_len.__setitem__(k, t.store_len[k] - 4)),               # last target bases not stored
             broken(lambda t: t.store_len.__setitem__(k, t.l_seq[k] + 2)))                   # window longer than the read
    c = GpuCaller(prm, 0)
    try:
        c.call(good, loci)                                                                    # the unbroken one is fine
        for t in cases:
            with pytest.raises(RuntimeError, match="stored window"):
                c.call(t, loci)
    finally:
        c.close()


def _plant_ambiguity_codes(soa, chrom_idx, p, codes):
    """Rewrite the base at reference position ``p`` in every fully matched read of the deepest barcode there: fragment k of
    that barcode shows BAM nibble codes[k] (IUPAC ambiguity codes and N: each a distinct 'dynamic' allele), quality 37."""
    import numpy as np
    ends = soa.ref_end()
    plain = (soa.n_cigar == 1) & ((soa.cigar[soa.cigar_off] & 15) == 0)
    cover = np.flatnonzero((soa.ref_id == chrom_idx) & (soa.pos <= p) & (ends > p) & plain)
    umis, counts = np.unique(soa.umi[cover], return_counts=True)
    order = np.argsort(-counts)
    for u in umis[order]:
        reads = cover[soa.umi[cover] == u]
        frags = list(dict.fromkeys(soa.frag_id[reads].tolist()))
        if len(frags) >= len(codes):
            break
    else:
        raise AssertionError("no barcode with %d fragments over the locus" % len(codes))
    seq, qual = soa.seq.copy(), soa.qual.copy()
    for k, f in enumerate(frags[:len(codes)]):
        for r in reads[soa.frag_id[reads] == f]:
            q = int(p - soa.pos[r])
            b = int(soa.seq_off[r]) + (q >> 1)
            seq[b] = (seq[b] & 0x0F) | (codes[k] << 4) if q % 2 == 0 else (seq[b] & 0xF0) | codes[k]
            qual[int(soa.qual_off[r]) + q] = 37
    soa.seq, soa.qual = seq, qual
    return soa


@pytest.mark.parametrize("codes", [(15, 3, 5), (3, 5, 9, 10, 6, 12), (3, 5, 6, 7, 9, 10, 11, 12, 13, 14)])
def test_barcode_with_many_dynamic_alleles_at_one_locus(codes):
    """One barcode showing three / six / TEN distinct non-ACGT alleles at one locus (N and IUPAC codes stand in for the distinct
    insertion / deletion alleles a homopolymer collects): six dynamic alleles of a barcode live in shared memory, further ones in
    a spill record in global memory (up to 27); canonical order = ascending allele key.  calProb over 7 / 10 / 14 alleles must
    match the oracle.  The 10-allele case also starts with a one-record spill pool, so the pool has to grow (GF_SPILL_FULL)."""
    from helpers import run_case
    ivs = [("chr1", 1000, 1040)]
    spec = SynthSpec(umis_per_locus=25, rpb=9.0, snv_every=0, n_frac=0.0, softclip_frac=0.0, lowmapq_frac=0.0)

    def plant(s):
        s = _plant_ambiguity_codes(s, 0, 1020, codes)
        if len(codes) > 6:                                   # a second overflowing barcode at another locus: two spill records needed
            s = _plant_ambiguity_codes(s, 0, 1030, codes[::-1][:8])
        return s
    go = lambda: run_case(ivs, spec, VcParams(mtDepth=25, rpb=9.0), seed=83, mutate=plant)
    problems, stats, _ = _with_env("SMC_SPILL_CAP0", "1", go) if len(codes) > 6 else go()
    print(stats)
    assert stats["n_dyn"] >= len(codes)
    assert not problems, "\n".join(problems)


def _fisher_tables(n, seed, scales=(6, 60, 3000, 100000), near_sym_scale=30000):
    """Random 2x2 tables at smCounter's scales: tiny, panel-sized, 1e5-cell, exactly symmetric margins (mirrored outcomes tie with
    the observed one), near-symmetric ones, and degenerate rows / columns."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for scale in scales:
        t = rng.integers(0, scale, size=(n // 8, 4))
        out.append(t)
        u = rng.integers(0, scale, size=(n // 16, 4))                      # a strongly skewed second row: the SB / R1CP shape
        u[:, 2:] = rng.integers(0, max(2, scale // 50), size=(len(u), 2))
        out.append(u)
    sym = rng.integers(0, 400, size=(n // 8, 4))
    sym[:, 3] = np.maximum(sym[:, 0] + sym[:, 2] - sym[:, 1], 0)           # a + c == b + d when possible
    out.append(sym)
    sym2 = rng.integers(0, near_sym_scale, size=(n // 16, 4))
    sym2[:, 1] = sym2[:, 0] + rng.integers(-2, 3, size=len(sym2)); sym2[:, 3] = sym2[:, 2] + rng.integers(-2, 3, size=len(sym2))
    out.append(np.maximum(sym2, 0))
    deg = rng.integers(0, 50, size=(n // 16, 4))
    deg[np.arange(len(deg)), rng.integers(0, 4, size=len(deg))] = 0
    deg[::3, :2] = 0; deg[1::7, [0, 2]] = 0
    out.append(deg)
    out.append(np.array([[3, 0, 0, 3], [0, 5, 5, 0], [7, 0, 3, 0], [0, 0, 3, 4], [10, 20, 25, 15], [30000, 30100, 150, 50], [1, 1, 1, 1],
                         [0, 0, 0, 0], [45000, 44000, 12, 90], [2, 99999, 99999, 2]]))
    return np.concatenate(out).astype(np.int32)


def _scipy_fisher(t):
    import scipy.stats
    r = scipy.stats.fisher_exact([[int(t[0]), int(t[1])], [int(t[2]), int(t[3])]])
    return float(r[0]), float(r[1])


def _legacy_fisher(t):
    from oracle import smcounter_oracle as orc
    return orc.fisher_exact_legacy([[int(t[0]), int(t[1])], [int(t[2]), int(t[3])]])


P_FLOOR = 1e-280     # below this the recurrence weights underflow relative to the mode: only "p is that small" is asserted


def _fisher_diff(tables, got_p, got_or, want, p_rtol):
    import math
    from helpers import rel_err
    bad, worst, n_tiny = [], 0.0, 0
    for t, gp, go, (wo, wp) in zip(tables, got_p, got_or, want):
        same_or = (math.isnan(go) and math.isnan(wo)) or go == wo or rel_err(float(go), wo) <= 1e-12
        if wp < P_FLOOR:
            n_tiny += 1
            ok_p = float(gp) < 1e-270
        else:
            e = rel_err(float(gp), wp)
            worst = max(worst, e if math.isfinite(e) else 1.0)
            ok_p = e <= p_rtol
        if not ok_p or not same_or:
            bad.append((t.tolist(), float(gp), wp, float(go), wo))
    return bad, worst, n_tiny


def test_fisher_kernel_against_scipy_on_100k_tables():
    """k_fisher's arithmetic (smc_fisher_exact: the same device function smc_call_batch uses for SB / R1CP / R2CP / PrimerCP,
    smCounter.py:215-266) against scipy.stats.fisher_exact on 100 000 tables: p within 1e-9 relative, odds ratio identical
    (inf / nan included).  The pipeline itself only exercises a few dozen tables per parity case.  For p < 1e-280 -- 275 orders of
    magnitude below the filters' cut-offs (1e-3, 1e-5) -- only the smallness is asserted: the mode-relative weights underflow there."""
    import multiprocessing
    import os
    from smcounter_b200.caller import GpuCaller
    tables = _fisher_tables(100000, seed=1)
    assert len(tables) >= 90000
    c = GpuCaller(VcParams(mtDepth=100, rpb=3.0), 0)
    p, o = c.fisher_exact(tables)
    c.close()
    with multiprocessing.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
        want = pool.map(_scipy_fisher, tables, chunksize=512)
    bad, worst, n_tiny = _fisher_diff(tables, p, o, want, 1e-9)
    print("fisher: %d tables (%d with p < %g), worst relative p error above that %.2e, %d failures" % (len(tables), n_tiny, P_FLOOR, worst, len(bad)))
    assert not bad, bad[:5]


def test_fisher_kernel_legacy_scipy_semantics():
    """fisherLegacy = 1: the two-sided p of scipy <= 1.6 (epsilon = 1 - 1e-4), against the oracle's literal restatement of that
    algorithm; includes near-symmetric tables where the two eras differ by one mirrored outcome."""
    import multiprocessing
    import os
    from smcounter_b200.caller import GpuCaller
    tables = _fisher_tables(2400, seed=2, scales=(6, 60, 1500), near_sym_scale=4000)     # (the old algorithm is slow on huge tables)
    c = GpuCaller(VcParams(mtDepth=100, rpb=3.0, fisherLegacy=1), 0)
    p, o = c.fisher_exact(tables)
    c.close()
    c0 = GpuCaller(VcParams(mtDepth=100, rpb=3.0), 0)
    p0, _ = c0.fisher_exact(tables)
    c0.close()
    with multiprocessing.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
        want = pool.map(_legacy_fisher, tables, chunksize=128)
    bad, worst, _ = _fisher_diff(tables, p, o, want, 1e-9)
    n_era = int((abs(p - p0) > 1e-12 * abs(p0)).sum())
    print("legacy fisher: %d tables, worst relative p error %.2e, %d outside 1e-9; %d tables where the two scipy eras differ" % (len(tables), worst, len(bad), n_era))
    assert not bad, bad[:5]
    assert n_era > 0


def _fuzz_seeds():
    """Seeds of the differential fuzz; SMC_FUZZ_SEEDS=a:b runs a wider sweep by hand (e.g. 200:300)."""
    import os
    spec = os.environ.get("SMC_FUZZ_SEEDS")
    if spec:
        a, b = spec.split(":")
        return list(range(int(a), int(b)))
    return list(range(101, 113))


@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_randomised_parameters_and_panels(seed):
    """Differential fuzz: panel shape, read-error knobs and every vc() parameter drawn from a seeded generator; the CUDA path
    (alternating plain / packed / target-trimmed encodings and chunked uploads) against the oracle, field by field."""
    from helpers import run_case
    ivs, spec, prm = fuzz_case(seed)
    enc = seed % 4
    gpu_mutate = (None, lambda s: s.repack(), lambda s: s.trim_to_targets(ivs), lambda s: s.trim_to_targets(ivs).compact())[enc]
    chunks = "1" if seed % 2 else "3"
    problems, stats, _ = _with_env("SMC_PIPE_CHUNKS", chunks, lambda: run_case(ivs, spec, prm, seed=seed, gpu_mutate=gpu_mutate))
    print(seed, ivs, spec, prm, stats)
    assert not problems, "\n".join(problems)
