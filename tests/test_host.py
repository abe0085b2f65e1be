"""CPU tests of the host side of the product package (no compute calls): formatting with Python-2 semantics, target
list, FASTA access, SoA containers, repeat filters, writers, down-sampling emulation, sharding, BAM round trip.
Where the oracle restates the same reference code independently, the two implementations are checked against each
other; the writers are also checked against the reference's committed example outputs."""
import math
import os
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import py2compat
from oracle import smcounter_oracle as orc
from smcounter_b200 import downsample, repeats, rows, shard, soa, targets, writers
from smcounter_b200.caller import UmiKeep, VcParams
from smcounter_b200.fasta import FastaFile, SparseRef
from smcounter_b200.synth import SynthSpec, make_panel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------------- Py2 formatting
@settings(max_examples=300, deadline=None)
@given(st.integers(0, 200000), st.integers(1, 200000), st.sampled_from([2, 4]))
def test_py2round_matches_oracle_on_ratios(a, b, nd):
    x = 1.0 * a / b
    assert rows.py2round(x, nd) == py2compat.py2round(x, nd)


@pytest.mark.parametrize("x,nd,want", [(0.03125, 4, 0.0313), (0.00005, 4, 0.0001), (2.675, 2, 2.67), (0.125, 2, 0.13),
                                        (1.0, 4, 1.0), (0.0, 2, 0.0), (10892.584999999999, 2, 10892.58)])
def test_py2round_cases(x, nd, want):
    assert rows.py2round(x, nd) == want == py2compat.py2round(x, nd)


@settings(max_examples=200, deadline=None)
@given(st.floats(min_value=0.0, max_value=1e6, allow_nan=False))
def test_py2str_matches_oracle(x):
    assert rows.py2str(x) == py2compat.py2str(x)
    assert rows.py2str(round(x, 4)) == py2compat.py2str(round(x, 4))


def test_convert_to_vcf():
    for origRef, origAlt in (("A", "C"), ("G", "DEL"), ("T", "INS|T|TAC"), ("C", "DEL|CAG|C"), ("A", "N"), ("A", "XYZ")):
        assert rows.convert_to_vcf(origRef, origAlt) == orc.convert_to_vcf(origRef, origAlt)
    assert rows.convert_to_vcf("T", "INS|T|TAC") == ("T", "TAC", "INDEL")
    assert rows.convert_to_vcf("G", "DEL") == ("G", "DEL", "SDEL")


def test_hp_window_holds_every_base_the_reference_fetches():
    """rows.hp_window(): the one reference window per candidate that smc_hp_lowcomp gets must contain all six fetch() ranges
    of isHPorLowComp() (smCounter.py:127-129, 143-145), also at the contig ends."""
    rng = random.Random(3)
    seq = "".join(rng.choice("ACGTacgt") for _ in range(120))
    ref = orc.DictFasta({"c": seq})
    n = len(seq)
    for hp in (3, 8, 10):
        for pos0 in (0, 1, 5, 2 * hp, 60, n - 2 * hp - 1, n - 3, n - 1):
            for (r, a) in (("A", "C"), ("A", "ATT"), ("ACGT", "A")):
                win, wpos = rows.hp_window("c", pos0, hp, r, a, ref)
                w0 = pos0 - wpos
                assert w0 == max(0, pos0 - 2 * hp) and win == seq[w0:min(n, pos0 + max(len(r), len(a)) + 2 * hp)].upper()
                for flank in (hp, 2 * hp):
                    for b in (r, a):
                        lo, hi = pos0 + len(b), min(pos0 + len(b) + flank, n)
                        assert ref.fetch("c", max(0, pos0 - flank), pos0).upper() == win[max(0, wpos - flank):wpos]
                        assert ref.fetch("c", lo, hi).upper() == (win[lo - w0:hi - w0] if hi > lo else "")


# ---------------------------------------------------------------------------------------------- targets / fasta / SoA
def test_loc_list_and_build_loci(tmp_path):
    lines = ["track name=x\n", "chr2\t10\t14\textra\n", "chr1\t5\t8\n", "chr2\t12\t16\n", "chr1\t7\t7\n"]
    ivs = targets.intervals_from_bed_lines(lines)
    assert ivs == [("chr2", 10, 14), ("chr1", 5, 8), ("chr2", 12, 16), ("chr1", 7, 7)]
    assert targets.loc_list(ivs) == orc.loci_from_bed(lines)
    ref = SparseRef({"chr1": 100, "chr2": 100})
    ref.add_window("chr1", 0, np.frombuffer(b"ACGTACGTACGTACGT", dtype=np.uint8))
    loci, order = targets.build_loci(ivs, ["chr1", "chr2"], ref)
    assert loci.n == 3 + 6 and len(order) == 4 + 3 + 4
    keys = (loci.ref_id.astype(np.int64) << 32) | loci.pos0
    assert np.all(np.diff(keys) > 0)
    ll = targets.loc_list(ivs)
    for k, (c, p) in enumerate(ll):
        i = order[k]
        assert ["chr1", "chr2"][loci.ref_id[i]] == c and loci.pos0[i] + 1 == int(p)
    assert bytes(loci.ref_base[:3]) == b"CGT" and bytes(loci.ref_base[3:]) == b"NNNNNN"


def test_fasta_reader(tmp_path):
    p = tmp_path / "r.fa"
    p.write_text(">c1 desc\nACGTAC\nGTACGT\nAC\n>c2\nTTTT\n")
    fa = FastaFile(str(p))
    assert fa.get_reference_length("c1") == 14 and fa.get_reference_length("c2") == 4
    full = "ACGTACGTACGTAC"
    for s in range(14):
        for e in range(s, 18):
            assert fa.fetch("c1", s, e) == full[s:min(e, 14)]
    assert fa.fetch("c2", 1, 3) == "TT"
    with pytest.raises(ValueError):
        fa.fetch("c1", -1, 3)
    # with a .fai on disk
    (tmp_path / "r.fa.fai").write_text("c1\t14\t9\t6\t7\nc2\t4\t30\t4\t5\n")
    fb = FastaFile(str(p))
    assert fb.fetch("c1", 5, 13) == full[5:13] and fb.fetch("c2", 0, 99) == "TTTT"


def test_soa_roundtrip_and_select():
    spec = SynthSpec(umis_per_locus=12, rpb=2.0, snv_every=20, snv_vaf=0.3, indel_every=25, indel_vaf=0.3, softclip_frac=0.3)
    s, refs, _ = make_panel([("chr1", 300, 340), ("chr2", 100, 120)], spec, seed=4)
    recs = soa.soa_to_records(s, orc.Read)
    s2 = soa.records_to_soa(recs, s.chroms)
    for f in ("ref_id", "pos", "flag", "mapq", "nm", "l_seq", "seq_off", "qual_off", "cigar_off", "n_cigar", "umi", "seq", "qual", "cigar"):
        assert np.array_equal(getattr(s, f), getattr(s2, f)), f
    # frag ids are renumbered by first appearance: same partition
    assert len(set(zip(s.frag_id.tolist(), s2.frag_id.tolist()))) == len(set(s.frag_id.tolist()))
    idx = np.arange(0, s.n, 3)
    sub = s.select(idx)
    r_sub = soa.soa_to_records(sub, orc.Read)
    assert [r[2:] for r in r_sub] == [recs[i][2:] for i in idx]
    assert np.array_equal(s.ref_end()[idx], sub.ref_end())
    assert all(orc.reference_end(r) == e for r, e in zip(recs, s.ref_end()))


def test_umi_code_injective():
    tab = {}
    codes = {}
    for bc in ["", "A", "AA", "C", "ACGT", "TTTT", "ACGTN", "A" * 31, "A" * 32, "ACGTACGTACGTACGTACGTACGTACGTACGTACGT", "acgt"]:
        c = soa.umi_code(bc, tab)
        assert c not in codes.values() or codes.get(bc) == c
        codes[bc] = c
    assert soa.umi_string(codes["ACGT"]) == "ACGT" and soa.umi_string(codes["A" * 31]) == "A" * 31
    assert soa.umi_code("ACGTN", tab) == codes["ACGTN"]


def test_umikeep_layout():
    k = UmiKeep({7: [5, 3, 9], 2: [8]})
    assert k.locus.tolist() == [2, 7] and k.off.tolist() == [0, 1, 4] and k.umi.tolist() == [8, 3, 5, 9]
    assert VcParams(mtDepth=3612, rpb=8.6).ds == 7224 and VcParams(mtDepth=10, rpb=1.0, maxMT=7).ds == 7


# ---------------------------------------------------------------------------------------------- repeats / writers
def _rand_bed(rng, chroms, n, lo, hi, names=None):
    out = []
    for _ in range(n):
        c = rng.choice(chroms)
        s = rng.randrange(lo, hi)
        e = s + rng.randrange(1, 40)
        out.append((c, str(s), str(e)) + ((rng.choice(names),) if names else ()))
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_repeat_filters_match_oracle(seed):
    rng = random.Random(seed)
    chroms = ["chr1", "chr2", "1"]
    target = sorted(_rand_bed(rng, chroms[:2], 12, 0, 600), key=lambda r: (r[0], int(r[1])))
    trf = sorted(_rand_bed(rng, chroms, 25, 0, 600), key=lambda r: (r[0], int(r[1])))
    rm = sorted(_rand_bed(rng, chroms, 25, 0, 600, ["Simple_repeat", "Low_complexity", "Satellite", "L1"]), key=lambda r: (r[0], int(r[1])))
    out_rows = []
    for (c, s, e) in target:
        for p in range(int(s), int(e)):
            f = [c, str(p + 1), "A", rng.choice(["C", "DEL", "ACT"]), "SNP"] + ["1"] * 5 + \
                [orc.py2str(orc.py2round(rng.random() * 12, 2))] + ["1"] * 3 + ["0.5"] + ["0"] * 29 + [rng.choice([";", ";LSM;", ";LM;LSM;"])]
            out_rows.append("\t".join(f))
    out_rows.append("\t".join(["chr1", "9", "G"] + [""] * 41 + ["Zero_Coverage"]))
    trf_r, rm_r = repeats.build_repeat_regions(target, trf, rm)
    got = repeats.apply_repeat_filters(out_rows, trf_r, rm_r)
    want = orc.repeat_filter_rows(out_rows, target, trf, rm)
    assert got == want
    tags = {t for r in got for t in r.split("\t")[-1].split(";")}
    assert "RepT" in tags and ({"RepS", "LowC", "SL", "Other_Repeat"} & tags)
    assert got[-1].endswith("Zero_Coverage")


def test_bed_ops():
    rows_ = [("c", "1", "5", "x"), ("c", "5", "9", "y"), ("c", "20", "30", "x"), ("c", "25", "27", "x"), ("d", "0", "3", "z")]
    assert repeats.bed_merge(rows_) == [("c", 1, 9), ("c", 20, 30), ("d", 0, 3)]                    # book-ended features merge
    assert repeats.bed_merge(rows_, True) == [("c", 1, 9, "x,y"), ("c", 20, 30, "x"), ("d", 0, 3, "z")]
    assert repeats.bed_intersect([("c", 0, 25, "k")], [("c", 1, 9), ("c", 20, 30)]) == [("c", 1, 9, "k"), ("c", 20, 25, "k")]
    assert repeats.bed_sort([("d", 0, 3), ("c", 20, 30), ("c", 1, 9)]) == [("c", 1, 9), ("c", 20, 30), ("d", 0, 3)]


def test_writers_regenerate_golden_files(tmp_path):
    with open(os.path.join(GOLD, "example.smCounter.all.txt")) as fh:
        body = fh.read().split("\n")[1:-1]
    thr = writers.write_outputs(body, str(tmp_path / "example"), 3612, 0)
    assert thr == 58 == writers.pi_threshold(3612) and writers.pi_threshold(3612, 40) == 40
    for ext in ("all.txt", "cut.txt", "cut.vcf"):
        got = (tmp_path / ("example.smCounter." + ext)).read_text()
        want = open(os.path.join(GOLD, "example.smCounter." + ext)).read()
        if ext == "cut.vcf":        # the sample column is named after outPrefix (a path here)
            got = got.replace(str(tmp_path / "example"), "example")
        assert got == want, ext
    assert writers.vcf_header("s1") == orc.vcf_header("s1")
    # genotype hack branches (smCounter.py:868-882)
    base = body[299].split("\t")
    for alt, vmf, chrom, gt, ad in (("C,T", "0.5", "chr1", "1/2", ",1"), ("C", "0.97", "chr1", "1/1", ""), ("C", "0.5", "chrY", "1", ""),
                                    ("C", "0.5", "chr1", "0/1", "")):
        f = list(base)
        f[0], f[3], f[10], f[14] = chrom, alt, "99.0", vmf
        vcf, short = writers.called_lines(f, 58)
        assert vcf.rstrip("\n").split("\t")[-1].startswith(gt + ":") and short.split("\t")[3] == alt
        assert vcf.rstrip("\n").split("\t")[-1].split(":")[1].endswith(f[13] + ad)
    f = list(base); f[3] = "DEL"; f[10] = "99.0"
    assert writers.called_lines(f, 58) is None


# ---------------------------------------------------------------------------------------------- down-sampling emulation
def test_downsample_emulation_matches_oracle_restatement():
    assert downsample.py2_string_hash("a") == 12416037344 == py2compat.py2hash("a")
    rng = random.Random(11)
    for n, k in ((30, 12), (300, 120), (2000, 1500), (5000, 600)):
        bcs = ["".join(rng.choice("ACGT") for _ in range(12)) for _ in range(n)]
        bcs = list(dict.fromkeys(bcs))
        for s in bcs[:50]:
            assert downsample.py2_string_hash(s) == py2compat.py2hash(s)
        assert downsample.py2_dict_key_order(bcs) == py2compat.py2_dict_order(bcs)
        pop = downsample.py2_dict_key_order(bcs)
        assert downsample.py2_seeded_sample("41245237", pop, k) == py2compat.py2_sample("41245237", pop, k)
    assert downsample.py2_dict_key_order(["A", "T", "G", "C", "DEL", "N"]) == ["A", "C", "G", "N", "DEL", "T"]


# ---------------------------------------------------------------------------------------------- sharding
def test_shard_plan_covers_everything_and_balances():
    spec = SynthSpec(umis_per_locus=10, rpb=2.0, depth_sigma=0.8)
    ivs = [("chr1", 1000 + 400 * i, 1000 + 400 * i + 30 + 7 * (i % 5)) for i in range(12)] + [("chr2", 500, 520), ("chr2", 510, 530)]
    s, refs, _ = make_panel(ivs, spec, seed=9)
    w = shard.estimate_interval_events(s, ivs, s.chroms)
    assert w.shape == (len(ivs),) and (w > 0).all()
    # brute force check of one weight
    ends = s.ref_end()
    c, a, b = ivs[3]
    m = (s.ref_id == s.chroms.index(c))
    ov = np.minimum(ends[m], b) - np.maximum(s.pos[m].astype(np.int64), a)
    assert w[3] == ov[ov > 0].sum()
    for n in (1, 2, 3, 8):
        plan = shard.plan_shards(s, ivs, s.chroms, n)
        assert len(plan) == n
        allk = sorted(k for idxs, _ in plan for k in idxs)
        assert allk == list(range(len(ivs)))
        assert all(idxs == sorted(idxs) for idxs, _ in plan)
        if n == 2:
            loads = [sum(w[k] for k in idxs) for idxs, _ in plan]
            assert max(loads) / sum(loads) < 0.62
    plan = shard.plan_shards(s, ivs, s.chroms, 3)
    fake = [["%d:%d" % (k, j) for k in idxs for j in range(ivs[k][2] - ivs[k][1])] for idxs, _ in plan]
    merged = shard.interleave_rows(plan, ivs, fake)
    assert merged == ["%d:%d" % (k, j) for k in range(len(ivs)) for j in range(ivs[k][2] - ivs[k][1])]
    idx = shard.reads_for_intervals(s, [ivs[0], ivs[12]], s.chroms)
    assert len(idx) and np.all(np.diff(idx) > 0)
    for i in range(s.n):
        touches = any(s.chroms[s.ref_id[i]] == c and s.pos[i] < b and ends[i] > a for (c, a, b) in (ivs[0], ivs[12]))
        assert touches == (i in set(idx.tolist()))


def test_repack_puts_payload_in_read_order():
    """ReadsSoA.repack(): same reads, bases / qualities / CIGARs stored in read order (the layout a BAM decode delivers
    and the one smc_call_batch can overlap with compute); merge_soas() output already has it."""
    import numpy as np
    from oracle import smcounter_oracle as orc
    from smcounter_b200.soa import soa_to_records
    from smcounter_b200.synth import SynthSpec, make_panel, make_panel_mp
    ivs = [("chr1", 1000, 1100), ("chr2", 300, 380), ("chr1", 5000, 5060)]
    spec = SynthSpec(umis_per_locus=30, rpb=3.0, indel_every=30, indel_vaf=0.1, softclip_frac=0.2)
    soa, _, _ = make_panel(ivs, spec, seed=5)
    rev = soa.select(np.arange(soa.n))
    # scramble: store the payload back to front
    order = np.arange(soa.n)[::-1]
    sb = (soa.l_seq.astype(np.int64) + 1) // 2
    so = np.zeros(soa.n, np.int64); so[order] = np.concatenate(([0], np.cumsum(sb[order])))[:-1]
    seq = np.zeros_like(soa.seq)
    for r in range(soa.n):
        seq[so[r]:so[r] + sb[r]] = soa.seq[soa.seq_off[r]:soa.seq_off[r] + sb[r]]
    rev.seq, rev.seq_off = seq, so
    packed = rev.repack(block=37)
    a, b = soa_to_records(soa, orc.Read), soa_to_records(packed, orc.Read)
    assert all(x.seq == y.seq and x.qual == y.qual and x.cigar == y.cigar for x, y in zip(a, b))
    for f in ("seq_off", "qual_off", "cigar_off"):
        assert (np.diff(getattr(packed, f)) >= 0).all()
    merged, _, _ = make_panel_mp(ivs, spec, seed=5, workers=2)
    for f in ("seq_off", "qual_off", "cigar_off", ):
        assert (np.diff(getattr(merged, f)) >= 0).all()
    assert (np.diff(merged.ref_id.astype(np.int64) << 32 | merged.pos) >= 0).all()


def test_trim_to_targets_keeps_exactly_the_target_window():
    """ReadsSoA.trim_to_targets(): stored window = query bases from the first to the last target position of a plain read
    (even start), whole read otherwise; stored bytes identical to the same window of the untrimmed payload."""
    import numpy as np
    from smcounter_b200.synth import SynthSpec, make_panel
    ivs = [("chr1", 1000, 1100), ("chr2", 300, 380), ("chr1", 5000, 5060)]
    soa, _, _ = make_panel(ivs, SynthSpec(umis_per_locus=30, rpb=3.0, indel_every=30, indel_vaf=0.1, softclip_frac=0.3), seed=5)
    t = soa.trim_to_targets(ivs)
    assert t.packed and t.is_packed() and t.seq.nbytes + t.qual.nbytes < 0.7 * (soa.seq.nbytes + soa.qual.nbytes)
    ends = soa.ref_end()
    cidx = {c: i for i, c in enumerate(soa.chroms)}
    n_whole = 0
    for r in range(soa.n):
        lo, ln, L = int(t.store_lo[r]), int(t.store_len[r]), int(soa.l_seq[r])
        assert lo % 2 == 0 and 0 <= lo and lo + ln <= L
        assert (t.qual[t.qual_off[r]:t.qual_off[r] + ln] == soa.qual[soa.qual_off[r] + lo:soa.qual_off[r] + lo + ln]).all()
        a = t.seq[t.seq_off[r]:t.seq_off[r] + (ln + 1) // 2]
        b = soa.seq[soa.seq_off[r] + lo // 2:soa.seq_off[r] + lo // 2 + (ln + 1) // 2]
        assert (a[:ln // 2] == b[:ln // 2]).all() and (ln % 2 == 0 or (a[-1] >> 4) == (b[-1] >> 4))
        cig = soa.cigar[soa.cigar_off[r]:soa.cigar_off[r] + soa.n_cigar[r]]
        plain = all((int(w) & 15) in (0, 4) for w in cig) and sum((int(w) & 15) == 0 for w in cig) == 1
        if not plain:
            assert lo == 0 and ln == L
            n_whole += 1
            continue
        left_sp = int(cig[0]) >> 4 if (int(cig[0]) & 15) == 4 else 0
        tp = [p for (c, s, e) in ivs if cidx[c] == soa.ref_id[r] for p in range(max(s, int(soa.pos[r])), min(e, int(ends[r])))]
        if tp:
            assert lo == ((min(tp) - int(soa.pos[r]) + left_sp) & ~1) and lo + ln == max(tp) - int(soa.pos[r]) + left_sp + 1
        else:
            assert ln == 0
    assert n_whole > 0


def test_read_locator_and_batch_plan():
    """shard.ReadLocator.select() == brute-force overlap scan (coordinate-sorted and shuffled reads); plan_batches() cuts the
    BED order into consecutive batches under the given limits and keeps every interval exactly once."""
    import numpy as np
    from smcounter_b200.shard import ReadLocator, plan_batches, reads_for_intervals
    from smcounter_b200.synth import SynthSpec, make_panel
    ivs = [("chr1", 1000, 1100), ("chr2", 300, 380), ("chr1", 5000, 5060), ("chr1", 1080, 1150), ("chr3", 10, 11), ("chr2", 900, 905)]
    soa, _, _ = make_panel(ivs[:4], SynthSpec(umis_per_locus=30, rpb=3.0, indel_every=30, indel_vaf=0.1), seed=5)
    chroms = soa.chroms + ["chr3"]
    ends, starts = soa.ref_end(), soa.pos.astype(np.int64)

    def brute(s_, sel):
        keep = np.zeros(s_.n, dtype=bool)
        e_, st_ = s_.ref_end(), s_.pos.astype(np.int64)
        for (c, s, e) in sel:
            if c in s_.chroms:
                keep |= (s_.ref_id == s_.chroms.index(c)) & (st_ < e) & (e_ > s)
        return np.flatnonzero(keep)

    loc = ReadLocator(soa, chroms)
    assert loc.sorted
    for sel in (ivs, ivs[1:2], ivs[4:], [("chr1", 0, 10)], [("chr1", 1099, 1100)], []):
        assert np.array_equal(loc.select(sel), brute(soa, sel))
        assert np.array_equal(reads_for_intervals(soa, sel, chroms), brute(soa, sel))
    shuffled = soa.select(np.random.default_rng(1).permutation(soa.n))
    loc2 = ReadLocator(shuffled, chroms)
    assert not loc2.sorted and np.array_equal(loc2.select(ivs), brute(shuffled, ivs))
    idxs = list(range(len(ivs)))
    one = plan_batches(loc, ivs, idxs)
    assert one == [idxs]
    small = plan_batches(loc, ivs, idxs, max_loci=120)
    assert [k for b in small for k in b] == idxs and len(small) >= 3
    assert all(sum(ivs[k][2] - ivs[k][1] for k in b) <= 120 or len(b) == 1 for b in small)
    by_reads = plan_batches(loc, ivs, idxs, max_reads=loc.count_upper(*ivs[0]) + 1)
    assert [k for b in by_reads for k in b] == idxs and len(by_reads) >= 2


def test_bench_rank_batches_partition_the_panel():
    """bench.rank_batches(): at N ranks the seeded panel subset is split into disjoint interval groups that cover it (the product's
    shard.assign_intervals), every rank's share is cut into its distinct batches in BED order, and the reference arm and the CUDA
    arm print the same ``config``."""
    import importlib.util
    import os
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    args = types.SimpleNamespace(intervals=3, batches=2, seed=1, whole_reads=False, no_compact=False, workload="cfg2", gpus=3)
    world = 3
    want = bench.workload_intervals(args, world)
    assert len(want) == 3 * 2 * 3 and want == sorted(want, key=want.index)
    got = []
    for r in range(world):
        groups = bench.rank_batches(args, r, world)
        assert len(groups) == args.batches and all(groups)
        for g in groups:
            got.extend(g)
    assert sorted(got) == sorted(want) and len(got) == len(set(got))
    assert [iv for g in bench.rank_batches(args, 0, 1) for iv in g] == bench.workload_intervals(args, 1)
    cfg = bench.config_dict(args, world)
    assert cfg["loci_per_step"] == sum(e - s for (_, s, e) in want) and cfg["batches_per_step"] == 2
    for wl in bench.WORKLOADS:
        a2 = types.SimpleNamespace(intervals=bench.WORKLOADS[wl][3], batches=2, seed=1, workload=wl)
        ivs = bench.workload_intervals(a2, 2)
        assert len(ivs) == a2.intervals * 4 and len(set(ivs)) == len(ivs)
        assert bench.vc_params(a2).mtDepth == bench.WORKLOADS[wl][2]["mtDepth"]


def test_vectorised_float_columns_equal_the_scalar_py2_formatting():
    """rows._f4_column / _f2_column (numpy, table lookup) against the scalar py2round + repr path, including exact decimal
    ties (multiples of 1/32, odd multiples of 1/20000 and 1/200) where Python-2 rounds half away from zero."""
    import numpy as np
    rng = np.random.default_rng(0)
    den = rng.integers(1, 40000, size=60000)
    num = (rng.random(60000) * den).astype(np.int64)
    num[:2000] = np.arange(2000); den[:2000] = 32
    num[2000:4000] = np.arange(1, 4001, 2)[:2000]; den[2000:4000] = 20000
    num[4000:4010] = 0; den[4010:4020] = 0; num[4010:4020] = 0
    got = rows._f4_column(num, den)
    want = [rows._f4(int(a), int(b)) if b > 0 else "0.0" for a, b in zip(num, den)]
    assert got == want and "0.0313" in got and "1.0" in got
    x = np.concatenate([rng.random(40000) * 300, rng.random(20000) * 1e5, np.arange(0, 2000) / 8.0, np.arange(1, 4001, 2) / 200.0,
                        [0.0, 0.005, 0.015, 16.0, 1e9, 2.5e10]])
    assert rows._f2_column(x) == [rows._f2(float(v)) for v in x]
    for v in (0.125, 2.675, 1.005, 0.5, 105.585):
        assert rows._f2(v) == py2compat.py2str(py2compat.py2round(v, 2))


def test_missing_repeat_track_aborts_like_the_reference(tmp_path, capsys):
    """smCounter.py:700-710 runs bedtools on both tracks under check_call: a missing file aborts the run.  The product must
    not write PASS rows for variants inside repeats because a path was mistyped; opting out is explicit ('none')."""
    from smcounter_b200 import smCounter
    with pytest.raises(IOError, match="bedTandemRepeats"):
        smCounter._read_track(str(tmp_path / "nope.bed"), "--bedTandemRepeats", 3)
    with pytest.raises(IOError, match="bedRepeatMaskerSubset"):
        smCounter._read_track("/qgen/home/xuc/UCSC/SR_LC_SL.nochr.bed", "--bedRepeatMaskerSubset", 4)
    assert smCounter._read_track("none", "--bedTandemRepeats", 3) == []
    assert smCounter._read_track("", "--bedRepeatMaskerSubset", 4) == []
    assert "disabled" in capsys.readouterr().out
    (tmp_path / "rm.bed").write_text("chr1\t5\t9\tSimple_repeat\textra\nchr2 1 2 Satellite\n")
    assert smCounter._read_track(str(tmp_path / "rm.bed"), "--bedRepeatMaskerSubset", 4) == [("chr1", "5", "9", "Simple_repeat"), ("chr2", "1", "2", "Satellite")]


def _fabricated_results(n, seed):
    """Random per-locus device results (no GPU): every code path of the row formatter -- zero coverage, bi-allelic pairs, dynamic
    alleles, all FILTER bits, exact decimal ties in the rounded columns."""
    from smcounter_b200 import _ffi
    from smcounter_b200.caller import LocusResults
    rng = np.random.default_rng(seed)
    chroms = ["chr1", "chr2", "chrY"]
    ref_id = np.sort(rng.integers(0, 3, n)).astype(np.int32)
    pos0 = np.zeros(n, np.int32)
    for c in range(3):
        m = ref_id == c
        pos0[m] = np.sort(rng.choice(100000, int(m.sum()), replace=False))
    loci = soa.Loci(ref_id, pos0, np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy())
    res = LocusResults(n, 64)
    res.loc[:] = 0
    cvg = rng.integers(1, 120000, n)
    used = rng.integers(1, 8000, n)
    res.loc[_ffi.L_CVG], res.loc[_ffi.L_USEDMT] = cvg, used
    res.loc[_ffi.L_ALLFRAG], res.loc[_ffi.L_ALLMT], res.loc[_ffi.L_USEDFRAG] = rng.integers(0, 90000, n), rng.integers(1, 9000, n), rng.integers(0, 9000, n)
    for k in (_ffi.L_MT3, _ffi.L_MT5, _ffi.L_MT7, _ffi.L_MT10):
        res.loc[k] = rng.integers(0, 5000, n)
    res.loc[_ffi.L_STATUS] = np.where(rng.random(n) < 0.02, 1, 0)
    for a in range(5):
        res.cnt[a, _ffi.C_ALLELE] = (cvg * rng.random(n) * rng.choice([0, 0.001, 0.3, 1], n)).astype(np.int32)
        res.cnt[a, _ffi.C_MT] = (used * rng.random(n) * rng.choice([0, 0.01, 0.5, 1], n)).astype(np.int32)
        res.cnt[a, _ffi.C_STRONG] = rng.integers(0, 50, n)
        res.pi[a] = np.where(rng.random(n) < 0.3, 0.0, rng.random(n) * rng.choice([0.01, 1, 30, 20000], n))
    res.pi[0, :50] = np.arange(50) * 0.125 + 0.005                    # x.xx5 values that are exact in binary: round-half-away cases
    res.pi[1, :50] = np.arange(50) / 8.0 + 0.625
    res.cnt[0, _ffi.C_ALLELE, :64] = np.arange(64)                     # k / 32 fractions: exact ties at 4 decimals (1/32 = 0.03125)
    res.loc[_ffi.L_CVG, :64] = 32 * np.arange(1, 65)
    res.loc[_ffi.L_STATUS, :64] = 0
    nd = 40
    res.n_dyn = nd
    res.dyn_cnt[:nd] = rng.integers(0, 300, (nd, 13))
    res.dyn_pi[:nd] = rng.random(nd) * 80
    res.alt_allele[:] = rng.choice([0, 1, 2, 3, 4], n)
    res.second_allele[:] = rng.choice([0, 1, 3, 4], n)
    dynloc = rng.choice(n, nd, replace=False)
    res.alt_allele[dynloc] = 5 + np.arange(nd)
    res.biallelic[:] = (rng.random(n) < 0.1).astype(np.uint8)
    res.second_allele[dynloc[:10]] = 5 + np.arange(10)[::-1]
    res.biallelic[dynloc[:10]] = 1
    bits = [_ffi.F_LM, _ffi.F_LSM, _ffi.F_DP, _ffi.F_SB, _ffi.F_LOWQ, _ffi.F_R1CP, _ffi.F_R2CP, _ffi.F_PRIMERCP, _ffi.F_HPGATE]

    def rbits():
        f = np.zeros(n, np.uint32)
        for b in bits:
            f |= np.where(rng.random(n) < 0.15, b, 0).astype(np.uint32)
        return np.where(rng.random(n) < 0.5, f | _ffi.F_EVALUATED, 0).astype(np.uint32)
    res.fl1[:], res.fl2[:] = rbits(), rbits()
    hp = {}
    for i in range(n):
        for cand, fl in ((0, res.fl1), (1, res.fl2)):
            if fl[i] & _ffi.F_HPGATE:
                hp[(i, cand)] = (bool(rng.random() < 0.5), bool(rng.random() < 0.5))
    return chroms, loci, res, hp, rng


def test_native_output_stage_equals_the_python_rows_filters_and_writers(monkeypatch):
    """csrc/smc_rows.cpp (the product's output stage) against the Python restatement kept for this purpose: format_rows(),
    repeats.apply_repeat_filters() and writers.render_outputs() -- rows, repeat tags, PASS / strip, cut.txt and VCF lines, byte
    for byte, on fabricated device results."""
    from smcounter_b200 import _ffi
    chroms, loci, res, hp, rng = _fabricated_results(12000, seed=5)
    names = ["N", "INS|A|ATT", "DEL|ACG|A", "R", "INS|C|CGGGGGGGGGGGG"]
    monkeypatch.setattr(rows.AlleleNamer, "name", lambda self, a: _ffi.FIXED_NAMES[a] if a < 5 else names[(a - 5) % 5])
    refs = SparseRef({c: 200000 for c in chroms})
    order = rng.permutation(loci.n)[:9000]
    want = rows.format_rows(res, None, loci, chroms, refs, 10, order, workers=1, hp_flags=hp)
    got = rows.emit_rows(res, None, loci, chroms, refs, 10, order, hp_flags=hp)
    assert got.rows() == want
    assert rows.emit_rows(res, None, loci, chroms, refs, 10, order, hp_flags=hp, threads=1).all == got.all
    assert sum(1 for r in want if r.endswith("Zero_Coverage")) > 50 and sum(1 for r in want if "," in r.split("\t")[3]) > 20
    trf_rows = [("chr1", str(s), str(s + int(rng.integers(5, 400)))) for s in sorted(rng.integers(0, 100000, 300).tolist())]
    rm_rows = sorted([("chr%d" % rng.integers(1, 3), str(s), str(s + int(rng.integers(5, 300))), str(rng.choice(["Simple_repeat", "Low_complexity", "Satellite", "L1"])))
                      for s in rng.integers(0, 100000, 500).tolist()], key=lambda r: (r[0], int(r[1])))
    trf, rm = repeats.build_repeat_regions([(c, "0", "100000") for c in chroms], trf_rows, rm_rows)
    fin = repeats.apply_repeat_filters(want, trf, rm)
    thr, a, c, v = writers.render_outputs(fin, "pfx", 3000, 0)
    em = rows.emit_rows(res, None, loci, chroms, refs, 10, order, hp_flags=hp, finalize=True, threshold=thr, trf=trf, rm=rm)
    assert "\t".join(rows.headerAll) + "\n" + em.all.decode() == a
    assert "\t".join(rows.headerVariants) + "\n" + em.cut.decode() == c
    assert writers.vcf_header("pfx") + em.vcf.decode() == v
    assert c.count("\n") > 500 and "RepT" in a and "RepS" in a and "1/2" in v and "1/1" in v
    # regrouping by row ranges (what the CLI does per BED interval)
    x = em.slices(100, 200)
    assert x[0] == "".join(l + "\n" for l in fin[100:200]).encode()
    # a locus that needs a down-sampling mask has no row: same error as the Python formatter
    res.loc[_ffi.L_STATUS, int(order[7])] = _ffi.ST_NEED_DOWNSAMPLE
    with pytest.raises(RuntimeError, match="Exception thrown in vc"):
        rows.emit_rows(res, None, loci, chroms, refs, 10, order, hp_flags=hp)


def test_native_batch_packer_equals_select_plus_compact():
    """smc_soa_pack (include/smc_soa.h): the reads of a batch gathered and written in the compact wire encodings by one
    native pass == ReadsSoA.select(idx).compact() made with numpy, array for array; any thread count; plain encodings too."""
    import numpy as np
    from smcounter_b200 import _bamio
    from smcounter_b200.synth import SynthSpec, make_panel
    ivs = [("chr1", 1000, 1400), ("chr2", 500, 800)]
    soa, _, _ = make_panel(ivs, SynthSpec(umis_per_locus=25, rpb=2.5, n_frac=0.01, indel_every=60, indel_vaf=0.2), seed=9)
    rng = np.random.default_rng(3)
    for plain in (soa.repack(), soa.trim_to_targets(ivs)):
        for idx in (None, np.flatnonzero(rng.random(plain.n) < 0.6), np.zeros(0, np.int64)):
            for threads, sbits in ((1, 8), (5, 8), (3, 16)):
                want = (plain if idx is None else plain.select(idx)).compact(scalar_bits_min=sbits)
                plain.__dict__.pop("_upload_codebook", None)
                _bamio.upload_codebook(plain)["scalar_bits"] = sbits          # reads of 150 bases fit 8 bits; the 16-bit path is forced
                got = _bamio.pack_upload(plain, idx, threads=threads)
                assert (got.scalar_bits, got.qual_bits, got.seq_bits) == (want.scalar_bits, want.qual_bits, want.seq_bits) == (sbits, 2, 2)
                for f in ("ref_id", "pos", "flag", "mapq", "nm", "l_seq", "n_cigar", "umi", "frag_id", "seq", "qual", "cigar", "qual_lut"):
                    a, b = getattr(got, f), getattr(want, f)
                    if got.n == 0 and f in ("qual_lut", "ref_id", "umi"):
                        assert len(a) == len(b) or f == "qual_lut"
                        continue                        # the codebook / the widths are the whole file's, whatever the batch holds
                    assert a.dtype == b.dtype and np.array_equal(a, b), f
                if plain.store_lo is not None:
                    assert np.array_equal(got.store_lo, want.store_lo) and np.array_equal(got.store_len, want.store_len)
                for a, b in zip(got.seq_exc, want.seq_exc):
                    assert a.dtype == b.dtype and np.array_equal(a, b)
                if got.n:
                    r = got.n // 2
                    L = int(got.stored_len()[r])
                    assert got.compact_bases(r, 0, L) == want.compact_bases(r, 0, L)
    # 4-bit bases and a codebook that does not fit 4 bits
    soa2, _, _ = make_panel(ivs[:1], SynthSpec(umis_per_locus=10, rpb=2.0, q_values=tuple(range(10, 30)), q_probs=tuple([0.05] * 20)), seed=4)
    plain = soa2.repack()
    got, want = _bamio.pack_upload(plain, None, seq_bits=4), plain.compact(seq_bits_wanted=4)
    assert got.qual_bits == want.qual_bits == 8 and got.seq_bits == 4
    assert np.array_equal(got.seq, want.seq) and np.array_equal(got.qual, want.qual) and np.array_equal(got.seq_off, plain.seq_off)
    # a read index out of order is refused
    import pytest
    with pytest.raises(RuntimeError):
        _bamio.pack_upload(plain, np.array([3, 2], np.int64))


def test_native_order_stats_equal_numpy():
    """smc_soa_order_stats (include/smc_soa.h): reference ends, coordinate-order check and longest span of the interval locator."""
    import numpy as np
    from smcounter_b200 import _bamio
    from smcounter_b200.shard import ReadLocator
    from smcounter_b200.synth import SynthSpec, make_panel
    ivs = [("chr1", 1000, 1300), ("chr2", 200, 420)]
    soa, _, _ = make_panel(ivs, SynthSpec(umis_per_locus=20, rpb=2.0, indel_every=40, indel_vaf=0.3, softclip_frac=0.3), seed=2)
    soa = soa.repack()
    ends, is_sorted, span = _bamio.order_stats_native(soa, threads=3)
    want = soa._ref_end_compute()
    assert np.array_equal(ends, want) and is_sorted and span == int((want - soa.pos).max())
    loc = ReadLocator(soa, soa.chroms)
    assert loc.sorted and loc.max_span == span
    # out of order: swap two reads
    perm = np.arange(soa.n)
    perm[[5, 50]] = perm[[50, 5]]
    shuffled = soa.select(np.sort(perm))             # select() keeps order: build the disorder by hand on pos instead
    shuffled.pos = shuffled.pos.copy()
    shuffled.pos[7] = shuffled.pos[6] - 1 if shuffled.ref_id[7] == shuffled.ref_id[6] else shuffled.pos[7]
    _, is_sorted2, _ = _bamio.order_stats_native(shuffled, threads=2)
    key = (shuffled.ref_id.astype(np.int64) << 32) | shuffled.pos.astype(np.int64)
    assert is_sorted2 == bool(np.all(key[1:] >= key[:-1]))
