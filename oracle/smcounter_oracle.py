"""CPU oracle: a Python-3 restatement of smCounter's per-locus calling path (TEST INFRASTRUCTURE ONLY).

This file is the *checker* for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (cpu_baseline / ``--impl reference`` legs) may import it; the product package
``smcounter_b200`` never does and fails loudly when its CUDA library is missing.

It follows, function by function, /root/reference/smCounter.py (Python 2.7, cannot run here):

    calProb            smCounter.py:26-98      -> cal_prob()
    convertToVcf       smCounter.py:103-117    -> convert_to_vcf()
    isHPorLowComp      smCounter.py:122-177    -> is_hp_or_low_comp()
    filterVariants     smCounter.py:182-269    -> filter_variants()
    vc                 smCounter.py:274-600    -> vc()
    vc_wrapper         smCounter.py:605-611    -> vc_wrapper()
    main (post-proc)   smCounter.py:675-680, 699-901 -> loci_from_bed(), repeat_filter_rows(), write_outputs()

Third-party behaviour the reference leans on, restated here because the dependency is absent:

    pysam / htslib pileup (smCounter.py:316-317): ``pileup_column()`` restates htslib ``resolve_cigar2``
        (which op covers the column, query position, the peeked ``indel`` length, ``is_del``).  Version unpinned
        in the reference (README: "pysam"), so this part is **parity unpinned**.
    scipy.stats.fisher_exact (smCounter.py:215,238,248,260): the container's scipy (1.18.1) is called directly.
    CPython 2.7 ``round`` / ``str`` / dict order / ``random.sample``: oracle/py2compat.py.

Pinned by: mt_depths_lod.R:4-5 (PI of an 8-pair barcode with 7 concordant pairs = 3.5), the PI = 0.3
indel rows and the 2 000-row invariants of example/example.smCounter.all.txt (tests/test_oracle_golden.py).
The BAM/FASTA of the example are not distributed with the reference, so the full-file diff cannot be run:
**parity unpinned** for pysam column membership, Py2 dict order and Py2 sampling.

Canonical orders (where the reference depends on CPython-2 hash order):
  * fragments of a barcode are visited in ascending *first appearance in the BAM* of the fragment
    (``frag_order``); products in cal_prob() therefore have a defined multiplication sequence;
  * per-barcode alleles are visited A, C, DEL, T, G, N (the slot order of a Py2 8-slot table) and then any
    other allele string sorted by ``allele_sort_key``;
  * the prediction index of an allele is the *exactly rounded* sum of its per-barcode terms
    (``exact_sum``; identical to math.fsum up to 2**-108 truncation), which is independent of barcode
    order -- the reference's own summation order is CPython-2 dict order and unknowable here;
  * ties in the descending sort of finalDict (smCounter.py:534) are broken in Py2 dict order of the key
    set inserted canonically (A, T, G, C first, then others); see ``final_key_order``.
"""
from __future__ import annotations

import math
import os
import traceback
from collections import defaultdict, namedtuple

from .py2compat import py2round, py2str, py2_sample, py2_dict_order

pcr_no_error = 1.0 - 3e-5          # smCounter.py:20
atgc = ("A", "T", "G", "C")        # smCounter.py:21

# ----------------------------------------------------------------------------------------------
# record model (pysam-free).  One entry per BAM record, list kept in BAM (coordinate) order.
# ----------------------------------------------------------------------------------------------
Read = namedtuple("Read", "qname chrom pos flag mapq nm cigar seq qual")
# cigar: list of (op, length) with BAM op codes M0 I1 D2 N3 S4 H5 P6 =7 X8; seq: str; qual: sequence of int
# nm: int or None (no NM tag)

_REF_CONSUMING = (0, 2, 3, 7, 8)
_QUERY_CONSUMING = (0, 1, 4, 7, 8)


def reference_end(read):
    return read.pos + sum(l for (op, l) in read.cigar if op in _REF_CONSUMING)


class ReadIndex:
    """Reads grouped by contig in BAM order, with reference ends precomputed, for column lookups."""

    def __init__(self, reads):
        self.by_chrom = defaultdict(list)
        frag_order = {}
        for i, r in enumerate(reads):
            if r.flag & 0x4:
                continue  # htslib never piles up unmapped reads
            parts = r.qname.split(":")
            key = (parts[-2], ":".join(parts[:-2]))
            if key not in frag_order:
                frag_order[key] = len(frag_order)
            self.by_chrom[r.chrom].append((r.pos, reference_end(r), i, r))
        self.frag_order = frag_order
        self.max_span = {c: max((e - s for (s, e, _, _) in v), default=0) for c, v in self.by_chrom.items()}
        self.starts = {c: [s for (s, _, _, _) in v] for c, v in self.by_chrom.items()}

    def column(self, chrom, p):
        """Reads covering 0-based position p, in BAM order."""
        import bisect

        v = self.by_chrom.get(chrom)
        if not v:
            return
        starts = self.starts[chrom]
        lo = bisect.bisect_left(starts, p - self.max_span[chrom])
        hi = bisect.bisect_right(starts, p)
        for j in range(lo, hi):
            s, e, _, r = v[j]
            if s <= p < e:
                yield r


def pileup_column(read, p):
    """htslib ``resolve_cigar2`` for one read at 0-based column ``p`` (must satisfy pos <= p < ref_end).

    Returns (query_position, indel, is_del).  ``indel`` is the length of the insertion (>0) or deletion (<0)
    that *follows* p when p is the last reference base of the covering op and the next CIGAR op is I or D
    (a P op followed by I's is summed as htslib does).  For a D/N op covering p: is_del = 1 and
    query_position is the query index of the next aligned base (htslib's qpos).
    """
    x = read.pos
    y = 0
    cig = read.cigar
    n = len(cig)
    for k in range(n):
        op, l = cig[k]
        if op in (0, 7, 8, 2, 3):
            if x <= p < x + l:
                indel = 0
                if p == x + l - 1 and k + 1 < n:
                    op2, l2 = cig[k + 1]
                    if op2 == 2:
                        indel = -l2
                    elif op2 == 1:
                        indel = l2
                    elif op2 == 6 and k + 2 < n:
                        l3 = 0
                        for kk in range(k + 2, n):
                            op3, ll = cig[kk]
                            if op3 == 1:
                                l3 += ll
                            elif op3 in (2, 0, 3, 7, 8):
                                break
                        if l3 > 0:
                            indel = l3
                if op in (2, 3):
                    return y, indel, True
                return y + (p - x), indel, False
            x += l
            if op in (0, 7, 8):
                y += l
        elif op in (1, 4):
            y += l
    raise ValueError("column not covered by read")


# ----------------------------------------------------------------------------------------------
# canonical orders
# ----------------------------------------------------------------------------------------------
_FIXED_ORDER = {"A": 0, "C": 1, "DEL": 2, "T": 3, "G": 4, "N": 5}
_NT_CODE = {"=": 0, "A": 1, "C": 2, "M": 3, "G": 4, "R": 5, "S": 6, "V": 7, "T": 8, "W": 9, "Y": 10, "H": 11,
            "K": 12, "D": 13, "B": 14, "N": 15}


def allele_sort_key(a):
    """Canonical order of allele strings: A, C, DEL, T, G, then everything else by a structural key.

    The first six follow the slot order of a Py2 8-slot dict/set.  Other alleles
    (``N``/IUPAC bases, ``INS|s|s+ins``, ``DEL|s+del|s``) sort by (kind, site code, length, BAM nibbles),
    which is the order of the library's 64-bit allele signature.
    """
    if a in _FIXED_ORDER and a != "N":
        return (0, _FIXED_ORDER[a])
    if len(a) == 1:
        return (1, 0, _NT_CODE.get(a, 16), 0, ())
    kind, left, right = a.split("|")
    if kind == "INS":
        ins = right[1:]
        return (1, 1, _NT_CODE.get(left, 16), len(ins), tuple(_NT_CODE.get(c, 16) for c in ins))
    return (1, 2, _NT_CODE.get(right, 16), len(left) - 1, ())


def final_key_order(keys):
    """Tie-break order of finalDict keys for the stable sort at smCounter.py:534.

    Base keys follow the exact Py2 slot order (<=5 keys: A, C, DEL, T, G, N in an 8-slot table; >=6 keys: the
    32-slot order A, C, G, N, DEL, T); indel allele strings follow in ``allele_sort_key`` order (their true
    Py2 position depends on a string hash of reference bases; not reproduced -- documented deviation).
    """
    keys = list(keys)
    fixed = [k for k in keys if k in _FIXED_ORDER]
    other = sorted((k for k in keys if k not in _FIXED_ORDER), key=allele_sort_key)
    if len(keys) <= 5:
        fixed.sort(key=lambda k: _FIXED_ORDER[k])
    else:
        order32 = {"A": 0, "C": 1, "G": 2, "N": 3, "DEL": 4, "T": 5}
        fixed.sort(key=lambda k: order32[k])
    return fixed + other


# ----------------------------------------------------------------------------------------------
# exact (order-independent) summation of per-barcode prediction-index terms
# ----------------------------------------------------------------------------------------------
_LIMB_SHIFT = 108


def exact_sum(values):
    """Correctly rounded sum of non-negative doubles < 2**20 whose bits below 2**-108 are dropped.

    Every term here is ``-log10(1-p)`` in {0} U [2**-55, 16]; a 53-bit mantissa then never reaches below
    2**-108, so the truncation is void and the result equals math.fsum(values).  Integer addition is
    associative, so the CUDA path can accumulate the same fixed-point integer in any order and round once.
    """
    acc = 0
    for v in values:
        if v == 0.0:
            continue
        m, e = math.frexp(v)            # v = m * 2**e, 0.5 <= m < 1
        mi = int(m * (1 << 53))         # exact 53-bit integer
        sh = e - 53 + _LIMB_SHIFT
        acc += (mi << sh) if sh >= 0 else (mi >> -sh)
    return acc / (1 << _LIMB_SHIFT)     # int/int true division is correctly rounded in CPython


# ----------------------------------------------------------------------------------------------
# calProb (smCounter.py:26-98)
# ----------------------------------------------------------------------------------------------
def cal_prob(frags, mtDrop):
    """``frags``: list of [base, prob, pairOrder] in canonical fragment order (== oneBC.values()).

    Returns dict allele -> posterior, with keys in canonical allele order.
    """
    out = {}
    if len(frags) <= mtDrop:                                            # :28-32
        for c in atgc:
            out[c] = 0.0
        return out
    exist = sorted({f[0] for f in frags}, key=allele_sort_key)          # :47
    uniq = list(exist)                                                  # :48-54
    if len(uniq) < 4:
        for b in atgc:
            if b not in uniq:
                uniq.append(b)
                if len(uniq) == 4:
                    break
    uniq.sort(key=allele_sort_key)
    prodP = {b: 1.0 for b in uniq}                                      # :59-60
    cnt = defaultdict(int)
    rightP = 1.0
    for f in frags:                                                     # :62-77
        base, prob, order = f[0], f[1], f[2]
        if order != "Paired":
            prob = 0.1
        prodP[base] *= 1.0 - prob
        cnt[base] += 1
        for ch in uniq:
            if ch != base:
                prodP[ch] *= prob
        rightP *= 1.0 - prob
    pcrP = {}
    for ch in uniq:                                                     # :79-81
        ratio = (cnt[ch] + 0.5) / (len(frags) + 0.5 * len(uniq))
        pcrP[ch] = 10.0 ** (-6.0 * ratio)
    tmpOut = {}
    sumP = 0.0
    for key in uniq:                                                    # :83-93
        if key in exist:
            tmpOut[key] = pcr_no_error * prodP[key] + rightP * min(pcrP[ch] for ch in uniq if ch != key)
        else:
            t = rightP
            for ch in exist:
                if ch != key:
                    t *= pcrP[ch]
            tmpOut[key] = t
        sumP += tmpOut[key]
    for key in uniq:                                                    # :95-96
        out[key] = 0.0 if sumP <= 0 else tmpOut[key] / sumP
    return out


# ----------------------------------------------------------------------------------------------
# convertToVcf (smCounter.py:103-117)
# ----------------------------------------------------------------------------------------------
def convert_to_vcf(origRef, origAlt):
    vtype = "."
    ref = origRef
    alt = origAlt
    if len(origAlt) == 1:
        vtype = "SNP"
    elif origAlt == "DEL":
        vtype = "SDEL"
    else:
        vals = origAlt.split("|")
        if vals[0] in ("DEL", "INS"):
            vtype = "INDEL"
            ref = vals[1]
            alt = vals[2]
    return (ref, alt, vtype)


# ----------------------------------------------------------------------------------------------
# reference genome access (stands in for pysam.FastaFile)
# ----------------------------------------------------------------------------------------------
class DictFasta:
    """In-memory FASTA: {chrom: sequence}.  fetch() clips ``end`` at the contig length like pysam."""

    def __init__(self, seqs):
        self.seqs = seqs

    def fetch(self, reference, start, end):
        s = self.seqs[reference]
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        return s[start:min(end, len(s))]

    def get_reference_length(self, reference):
        return len(self.seqs[reference])


# ----------------------------------------------------------------------------------------------
# isHPorLowComp (smCounter.py:122-177)
# ----------------------------------------------------------------------------------------------
def is_hp_or_low_comp(chrom, pos, length, refb, altb, refs):
    chromLength = refs.get_reference_length(chrom)
    pos0 = int(pos) - 1
    Lseq = refs.fetch(chrom, max(0, pos0 - length), pos0).upper()
    Rseq_ref = refs.fetch(chrom, pos0 + len(refb), min(pos0 + len(refb) + length, chromLength)).upper()
    Rseq_alt = refs.fetch(chrom, pos0 + len(altb), min(pos0 + len(altb) + length, chromLength)).upper()
    refSeq = Lseq + refb + Rseq_ref
    altSeq = Lseq + altb + Rseq_alt
    homop = any(refSeq.find(c * length) >= 0 or altSeq.find(c * length) >= 0 for c in "ATGC")

    len2 = 2 * length
    LseqLC = refs.fetch(chrom, max(0, pos0 - len2), pos0).upper()
    Rseq_refLC = refs.fetch(chrom, pos0 + len(refb), min(pos0 + len(refb) + len2, chromLength)).upper()
    Rseq_altLC = refs.fetch(chrom, pos0 + len(altb), min(pos0 + len(altb) + len2, chromLength)).upper()
    refSeqLC = LseqLC + refb + Rseq_refLC
    altSeqLC = LseqLC + altb + Rseq_altLC
    lowcomp = False
    for seq in (refSeqLC, altSeqLC):
        if lowcomp:
            break
        for i in range(len(seq) - len2):
            sub = seq[i:i + len2]
            counts = sorted((sub.count("A"), sub.count("T"), sub.count("G"), sub.count("C")), reverse=True)
            if 1.0 * (counts[0] + counts[1]) / len2 >= 0.99:
                lowcomp = True
                break
    return (homop, lowcomp)


# ----------------------------------------------------------------------------------------------
# filterVariants (smCounter.py:182-269)
# ----------------------------------------------------------------------------------------------
def fisher_exact(table):
    import scipy.stats

    res = scipy.stats.fisher_exact(table)
    return float(res[0]), float(res[1])


def fisher_exact_legacy(table):
    """``scipy.stats.fisher_exact(table)`` (two-sided) as scipy <= 1.6 computed it -- the scipy of the reference's day (the
    reference pins no version): ``epsilon = 1 - 1e-4`` in the "pexact is the mode" shortcut and in the search for the
    opposite tail.  Restated from scipy/stats/stats.py of that era over today's hypergeom pmf / cdf / sf."""
    import numpy as np
    from scipy.stats import hypergeom
    c = np.asarray(table, dtype=np.int64)
    if 0 in c.sum(axis=0) or 0 in c.sum(axis=1):
        return float("nan"), 1.0
    oddsratio = c[0, 0] * c[1, 1] / float(c[1, 0] * c[0, 1]) if c[1, 0] > 0 and c[0, 1] > 0 else float("inf")
    n1, n2, n = int(c[0, 0] + c[0, 1]), int(c[1, 0] + c[1, 1]), int(c[0, 0] + c[1, 0])
    a = int(c[0, 0])
    pmf = lambda x: float(hypergeom.pmf(x, n1 + n2, n1, n))

    def binary_search(side):
        if side == "upper":
            minval, maxval = mode, n
        else:
            minval, maxval = 0, mode
        guess = -1
        while maxval - minval > 1:
            if maxval == minval + 1 and guess == minval:
                guess = maxval
            else:
                guess = (maxval + minval) // 2
            pguess = pmf(guess)
            ng = guess - 1 if side == "upper" else guess + 1
            if pguess <= pexact < pmf(ng):
                break
            elif pguess < pexact:
                maxval = guess
            else:
                minval = guess
        if guess == -1:
            guess = minval
        if side == "upper":
            while guess > 0 and pmf(guess) < pexact * epsilon:
                guess -= 1
            while pmf(guess) > pexact / epsilon:
                guess += 1
        else:
            while pmf(guess) < pexact * epsilon:
                guess += 1
            while guess > 0 and pmf(guess) > pexact / epsilon:
                guess -= 1
        return guess

    mode = int(float((n + 1) * (n1 + 1)) / (n1 + n2 + 2))
    pexact = pmf(a)
    pmode = pmf(mode)
    epsilon = 1 - 1e-4
    if abs(pexact - pmode) / max(pexact, pmode) <= 1 - epsilon:
        return oddsratio, 1.0
    elif a < mode:
        plower = float(hypergeom.cdf(a, n1 + n2, n1, n))
        if pmf(n) > pexact / epsilon:
            return oddsratio, plower
        guess = binary_search("upper")
        pvalue = plower + float(hypergeom.sf(guess - 1, n1 + n2, n1, n))
    else:
        pupper = float(hypergeom.sf(a - 1, n1 + n2, n1, n))
        if pmf(0) > pexact / epsilon:
            return oddsratio, pupper
        guess = binary_search("lower")
        pvalue = pupper + float(hypergeom.cdf(guess, n1 + n2, n1, n))
    return oddsratio, min(pvalue, 1.0)


def filter_variants(ref, alt, vtype, origAlt, origRef, usedMT, strongMTCnt, chrom, pos, hpLen, refs, MTCnt,
                    alleleCnt, cvg, discordPairCnt, concordPairCnt, reverseCnt, forwardCnt, lowQReads,
                    r1BcEndPos, r2BcEndPos, r2PrimerEndPos, primerDist, dbg=None):
    """Returns the FILTER accumulator string (';' = pass).  ``dbg`` (dict) receives every intermediate."""
    if dbg is None:
        dbg = {}
    fltr = ";"
    if usedMT < 5:
        fltr += "LM;"
    if strongMTCnt[origAlt] < 2:
        fltr += "LSM;"
    (isHomopolymer, isLowComplexity) = is_hp_or_low_comp(chrom, pos, hpLen, ref, alt, refs)
    dbg["hp"], dbg["lowc"] = isHomopolymer, isLowComplexity
    if isHomopolymer and 1.0 * MTCnt[origAlt] / usedMT < 0.99:
        fltr += "HP;"
    if isLowComplexity and 1.0 * MTCnt[origAlt] / usedMT < 0.99:
        fltr += "LowC;"
    af_alt = 100.0 * alleleCnt[origAlt] / cvg
    pairs = discordPairCnt[origAlt] + concordPairCnt[origAlt]
    if pairs >= 1000 and 1.0 * discordPairCnt[origAlt] / pairs >= 0.5:
        fltr += "DP;"
    elif af_alt <= 60.0:
        refR = reverseCnt[origRef]
        refF = forwardCnt[origRef]
        altR = reverseCnt[origAlt]
        altF = forwardCnt[origAlt]
        oddsRatio, pvalue = fisher_exact([[refR, refF], [altR, altF]])
        dbg["sb"] = ((refR, refF, altR, altF), oddsRatio, pvalue)
        if pvalue < 0.00001 and (oddsRatio >= 50 or oddsRatio <= 1.0 / 50):
            fltr += "SB;"
    if vtype == "SNP" and origAlt in alleleCnt.keys() and origAlt in lowQReads.keys():
        bqAlt = 1.0 * lowQReads[origAlt] / alleleCnt[origAlt]
    else:
        bqAlt = 0.0
    if bqAlt > 0.4:
        fltr += "LowQ;"
    if vtype == "SNP":
        endBase = 20
        for tag, pool in (("r1", r1BcEndPos), ("r2", r2BcEndPos)):
            refLeEnd = sum(d <= endBase for d in pool[origRef])
            refGtEnd = len(pool[origRef]) - refLeEnd
            altLeEnd = sum(d <= endBase for d in pool[origAlt])
            altGtEnd = len(pool[origAlt]) - altLeEnd
            oddsRatio, pvalue = fisher_exact([[refLeEnd, refGtEnd], [altLeEnd, altGtEnd]])
            dbg[tag] = ((refLeEnd, refGtEnd, altLeEnd, altGtEnd), oddsRatio, pvalue)
            if pvalue < 0.001 and oddsRatio < 0.05 and af_alt <= 60.0:
                fltr += "R1CP;" if tag == "r1" else "R2CP;"
        endBase = primerDist
        refLeEnd = sum(d <= endBase for d in r2PrimerEndPos[origRef])
        refGtEnd = len(r2PrimerEndPos[origRef]) - refLeEnd
        altLeEnd = sum(d <= endBase for d in r2PrimerEndPos[origAlt])
        altGtEnd = len(r2PrimerEndPos[origAlt]) - altLeEnd
        oddsRatio, pvalue = fisher_exact([[refLeEnd, refGtEnd], [altLeEnd, altGtEnd]])
        dbg["primer"] = ((refLeEnd, refGtEnd, altLeEnd, altGtEnd), oddsRatio, pvalue)
        if altLeEnd + altGtEnd > 0:
            if 1.0 * altLeEnd / (altLeEnd + altGtEnd) >= 0.98 or (pvalue < 0.001 and oddsRatio < 1.0 / 20):
                fltr += "PrimerCP;"
    return fltr


# ----------------------------------------------------------------------------------------------
# vc (smCounter.py:274-600)
# ----------------------------------------------------------------------------------------------
def vc(index, chrom, pos, minBQ, minMQ, mtDepth, rpb, hpLen, mismatchThr, mtDrop, maxMT, primerDist, refs,
       keep_umis=None, detail=None, sampler="py2"):
    """One locus.  ``index``: ReadIndex (stands in for the BAM); ``pos``: 1-based position as a *str*.

    ``keep_umis``: optional explicit set of barcodes to keep when down-sampling fires (else the Py2
    ``random.sample`` restatement is used).  ``detail``: optional dict that receives every integer tally
    and float the CUDA kernels are diffed against.
    Returns the 45-field tab-joined row exactly as the reference's vc() does.
    """
    cvg = 0
    bcDict = {}                       # BC -> {readid -> [base, prob, pairOrder, frag_order]}  (insertion ordered)
    allBcDict = defaultdict(list)
    alleleCnt = defaultdict(int)
    MTCnt = defaultdict(int)
    r1BcEndPos = defaultdict(list)
    r2BcEndPos = defaultdict(list)
    r2PrimerEndPos = defaultdict(list)
    MT3Cnt = MT5Cnt = MT7Cnt = MT10Cnt = 0
    strongMTCnt = defaultdict(int)
    forwardCnt = defaultdict(int)
    reverseCnt = defaultdict(int)
    concordPairCnt = defaultdict(int)
    discordPairCnt = defaultdict(int)
    lowQReads = defaultdict(int)

    if rpb < 1.5:                                                        # :303-308
        smt = 2.0
    elif rpb < 3.0:
        smt = 3.0
    else:
        smt = 4.0

    p1 = int(pos)
    origRef = refs.fetch(chrom, p1 - 1, p1).upper()                      # :311-313

    pairOrder = None
    for read in index.column(chrom, p1 - 1):                             # :316-317
        qpos, indel, is_del = pileup_column(read, p1 - 1)
        qnameSplit = read.qname.split(":")                               # :319-325
        readid = ":".join(qnameSplit[:-2])
        BC = qnameSplit[-2]
        mq = read.mapq
        NM = read.nm if read.nm is not None else 0                       # :329-334
        nIndel = 0                                                       # :336-349
        leftSP = 0
        for cigarOrder, (op, value) in enumerate(read.cigar, 1):
            if op == 1 or op == 2:
                nIndel += value
            if cigarOrder == 1 and op == 4:
                leftSP = value
        mismatch = max(0, NM - nIndel)                                   # :352
        readLen = len(read.seq)                                          # :354 query_length
        mismatchPer100b = 100.0 * mismatch / readLen if readLen > 0 else 0.0
        if read.flag & 0x40:                                             # :359-362
            pairOrder = "R1"
        if read.flag & 0x80:
            pairOrder = "R2"
        if pairOrder is None:
            raise NameError("pairOrder referenced before assignment (read is neither read1 nor read2)")
        is_reverse = bool(read.flag & 0x10)
        cvg += 1                                                         # :368
        qal = readLen - soft_clip_total(read.cigar)                      # query_alignment_length

        if indel > 0:                                                    # :371-389
            site = read.seq[qpos]
            inserted = read.seq[qpos + 1: qpos + 1 + indel]
            base = "INS|" + site + "|" + site + inserted
            bq = read.qual[qpos]
            incCond = bq >= minBQ and mq >= minMQ and mismatchPer100b <= mismatchThr
            alleleCnt[base] += 1
            if is_reverse:
                reverseCnt[base] += 1
            else:
                forwardCnt[base] += 1
        elif indel < 0:                                                  # :392-411
            site = read.seq[qpos]
            deleted = refs.fetch(chrom, p1, p1 + abs(indel)).upper()
            base = "DEL|" + site + deleted + "|" + site
            bq = read.qual[qpos]
            incCond = bq >= minBQ and mq >= minMQ and mismatchPer100b <= mismatchThr
            alleleCnt[base] += 1
            if is_reverse:
                reverseCnt[base] += 1
            else:
                forwardCnt[base] += 1
        else:
            if is_del:                                                   # :416-421
                base = "DEL"
                bq = minBQ
                incCond = bq >= minBQ and mq >= minMQ and mismatchPer100b <= mismatchThr
            else:                                                        # :423-457
                base = read.seq[qpos]
                bq = read.qual[qpos]
                if bq < minBQ:
                    lowQReads[base] += 1
                incCond = bq >= minBQ and mq >= minMQ and mismatchPer100b <= mismatchThr
                if pairOrder == "R1":
                    if is_reverse:
                        distToBcEnd = qal - (qpos - leftSP)
                    else:
                        distToBcEnd = qpos - leftSP
                    if incCond:
                        r1BcEndPos[base].append(distToBcEnd)
                if pairOrder == "R2":
                    if is_reverse:
                        distToBcEnd = qpos - leftSP
                        distToPrimerEnd = qal - (qpos - leftSP)
                    else:
                        distToBcEnd = qal - (qpos - leftSP)
                        distToPrimerEnd = qpos - leftSP
                    if incCond:
                        r2BcEndPos[base].append(distToBcEnd)
                        r2PrimerEndPos[base].append(distToPrimerEnd)
                if is_reverse:
                    reverseCnt[base] += 1
                else:
                    forwardCnt[base] += 1
            alleleCnt[base] += 1                                         # :459

        if readid not in allBcDict[BC]:                                  # :463-464
            allBcDict[BC].append(readid)

        if incCond:                                                      # :467-479
            frs = bcDict.setdefault(BC, {})
            if readid not in frs:
                prob = pow(10.0, -bq / 10.0)
                frs[readid] = [base, prob, pairOrder, index.frag_order[(BC, readid)]]
            elif base == frs[readid][0] or base in ["N", "*"]:
                frs[readid][1] = max(pow(10.0, -bq / 10.0), frs[readid][1])
                frs[readid][2] = "Paired"
                if base == frs[readid][0]:
                    concordPairCnt[base] += 1
            else:
                del frs[readid]
                discordPairCnt[base] += 1

    allMT = len(allBcDict)                                               # :482-483
    allFrag = sum(len(allBcDict[bc]) for bc in allBcDict)
    ds = maxMT if maxMT > 0 else int(py2round(2.0 * mtDepth, 0))         # :486
    usedMT = min(ds, len(bcDict))                                        # :489

    if detail is not None:
        detail.update(dict(cvg=cvg, allMT=allMT, allFrag=allFrag, nBC=len(bcDict), ds=ds, origRef=origRef,
                           bcKeysAll=list(bcDict.keys())))

    if usedMT == 0:                                                      # :492-494
        return "\t".join([chrom, pos, origRef] + [""] * 41 + ["Zero_Coverage"])

    if len(bcDict) > ds:                                                 # :496-500
        if keep_umis is not None:
            bcKeys = [bc for bc in bcDict if bc in keep_umis]
            assert len(bcKeys) == ds
        elif sampler == "py2":
            population = py2_dict_order(list(bcDict.keys()))
            bcKeys = py2_sample(pos, population, ds)
        else:
            raise ValueError("down-sampling needed but no sampler")
    else:
        bcKeys = list(bcDict.keys())
    usedFrag = sum(len(bcDict[bc]) for bc in bcKeys)                     # :501

    piTerms = defaultdict(list)        # finalDict, kept as term lists for the exact sum
    keyFirstSeen = []
    for bc in bcKeys:                                                    # :506-532
        frags = sorted(bcDict[bc].values(), key=lambda f: f[3])          # canonical fragment order
        bcProb = cal_prob(frags, mtDrop)
        predIndex = {}
        for ch in bcProb:
            x = 1.0 - bcProb[ch]
            log10P = -math.log10(x) if x > 0.0 else 16.0
            predIndex[ch] = log10P
            if ch not in piTerms:
                keyFirstSeen.append(ch)
            piTerms[ch].append(log10P)
        mx = max(predIndex.values())
        max_base = [x for x in predIndex if predIndex[x] == mx]
        if len(max_base) == 1:
            cons = max_base[0]
            MTCnt[cons] += 1
            if predIndex[cons] > smt:
                strongMTCnt[cons] += 1
        elif len(frags) == 1:
            cons = frags[0][0]
            MTCnt[cons] += 1
        nf = len(frags)
        if nf >= 3:
            MT3Cnt += 1
        if nf >= 5:
            MT5Cnt += 1
        if nf >= 7:
            MT7Cnt += 1
        if nf >= 10:
            MT10Cnt += 1

    finalDict = defaultdict(float)
    for ch in final_key_order(piTerms.keys()):
        finalDict[ch] = exact_sum(piTerms[ch])

    sortedList = sorted(finalDict.items(), key=lambda kv: kv[1], reverse=True)   # :534 (stable)
    maxBase, maxPI = sortedList[0]
    secondMaxBase, secondMaxPI = sortedList[1]
    origAlt = secondMaxBase if maxBase == origRef else maxBase           # :541-542
    altPI = secondMaxPI if maxBase == origRef else maxPI
    (ref, alt, vtype) = convert_to_vcf(origRef, origAlt)                 # :545

    fa = (usedMT, strongMTCnt, chrom, pos, hpLen, refs, MTCnt, alleleCnt, cvg, discordPairCnt, concordPairCnt,
          reverseCnt, forwardCnt, lowQReads, r1BcEndPos, r2BcEndPos, r2PrimerEndPos, primerDist)
    dbg1, dbg2 = {}, {}
    fltr = ";"                                                           # :548-550
    if altPI >= 5 and vtype in ("SNP", "INDEL"):
        fltr = filter_variants(ref, alt, vtype, origAlt, origRef, *fa, dbg=dbg1)

    mfAlt = 1.0 * MTCnt[maxBase] / usedMT                                # :553-573
    mfAlt2 = 1.0 * MTCnt[secondMaxBase] / usedMT
    firstAlt = origAlt
    biallelic = False
    if maxBase != origRef and secondMaxBase != origRef and mfAlt >= 0.45 and mfAlt2 >= 0.45:
        biallelic = True
        origAlt2 = secondMaxBase
        (ref2, alt2, vtype2) = convert_to_vcf(origRef, origAlt2)
        fltr2 = ";"
        if secondMaxPI >= 5 and vtype2 in ("SNP", "INDEL"):
            fltr2 = filter_variants(ref2, alt2, vtype2, origAlt2, origRef, *fa, dbg=dbg2)
        if fltr == ";" and fltr2 == ";":
            alt = alt + "," + alt2
            vtype = vtype.lower() + "," + vtype2.lower()
        elif fltr != ";" and fltr2 == ";":
            alt = alt2
            fltr = fltr2
            origAlt = origAlt2

    frac_alt = py2round((1.0 * alleleCnt[origAlt] / cvg), 4)             # :576-600
    frac_A = py2round((1.0 * alleleCnt["A"] / cvg), 4)
    frac_T = py2round((1.0 * alleleCnt["T"] / cvg), 4)
    frac_G = py2round((1.0 * alleleCnt["G"] / cvg), 4)
    frac_C = py2round((1.0 * alleleCnt["C"] / cvg), 4)
    fracs = (alleleCnt["A"], alleleCnt["T"], alleleCnt["G"], alleleCnt["C"], frac_A, frac_T, frac_G, frac_C)
    MT_f_alt = py2round((1.0 * MTCnt[origAlt] / usedMT), 4)
    MT_f_A = py2round((1.0 * MTCnt["A"] / usedMT), 4)
    MT_f_T = py2round((1.0 * MTCnt["T"] / usedMT), 4)
    MT_f_G = py2round((1.0 * MTCnt["G"] / usedMT), 4)
    MT_f_C = py2round((1.0 * MTCnt["C"] / usedMT), 4)
    MTs = (MT3Cnt, MT5Cnt, MT7Cnt, MT10Cnt, MTCnt["A"], MTCnt["T"], MTCnt["G"], MTCnt["C"], MT_f_A, MT_f_T,
           MT_f_G, MT_f_C)
    strongMT = (strongMTCnt["A"], strongMTCnt["T"], strongMTCnt["G"], strongMTCnt["C"])
    predIdx = (py2round(finalDict["A"], 2), py2round(finalDict["T"], 2), py2round(finalDict["G"], 2),
               py2round(finalDict["C"], 2))
    outvec = [chrom, pos, ref, alt, vtype, cvg, allFrag, allMT, usedFrag, usedMT,
              py2round(finalDict[origAlt], 2), alleleCnt[origAlt], frac_alt, MTCnt[origAlt], MT_f_alt,
              strongMTCnt[origAlt]]
    outvec.extend(fracs)
    outvec.extend(MTs)
    outvec.extend(strongMT)
    outvec.extend(predIdx)
    outvec.append(fltr)

    if detail is not None:
        alleles = sorted(set(alleleCnt) | set(piTerms), key=allele_sort_key)
        detail.update(dict(
            usedMT=usedMT, usedFrag=usedFrag, MT3=MT3Cnt, MT5=MT5Cnt, MT7=MT7Cnt, MT10=MT10Cnt,
            alleles=alleles, keys=[k for k, _ in sortedList], bcKeys=list(bcKeys),
            alleleCnt=dict(alleleCnt), forwardCnt=dict(forwardCnt), reverseCnt=dict(reverseCnt),
            lowQReads=dict(lowQReads), concord=dict(concordPairCnt), discord=dict(discordPairCnt),
            MTCnt=dict(MTCnt), strongMTCnt=dict(strongMTCnt),
            PI={k: finalDict[k] for k in piTerms},
            r1Le={k: sum(d <= 20 for d in v) for k, v in r1BcEndPos.items()},
            r1Tot={k: len(v) for k, v in r1BcEndPos.items()},
            r2Le={k: sum(d <= 20 for d in v) for k, v in r2BcEndPos.items()},
            r2Tot={k: len(v) for k, v in r2BcEndPos.items()},
            r2PLe={k: sum(d <= primerDist for d in v) for k, v in r2PrimerEndPos.items()},
            maxBase=maxBase, secondMaxBase=secondMaxBase, firstAlt=firstAlt, origAlt=origAlt, altPI=altPI,
            secondMaxPI=secondMaxPI, vtype=vtype, biallelic=biallelic, fltr=fltr, filt1=dbg1, filt2=dbg2))
    return "\t".join(py2str(x) for x in outvec)


def soft_clip_total(cigar):
    """Soft-clipped bases at both ends (query_length - query_alignment_length)."""
    n = 0
    i = 0
    while i < len(cigar) and cigar[i][0] in (4, 5):
        if cigar[i][0] == 4:
            n += cigar[i][1]
        i += 1
    j = len(cigar) - 1
    while j >= i and cigar[j][0] in (4, 5):
        if cigar[j][0] == 4:
            n += cigar[j][1]
        j -= 1
    return n


def vc_wrapper(*args, **kw):                                             # smCounter.py:605-611
    try:
        output = vc(*args, **kw)
    except Exception:
        output = "Exception thrown!\n" + traceback.format_exc()
    return output


# ----------------------------------------------------------------------------------------------
# main(): locus list, repeat filters, writers (smCounter.py:675-680, 699-901)
# ----------------------------------------------------------------------------------------------
headerAll = ('CHROM', 'POS', 'REF', 'ALT', 'TYPE', 'DP', 'FR', 'MT', 'UFR', 'UMT', 'PI', 'VDP', 'VAF', 'VMT',
             'VMF', 'VSM', 'DP_A', 'DP_T', 'DP_G', 'DP_C', 'AF_A', 'AF_T', 'AF_G', 'AF_C', 'MT_3RPM', 'MT_5RPM',
             'MT_7RPM', 'MT_10RPM', 'UMT_A', 'UMT_T', 'UMT_G', 'UMT_C', 'UMF_A', 'UMF_T', 'UMF_G', 'UMF_C',
             'VSM_A', 'VSM_T', 'VSM_G', 'VSM_C', 'PI_A', 'PI_T', 'PI_G', 'PI_C', 'FILTER')
headerVariants = ('CHROM', 'POS', 'REF', 'ALT', 'TYPE', 'DP', 'MT', 'UMT', 'PI', 'THR', 'VMT', 'VMF', 'VSM',
                  'FILTER')


def loci_from_bed(bed_lines):                                            # :675-680
    locList = []
    for line in bed_lines:
        if not line.startswith("track "):
            (chrom, regionStart, regionEnd) = line.strip().split("\t")[0:3]
            for pos in range(int(regionStart), int(regionEnd)):
                locList.append((chrom, str(pos + 1)))
    return locList


# -- bedtools 2.25 restatement [third party, absent]: merge / sort / intersect on in-memory rows -------------
def bed_merge(rows, distinct_col4=False):
    """``bedtools merge [-c 4 -o distinct]`` on rows in file order (the reference does not pre-sort)."""
    out = []
    cur = None
    for r in rows:
        chrom, s, e = r[0], int(r[1]), int(r[2])
        name = r[3] if distinct_col4 else None
        if cur is not None and cur[0] == chrom and s <= cur[2]:
            if e > cur[2]:
                cur[2] = e
            if distinct_col4:
                cur[3].add(name)
        else:
            if cur is not None:
                out.append(cur)
            cur = [chrom, s, e, {name} if distinct_col4 else None]
    if cur is not None:
        out.append(cur)
    if distinct_col4:
        return [(c, s, e, ",".join(sorted(n))) for (c, s, e, n) in out]
    return [(c, s, e) for (c, s, e, _) in out]


def bed_sort(rows):
    return sorted(rows, key=lambda r: (r[0], int(r[1])))


def bed_intersect(a_rows, b_rows):
    """``bedtools intersect -a A -b B``: each A row clipped to every overlapping B row, A's extra columns kept."""
    by_chrom = defaultdict(list)
    for b in b_rows:
        by_chrom[b[0]].append((int(b[1]), int(b[2])))
    out = []
    for a in a_rows:
        s, e = int(a[1]), int(a[2])
        for (bs, be) in by_chrom.get(a[0], ()):
            lo, hi = max(s, bs), min(e, be)
            if lo < hi:
                out.append((a[0], lo, hi) + tuple(a[3:]))
    return out


def repeat_filter_rows(output, target_rows, trf_rows, rm_rows):             # :699-785
    bedRepeatMasker = bed_sort(bed_merge(rm_rows, distinct_col4=True))
    bedTarget = bed_sort(bed_merge(target_rows))
    rep1 = bed_sort(bed_intersect(trf_rows, bedTarget))
    rep2 = bed_sort(bed_intersect(bedRepeatMasker, bedTarget))
    trfRegions = defaultdict(list)
    for r in rep1:
        trfRegions[r[0]].append((int(r[1]), int(r[2]), "RepT;"))
    rmRegions = defaultdict(list)
    for (chrom, s, e, typeCodes) in rep2:
        repTypes = []
        for typeCode in typeCodes.split(","):
            if typeCode == "Simple_repeat":
                repTypes.append("RepS")
            elif typeCode == "Low_complexity":
                repTypes.append("LowC")
            elif typeCode == "Satellite":
                repTypes.append("SL")
            else:
                repTypes.append("Other_Repeat")
        rmRegions[chrom].append((int(s), int(e), ";".join(repTypes) + ";"))
    idx = {h: i for i, h in enumerate(headerAll)}
    out = list(output)
    for i in range(len(out)):
        lineList = out[i].split("\t")
        chromTr = lineList[idx["CHROM"]]
        altTr = lineList[idx["ALT"]]
        try:
            posTr = int(lineList[idx["POS"]])
        except ValueError:
            continue
        try:
            altMtFracTr = float(lineList[idx["VMF"]])
        except ValueError:
            continue
        try:
            pred = int(float(lineList[idx["PI"]]))
        except ValueError:
            pred = 0
        if pred >= 5 and altTr != "DEL":
            if altMtFracTr < 40:
                for (locL, locR, repType) in trfRegions[chromTr]:
                    if locL < posTr <= locR:
                        lineList[-1] += repType
                        break
            for (locL, locR, repType) in rmRegions[chromTr]:
                if locL < posTr <= locR:
                    lineList[-1] += repType
                    break
        lineList[-1] = "PASS" if lineList[-1] == ";" else lineList[-1].strip(";")
        out[i] = "\t".join(lineList)
    return out


def vcf_header(outPrefix):                                               # :788-817
    return (
        '##fileformat=VCFv4.2\n'
        '##reference=GRCh37\n'
        '##INFO=<ID=TYPE,Number=1,Type=String,Description="Variant type: SNP or INDEL">\n'
        '##INFO=<ID=DP,Number=1,Type=Integer,Description="Total read depth">\n'
        '##INFO=<ID=MT,Number=1,Type=Integer,Description="Total MT depth">\n'
        '##INFO=<ID=UMT,Number=1,Type=Integer,Description="Filtered MT depth">\n'
        '##INFO=<ID=PI,Number=1,Type=Float,Description="Variant prediction index">\n'
        '##INFO=<ID=THR,Number=1,Type=Integer,Description="Variant prediction index minimum threshold">\n'
        '##INFO=<ID=VMT,Number=1,Type=Integer,Description="Variant MT depth">\n'
        '##INFO=<ID=VMF,Number=1,Type=Float,Description="Variant MT fraction">\n'
        '##INFO=<ID=VSM,Number=1,Type=Integer,Description="Variant strong MT depth">\n'
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
        '##FORMAT=<ID=AD,Number=.,Type=Integer,Description="Filtered allelic MT depths for the ref and alt alleles">\n'
        '##FORMAT=<ID=VF,Number=1,Type=Float,Description="Variant MT fraction, same as VMF">\n'
        '##FILTER=<ID=RepT,Description="Variant in simple tandem repeat region, as defined by Tandem Repeats Finder">\n'
        '##FILTER=<ID=RepS,Description="Variant in simple repeat region, as defined by RepeatMasker">\n'
        '##FILTER=<ID=LowC,Description="Variant in low complexity region, as defined by RepeatMasker">\n'
        '##FILTER=<ID=SL,Description="Variant in micro-satelite region, as defined by RepeatMasker">\n'
        '##FILTER=<ID=HP,Description="Inside or flanked by homopolymer region">\n'
        '##FILTER=<ID=LM,Description="Low coverage (fewer than 5 MTs)">\n'
        '##FILTER=<ID=LSM,Description="Fewer than 2 strong MTs">\n'
        '##FILTER=<ID=SB,Description="Strand bias">\n'
        '##FILTER=<ID=LowQ,Description="Low base quality (mean < 22)">\n'
        '##FILTER=<ID=MM,Description="Too many genome reference mismatches in reads (default threshold is 6.5 per 100 bases)">\n'
        '##FILTER=<ID=DP,Description="Too many discordant read pairs">\n'
        '##FILTER=<ID=R1CP,Description="Variants are clustered at the end of R1 reads">\n'
        '##FILTER=<ID=R2CP,Description="Variants are clustered at the end of R2 reads">\n'
        '##FILTER=<ID=PrimerCP,Description="Variants are clustered immediately after the primer, possible enzyme initiation error">\n'
        + '\t'.join(('#CHROM', 'POS', 'ID', 'REF', 'ALT', 'QUAL', 'FILTER', 'INFO', 'FORMAT', outPrefix)) + '\n')


def write_outputs(output, outPrefix, mtDepth, threshold_arg=0):          # :819-901
    """Returns (threshold, all_txt, cut_txt, cut_vcf) as strings; ``output`` = rows after repeat_filter_rows."""
    idx = {h: i for i, h in enumerate(headerAll)}
    threshold = int(math.ceil(14.0 + 0.012 * mtDepth)) if threshold_arg == 0 else threshold_arg
    outAll = ["\t".join(headerAll) + "\n"]
    outVariants = ["\t".join(headerVariants) + "\n"]
    outVcf = [vcf_header(outPrefix)]
    for line in output:
        outAll.append(line + "\n")
        fields = line.split("\t")
        PI = fields[idx["PI"]]
        if len(PI) == 0:
            continue
        ALT = fields[idx["ALT"]]
        QUAL = str(int(float(PI)))
        if int(QUAL) >= threshold and ALT != "DEL":
            CHROM, POS, REF, TYPE = fields[idx["CHROM"]], fields[idx["POS"]], fields[idx["REF"]], fields[idx["TYPE"]]
            DP, MT, UMT = fields[idx["DP"]], fields[idx["MT"]], fields[idx["UMT"]]
            VMT, VMF, VSM, FILTER = fields[idx["VMT"]], fields[idx["VMF"]], fields[idx["VSM"]], fields[idx["FILTER"]]
            THR = str(threshold)
            INFO = ";".join(("TYPE=" + TYPE, "DP=" + DP, "MT=" + MT, "UMT=" + UMT, "PI=" + PI, "THR=" + THR,
                             "VMT=" + VMT, "VMF=" + VMF, "VSM=" + VSM))
            alts = ALT.split(",")
            if len(alts) == 2:
                genotype = "1/2"
            elif len(alts) != 1:
                raise Exception("error hacking genotype field for " + str(alts))
            elif CHROM == "chrY" or CHROM == "chrM":
                genotype = "1"
            elif float(VMF) > 0.95:
                genotype = "1/1"
            else:
                genotype = "0/1"
            REFMT = str(int(UMT) - int(VMT))
            AD = REFMT + "," + VMT
            if len(alts) == 2:
                AD = AD + ",1"
            FORMAT = "GT:AD:VF"
            SAMPLE = ":".join((genotype, AD, VMF))
            outVcf.append("\t".join((CHROM, POS, ".", REF, ALT, QUAL, FILTER, INFO, FORMAT, SAMPLE)) + "\n")
            outVariants.append("\t".join((CHROM, POS, REF, ALT, TYPE, DP, MT, UMT, PI, THR, VMT, VMF, VSM,
                                          FILTER)) + "\n")
    return threshold, "".join(outAll), "".join(outVariants), "".join(outVcf)


def run(reads, bed_lines, refs, *, mtDepth, rpb, minBQ=20, minMQ=30, hpLen=10, mismatchThr=6.0, mtDrop=0,
        maxMT=0, primerDist=2, threshold=0, outPrefix="out", trf_rows=(), rm_rows=(), keep_umis=None,
        details=None):
    """The whole of main() on in-memory inputs; returns (threshold, all_txt, cut_txt, cut_vcf)."""
    index = reads if isinstance(reads, ReadIndex) else ReadIndex(reads)
    locList = loci_from_bed(bed_lines)
    output = []
    for (chrom, pos) in locList:
        d = {} if details is not None else None
        k = None if keep_umis is None else keep_umis.get((chrom, pos))
        line = vc_wrapper(index, chrom, pos, minBQ, minMQ, mtDepth, rpb, hpLen, mismatchThr, mtDrop, maxMT,
                          primerDist, refs, keep_umis=k, detail=d)
        if line.startswith("Exception thrown!"):
            raise Exception("Exception thrown in vc() at location: " + str((chrom, pos)) + "\n" + line)
        output.append(line)
        if details is not None:
            details.append(d)
    target_rows = [l.strip().split("\t")[0:3] for l in bed_lines if not l.startswith("track ")]
    output = repeat_filter_rows(output, target_rows, list(trf_rows), list(rm_rows))
    return write_outputs(output, outPrefix, mtDepth, threshold)
