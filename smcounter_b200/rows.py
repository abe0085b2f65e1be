"""Host-side string work of the hot path: turns the device results into the reference's 45-column rows.

Covers what stays on the host by design:
  * allele strings ('INS|s|s+ins', 'DEL|s+del|s', smCounter.py:374,396) -- built from a representative read / the FASTA;
  * convertToVcf()            smCounter.py:103-117;
  * isHPorLowComp()           smCounter.py:122-177: the reference windows are cut here (hp_window), the test itself runs on
                              the device (smc_hp_lowcomp) for all candidates of a batch at once (device_hp_flags);
  * FILTER string assembly    smCounter.py:184-269 (device supplies the bits, order of tags as in the reference);
  * bi-allelic resolution     smCounter.py:553-573;
  * the output vector         smCounter.py:575-600 with Python-2 round()/str() semantics.
"""
from __future__ import annotations

from decimal import Decimal, ROUND_HALF_UP

import numpy as np

from . import _ffi
from ._ffi import (A_A, A_C, A_G, A_T, C_ALLELE, C_MT, C_STRONG, F_DP, F_EVALUATED, F_HPGATE, F_LM, F_LOWQ, F_LSM, F_PRIMERCP,
                   F_R1CP, F_R2CP, F_SB, FIXED_NAMES, K_BASE, K_DEL, K_INS, L_ALLFRAG, L_ALLMT, L_CVG, L_MT3, L_MT5, L_MT7,
                   L_MT10, L_STATUS, L_USEDFRAG, L_USEDMT, SMC_NFIXED, ST_BAD_MASK, ST_NEED_DOWNSAMPLE, ST_UMI_OVERFLOW,
                   ST_ZERO_COVERAGE)
from .soa import NT16

headerAll = ('CHROM', 'POS', 'REF', 'ALT', 'TYPE', 'DP', 'FR', 'MT', 'UFR', 'UMT', 'PI', 'VDP', 'VAF', 'VMT', 'VMF', 'VSM',
             'DP_A', 'DP_T', 'DP_G', 'DP_C', 'AF_A', 'AF_T', 'AF_G', 'AF_C', 'MT_3RPM', 'MT_5RPM', 'MT_7RPM', 'MT_10RPM',
             'UMT_A', 'UMT_T', 'UMT_G', 'UMT_C', 'UMF_A', 'UMF_T', 'UMF_G', 'UMF_C', 'VSM_A', 'VSM_T', 'VSM_G', 'VSM_C',
             'PI_A', 'PI_T', 'PI_G', 'PI_C', 'FILTER')                          # smCounter.py:743
headerVariants = ('CHROM', 'POS', 'REF', 'ALT', 'TYPE', 'DP', 'MT', 'UMT', 'PI', 'THR', 'VMT', 'VMF', 'VSM', 'FILTER')   # :744


def py2round(x: float, nd: int) -> float:
    """Python-2.7 round(): correctly rounded, exact decimal ties away from zero (Py3 rounds them to even)."""
    r = round(x, nd)
    s = x * (2 * 10 ** nd)
    if s == int(s) and int(s) & 1:       # x * 10^nd is an odd multiple of 1/2: possibly an exact halfway case -> decide on the
        return float(Decimal(x).quantize(Decimal(1).scaleb(-nd), rounding=ROUND_HALF_UP))   # exact binary value
    return r


def _f4(num: int, den: int) -> str:
    """py2str(py2round(1.0 * num / den, 4)): the fraction columns (VAF, VMF, AF_*, UMF_*) of smCounter.py:575-600.  repr() of a
    value rounded to <= 4 decimals is its Python-2 str() ('%.12g' with '.0' for integral values)."""
    if num == 0:
        return "0.0"
    return repr(py2round(1.0 * num / den, 4))


def _f2(x: float) -> str:
    """py2str(py2round(x, 2)) for the prediction-index columns (below 1e10, so repr() == '%.12g' form)."""
    if x == 0.0:
        return "0.0"
    r = py2round(x, 2)
    return repr(r) if abs(r) < 1e10 else py2str(r)


def py2str(v) -> str:
    """Python-2.7 str() of the value types that reach the output vector (smCounter.py:599)."""
    if isinstance(v, float):
        s = "%.12g" % v
        if "." not in s and "e" not in s and "n" not in s:
            s += ".0"
        return s
    if isinstance(v, (int, np.integer)):
        return "%d" % int(v)
    return str(v)


def convert_to_vcf(origRef: str, origAlt: str):
    """smCounter.py:103-117."""
    vtype, ref, alt = ".", origRef, origAlt
    if len(origAlt) == 1:
        vtype = "SNP"
    elif origAlt == "DEL":
        vtype = "SDEL"
    else:
        vals = origAlt.split("|")
        if vals[0] in ("DEL", "INS"):
            vtype = "INDEL"
            ref, alt = vals[1], vals[2]
    return ref, alt, vtype


def hp_window(chrom, pos0, hpLen, ref, alt, refs):
    """(window, position inside it) for smc_hp_lowcomp: upper-case reference[max(0, pos0 - 2*hpLen), min(contig length,
    pos0 + max(len(ref), len(alt)) + 2*hpLen)) -- every base isHPorLowComp() fetches (smCounter.py:127-129, 143-145)."""
    w0 = max(0, pos0 - 2 * hpLen)
    w1 = min(refs.get_reference_length(chrom), pos0 + max(len(ref), len(alt)) + 2 * hpLen)
    return refs.fetch(chrom, w0, w1).upper(), pos0 - w0


def device_hp_flags(caller, res, reads, loci, chroms, refs, hpLen):
    """{(locus index, candidate 0/1): (homopolymer, low complexity)} for every candidate whose FILTER evaluation can append
    HP / LowC (filterVariants() was entered and MTCnt[alt]/usedMT < 0.99, smCounter.py:195-203), computed by ONE
    smc_hp_lowcomp call on the device of ``caller`` (any live GpuCaller)."""
    namer = AlleleNamer(res, reads, loci, chroms, refs)
    keys, cands = [], []
    want = F_EVALUATED | F_HPGATE
    idx1 = np.flatnonzero((res.fl1[:loci.n] & want) == want)
    idx2 = np.flatnonzero((res.biallelic[:loci.n] != 0) & ((res.fl2[:loci.n] & want) == want))
    for cand, idx, alleles in ((0, idx1, res.alt_allele), (1, idx2, res.second_allele)):
        for i in idx:
            i = int(i)
            chrom = chroms[int(loci.ref_id[i])]
            ref, alt, _ = convert_to_vcf(chr(int(loci.ref_base[i])), namer.name(int(alleles[i])))
            win, wpos = hp_window(chrom, int(loci.pos0[i]), hpLen, ref, alt, refs)
            keys.append((i, cand))
            cands.append((win, wpos, ref, alt))
    flags = caller.hp_lowcomp(hpLen, cands)
    return {k: (bool(f & _ffi.HP_HOMOPOLYMER), bool(f & _ffi.HP_LOWCOMP)) for k, f in zip(keys, flags.tolist())}


class AlleleNamer:
    """Allele reference (0..4 fixed, 5+j dynamic row j) -> the reference's allele string."""

    def __init__(self, res, reads, loci, chroms, refs):
        self.res, self.reads, self.loci, self.chroms, self.refs = res, reads, loci, chroms, refs
        self._cache = {}

    def _read_bases(self, r, q0, n):
        rd = self.reads
        if getattr(rd, "store_lo", None) is not None:              # stored window: seq holds the bases from store_lo on
            q0 -= int(rd.store_lo[r])
        if getattr(rd, "seq_bits", 4) == 2:                         # compact bases: 2-bit codes + the list of non-ACGT bases
            return "".join(NT16[v] for v in rd.compact_bases(r, q0, n))
        so = int(rd.seq_off[r])
        out = []
        for q in range(q0, q0 + n):
            b = int(rd.seq[so + (q >> 1)])
            out.append(NT16[(b & 15) if (q & 1) else (b >> 4)])
        return "".join(out)

    def name(self, a: int) -> str:
        if a < SMC_NFIXED:
            return FIXED_NAMES[a]
        j = a - SMC_NFIXED
        if j in self._cache:
            return self._cache[j]
        res = self.res
        kind, site, ln = int(res.dyn_kind[j]), NT16[int(res.dyn_site[j])], int(res.dyn_len[j])
        if kind == K_BASE:
            s = site
        elif kind == K_INS:                                                     # smCounter.py:372-374
            ins = self._read_bases(int(res.dyn_rep_read[j]), int(res.dyn_rep_qpos[j]) + 1, ln)
            s = "INS|" + site + "|" + site + ins
        else:                                                                   # smCounter.py:393-396
            i = int(res.dyn_locus[j])
            chrom = self.chroms[int(self.loci.ref_id[i])]
            p1 = int(self.loci.pos0[i]) + 1
            deleted = self.refs.fetch(chrom, p1, p1 + ln).upper()
            s = "DEL|" + site + deleted + "|" + site
        self._cache[j] = s
        return s


_FILTER_TAGS = ((F_LM, "LM;"), (F_LSM, "LSM;"))
_FILTER_TAGS2 = ((F_DP, "DP;"), (F_SB, "SB;"), (F_LOWQ, "LowQ;"), (F_R1CP, "R1CP;"), (F_R2CP, "R2CP;"), (F_PRIMERCP, "PrimerCP;"))


def _filter_string(bits, hp_lc):
    """FILTER accumulator of filterVariants() (';' = nothing fired), tags in the reference's order.  ``hp_lc`` = the
    device's isHPorLowComp() result for this candidate (present whenever F_HPGATE is set)."""
    if not (bits & F_EVALUATED):
        return ";"
    f = ";"
    for b, t in _FILTER_TAGS:
        if bits & b:
            f += t
    if bits & F_HPGATE:                                                        # smCounter.py:195-203
        if hp_lc is None:
            raise RuntimeError("format_rows: no device HP/LowC flags for a candidate that needs them (pass hp_flags=device_hp_flags(...))")
        hp, lc = hp_lc
        if hp:
            f += "HP;"
        if lc:
            f += "LowC;"
    for b, t in _FILTER_TAGS2:
        if bits & b:
            f += t
    return f


_F4_TABLE = None


def _f4_table():
    """str of every value k / 10000 for k = 0 .. 10000: the fraction columns after rounding to 4 decimals."""
    global _F4_TABLE
    if _F4_TABLE is None:
        _F4_TABLE = [repr(k / 10000.0) for k in range(10001)]
    return _F4_TABLE


def _f4_column(num, den):
    """[_f4(n, d)] for integer arrays, vectorised: k = round(n / d * 1e4) is exact whenever the scaled value is not within
    1e-6 of a half (the double rounding of the scaling is ~1e-12); the few near-ties go through py2round()."""
    num = np.asarray(num, dtype=np.float64)
    den = np.asarray(den, dtype=np.float64)
    ok = den > 0
    y = np.where(ok, num / np.where(ok, den, 1.0), 0.0) * 1e4
    k = np.rint(y)
    near = ok & (np.abs(y - np.floor(y) - 0.5) < 1e-6)
    tab = _f4_table()
    out = [tab[j] for j in np.clip(k, 0, 10000).astype(np.int64).tolist()]
    for i in np.flatnonzero(near | (k > 10000) | (k < 0)).tolist():
        out[i] = _f4(int(num[i]), int(den[i])) if den[i] > 0 else "0.0"
    return out


def _f2_column(x):
    """[_f2(v)] for a float array (prediction indices), vectorised the same way."""
    x = np.asarray(x, dtype=np.float64)
    y = x * 100.0
    k = np.rint(y)
    near = (np.abs(y - np.floor(y) - 0.5) < 1e-6) | ~(np.abs(x) < 1e9)
    out = [repr(v) for v in (k / 100.0).tolist()]
    for i in np.flatnonzero(near).tolist():
        out[i] = _f2(float(x[i]))
    return out


def _float_columns(res, n):
    """The twelve per-base float columns (AF_*, UMF_*, PI_* in A, T, G, C order) as ready strings for every locus of the batch."""
    atgc = (A_A, A_T, A_G, A_C)
    return ([_f4_column(res.cnt[a, C_ALLELE, :n], res.loc[L_CVG, :n]) for a in atgc],
            [_f4_column(res.cnt[a, C_MT, :n], res.loc[L_USEDMT, :n]) for a in atgc],
            [_f2_column(res.pi[a, :n]) for a in atgc])


_POOL_JOB = None          # inputs of the forked formatting workers (inherited, never pickled)
PARALLEL_MIN_ROWS = 60000        # below this the fork + pickle overhead exceeds the ~10 us per row of the inline loop


def _format_slice(bounds):
    res, reads, loci, chroms, refs, hpLen, order, hp_flags, cols = _POOL_JOB
    try:
        return format_rows(res, reads, loci, chroms, refs, hpLen, order[bounds[0]:bounds[1]], workers=1, hp_flags=hp_flags, _cols=cols)
    except RuntimeError as e:                  # plain message: survives pickling back to the parent
        return e


def format_rows(res, reads, loci, chroms, refs, hpLen, locus_order=None, workers=None, hp_flags=None, _cols=None):
    """The 45-field rows of vc() (smCounter.py:575-600) for ``locus_order`` (indices into loci; default all, in order).

    Row formatting is per-locus string work, as independent as the reference's per-locus workers (smCounter.py:683-685):
    from PARALLEL_MIN_ROWS rows up it is fanned out over forked worker processes (``workers``: default = host cores, 1 =
    inline); the device results are inherited through fork(), only the finished strings travel back.

    ``hp_flags``: device_hp_flags(...) of the same results (the HP / LowC bits of smCounter.py:195-203 are computed on the
    device, like every other filter bit); required as soon as a candidate reaches that test.

    Raises RuntimeError for loci the device flagged as needing a down-sampling mask or as unsupported.
    """
    global _POOL_JOB
    if hp_flags is None:
        hp_flags = {}
    import os
    n_rows = loci.n if locus_order is None else len(locus_order)
    if workers is None:
        workers = os.cpu_count() or 1
    import threading
    if workers > 1 and n_rows >= PARALLEL_MIN_ROWS and threading.current_thread() is threading.main_thread() and hasattr(os, "fork"):
        import multiprocessing as mp
        order = np.arange(loci.n) if locus_order is None else np.asarray(locus_order)
        nchunk = max(2, min(n_rows // 8192, workers * 2))
        cuts = [(n_rows * k) // nchunk for k in range(nchunk + 1)]
        _POOL_JOB = (res, reads, loci, chroms, refs, hpLen, order, hp_flags, _float_columns(res, loci.n))   # columns once, inherited
        try:
            with mp.get_context("fork").Pool(min(workers, nchunk)) as pool:
                parts = pool.map(_format_slice, list(zip(cuts[:-1], cuts[1:])), chunksize=1)
        finally:
            _POOL_JOB = None
        rows = []
        for p in parts:
            if isinstance(p, Exception):
                raise p
            rows.extend(p)
        return rows
    namer = AlleleNamer(res, reads, loci, chroms, refs)
    n = loci.n
    order = range(n) if locus_order is None else np.asarray(locus_order).tolist()
    # plain Python lists: element access on numpy arrays costs more than the formatting itself
    loc, cnt, pi = res.loc[:, :n].tolist(), res.cnt[:, :, :n].tolist(), res.pi[:, :n].tolist()
    ref_id, pos0, ref_base = loci.ref_id.tolist(), loci.pos0.tolist(), loci.ref_base.tolist()
    alt_allele, second_allele = res.alt_allele[:n].tolist(), res.second_allele[:n].tolist()
    fl1, fl2, biallelic = res.fl1[:n].tolist(), res.fl2[:n].tolist(), res.biallelic[:n].tolist()
    l_cvg, l_allfrag, l_allmt, l_usedfrag, l_usedmt, l_status = (loc[k] for k in (L_CVG, L_ALLFRAG, L_ALLMT, L_USEDFRAG, L_USEDMT, L_STATUS))
    l_mt3, l_mt5, l_mt7, l_mt10 = (loc[k] for k in (L_MT3, L_MT5, L_MT7, L_MT10))
    ATGC = (A_A, A_T, A_G, A_C)
    c_allele = [cnt[a][C_ALLELE] for a in ATGC]
    c_mt = [cnt[a][C_MT] for a in ATGC]
    c_strong = [cnt[a][C_STRONG] for a in ATGC]
    af_s, umf_s, pi_s = _cols if _cols is not None else _float_columns(res, n)
    bad_status = ST_NEED_DOWNSAMPLE | ST_UMI_OVERFLOW | ST_BAD_MASK
    zero_tail = "\t" * 42 + "Zero_Coverage"
    rows = []
    for i in order:
        chrom = chroms[ref_id[i]]
        pos = "%d" % (pos0[i] + 1)
        origRef = chr(ref_base[i])
        status = l_status[i]
        if status & bad_status:
            raise RuntimeError("Exception thrown in vc() at location: %s (device status 0x%x)" % ((chrom, pos), status))
        if status & ST_ZERO_COVERAGE:                                           # smCounter.py:492-494: 3 fields + 41 blanks + tag
            rows.append(chrom + "\t" + pos + "\t" + origRef + zero_tail)
            continue
        cvg, usedMT = l_cvg[i], l_usedmt[i]
        a1 = alt_allele[i]
        origAlt = FIXED_NAMES[a1] if a1 < SMC_NFIXED else namer.name(a1)
        ref, alt, vtype = convert_to_vcf(origRef, origAlt)
        fltr = _filter_string(fl1[i], hp_flags.get((i, 0))) if fl1[i] & F_EVALUATED else ";"
        alt_ref = a1
        if biallelic[i]:                                                        # smCounter.py:555-573
            a2 = second_allele[i]
            origAlt2 = namer.name(a2)
            ref2, alt2, vtype2 = convert_to_vcf(origRef, origAlt2)
            fltr2 = _filter_string(fl2[i], hp_flags.get((i, 1)))
            if fltr == ";" and fltr2 == ";":
                alt = alt + "," + alt2
                vtype = vtype.lower() + "," + vtype2.lower()
            elif fltr != ";" and fltr2 == ";":
                alt = alt2
                fltr = fltr2
                alt_ref = a2
        if alt_ref < SMC_NFIXED:          # counters / PI of the reported allele (fixed slot or dynamic-allele row)
            ca = cnt[alt_ref]
            v_dp, v_mt, v_sm, v_pi = ca[C_ALLELE][i], ca[C_MT][i], ca[C_STRONG][i], pi[alt_ref][i]
        else:
            j = alt_ref - SMC_NFIXED
            v_dp, v_mt, v_sm = (int(res.dyn_cnt[j, c]) for c in (C_ALLELE, C_MT, C_STRONG))
            v_pi = float(res.dyn_pi[j])
        a0, a1c, a2c, a3c = c_allele[0][i], c_allele[1][i], c_allele[2][i], c_allele[3][i]
        m0, m1, m2, m3 = c_mt[0][i], c_mt[1][i], c_mt[2][i], c_mt[3][i]
        rows.append("%s\t%s\t%s\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%s\t%d\t%s\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%s\t%s\t%d\t%d\t%d\t%d\t"
                    "%d\t%d\t%d\t%d\t%s\t%s\t%s\t%s\t%d\t%d\t%d\t%d\t%s\t%s\t%s\t%s\t%s" % (
                        chrom, pos, ref, alt, vtype, cvg, l_allfrag[i], l_allmt[i], l_usedfrag[i], usedMT,
                        _f2(v_pi), v_dp, _f4(v_dp, cvg), v_mt, _f4(v_mt, usedMT), v_sm,
                        a0, a1c, a2c, a3c, af_s[0][i], af_s[1][i], af_s[2][i], af_s[3][i],
                        l_mt3[i], l_mt5[i], l_mt7[i], l_mt10[i],
                        m0, m1, m2, m3, umf_s[0][i], umf_s[1][i], umf_s[2][i], umf_s[3][i],
                        c_strong[0][i], c_strong[1][i], c_strong[2][i], c_strong[3][i],
                        pi_s[0][i], pi_s[1][i], pi_s[2][i], pi_s[3][i], fltr))
    return rows


# ----------------------------------------------------------------------------------------------------------------------
# native output stage (libsmc_bamio.so: csrc/smc_rows.cpp, include/smc_rows.h)
# ----------------------------------------------------------------------------------------------------------------------
class EmittedRows:
    """Text produced by smc_rows_emit for one batch: ``all`` (45-column rows) and, when finalised, ``cut`` / ``vcf`` (the rows of the
    called variants), each as bytes plus per-row offsets so that rows can be regrouped by BED interval without re-parsing."""

    def __init__(self, all_b, all_off, cut_b, cut_off, vcf_b, vcf_off):
        self.all, self.all_off, self.cut, self.cut_off, self.vcf, self.vcf_off = all_b, all_off, cut_b, cut_off, vcf_b, vcf_off

    def rows(self):
        """The rows as a list of str (without the trailing newline) -- the form format_rows() returns."""
        s = self.all.decode()
        return s.split("\n")[:-1] if s else []

    def slices(self, r0, r1):
        """(all, cut, vcf) bytes of rows [r0, r1)."""
        return (self.all[self.all_off[r0]:self.all_off[r1]], self.cut[self.cut_off[r0]:self.cut_off[r1]], self.vcf[self.vcf_off[r0]:self.vcf_off[r1]])


def _needed_dyn_names(res, n, namer):
    """Names of the dynamic-allele rows that are an ALT candidate of some locus (the only ones a row can mention)."""
    need = set()
    a1 = res.alt_allele[:n]
    need.update(int(a) - SMC_NFIXED for a in a1[a1 >= SMC_NFIXED].tolist())
    bi = res.biallelic[:n] != 0
    a2 = res.second_allele[:n][bi]
    need.update(int(a) - SMC_NFIXED for a in a2[a2 >= SMC_NFIXED].tolist())
    nd = int(res.n_dyn)
    off = np.zeros(nd + 1, dtype=np.int64)
    parts = []
    for j in sorted(need):
        s = namer.name(SMC_NFIXED + j).encode()
        parts.append((j, s))
    lens = np.zeros(nd, dtype=np.int64)
    for j, s in parts:
        lens[j] = len(s)
    off[1:] = np.cumsum(lens)
    return b"".join(s for _, s in parts), off


def emit_rows(res, reads, loci, chroms, refs, hpLen, locus_order=None, hp_flags=None, finalize=False, threshold=0, trf=None, rm=None,
              threads=0) -> EmittedRows:
    """The rows of a batch through the native output stage: what format_rows() produces (finalize=False), or -- finalize=True --
    what format_rows() + repeats.apply_repeat_filters() + writers.called_lines() produce, in one threaded pass over the device
    results.  ``trf`` / ``rm``: the dicts of repeats.build_repeat_regions()."""
    import ctypes as C
    from . import _bamio
    lib = _bamio.load()
    n = loci.n
    namer = AlleleNamer(res, reads, loci, chroms, refs)
    names, name_off = _needed_dyn_names(res, n, namer)
    hp1 = np.zeros(max(n, 1), np.uint8); hp2 = np.zeros(max(n, 1), np.uint8)
    for (i, cand), (hp, lc) in (hp_flags or {}).items():
        (hp1 if cand == 0 else hp2)[i] = 128 | (1 if hp else 0) | (2 if lc else 0)
    order = None if locus_order is None else np.ascontiguousarray(locus_order, dtype=np.int64)
    cidx = {c: k for k, c in enumerate(chroms)}

    def flat(regions):
        ch, lo, hi, tags = [], [], [], []
        for c in chroms:
            for (a, b, t) in (regions or {}).get(c, ()):
                ch.append(cidx[c]); lo.append(a); hi.append(b); tags.append(t.encode())
        toff = np.zeros(len(tags) + 1, dtype=np.int64)
        if tags:
            toff[1:] = np.cumsum([len(t) for t in tags])
        return np.asarray(ch, np.int32), np.asarray(lo, np.int64), np.asarray(hi, np.int64), b"".join(tags), toff

    t_ch, t_lo, t_hi, _, _ = flat(trf)
    r_ch, r_lo, r_hi, r_tags, r_toff = flat(rm)
    a = smc = _bamio.smc_rows_in()
    p = lambda x: None if x is None else x.ctypes.data
    keep = [np.ascontiguousarray(x) for x in (loci.ref_id, loci.pos0, loci.ref_base, res.loc, res.cnt, res.pi, res.alt_allele, res.second_allele,
                                              res.fl1, res.fl2, res.biallelic, res.dyn_cnt, res.dyn_pi)]
    # the 2-D result arrays are allocated for max(n, 1) loci: the stride is their second dimension
    stride = res.loc.shape[1]
    a.n_loci, a.n_rows, a.order = stride, (n if order is None else len(order)), p(order)
    a.ref_id, a.pos0, a.ref_base = p(keep[0]), p(keep[1]), p(keep[2])
    carr = (C.c_char_p * max(len(chroms), 1))(*[c.encode() for c in chroms])
    a.n_chroms, a.chroms = len(chroms), carr
    a.loc, a.cnt, a.pi, a.alt_allele, a.second_allele, a.fl1, a.fl2, a.biallelic = (p(k) for k in keep[3:11])
    a.n_dyn, a.dyn_cnt, a.dyn_pi = int(res.n_dyn), p(keep[11]), p(keep[12])
    names_buf = C.create_string_buffer(names, len(names) + 1)
    a.dyn_names, a.dyn_name_off = C.addressof(names_buf), p(name_off)
    a.hp1, a.hp2 = p(hp1), p(hp2)
    a.finalize, a.threshold = int(bool(finalize)), int(threshold)
    a.n_trf, a.trf_chrom, a.trf_lo, a.trf_hi = len(t_ch), p(t_ch), p(t_lo), p(t_hi)
    tags_buf = C.create_string_buffer(r_tags, len(r_tags) + 1)
    a.n_rm, a.rm_chrom, a.rm_lo, a.rm_hi, a.rm_tags, a.rm_tag_off = len(r_ch), p(r_ch), p(r_lo), p(r_hi), C.addressof(tags_buf), p(r_toff)
    a.threads = int(threads)
    if loci.ref_id.shape[0] < stride and n > 0:
        raise ValueError("emit_rows: loci arrays shorter than the result stride")
    out = _bamio.smc_rows_out()
    rc = lib.smc_rows_emit(C.byref(a), C.byref(out))
    if rc != 0:
        k = int(out.bad_row)
        i = int(k if order is None or k < 0 else order[k])
        where = (chroms[int(loci.ref_id[i])], "%d" % (int(loci.pos0[i]) + 1)) if 0 <= i < n else "?"
        if rc == -2:
            raise RuntimeError("Exception thrown in vc() at location: %s (device status 0x%x)" % (where, int(out.bad_status)))
        if rc == -3:
            raise RuntimeError("format_rows: no device HP/LowC flags for a candidate that needs them (pass hp_flags=device_hp_flags(...))")
        raise RuntimeError("smc_rows_emit failed (%d) at %s" % (rc, where))
    try:
        nr = a.n_rows

        def take(buf, off):
            o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(nr + 1,)).copy()
            return C.string_at(buf, int(o[-1])), o
        all_b, all_off = take(out.all, out.all_off)
        cut_b, cut_off = take(out.cut, out.cut_off)
        vcf_b, vcf_off = take(out.vcf, out.vcf_off)
    finally:
        lib.smc_rows_free(C.byref(out))
    return EmittedRows(all_b, all_off, cut_b, cut_off, vcf_b, vcf_off)
