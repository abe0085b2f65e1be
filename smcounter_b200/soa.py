"""Flat structure-of-arrays buffers that cross the C-ABI (include/smc_b200.h: smc_reads_soa, smc_loci).

One entry per BAM record, in BAM (coordinate) order -- the index of a read IS its pileup order, which the
reference's fragment merge depends on (smCounter.py:467-479).  Per-read identity fields replace the
qname parsing at smCounter.py:319-325: ``umi`` is an injective 64-bit code of the barcode string and
``frag_id`` the dictionary id of (barcode, readid), assigned in order of first appearance in the BAM.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# BAM 4-bit base codes ("=ACMGRSVTWYHKDBN")
NT16 = "=ACMGRSVTWYHKDBN"
_NT16_LUT = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(NT16):
    _NT16_LUT[ord(_c)] = _i
    _NT16_LUT[ord(_c.lower())] = _i
_NT16_CHARS = np.frombuffer(NT16.encode(), dtype=np.uint8)

CIGAR_OPS = "MIDNSHP=X"


@dataclass
class ReadsSoA:
    ref_id: np.ndarray      # int32   contig index into ``chroms``
    pos: np.ndarray         # int32   0-based leftmost reference position
    flag: np.ndarray        # uint16  BAM flag (0x10 reverse, 0x40 read1, 0x80 read2, 0x4 unmapped)
    mapq: np.ndarray        # uint8
    nm: np.ndarray          # int32   NM tag, 0 if absent (smCounter.py:329-334)
    l_seq: np.ndarray       # int32   query length incl. soft clips
    seq_off: np.ndarray     # int64   byte offset of the read's packed bases in ``seq``
    qual_off: np.ndarray    # int64   byte offset of the read's qualities in ``qual``
    cigar_off: np.ndarray   # int64   index of the read's first CIGAR word in ``cigar``
    n_cigar: np.ndarray     # uint16
    umi: np.ndarray         # uint64  injective barcode code
    frag_id: np.ndarray     # uint32  (barcode, readid) id, first-appearance order
    seq: np.ndarray         # uint8   4-bit BAM nibbles, high nibble first, each read byte-aligned
    qual: np.ndarray        # uint8   phred
    cigar: np.ndarray       # uint32  len<<4 | op
    chroms: list = field(default_factory=list)
    umi_names: dict | None = None   # optional code -> barcode string (host side only)
    packed: bool = False            # payloads stored back to back in read order: the offsets need not cross the ABI (NULL)
    store_lo: np.ndarray | None = None    # int32, optional stored window (include/smc_b200.h): first stored query base (even) ...
    store_len: np.ndarray | None = None   # int32  ... and number of stored bases; None = reads stored whole
    # compact wire encodings (include/smc_b200.h, ABI v3); a compacted SoA is an upload format, not a working one
    scalar_bits: int = 32                 # 16 / 8: nm / l_seq / store_lo / store_len are uint16 / uint8 arrays
    qual_bits: int = 8                    # 4 / 2: ``qual`` holds qual_bits-wide codes (low bits first, reads byte aligned) ...
    qual_lut: np.ndarray | None = None    # ... and this is the phred value of each code
    seq_bits: int = 4                     # 2: ``seq`` holds A C G T = 0..3 (low bits first, reads byte aligned) ...
    seq_exc: tuple | None = None          # ... and (read u32, base index u32, BAM nibble u8) of every other base

    @property
    def n(self) -> int:
        return int(self.ref_id.shape[0])

    def nbytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in
                   ("ref_id", "pos", "flag", "mapq", "nm", "l_seq", "seq_off", "qual_off", "cigar_off", "n_cigar",
                    "umi", "frag_id", "seq", "qual", "cigar")) + (0 if self.store_lo is None else self.store_lo.nbytes + self.store_len.nbytes) \
            + (0 if self.seq_exc is None else sum(a.nbytes for a in self.seq_exc))

    def stored_len(self) -> np.ndarray:
        """Stored bases per read (== l_seq unless the SoA was trimmed to its targets)."""
        return (self.l_seq if self.store_len is None else self.store_len).astype(np.int64)

    def select(self, idx: np.ndarray) -> "ReadsSoA":
        """Sub-batch with the reads ``idx`` (ascending), variable-length payloads re-packed."""
        if self.qual_bits != 8 or self.scalar_bits != 32 or self.seq_bits != 4:
            raise ValueError("a compacted SoA is an upload format: select / trim before compact()")
        idx = np.asarray(idx, dtype=np.int64)
        l_seq = self.stored_len()[idx]
        sb = (l_seq + 1) // 2
        nc = self.n_cigar[idx].astype(np.int64)
        new_seq_off = np.concatenate(([0], np.cumsum(sb)))[:-1]
        new_qual_off = np.concatenate(([0], np.cumsum(l_seq)))[:-1]
        new_cig_off = np.concatenate(([0], np.cumsum(nc)))[:-1]

        def gather(src, offs, lens):
            tot = int(lens.sum())
            if tot == 0:
                return np.zeros(0, dtype=src.dtype)
            starts = np.repeat(offs - np.concatenate(([0], np.cumsum(lens)))[:-1], lens)
            return src[starts + np.arange(tot, dtype=np.int64)]

        return ReadsSoA(
            ref_id=self.ref_id[idx], pos=self.pos[idx], flag=self.flag[idx], mapq=self.mapq[idx], nm=self.nm[idx],
            l_seq=self.l_seq[idx], seq_off=new_seq_off, qual_off=new_qual_off, cigar_off=new_cig_off,
            n_cigar=self.n_cigar[idx], umi=self.umi[idx],
            # ids stay dense (< n reads, include/smc_b200.h) and keep their relative order = the fragment order inside a barcode
            frag_id=np.unique(self.frag_id[idx], return_inverse=True)[1].astype(np.uint32) if len(idx) else self.frag_id[idx],
            seq=gather(self.seq, self.seq_off[idx], sb), qual=gather(self.qual, self.qual_off[idx], l_seq),
            cigar=gather(self.cigar, self.cigar_off[idx], nc), chroms=self.chroms, umi_names=self.umi_names, packed=True,
            store_lo=None if self.store_lo is None else self.store_lo[idx], store_len=None if self.store_len is None else self.store_len[idx])

    def repack(self, block: int = 1 << 18) -> "ReadsSoA":
        """Same reads, with bases / qualities / CIGARs stored in read order (what a BAM decode produces).  libsmc_b200
        accepts any layout, but only a read-ordered payload lets smc_call_batch overlap the upload with the kernels."""
        l_seq = self.stored_len()
        sb = (l_seq + 1) // 2
        nc = self.n_cigar.astype(np.int64)
        out = {}
        for name, src, offs, lens in (("seq", self.seq, self.seq_off, sb), ("qual", self.qual, self.qual_off, l_seq),
                                      ("cigar", self.cigar, self.cigar_off, nc)):
            new_off = np.concatenate(([0], np.cumsum(lens)))
            dst = np.empty(int(new_off[-1]), dtype=src.dtype)
            for a in range(0, self.n, block):
                b = min(self.n, a + block)
                ln = lens[a:b]
                tot = int(ln.sum())
                if tot:
                    starts = np.repeat(offs[a:b] - (new_off[a:b] - new_off[a]), ln)
                    dst[new_off[a]:new_off[b]] = src[starts + np.arange(tot, dtype=np.int64)]
            out[name] = (dst, new_off[:-1].copy())
        return ReadsSoA(ref_id=self.ref_id, pos=self.pos, flag=self.flag, mapq=self.mapq, nm=self.nm, l_seq=self.l_seq,
                        seq_off=out["seq"][1], qual_off=out["qual"][1], cigar_off=out["cigar"][1], n_cigar=self.n_cigar, umi=self.umi,
                        frag_id=self.frag_id, seq=out["seq"][0], qual=out["qual"][0], cigar=out["cigar"][0], chroms=self.chroms,
                        umi_names=self.umi_names, packed=True, store_lo=self.store_lo, store_len=self.store_len)

    def trim_to_targets(self, intervals, block: int = 1 << 18) -> "ReadsSoA":
        """Same reads with only the bases a pileup over ``intervals`` can see kept in seq / qual (store_lo / store_len of
        include/smc_b200.h): for a read that is one plain aligned run, the query bases from its first to its last target
        position (start rounded down to an even base); every other read is kept whole.  Payload packed in read order."""
        if self.store_lo is not None:
            raise ValueError("already trimmed")
        n = self.n
        cidx = {c: i for i, c in enumerate(self.chroms)}
        l_seq = self.l_seq.astype(np.int64)
        ops = self.cigar & 0xF
        first = self.cigar[np.minimum(self.cigar_off, max(len(self.cigar) - 1, 0))] if len(self.cigar) else np.zeros(n, np.uint32)
        is_ref = np.isin(ops, (0, 7, 8))
        bad = np.isin(ops, (1, 2, 3, 5, 6))
        cs_ref = np.concatenate(([0], np.cumsum(is_ref)))
        cs_bad = np.concatenate(([0], np.cumsum(bad)))
        a, b = self.cigar_off, self.cigar_off + self.n_cigar.astype(np.int64)
        simple = ((cs_ref[b] - cs_ref[a]) == 1) & ((cs_bad[b] - cs_bad[a]) == 0)
        left_sp = np.where((self.n_cigar > 0) & ((first & 0xF) == 4), (first >> 4).astype(np.int64), 0)
        start = self.pos.astype(np.int64)
        end = self.ref_end()
        # first / last target position inside [start, end) per read: over the sorted unique target positions of its contig
        p_lo = np.full(n, -1, dtype=np.int64)
        p_hi = np.full(n, -1, dtype=np.int64)
        for c in set(iv[0] for iv in intervals):
            if c not in cidx:
                continue
            pos = np.unique(np.concatenate([np.arange(s, e, dtype=np.int64) for (cc, s, e) in intervals if cc == c and e > s] or
                                           [np.zeros(0, np.int64)]))
            if not len(pos):
                continue
            m = np.flatnonzero(self.ref_id == cidx[c])
            i0 = np.searchsorted(pos, start[m], side="left")
            i1 = np.searchsorted(pos, end[m], side="left")
            has = i1 > i0
            p_lo[m[has]] = pos[i0[has]]
            p_hi[m[has]] = pos[i1[has] - 1]
        covered = p_lo >= 0
        q_lo = np.where(simple & covered, (p_lo - start + left_sp) & ~1, 0)
        q_hi = np.where(simple & covered, p_hi - start + left_sp + 1, np.where(simple, 0, l_seq))
        q_lo = np.where(simple, q_lo, 0)
        store_lo = q_lo.astype(np.int32)
        store_len = np.maximum(q_hi - q_lo, 0).astype(np.int32)
        # payloads: seq bytes [store_lo/2, ...), qual bytes [store_lo, ...)
        sl = store_len.astype(np.int64)
        out = {}
        for name, src, offs, lens in (("seq", self.seq, self.seq_off + q_lo // 2, (sl + 1) // 2), ("qual", self.qual, self.qual_off + q_lo, sl),
                                      ("cigar", self.cigar, self.cigar_off, self.n_cigar.astype(np.int64))):
            new_off = np.concatenate(([0], np.cumsum(lens)))
            dst = np.empty(int(new_off[-1]), dtype=src.dtype)
            for x in range(0, n, block):
                y = min(n, x + block)
                ln = lens[x:y]
                tot = int(ln.sum())
                if tot:
                    starts = np.repeat(offs[x:y] - (new_off[x:y] - new_off[x]), ln)
                    dst[new_off[x]:new_off[y]] = src[starts + np.arange(tot, dtype=np.int64)]
            out[name] = (dst, new_off[:-1].copy())
        # an odd stored length leaves the low nibble of the last byte holding the next (unstored) base: harmless, never read
        return ReadsSoA(ref_id=self.ref_id, pos=self.pos, flag=self.flag, mapq=self.mapq, nm=self.nm, l_seq=self.l_seq,
                        seq_off=out["seq"][1], qual_off=out["qual"][1], cigar_off=out["cigar"][1], n_cigar=self.n_cigar, umi=self.umi,
                        frag_id=self.frag_id, seq=out["seq"][0], qual=out["qual"][0], cigar=out["cigar"][0], chroms=self.chroms,
                        umi_names=self.umi_names, packed=True, store_lo=store_lo, store_len=store_len)

    def compact(self, block: int = 1 << 18, seq_bits_wanted: int = 2, scalar_bits_min: int = 8) -> "ReadsSoA":
        """The same reads in the compact upload encodings of include/smc_b200.h (ABI v3): 8- or 16-bit nm / l_seq / store_lo /
        store_len when every value fits, and 2- or 4-bit quality codes when the batch shows at most 4 / 16 distinct
        qualities (sequencers that bin qualities; a batch with more keeps one byte per base), 2-bit bases with a side list for
        the non-ACGT ones (`seq_bits_wanted=4` keeps BAM's nibbles).  Needs a packed layout (repack() / trim_to_targets() /
        select()).  On the cfg-2 panel batch: 157 -> 72 bytes per read over PCIe."""
        if self.qual_bits != 8 or self.scalar_bits != 32 or self.seq_bits != 4 or self.ref_id.dtype != np.int32 or self.umi.dtype != np.uint64:
            raise ValueError("already compact")
        src = self if self.packed else self.repack()
        kw = {f: getattr(src, f) for f in ("ref_id", "pos", "flag", "mapq", "seq_off", "qual_off", "cigar_off", "n_cigar", "umi", "frag_id",
                                            "seq", "cigar", "chroms", "umi_names")}
        scal = {"nm": src.nm, "l_seq": src.l_seq, "store_lo": src.store_lo, "store_len": src.store_len}
        top = max([0] + [int(a.max()) for a in scal.values() if a is not None and len(a)])
        low = min([0] + [int(a.min()) for a in scal.values() if a is not None and len(a)])
        scalar_bits = max(32 if (low < 0 or top >= 65536) else 16 if top >= 256 else 8, scalar_bits_min)
        if scalar_bits != 32:
            scal = {k: (None if a is None else a.astype(np.uint16 if scalar_bits == 16 else np.uint8)) for k, a in scal.items()}
        present = np.flatnonzero(np.bincount(src.qual, minlength=256)) if len(src.qual) else np.zeros(0, np.int64)
        bits = 2 if len(present) <= 4 else 4 if len(present) <= 16 else 8
        lens = src.stored_len()
        uoff = np.concatenate(([0], np.cumsum(lens)))
        qual, lut = src.qual, None
        if bits != 8:
            lut = np.zeros(1 << bits, np.uint8)
            lut[:len(present)] = present
            code_of = np.zeros(256, np.uint8)
            code_of[present] = np.arange(len(present), dtype=np.uint8)
            qual = _pack_codes(lambda a, b: code_of[src.qual[uoff[a]:uoff[b]]], lens, bits, block)
        # bases: 2 bits each (A C G T = 0 1 2 3); everything else (N, IUPAC codes, '=') goes to a side list of
        # (read, base index, 4-bit code) and travels as code 0
        seq, seq_bits, exc = src.seq, 4, None
        if seq_bits_wanted == 2:
            soff = np.concatenate(([0], np.cumsum((lens + 1) // 2)))
            code2 = np.zeros(16, np.uint8)
            code2[[1, 2, 4, 8]] = (0, 1, 2, 3)
            plain = np.zeros(16, bool)
            plain[[1, 2, 4, 8]] = True
            exc_read, exc_pos, exc_nib = [], [], []

            def base_codes(a, b):
                by = src.seq[soff[a]:soff[b]]
                nib = np.empty(2 * len(by), np.uint8)
                nib[0::2] = by >> 4
                nib[1::2] = by & 15
                ln = lens[a:b]
                tot = int(ln.sum())
                # nibble index of base i of read r inside this block: 2 * (soff[r] - soff[a]) + i
                rel = np.repeat(2 * (soff[a:b] - soff[a]) - (uoff[a:b] - uoff[a]), ln) + np.arange(tot, dtype=np.int64)
                nb = nib[rel]
                odd = np.flatnonzero(~plain[nb])
                if len(odd):
                    rd = np.searchsorted(uoff[a:b + 1] - uoff[a], odd, side="right") - 1
                    exc_read.append((rd + a).astype(np.uint32))
                    exc_pos.append((odd - (uoff[a:b] - uoff[a])[rd]).astype(np.uint32))
                    exc_nib.append(nb[odd])
                return code2[nb]
            seq = _pack_codes(base_codes, lens, 2, block)
            seq_bits = 2
            cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
            exc = (cat(exc_read, np.uint32), cat(exc_pos, np.uint32), cat(exc_nib, np.uint8))
        kw["seq"] = seq
        # reference indices below 256 travel as bytes, barcode codes below 2^32 (barcodes of <= 15 nt) as 32-bit words
        if src.n and int(src.ref_id.min()) >= 0 and int(src.ref_id.max()) < 256:
            kw["ref_id"] = src.ref_id.astype(np.uint8)
        if src.n and int(src.umi.max()) < (1 << 32):
            kw["umi"] = src.umi.astype(np.uint32)
        out = ReadsSoA(nm=scal["nm"], l_seq=scal["l_seq"], store_lo=scal["store_lo"], store_len=scal["store_len"], qual=qual, packed=True,
                        scalar_bits=scalar_bits, qual_bits=bits, qual_lut=lut, seq_bits=seq_bits, seq_exc=exc, **kw)
        if seq_bits == 2:            # what compact_bases() looks reads up with (allele names of insertions), made here once
            out.__dict__["_compact_memo"] = (np.concatenate(([0], np.cumsum((lens + 3) // 4))),
                                             (exc[0].astype(np.uint64) << np.uint64(32)) | exc[1].astype(np.uint64))
        return out

    def compact_bases(self, r: int, q0: int, n: int) -> list:
        """BAM nibble codes of stored bases [q0, q0 + n) of read r of a seq_bits == 2 SoA (the offsets and the sorted
        exception keys are memoised per instance: one cumsum over the reads the first time)."""
        memo = self.__dict__.get("_compact_memo")
        if memo is None:
            poff = np.concatenate(([0], np.cumsum((self.stored_len() + 3) // 4)))
            e = self.seq_exc
            keys = np.zeros(0, np.uint64) if e is None else (e[0].astype(np.uint64) << np.uint64(32)) | e[1].astype(np.uint64)
            self.__dict__["_compact_memo"] = memo = (poff, keys)
        poff, keys = memo
        so = int(poff[r])
        out = [1 << ((int(self.seq[so + (q >> 2)]) >> ((q & 3) * 2)) & 3) for q in range(q0, q0 + n)]
        if len(keys):
            lo = int(np.searchsorted(keys, np.uint64((r << 32) | max(q0, 0))))
            while lo < len(keys) and int(keys[lo]) < ((r << 32) | (q0 + n)):
                out[int(keys[lo]) - ((r << 32) | q0)] = int(self.seq_exc[2][lo])
                lo += 1
        return out

    def is_packed(self) -> bool:
        """True when every payload is stored back to back in read order (checked, O(n))."""
        l = self.stored_len()
        for off, lens, tot in ((self.seq_off, (l + 1) // 2, self.seq.shape[0]), (self.qual_off, l, self.qual.shape[0]),
                               (self.cigar_off, self.n_cigar.astype(np.int64), self.cigar.shape[0])):
            ends = np.cumsum(lens)
            if self.n and (not np.array_equal(off[1:], ends[:-1]) or off[0] != 0 or ends[-1] != tot):
                return False
            if not self.n and tot:
                return False
        return True

    def ref_end(self) -> np.ndarray:
        """0-based exclusive reference end of every read (host-side helper for sharding); memoised per instance."""
        cached = self.__dict__.get("_ref_end")
        if cached is not None and len(cached) == self.n:
            return cached
        self.__dict__["_ref_end"] = out = self._ref_end_compute()
        return out

    def _ref_end_compute(self) -> np.ndarray:
        if self.qual_bits == 8 and self.scalar_bits == 32 and self.seq_bits == 4 and self.n > (1 << 16):
            try:                                    # big plain SoAs: the threaded native pass (include/smc_soa.h)
                from ._bamio import ref_end_native
                return ref_end_native(self)
            except ImportError:
                pass
        ops = self.cigar & 0xF
        lens = (self.cigar >> 4).astype(np.int64)
        consumes = np.isin(ops, (0, 2, 3, 7, 8))
        w = np.where(consumes, lens, 0)
        cs = np.concatenate(([0], np.cumsum(w)))
        a = self.cigar_off
        b = self.cigar_off + self.n_cigar.astype(np.int64)
        return (self.pos.astype(np.int64) + cs[b] - cs[a]).astype(np.int64)


def _pack_codes(codes_of, lens: np.ndarray, bits: int, block: int) -> np.ndarray:
    """Fold per-base codes of `bits` bits into bytes, every read padded to whole bytes (first base in the low bits).
    codes_of(a, b) returns the codes of reads [a, b) back to back."""
    per = 8 // bits
    n = len(lens)
    poff = np.concatenate(([0], np.cumsum((lens * bits + 7) // 8)))
    uoff = np.concatenate(([0], np.cumsum(lens)))
    out = np.zeros(int(poff[-1]), np.uint8)
    for a in range(0, n, block):
        b = min(n, a + block)
        codes = codes_of(a, b)
        ln = lens[a:b]
        tot = int(ln.sum())
        if not tot:
            continue
        padded = np.zeros(int(poff[b] - poff[a]) * per, np.uint8)
        dst = np.repeat((poff[a:b] - poff[a]) * per - (uoff[a:b] - uoff[a]), ln) + np.arange(tot, dtype=np.int64)
        padded[dst] = codes
        m = padded.reshape(-1, per)
        acc = np.zeros(len(m), np.uint8)
        for k in range(per):
            acc |= m[:, k] << np.uint8(k * bits)
        out[poff[a]:poff[b]] = acc
    return out


@dataclass
class Loci:
    """Unique target positions, sorted by (ref_id, pos0) -- include/smc_b200.h: smc_loci."""
    ref_id: np.ndarray      # int32
    pos0: np.ndarray        # int32  0-based
    ref_base: np.ndarray    # uint8  ASCII upper-case reference base (smCounter.py:311-313)

    @property
    def n(self) -> int:
        return int(self.ref_id.shape[0])


def umi_code(bc: str, table: dict) -> int:
    """Injective 64-bit code for a barcode string: 2-bit pack with a length sentinel when the barcode is
    <= 31 nt of ACGT, else a dictionary id with the top bit set."""
    if len(bc) <= 31:
        v = 1
        for ch in bc:
            k = "ACGT".find(ch)
            if k < 0:
                break
            v = (v << 2) | k
        else:
            return v
    key = ("#", bc)
    if key not in table:
        table[key] = (1 << 63) | len(table)
    return table[key]


def umi_string(code: int, names: dict | None = None) -> str:
    if names is not None and code in names:
        return names[code]
    if code >> 63:
        return "U%x" % (code & ((1 << 63) - 1))
    s = []
    while code > 1:
        s.append("ACGT"[code & 3])
        code >>= 2
    return "".join(reversed(s))


def umi_strings_bulk(codes: np.ndarray) -> list:
    """umi_string() for an array of 2-bit packed codes (top bit clear), vectorised per barcode length."""
    codes = np.asarray(codes, dtype=np.uint64)
    out = [None] * len(codes)
    if len(codes) == 0:
        return out
    nbits = np.zeros(len(codes), dtype=np.int64)           # position of the length sentinel bit
    v = codes.copy()
    for sh in (32, 16, 8, 4, 2, 1):
        m = (v >> np.uint64(sh)) != 0
        nbits[m] += sh
        v[m] >>= np.uint64(sh)
    length = nbits // 2
    for L in np.unique(length):
        idx = np.flatnonzero(length == L)
        if L == 0:
            for i in idx:
                out[i] = ""
            continue
        shifts = (2 * np.arange(L - 1, -1, -1)).astype(np.uint64)
        digits = ((codes[idx, None] >> shifts[None, :]) & np.uint64(3)).astype(np.intp)
        chars = np.frombuffer(b"ACGT", dtype=np.uint8)[digits]
        strs = np.ascontiguousarray(chars).view("S%d" % L).ravel()
        for i, b in zip(idx, strs):
            out[i] = b.decode()
    return out


def pack_seq(seq: str) -> np.ndarray:
    codes = _NT16_LUT[np.frombuffer(seq.encode(), dtype=np.uint8)]
    if len(codes) & 1:
        codes = np.concatenate((codes, np.zeros(1, dtype=np.uint8)))
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def records_to_soa(records, chroms=None) -> ReadsSoA:
    """Build the SoA from pysam-free records (qname chrom pos flag mapq nm cigar seq qual), BAM order.

    Restates the identity parsing of smCounter.py:319-325: ``readid = ':'.join(parts[:-2])``, ``BC = parts[-2]``.
    """
    if chroms is None:
        chroms = []
        for r in records:
            if r.chrom not in chroms:
                chroms.append(r.chrom)
    cidx = {c: i for i, c in enumerate(chroms)}
    n = len(records)
    ref_id = np.zeros(n, np.int32); pos = np.zeros(n, np.int32); flag = np.zeros(n, np.uint16)
    mapq = np.zeros(n, np.uint8); nm = np.zeros(n, np.int32); l_seq = np.zeros(n, np.int32)
    seq_off = np.zeros(n, np.int64); qual_off = np.zeros(n, np.int64); cigar_off = np.zeros(n, np.int64)
    n_cigar = np.zeros(n, np.uint16); umi = np.zeros(n, np.uint64); frag_id = np.zeros(n, np.uint32)
    seqs, quals, cigs = [], [], []
    so = qo = co = 0
    umitab, fragtab, names = {}, {}, {}
    for i, r in enumerate(records):
        parts = r.qname.split(":")
        bc = parts[-2]
        readid = ":".join(parts[:-2])
        code = umi_code(bc, umitab)
        names[code] = bc
        fk = (bc, readid)
        if fk not in fragtab:
            fragtab[fk] = len(fragtab)
        ref_id[i] = cidx[r.chrom]; pos[i] = r.pos; flag[i] = r.flag; mapq[i] = r.mapq
        nm[i] = 0 if r.nm is None else r.nm
        l_seq[i] = len(r.seq); seq_off[i] = so; qual_off[i] = qo; cigar_off[i] = co; n_cigar[i] = len(r.cigar)
        umi[i] = code; frag_id[i] = fragtab[fk]
        ps = pack_seq(r.seq)
        seqs.append(ps); so += len(ps)
        quals.append(np.asarray(r.qual, dtype=np.uint8)); qo += len(r.seq)
        cigs.append(np.asarray([(l << 4) | op for (op, l) in r.cigar], dtype=np.uint32)); co += len(r.cigar)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return ReadsSoA(ref_id, pos, flag, mapq, nm, l_seq, seq_off, qual_off, cigar_off, n_cigar, umi, frag_id,
                    cat(seqs, np.uint8), cat(quals, np.uint8), cat(cigs, np.uint32), list(chroms), names)


def soa_to_records(soa: ReadsSoA, record_type=None):
    """Inverse of records_to_soa (used by tests to feed the same data to the oracle).  qname is synthesised as
    ``F<frag_id>:<barcode>:x`` so that (barcode, readid) round-trips."""
    from collections import namedtuple
    R = record_type or namedtuple("Read", "qname chrom pos flag mapq nm cigar seq qual")
    if soa.store_lo is not None:
        raise ValueError("soa_to_records needs whole reads (this SoA was trimmed to its targets)")
    out = []
    for i in range(soa.n):
        L = int(soa.l_seq[i])
        so = int(soa.seq_off[i])
        b = soa.seq[so: so + (L + 1) // 2]
        nib = np.empty(2 * len(b), np.uint8)
        nib[0::2] = b >> 4
        nib[1::2] = b & 15
        seq = _NT16_CHARS[nib[:L]].tobytes().decode()
        qo = int(soa.qual_off[i])
        qual = soa.qual[qo: qo + L].tolist()
        co = int(soa.cigar_off[i])
        cig = [(int(w) & 15, int(w) >> 4) for w in soa.cigar[co: co + int(soa.n_cigar[i])]]
        bc = umi_string(int(soa.umi[i]), soa.umi_names)
        out.append(R("F%d:%s:x" % (int(soa.frag_id[i]), bc), soa.chroms[int(soa.ref_id[i])], int(soa.pos[i]),
                     int(soa.flag[i]), int(soa.mapq[i]), int(soa.nm[i]), cig, seq, qual))
    return out
