"""Target loci: BED -> the reference's locus list (smCounter.py:675-680) -> the sorted unique ``Loci`` batch that
crosses the C-ABI, plus the map back to BED order (rows are emitted in BED order, duplicates preserved)."""
from __future__ import annotations

import numpy as np

from .soa import Loci


def intervals_from_bed_lines(lines):
    """[(chrom, start, end)] in file order; 'track ' lines skipped exactly as smCounter.py:677."""
    out = []
    for line in lines:
        if not line.startswith("track "):
            chrom, s, e = line.strip().split("\t")[0:3]
            out.append((chrom, int(s), int(e)))
    return out


def loc_list(intervals):
    """smCounter.py:679-680: every position of [start, end) as (chrom, str(pos+1))."""
    return [(c, str(p + 1)) for (c, s, e) in intervals for p in range(s, e)]


def build_loci(intervals, chroms, refs):
    """Returns (Loci sorted by (ref_id, pos0) without duplicates, bed_order) where bed_order[k] is the index into
    Loci of the k-th row of the reference's locList."""
    cidx = {c: i for i, c in enumerate(chroms)}
    rid, pos = [], []
    for (c, s, e) in intervals:
        if e > s:
            rid.append(np.full(e - s, cidx[c], dtype=np.int64))
            pos.append(np.arange(s, e, dtype=np.int64))
    if not rid:
        return Loci(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8)), np.zeros(0, np.int64)
    rid = np.concatenate(rid)
    pos = np.concatenate(pos)
    key = (rid << 32) | pos
    ukey, inverse = np.unique(key, return_inverse=True)
    u_rid = (ukey >> 32).astype(np.int32)
    u_pos = (ukey & 0xFFFFFFFF).astype(np.int32)
    base = np.full(len(ukey), ord("N"), dtype=np.uint8)
    # reference bases, fetched per run of consecutive positions (origRef, smCounter.py:311-313)
    brk = np.flatnonzero((np.diff(u_rid) != 0) | (np.diff(u_pos) != 1)) + 1
    starts = np.concatenate(([0], brk))
    ends = np.concatenate((brk, [len(ukey)]))
    for a, b in zip(starts, ends):
        chrom = chroms[int(u_rid[a])]
        s = refs.fetch(chrom, int(u_pos[a]), int(u_pos[b - 1]) + 1).upper()
        arr = np.frombuffer(s.encode(), dtype=np.uint8)
        base[a:a + len(arr)] = arr
    return Loci(u_rid, u_pos, base), inverse.astype(np.int64)
