"""Sharding of the target across GPUs: loci are independent (vc() takes one position and
shares nothing, smCounter.py:274,684), so BED intervals are split into per-GPU groups balanced by estimated pileup
depth and each GPU receives only the reads that overlap its intervals.  No collective is needed: the host
concatenates the per-locus rows in BED order (the reference gathers ``p.get()`` in submission order, :685).
"""
from __future__ import annotations

import numpy as np

from .soa import Loci, ReadsSoA


def estimate_interval_events(reads: ReadsSoA, intervals, chroms, locator: "ReadLocator | None" = None) -> np.ndarray:
    """Estimated pileup read-events per interval = sum over reads of the overlap length with the interval.  Reads in BAM
    coordinate order: only the reads that can reach the interval are looked at (ReadLocator.candidates)."""
    loc = locator if locator is not None else ReadLocator(reads, chroms)
    out = np.zeros(len(intervals), dtype=np.float64)
    if loc.sorted:
        for k, (c, s, e) in enumerate(intervals):
            lo, hi = loc.candidates(c, s, e)
            if hi > lo:
                ov = np.minimum(loc.ends[lo:hi], e) - np.maximum(loc.starts[lo:hi], s)
                out[k] = float(ov[ov > 0].sum())
        return out
    for k, (c, s, e) in enumerate(intervals):
        r = loc.cidx.get(c)
        if r is None or e <= s:
            continue
        m = reads.ref_id == r
        ov = np.minimum(loc.ends[m], e) - np.maximum(loc.starts[m], s)
        out[k] = float(ov[ov > 0].sum())
    return out


def assign_intervals(weights, n_shards: int):
    """Greedy longest-processing-time assignment; returns shard index per interval (deterministic)."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(n_shards, dtype=np.float64)
    shard = np.zeros(len(w), dtype=np.int64)
    for k in order:
        g = int(np.argmin(load))
        shard[k] = g
        load[g] += w[k] + 1.0          # +1 so that empty intervals still spread out
    return shard, load


class ReadLocator:
    """Finds the reads whose reference span touches an interval.  For reads in BAM coordinate order (checked once) this is a
    binary search per interval over each contig's start positions plus an exact end test on the few candidates; any other
    order falls back to a full scan per interval."""

    def __init__(self, reads: ReadsSoA, chroms):
        self.reads = reads
        self.cidx = {c: i for i, c in enumerate(chroms)}
        stats = None
        if reads.n > (1 << 16) and reads.qual_bits == 8 and reads.scalar_bits == 32 and reads.seq_bits == 4 and reads.__dict__.get("_ref_end") is None:
            try:                                # big plain SoAs: reference ends, order check and longest span in one native pass
                from ._bamio import order_stats_native
                stats = order_stats_native(reads)
                reads.__dict__["_ref_end"] = stats[0]
            except ImportError:
                stats = None
        self.starts = reads.pos                 # int32 is fine for searchsorted and comparisons
        self.ends = reads.ref_end()
        if stats is not None:
            self.sorted, self.max_span = stats[1], stats[2]
        else:
            key = (reads.ref_id.astype(np.int64) << 32) | self.starts.astype(np.int64)
            self.sorted = bool(reads.n == 0 or np.all(key[1:] >= key[:-1]))
            self.max_span = int((self.ends - self.starts).max()) if reads.n else 0
        self.block = {}
        if self.sorted and reads.n:
            rid = reads.ref_id
            cuts = np.flatnonzero(rid[1:] != rid[:-1]) + 1
            for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [reads.n]))):
                self.block[int(rid[a])] = (int(a), int(b))

    def candidates(self, chrom, s, e):
        """(lo, hi): index range that contains every read of ``chrom`` starting in [s - max_span, e)."""
        r = self.cidx.get(chrom)
        if r is None or e <= s or r not in self.block:
            return 0, 0
        a, b = self.block[r]
        st = self.starts[a:b]
        return a + int(np.searchsorted(st, s - self.max_span, side="left")), a + int(np.searchsorted(st, e, side="left"))

    def count_upper(self, chrom, s, e) -> int:
        lo, hi = self.candidates(chrom, s, e)
        return hi - lo

    def select(self, intervals) -> np.ndarray:
        """Ascending indices of the reads that touch one of ``intervals`` (BAM order is kept)."""
        n = self.reads.n
        if self.sorted:
            # only the index span the candidates of these intervals cover is looked at (a shard of a big file is a small part of it)
            spans = [(lo, hi, s) for (lo, hi), s in ((self.candidates(c, s, e), s) for (c, s, e) in intervals) if hi > lo]
            if not spans:
                return np.zeros(0, dtype=np.int64)
            base, top = min(t[0] for t in spans), max(t[1] for t in spans)
            keep = np.zeros(top - base, dtype=bool)
            for (lo, hi, s) in spans:
                keep[lo - base:hi - base] |= self.ends[lo:hi] > s
            return np.flatnonzero(keep) + base
        keep = np.zeros(n, dtype=bool)
        for (c, s, e) in intervals:
            r = self.cidx.get(c)
            if r is None or e <= s:
                continue
            keep |= (self.reads.ref_id == r) & (self.starts < e) & (self.ends > s)
        return np.flatnonzero(keep)


def reads_for_intervals(reads: ReadsSoA, intervals, chroms) -> np.ndarray:
    """Ascending indices of the reads whose reference span touches one of ``intervals`` (BAM order is kept)."""
    return ReadLocator(reads, chroms).select(intervals)


def plan_batches(locator: ReadLocator, intervals, idxs, max_payload_bytes: int = 1 << 30, max_loci: int = 1 << 21,
                 max_reads: int = 1 << 27):
    """Cut the interval indices ``idxs`` (BED order) into consecutive batches that respect the per-batch limits of
    libsmc_b200 (include/smc_b200.h: < 4 GiB of bases / qualities, <= 2^30 reads, <= 4 194 302 loci) with a wide margin,
    from an upper estimate of the reads per interval.  A whole panel or exome goes through one GPU as a stream of batches."""
    reads = locator.reads
    per_read = (float(reads.seq.nbytes + reads.qual.nbytes) / reads.n) if reads.n else 0.0
    batches, cur, loci, nreads = [], [], 0, 0
    for k in idxs:
        c, s, e = intervals[k]
        n_loci = max(0, e - s)
        cnt = locator.count_upper(c, s, e) if locator.sorted else reads.n
        if cur and (loci + n_loci > max_loci or (nreads + cnt) * per_read > max_payload_bytes or nreads + cnt > max_reads):
            batches.append(cur)
            cur, loci, nreads = [], 0, 0
        cur.append(k)
        loci += n_loci
        nreads += cnt
    if cur:
        batches.append(cur)
    return batches


def split_long_intervals(intervals, locator: ReadLocator | None = None, max_loci: int = 1 << 20, max_payload_bytes: int = 1 << 30,
                         max_reads: int = 1 << 27):
    """BED intervals no single one of which exceeds the per-batch limits: an interval with more positions than ``max_loci`` (a
    whole-chromosome line), or -- when a ``locator`` over coordinate-sorted reads is given -- with more overlapping reads than
    a batch may carry, is cut into consecutive sub-intervals (halved until each piece fits or is one position wide).  The
    reference takes any interval size (it works per position, smCounter.py:675-680); rows are per position, so the pieces'
    rows concatenate to the interval's."""
    per_read = 0.0
    if locator is not None and locator.reads.n:
        per_read = float(locator.reads.seq.nbytes + locator.reads.qual.nbytes) / locator.reads.n

    def too_big(c, s, e):
        if e - s > max_loci:
            return True
        if locator is None or not locator.sorted or e - s <= 1:
            return False
        cnt = locator.count_upper(c, s, e)
        return cnt > max_reads or cnt * per_read > max_payload_bytes

    out = []
    stack = list(reversed([tuple(iv) for iv in intervals]))
    while stack:
        c, s, e = stack.pop()
        if e - s > 1 and too_big(c, s, e):
            m = (s + e) // 2
            stack.append((c, m, e))
            stack.append((c, s, m))
        else:
            out.append((c, s, e))
    return out


def subset_loci(loci: Loci, intervals, chroms):
    """Indices (into ``loci``) of the unique loci that fall inside ``intervals``."""
    cidx = {c: i for i, c in enumerate(chroms)}
    keep = np.zeros(loci.n, dtype=bool)
    for (c, s, e) in intervals:
        r = cidx.get(c)
        if r is None:
            continue
        keep |= (loci.ref_id == r) & (loci.pos0 >= s) & (loci.pos0 < e)
    return np.flatnonzero(keep)


def plan_shards(reads: ReadsSoA, intervals, chroms, n_shards: int, locator: "ReadLocator | None" = None):
    """[(interval indices, estimated events)] per shard; intervals keep their BED order inside a shard."""
    if n_shards <= 1:
        return [(list(range(len(intervals))), 0.0)]          # nothing to balance: skip the estimate
    w = estimate_interval_events(reads, intervals, chroms, locator)
    shard, load = assign_intervals(w, n_shards)
    return [([int(k) for k in np.flatnonzero(shard == g)], float(load[g])) for g in range(n_shards)]


def interleave_rows(plan, intervals, shard_rows):
    """Per-shard row lists (each in the BED order of the shard's own intervals) -> one list in the BED order of
    ``intervals`` (what the reference's in-order ``p.get()`` gather produces, smCounter.py:685)."""
    per_interval = {}
    for g, (idxs, _) in enumerate(plan):
        o = 0
        for k in idxs:
            n = max(0, intervals[k][2] - intervals[k][1])
            per_interval[k] = shard_rows[g][o:o + n]
            o += n
    out = []
    for k in range(len(intervals)):
        out.extend(per_interval.get(k, ()))
    return out
