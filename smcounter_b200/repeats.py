"""Repeat-region filters of main() (reference smCounter.py:699-785), host side.

The reference shells out to bedtools 2.25 (``merge -c 4 -o distinct``, ``merge``, ``intersect -a -b``, ``sort``;
smCounter.py:700-710) and then scans the two resulting region lists per output row (:752-785).  bedtools is a
third-party binary that is not part of the reference tree; its three operations are restated here on in-memory
rows with numpy, and the per-row scan is replaced by a first-match lookup with identical results:

  * ``merge`` walks the file in the order given (the reference does NOT pre-sort its inputs, :702,:706) and merges a
    row into the current block when it is on the same chromosome and starts at or before the block end
    (overlapping *and* book-ended features); ``-c 4 -o distinct`` joins the distinct column-4 values with ',' in
    lexicographic order;
  * ``intersect -a A -b B`` emits A clipped to every overlapping B feature, keeping A's extra columns;
  * ``sort`` orders by chromosome (lexicographic) then start.

Row filter (:752-785): rows whose printed PI truncates to >= 5 and whose ALT is not 'DEL' get the tag of the FIRST
region (in sorted order) with ``locL < pos <= locR`` from the TRF list ("RepT;", guarded by ``VMF < 40`` which is
always true because VMF is a fraction -- kept as is, :772) and from the RepeatMasker list (RepS / LowC / SL /
Other_Repeat).  Finally ``';'`` becomes ``PASS`` and other values lose their leading/trailing ';' (:784).
"""
from __future__ import annotations

from collections import defaultdict


def read_bed_rows(path, ncols=None):
    """Whitespace-split rows of a BED file (the reference uses ``line.strip().split()``, smCounter.py:715,722)."""
    rows = []
    with open(path, "r") as fh:
        for line in fh:
            vals = line.strip().split()
            if not vals or line.startswith("track ") or line.startswith("#") or line.startswith("browser "):
                continue
            rows.append(tuple(vals if ncols is None else vals[:ncols]))
    return rows


def bed_merge(rows, distinct_col4=False):
    out = []
    cur = None
    for r in rows:
        chrom, s, e = r[0], int(r[1]), int(r[2])
        if cur is not None and cur[0] == chrom and s <= cur[2]:
            if e > cur[2]:
                cur[2] = e
            if distinct_col4:
                cur[3].add(r[3])
        else:
            if cur is not None:
                out.append(cur)
            cur = [chrom, s, e, {r[3]} if distinct_col4 else None]
    if cur is not None:
        out.append(cur)
    if distinct_col4:
        return [(c, s, e, ",".join(sorted(n))) for (c, s, e, n) in out]
    return [(c, s, e) for (c, s, e, _) in out]


def bed_sort(rows):
    return sorted(rows, key=lambda r: (r[0], int(r[1])))


def bed_intersect(a_rows, b_rows):
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


_RM_TAG = {"Simple_repeat": "RepS", "Low_complexity": "LowC", "Satellite": "SL"}


def build_repeat_regions(target_rows, trf_rows, rm_rows):
    """(trfRegions, rmRegions): chrom -> [(locL, locR, tag)] in the order the reference's scan sees them."""
    bedRepeatMasker = bed_sort(bed_merge(rm_rows, distinct_col4=True))                        # :702
    bedTarget = bed_sort(bed_merge(target_rows))                                              # :706
    rep1 = bed_sort(bed_intersect(trf_rows, bedTarget))                                       # :709
    rep2 = bed_sort(bed_intersect(bedRepeatMasker, bedTarget))                                # :710
    trf = defaultdict(list)
    for r in rep1:                                                                            # :713-717
        trf[r[0]].append((int(r[1]), int(r[2]), "RepT;"))
    rm = defaultdict(list)
    for (chrom, s, e, typeCodes) in rep2:                                                     # :720-734
        tags = [_RM_TAG.get(t, "Other_Repeat") for t in typeCodes.split(",")]
        rm[chrom].append((int(s), int(e), ";".join(tags) + ";"))
    return trf, rm


def _first_hit(regions, pos):
    for (locL, locR, tag) in regions:
        if locL < pos <= locR:
            return tag
    return None


def apply_repeat_filters(rows, trf, rm, idx_chrom=0, idx_pos=1, idx_alt=3, idx_pi=10, idx_vmf=14):
    """smCounter.py:752-785 on the 45-field rows (tab-joined strings); returns the new list."""
    out = []
    for line in rows:
        f = line.split("\t")
        try:
            pos = int(f[idx_pos])
            vmf = float(f[idx_vmf])
        except ValueError:                      # zero-coverage rows: left untouched (:757-764)
            out.append(line)
            continue
        try:
            pred = int(float(f[idx_pi]))
        except ValueError:
            pred = 0
        if pred >= 5 and f[idx_alt] != "DEL":
            if vmf < 40:
                tag = _first_hit(trf.get(f[idx_chrom], ()), pos)
                if tag:
                    f[-1] += tag
            tag = _first_hit(rm.get(f[idx_chrom], ()), pos)
            if tag:
                f[-1] += tag
        f[-1] = "PASS" if f[-1] == ";" else f[-1].strip(";")
        out.append("\t".join(f))
    return out
