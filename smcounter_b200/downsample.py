"""Barcode down-sampling (reference smCounter.py:486-500), host side.

When a locus has more passing barcodes than ``ds = maxMT or round(2*mtDepth)`` the reference does
``random.seed(pos); bcKeys = random.sample(bcDict.keys(), ds)`` with ``pos`` a *str*.  Which barcodes survive therefore
depends on three CPython-2.7 behaviours, restated here from the CPython 2.7 sources (Objects/stringobject.c
string_hash, Objects/dictobject.c, Lib/random.py) because no Python 2 exists in this image:

  * the 64-bit string hash (no -R): seeds MT19937 through ``init_by_array`` and places keys in dict slots;
  * dict iteration order = slot order of an open-addressing table (perturbed probing, growth at 2/3 fill to the first
    power of two > 4*used, > 2*used beyond 50 000 entries), keys inserted in pileup (first-seen) order;
  * ``random.sample`` for these sizes: the pool algorithm ``j = int(random()*(n-i)); pick pool[j]; pool[j] = pool[n-i-1]``
    (or the selection-set variant when the population is much larger than k).

The device lists, for each flagged locus, the barcodes of bcDict with the BAM index of their first passing read
(``smc_list_barcodes``); the mask drawn here goes back to the device as ``smc_umi_keep``.  Unverifiable without a
Python 2 interpreter and the example BAM ("parity unpinned"); exercised against the oracle's
independent restatement in tests.
"""
from __future__ import annotations

import random as _random
from math import ceil, log

import numpy as np

from . import _ffi
from .soa import umi_string

_M64 = (1 << 64) - 1


def py2_string_hash(s: str) -> int:
    """CPython 2.7 ``hash(str)`` on a 64-bit build, as a signed integer."""
    if not s:
        return 0
    b = s.encode("latin-1", "replace")
    x = (b[0] << 7) & _M64
    for c in b:
        x = ((1000003 * x) & _M64) ^ c
    x ^= len(b)
    if x >= 1 << 63:
        x -= 1 << 64
    return -2 if x == -1 else x


def py2_dict_key_order(keys):
    """Iteration order of a CPython 2.7 dict after inserting the (distinct) ``keys`` in the given order."""
    size, used = 8, 0
    table = [None] * size

    def insert(tab, mask, key, h):
        i = h & mask
        perturb = h & _M64
        while tab[i] is not None:
            i = (5 * i + perturb + 1) & _M64
            perturb >>= 5
            i &= mask
        tab[i] = (key, h)

    for k in keys:
        h = py2_string_hash(k)
        insert(table, size - 1, k, h)
        used += 1
        if used * 3 >= size * 2:                                   # dictresize(mp, (used > 50000 ? 2 : 4) * used)
            minused = (2 if used > 50000 else 4) * used
            new = 8
            while new <= minused:
                new <<= 1
            old = table
            table = [None] * new
            size = new
            for e in old:
                if e is not None:
                    insert(table, size - 1, e[0], e[1])
    return [e[0] for e in table if e is not None]


def py2_seeded_sample(seed_str: str, population: list, k: int) -> list:
    """``random.seed(seed_str); random.sample(population, k)`` of CPython 2.7."""
    rng = _random.Random()
    rng.seed(py2_string_hash(seed_str) & _M64)          # Py2: init_by_array over (unsigned long)hash(str); Py3 int seed: same
    n = len(population)
    if not 0 <= k <= n:
        raise ValueError("sample larger than population")
    rnd = rng.random
    result = [None] * k
    setsize = 21
    if k > 5:
        setsize += 4 ** int(ceil(log(k * 3, 4)))
    if n <= setsize:
        pool = list(population)
        for i in range(k):
            j = int(rnd() * (n - i))
            result[i] = pool[j]
            pool[j] = pool[n - i - 1]
    else:
        selected = set()
        for i in range(k):
            j = int(rnd() * n)
            while j in selected:
                j = int(rnd() * n)
            selected.add(j)
            result[i] = population[j]
    return result


def draw_keep_masks(caller, res, reads, loci, chroms, prm):
    """For loci flagged SMC_ST_NEED_DOWNSAMPLE: list their barcodes on the device, draw the reference's sample on the
    host and return the UmiKeep mask (or None when no locus needs it)."""
    from .caller import UmiKeep
    flagged = np.flatnonzero(res.loc[_ffi.L_STATUS, :loci.n] & _ffi.ST_NEED_DOWNSAMPLE)
    if len(flagged) == 0:
        return None
    off, umis, first = caller.list_barcodes(flagged)
    ds = prm.ds
    mapping = {}
    names = reads.umi_names
    for k, i in enumerate(flagged):
        u = umis[off[k]:off[k + 1]]
        f = first[off[k]:off[k + 1]]
        order = np.argsort(f, kind="stable")                        # first-seen (pileup) order = dict insertion order
        bcs = [umi_string(int(c), names) for c in u[order]]
        codes = {bc: int(c) for bc, c in zip(bcs, u[order])}
        population = py2_dict_key_order(bcs)
        kept = py2_seeded_sample(str(int(loci.pos0[i]) + 1), population, ds)
        mapping[int(i)] = [codes[b] for b in kept]
    return UmiKeep(mapping)
