"""Python-2.7 semantics that leak into smCounter's results (TEST INFRASTRUCTURE ONLY).

This module is part of ``oracle/`` -- the CPU restatement of the reference used as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg.
Nothing in the product package may import it.

What is restated here (reference = /root/reference/smCounter.py, CPython 2.7 semantics):

* ``py2round``   -- ``round(x, n)`` rounds half away from zero on the exact binary value
                    (used at smCounter.py:576-593).
* ``py2str``     -- ``str(float)`` is ``'%.12g'`` with a forced ``.0`` (smCounter.py:599).
* ``py2hash``    -- 64-bit string hash (no -R), needed for dict order and ``random.seed(pos)``
                    (smCounter.py:497).
* ``Py2Dict``    -- insertion/iteration order of a CPython-2.7 dict (open addressing, perturb probing),
                    needed for ``finalDict.items()`` tie order (smCounter.py:534) and for
                    ``bcDict.keys()`` as the sampling population (smCounter.py:498).
* ``py2_sample`` -- ``random.seed(str); random.sample(list, k)`` of CPython 2.7 (smCounter.py:497-498).

Parity status: there is no Python 2 interpreter in the build container, so these are pinned only by the
published check value ``hash('a') == 12416037344`` and by the CPython source they restate: "parity unpinned"
for dict order / sampling.
"""
from __future__ import annotations

import random as _random
from decimal import Decimal, ROUND_HALF_UP

_MASK64 = (1 << 64) - 1


def py2round(x: float, ndigits: int) -> float:
    """CPython 2.7 ``round(x, ndigits)``: correctly rounded, ties away from zero."""
    if x != x or x in (float("inf"), float("-inf")):
        return x
    q = Decimal(1).scaleb(-ndigits)
    return float(Decimal(x).quantize(q, rounding=ROUND_HALF_UP))


def py2str(v) -> str:
    """CPython 2.7 ``str(v)`` for the value types smCounter prints (int, float, str)."""
    if isinstance(v, bool):
        return "True" if v else "False"
    if isinstance(v, int):
        return "%d" % v
    if isinstance(v, float):
        if v != v:
            return "nan"
        if v == float("inf"):
            return "inf"
        if v == float("-inf"):
            return "-inf"
        s = "%.12g" % v
        if "." not in s and "e" not in s and "n" not in s:
            s += ".0"
        return s
    return str(v)


def py2hash(s: str) -> int:
    """CPython 2.7 string hash on a 64-bit build without hash randomisation (signed result)."""
    if len(s) == 0:
        return 0
    x = (ord(s[0]) << 7) & _MASK64
    for ch in s:
        x = ((1000003 * x) & _MASK64) ^ ord(ch)
    x ^= len(s)
    x &= _MASK64
    if x >= 1 << 63:
        x -= 1 << 64
    if x == -1:
        x = -2
    return x


class Py2Dict:
    """Order-faithful model of a CPython 2.7 dict with str keys (Objects/dictobject.c).

    Only what iteration order needs: insertion, deletion (dummy slots), resize.  Values are kept so the
    class can stand in for ``finalDict`` / ``bcDict`` in the oracle's py2 ordering mode.
    """

    _DUMMY = object()
    PERTURB_SHIFT = 5
    MINSIZE = 8

    def __init__(self):
        self.size = self.MINSIZE
        self.table = [None] * self.size  # None = never used; _DUMMY = deleted; else [hash, key, value]
        self.fill = 0  # active + dummy
        self.used = 0  # active

    def _lookup(self, key, h):
        mask = self.size - 1
        i = h & mask
        ep = self.table[i]
        if ep is None:
            return i, None
        freeslot = None
        if ep is self._DUMMY:
            freeslot = i
        elif ep[1] == key:
            return i, ep
        perturb = h & _MASK64  # unsigned view of the hash
        while True:
            i = (i << 2) + i + perturb + 1
            ep = self.table[i & mask]
            if ep is None:
                return (freeslot if freeslot is not None else i & mask), None
            if ep is self._DUMMY:
                if freeslot is None:
                    freeslot = i & mask
            elif ep[1] == key:
                return i & mask, ep
            perturb >>= self.PERTURB_SHIFT

    def _resize(self, minused):
        newsize = self.MINSIZE
        while newsize <= minused:
            newsize <<= 1
        old = [ep for ep in self.table if ep is not None and ep is not self._DUMMY]
        self.size = newsize
        self.table = [None] * newsize
        self.fill = 0
        self.used = 0
        for ep in old:  # re-insert live entries in slot order (insertdict_clean)
            mask = self.size - 1
            h = ep[0]
            i = h & mask
            perturb = h & _MASK64
            while self.table[i & mask] is not None:
                i = (i << 2) + i + perturb + 1
                perturb >>= self.PERTURB_SHIFT
            self.table[i & mask] = ep
            self.fill += 1
            self.used += 1

    def __setitem__(self, key, value):
        h = py2hash(key)
        slot, ep = self._lookup(key, h)
        if ep is not None:
            ep[2] = value
            return
        n_used = self.used
        if self.table[slot] is None:
            self.fill += 1
        self.table[slot] = [h, key, value]
        self.used += 1
        if self.used > n_used and self.fill * 3 >= self.size * 2:
            self._resize((4 if self.used <= 50000 else 2) * self.used)

    def __getitem__(self, key):
        _, ep = self._lookup(key, py2hash(key))
        if ep is None:
            raise KeyError(key)
        return ep[2]

    def __contains__(self, key):
        return self._lookup(key, py2hash(key))[1] is not None

    def __delitem__(self, key):
        slot, ep = self._lookup(key, py2hash(key))
        if ep is None:
            raise KeyError(key)
        self.table[slot] = self._DUMMY
        self.used -= 1

    def __len__(self):
        return self.used

    def keys(self):
        return [ep[1] for ep in self.table if ep is not None and ep is not self._DUMMY]

    def items(self):
        return [(ep[1], ep[2]) for ep in self.table if ep is not None and ep is not self._DUMMY]

    def values(self):
        return [ep[2] for ep in self.table if ep is not None and ep is not self._DUMMY]


def py2_dict_order(keys_in_insertion_order):
    """Iteration order a Py2 dict would have after inserting ``keys`` in the given order."""
    d = Py2Dict()
    for k in keys_in_insertion_order:
        d[k] = None
    return d.keys()


def py2_sample(seed_str: str, population: list, k: int) -> list:
    """``random.seed(seed_str); random.sample(population, k)`` as CPython 2.7 computes it.

    Py2 seeds MT19937 with ``init_by_array`` over ``(unsigned long)hash(seed_str)``; Py3's
    ``random.seed(int)`` does the same for a non-negative int.  ``random.sample`` in 2.7 uses the pool
    algorithm when ``n <= setsize`` (true for every call smCounter makes: k >= n/2 ... ) and the
    selection-set algorithm otherwise; both are restated.
    """
    from math import ceil, log

    rng = _random.Random()
    rng.seed(py2hash(seed_str) & _MASK64)
    n = len(population)
    if not 0 <= k <= n:
        raise ValueError("sample larger than population")
    rnd = rng.random
    result = [None] * k
    setsize = 21
    if k > 5:
        setsize += 4 ** int(ceil(log(k * 3, 4)))
    if n <= setsize or hasattr(population, "keys"):
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
