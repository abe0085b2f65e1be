"""Stand-ins that let the reference's OWN code (/root/reference/smCounter.py) execute under Python 3
(TEST INFRASTRUCTURE ONLY -- part of ``oracle/``; nothing in the product package may import it).

``oracle/ref_build.py`` copies the reference script into ``oracle/_ref/`` (git-ignored build output) with two
one-token edits in ``main()``; every function on the hot path -- ``calProb`` (smCounter.py:26-98),
``isHPorLowComp`` (:122-177), ``filterVariants`` (:182-269), ``vc`` (:274-600), ``vc_wrapper`` (:605-611) --
runs *unmodified*.  What the script needs from its environment is supplied here:

  ``pysam``       AlignmentFile.pileup / PileupColumn / PileupRead / AlignedSegment / FastaFile over the
                  oracle's record model (``smcounter_oracle.Read``, ``pileup_column``).  Third party, absent
                  from the image: which reads enter a column and their ``indel / is_del / query_position``
                  are the oracle's restatement of htslib ``resolve_cigar2`` -> still "parity unpinned" for
                  that one layer; everything *above* it is the reference's code.
  containers      two flavours, chosen at load time:
                    order="py2"    ``defaultdict`` / ``set`` are order-faithful models of the CPython-2.7
                                   hash tables (py2compat.Py2Dict), so every iteration the reference makes --
                                   ``oneBC.values()`` (:62), ``prodP.keys()`` (:83), ``bcDict.keys()`` (:498),
                                   ``finalDict.items()`` (:534) -- visits the keys in the order Python 2.7
                                   would.  This is the flavour the parity tests diff the oracle against.
                    order="native" plain insertion-ordered dict/set with the Py2-only methods added
                                   (``iterkeys``, list-returning ``values``): fast, used for timing.
  ``round/str``   Python-2 ``round`` (half away from zero) and ``str(float)`` (``%.12g``), :576-599.
  ``random``      ``random.seed(str)`` + ``random.sample`` of CPython 2.7 (:497-498).
  ``subprocess``  the four ``bedtools`` pipelines of main() (:700-710) interpreted over
                  ``smcounter_oracle.bed_merge / bed_sort / bed_intersect`` (bedtools 2.25 is absent).
  ``multiprocessing``  optional in-process Pool (order="py2" keeps registries in this process).
"""
from __future__ import annotations

import collections
import shlex
import types

from . import py2compat
from . import smcounter_oracle as orc

# ------------------------------------------------------------------------------------------------------------
# registries: the reference opens files by path; tests register in-memory objects under made-up paths
# ------------------------------------------------------------------------------------------------------------
_BAMS = {}
_FASTAS = {}


def register_bam(path, reads_or_index):
    _BAMS[path] = reads_or_index if isinstance(reads_or_index, orc.ReadIndex) else orc.ReadIndex(reads_or_index)
    return path


def register_fasta(path, refs):
    _FASTAS[path] = refs
    return path


def clear_registries():
    _BAMS.clear()
    _FASTAS.clear()


# ------------------------------------------------------------------------------------------------------------
# pysam
# ------------------------------------------------------------------------------------------------------------
class AlignedSegment:
    """The attributes smCounter.py:319-365,372-447 reads from ``pileupRead.alignment``."""
    __slots__ = ("query_name", "mapping_quality", "tags", "cigar", "query_length", "is_read1", "is_read2", "is_reverse",
                 "query_sequence", "query_qualities", "query_alignment_length")

    def __init__(self, r):
        self.query_name = r.qname
        self.mapping_quality = r.mapq
        self.tags = [("NM", r.nm)] if r.nm is not None else []
        self.cigar = list(r.cigar)
        self.query_length = len(r.seq)
        self.is_read1 = bool(r.flag & 0x40)
        self.is_read2 = bool(r.flag & 0x80)
        self.is_reverse = bool(r.flag & 0x10)
        self.query_sequence = r.seq
        self.query_qualities = r.qual
        self.query_alignment_length = len(r.seq) - orc.soft_clip_total(r.cigar)


class PileupRead:
    __slots__ = ("alignment", "indel", "is_del", "query_position")

    def __init__(self, alignment, qpos, indel, is_del):
        self.alignment = alignment
        self.indel = indel
        self.is_del = 1 if is_del else 0
        self.query_position = None if is_del else qpos


class PileupColumn:
    def __init__(self, pileups, pos0):
        self.pileups = pileups
        self.reference_pos = pos0
        self.nsegments = len(pileups)


class AlignmentFile:
    def __init__(self, path, mode="rb"):
        if path not in _BAMS:
            raise IOError("pysam shim: no reads registered under %r (oracle.ref_shims.register_bam)" % (path,))
        self.index = _BAMS[path]
        # decoded-record views live with the registered reads (real pysam decodes per fetch in C; keeping them keeps the
        # shim's Python overhead out of the reference's measured time)
        if not hasattr(self.index, "_seg_cache"):
            self.index._seg_cache = {}
        self._seg = self.index._seg_cache

    def pileup(self, region=None, truncate=False, max_depth=8000, stepper="all", **kw):
        # smCounter.py:316 passes 'chrom:pos:pos'; the pysam of that era splits a region on [:-]
        chrom, start, end = region.rsplit(":", 2)
        assert truncate and stepper == "nofilter", "the shim restates only the call smCounter.py:316 makes"
        for p in range(int(start) - 1, int(end)):
            col = []
            for r in self.index.column(chrom, p):
                qpos, indel, is_del = orc.pileup_column(r, p)
                seg = self._seg.get(id(r))
                if seg is None:
                    seg = self._seg[id(r)] = AlignedSegment(r)
                col.append(PileupRead(seg, qpos, indel, is_del))
            if col:                      # htslib emits no column where no read is piled up
                yield PileupColumn(col, p)

    def close(self):
        pass


class FastaFile:
    def __init__(self, path):
        if path not in _FASTAS:
            raise IOError("pysam shim: no reference registered under %r (oracle.ref_shims.register_fasta)" % (path,))
        self.refs = _FASTAS[path]

    def fetch(self, reference=None, start=None, end=None):
        return self.refs.fetch(reference, start, end)

    def get_reference_length(self, reference):
        return self.refs.get_reference_length(reference)


def make_pysam_module():
    m = types.ModuleType("pysam")
    m.AlignmentFile = AlignmentFile
    m.FastaFile = FastaFile
    m.__doc__ = "oracle.ref_shims stand-in for pysam (absent from the image)"
    return m


# ------------------------------------------------------------------------------------------------------------
# containers, order="py2": CPython-2.7 hash-table models
# ------------------------------------------------------------------------------------------------------------
class Py2DefaultDict(py2compat.Py2Dict):
    """collections.defaultdict of CPython 2.7 for str keys: iteration in slot order, __missing__ inserts."""

    def __init__(self, default_factory=None):
        super().__init__()
        self.default_factory = default_factory

    def __getitem__(self, key):
        h = py2compat.py2hash(key)
        _, ep = self._lookup(key, h)
        if ep is not None:
            return ep[2]
        if self.default_factory is None:
            raise KeyError(key)
        v = self.default_factory()
        self[key] = v
        return v

    def __iter__(self):
        return iter(self.keys())

    iterkeys = py2compat.Py2Dict.keys
    itervalues = py2compat.Py2Dict.values
    iteritems = py2compat.Py2Dict.items

    def get(self, key, default=None):
        _, ep = self._lookup(key, py2compat.py2hash(key))
        return default if ep is None else ep[2]


class Py2Set:
    """CPython-2.7 ``set`` of str (Objects/setobject.c): same probing and growth rule as the dict."""

    def __init__(self, iterable=()):
        self._d = py2compat.Py2Dict()
        for k in iterable:
            self._d[k] = True

    def add(self, k):
        self._d[k] = True

    def __contains__(self, k):
        return k in self._d

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(self._d.keys())

    def __sub__(self, other):
        # set_difference: a new set filled by iterating self in slot order
        return Py2Set(k for k in self._d.keys() if k not in other)


# ------------------------------------------------------------------------------------------------------------
# containers, order="native": plain dict/set + the Python-2-only methods the script calls
# ------------------------------------------------------------------------------------------------------------
class NativeDefaultDict(collections.defaultdict):
    def iterkeys(self):
        return iter(list(self))

    def values(self):                     # smCounter.py:522 indexes .values()
        return list(collections.defaultdict.values(self))


# ------------------------------------------------------------------------------------------------------------
# round / str / random
# ------------------------------------------------------------------------------------------------------------
def py2_round(x, ndigits=0):
    return py2compat.py2round(float(x), ndigits)


_builtin_str = str


def py2_str(v=""):
    """``str`` as the script uses it: floats render the Python-2 way (smCounter.py:599), the rest is unchanged."""
    return py2compat.py2str(v) if isinstance(v, float) else _builtin_str(v)


class Py2Random:
    """``random.seed(pos); random.sample(keys, ds)`` (smCounter.py:497-498) as CPython 2.7 draws it."""

    def __init__(self):
        self._seed = None

    def seed(self, s):
        self._seed = s

    def sample(self, population, k):
        assert isinstance(self._seed, _builtin_str), "smCounter seeds with the position string"
        return py2compat.py2_sample(self._seed, list(population), k)


# ------------------------------------------------------------------------------------------------------------
# subprocess: the bedtools pipelines of smCounter.py:700-710
# ------------------------------------------------------------------------------------------------------------
def _read_bed(path, stdin_rows):
    if path == "-":
        return stdin_rows
    rows = []
    with open(path) as fh:
        for line in fh:
            if line.strip() and not line.startswith(("track ", "browser ", "#")):
                rows.append(line.rstrip("\n").split("\t"))
    return rows


def _bedtools(argv, stdin_rows):
    assert argv[0].endswith("bedtools"), argv
    sub = argv[1]
    opts = {}
    i = 2
    while i < len(argv):
        assert argv[i].startswith("-"), argv
        opts[argv[i]] = argv[i + 1]
        i += 2
    if sub == "merge":
        rows = _read_bed(opts["-i"], stdin_rows)
        if "-c" in opts:
            assert opts["-c"] == "4" and opts["-o"] == "distinct"
            return [list(map(_builtin_str, r)) for r in orc.bed_merge(rows, distinct_col4=True)]
        return [list(map(_builtin_str, r)) for r in orc.bed_merge(rows)]
    if sub == "sort":
        return [list(map(_builtin_str, r)) for r in orc.bed_sort(_read_bed(opts["-i"], stdin_rows))]
    if sub == "intersect":
        return [list(map(_builtin_str, r)) for r in orc.bed_intersect(_read_bed(opts["-a"], stdin_rows), _read_bed(opts["-b"], stdin_rows))]
    raise NotImplementedError("bedtools " + sub)


class SubprocessShim:
    CalledProcessError = RuntimeError

    @staticmethod
    def check_call(cmd, shell=False):
        assert shell and isinstance(cmd, _builtin_str)
        cmd, _, out_path = cmd.partition(">")
        rows = None
        for stage in cmd.split("|"):
            rows = _bedtools(shlex.split(stage), rows)
        with open(out_path.strip(), "w") as fh:
            for r in rows:
                fh.write("\t".join(r) + "\n")
        return 0


# ------------------------------------------------------------------------------------------------------------
# multiprocessing: in-process Pool (keeps the registries and the Py2 containers in one interpreter)
# ------------------------------------------------------------------------------------------------------------
class _Result:
    def __init__(self, v):
        self.v = v

    def get(self, timeout=None):
        return self.v


class InlinePool:
    def __init__(self, processes=None):
        pass

    def apply_async(self, fn, args=()):
        return _Result(fn(*args))

    def close(self):
        pass

    def join(self):
        pass


class InlineMultiprocessing:
    Pool = InlinePool
