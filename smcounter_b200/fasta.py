"""Reference-genome access for the host side (stands in for pysam.FastaFile, which the reference uses at
smCounter.py:124-129, 311-312, 394).  ``fetch(chrom, start, end)`` is 0-based half-open and clips ``end`` at the
contig length; both classes expose the two methods the caller needs and nothing else.
"""
from __future__ import annotations

import os

import numpy as np


class FastaFile:
    """Indexed FASTA reader (.fai is built in memory when missing).  Uncompressed FASTA only."""

    def __init__(self, path: str):
        self.path = path
        self.index = {}
        fai = path + ".fai"
        if os.path.exists(fai):
            with open(fai) as fh:
                for line in fh:
                    name, length, offset, linebases, linewidth = line.rstrip("\n").split("\t")[:5]
                    self.index[name] = (int(length), int(offset), int(linebases), int(linewidth))
        else:
            self._build_index()
        self.fh = open(path, "rb")

    def _build_index(self):
        with open(self.path, "rb") as fh:
            name = None
            length = offset = linebases = linewidth = 0
            pos = 0
            for line in fh:
                if line.startswith(b">"):
                    if name is not None:
                        self.index[name] = (length, offset, linebases, linewidth)
                    name = line[1:].split()[0].decode()
                    length = 0
                    offset = pos + len(line)
                    linebases = linewidth = 0
                else:
                    if linewidth == 0:
                        linewidth = len(line)
                        linebases = len(line.rstrip(b"\r\n"))
                    length += len(line.rstrip(b"\r\n"))
                pos += len(line)
            if name is not None:
                self.index[name] = (length, offset, linebases, linewidth)

    def get_reference_length(self, chrom: str) -> int:
        return self.index[chrom][0]

    def fetch(self, chrom: str, start: int, end: int) -> str:
        length, offset, linebases, linewidth = self.index[chrom]
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        end = min(end, length)
        if end <= start:
            return ""
        b0 = offset + (start // linebases) * linewidth + start % linebases
        b1 = offset + ((end - 1) // linebases) * linewidth + (end - 1) % linebases + 1
        raw = os.pread(self.fh.fileno(), b1 - b0, b0)        # positional read: no shared file offset (threads, forked workers)
        return raw.replace(b"\n", b"").replace(b"\r", b"").decode()

    def close(self):
        self.fh.close()


class SparseRef:
    """In-memory reference made of windows (synthetic panels): anything outside a window reads as 'N'."""

    def __init__(self, lengths: dict):
        self.lengths = dict(lengths)
        self.windows = {c: [] for c in lengths}    # chrom -> sorted list of (start, uint8 ASCII array)

    def add_window(self, chrom: str, start: int, bases: np.ndarray):
        self.windows.setdefault(chrom, []).append((int(start), np.asarray(bases, dtype=np.uint8)))
        self.windows[chrom].sort(key=lambda w: w[0])

    def get_reference_length(self, chrom: str) -> int:
        return self.lengths[chrom]

    def fetch_array(self, chrom: str, start: int, end: int) -> np.ndarray:
        end = min(end, self.lengths[chrom])
        out = np.full(max(0, end - start), ord("N"), dtype=np.uint8)
        for (ws, arr) in self.windows.get(chrom, ()):
            lo = max(start, ws)
            hi = min(end, ws + len(arr))
            if lo < hi:
                out[lo - start: hi - start] = arr[lo - ws: hi - ws]
        return out

    def fetch(self, chrom: str, start: int, end: int) -> str:
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        return self.fetch_array(chrom, start, end).tobytes().decode()
