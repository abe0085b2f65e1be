"""Optional run log (reference run_log.py:26-45): when --logFile is given, stdout/stderr are tee'd into
``<logFile>.run-log_<timestamp>.txt``.  Cosmetic; not part of the calling path."""
from __future__ import annotations

import datetime
import sys


class _Tee:
    def __init__(self, stream, fh):
        self.stream, self.fh = stream, fh

    def write(self, s):
        self.stream.write(s)
        if s.strip():
            self.fh.write("%s %s\n" % (datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S,%f")[:-3], s.rstrip("\n")))
            self.fh.flush()

    def flush(self):
        self.stream.flush()


def init(logFilePrefix):
    if not logFilePrefix:
        return None
    path = "%s.run-log_%s.txt" % (logFilePrefix, datetime.datetime.now().strftime("%Y.%m.%d_%H.%M.%S"))
    fh = open(path, "w")
    sys.stdout = _Tee(sys.stdout, fh)
    sys.stderr = _Tee(sys.stderr, fh)
    return path
