"""Recipe for ``oracle/_ref/``: the reference's own smCounter.py made runnable here (TEST INFRASTRUCTURE ONLY).

    python -m oracle.ref_build            # needs /root/reference; writes oracle/_ref/smCounter_ref.py + MANIFEST.json

The reference is a single Python-2.7 script.  The "build" reads it where it lies
(``/root/reference/smCounter.py``), applies the edits listed in ``EDITS`` -- both inside ``main()``, both the
Py2-only ``dict.iteritems`` spelled on a *plain* dict the shims cannot reach (smCounter.py:658, :667) -- and writes
the result to ``oracle/_ref/`` (git-ignored, like a compiled reference ``.so`` would be; it travels to the GPU box
with the snapshot, where /root/reference does not exist).  No reference source is committed.  Every function on
the hot path (calProb, isHPorLowComp, filterVariants, vc, vc_wrapper) is byte-for-byte the reference's: the
Python-2 behaviour they rely on is supplied from outside by ``oracle/ref_shims.py`` when the file is loaded.

``load(order=...)`` executes ``oracle/_ref/smCounter_ref.py`` with
  * ``pysam`` -> ref_shims (pileup / FastaFile over the oracle's record model),
  * ``defaultdict`` / ``set`` -> CPython-2.7 hash-table models (order="py2") or plain ones (order="native"),
  * ``round`` / ``str`` / ``random`` / ``subprocess`` (bedtools) / optionally ``multiprocessing`` -> ref_shims,
and returns the module: ``mod.vc(bam_path, chrom, pos_str, ...)`` then runs the reference's code.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SMC_REFERENCE_SRC", "/root/reference/smCounter.py")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_PY = os.path.join(OUT_DIR, "smCounter_ref.py")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")

# (line number in the reference, old text, new text): the complete list of source edits
EDITS = (
    (658, "in args.iteritems():", "in args.items():"),
    (667, "in vars(args).iteritems():", "in vars(args).items():"),
)
# lines of the hot path that must come through untouched (checked at build time)
HOT_RANGES = ((26, 98), (103, 117), (122, 177), (182, 269), (274, 600), (605, 611))


def build(force=False):
    """Returns the path of the generated module, or None when the reference tree is not present (GPU box)."""
    if not os.path.exists(REF_SRC):
        return OUT_PY if os.path.exists(OUT_PY) else None
    with open(REF_SRC, "rb") as fh:
        raw = fh.read()
    sha = hashlib.sha256(raw).hexdigest()
    if not force and os.path.exists(OUT_PY) and os.path.exists(MANIFEST):
        try:
            with open(MANIFEST) as fh:
                if json.load(fh).get("source_sha256") == sha:
                    return OUT_PY
        except (OSError, ValueError):
            pass
    lines = raw.decode("utf-8").split("\n")
    for (ln, old, new) in EDITS:
        assert not any(lo <= ln <= hi for (lo, hi) in HOT_RANGES), "edits must stay outside the hot path"
        assert lines[ln - 1].count(old) == 1, "reference line %d is not what the recipe expects: %r" % (ln, lines[ln - 1])
        lines[ln - 1] = lines[ln - 1].replace(old, new)
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(OUT_PY, "w") as fh:
        fh.write("\n".join(lines))
    with open(MANIFEST, "w") as fh:
        json.dump({"source": REF_SRC, "source_sha256": sha, "edits": [list(e) for e in EDITS],
                   "hot_path_lines_unmodified": [list(r) for r in HOT_RANGES]}, fh, indent=1)
    return OUT_PY


def available():
    return os.path.exists(OUT_PY) or os.path.exists(REF_SRC)


_LOADED = {}


def load(order="py2", inline_pool=True):
    """Execute the generated module with the Python-2 environment injected.  order: "py2" | "native"."""
    key = (order, inline_pool)
    if key in _LOADED:
        return _LOADED[key]
    from . import ref_shims as sh
    path = build()
    if path is None:
        raise FileNotFoundError("oracle/_ref is not built and %s is absent" % REF_SRC)
    with open(path) as fh:
        src = fh.read()
    name = "oracle._ref.smCounter_ref_%s%s" % (order, "_inline" if inline_pool else "")
    mod = types.ModuleType(name)
    mod.__file__ = path
    g = mod.__dict__
    # names the script resolves as builtins / globals
    g["round"] = sh.py2_round
    g["str"] = sh.py2_str
    if order == "py2":
        g["set"] = sh.Py2Set
    saved = {k: sys.modules.get(k) for k in ("pysam",)}
    sys.modules["pysam"] = sh.make_pysam_module()
    try:
        exec(compile(src, path, "exec"), g)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # names the script imported itself: rebind after the imports ran
    g["defaultdict"] = sh.Py2DefaultDict if order == "py2" else sh.NativeDefaultDict
    g["random"] = sh.Py2Random()
    g["subprocess"] = sh.SubprocessShim
    if inline_pool:
        g["multiprocessing"] = sh.InlineMultiprocessing
    sys.modules[name] = mod            # vc_wrapper must be picklable by reference for a real (forked) Pool
    _LOADED[key] = mod
    return mod


if __name__ == "__main__":
    p = build(force=True)
    print(p if p else "reference source not found at " + REF_SRC)
