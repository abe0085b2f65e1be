"""The C-ABI library (include/smc_b200.h): it loads, exports every declared symbol, the ctypes mirrors have the layout the
C compiler gives the header's structs, and -- on a box without a GPU -- it fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smc_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smc_[a-z_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from smcounter_b200 import build, _ffi
    build.build()
    return _ffi.load()


def test_exports_every_declared_symbol(lib):
    from smcounter_b200 import _ffi
    names = _declared_functions()
    assert set(names) == set(_ffi.EXPORTS), (names, _ffi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.smc_version() == 3
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(r"\bT %s\b" % n, out), n


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof/offsetof from gcc on the real header vs the ctypes mirrors in _ffi.py."""
    from smcounter_b200 import _ffi
    structs = {"smc_params": _ffi.smc_params, "smc_reads_soa": _ffi.smc_reads_soa, "smc_loci": _ffi.smc_loci,
               "smc_umi_keep": _ffi.smc_umi_keep, "smc_out": _ffi.smc_out, "smc_timings": _ffi.smc_timings}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smc_b200.h"', 'int main(void){']
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for f, _ in cls._fields_:
            assert int(got["%s.%s" % (name, f)]) == getattr(cls, f).offset, (name, f)
    # constants mirrored in Python
    hdr = open(HEADER).read()
    for cname, val in (("SMC_NFIXED", _ffi.SMC_NFIXED), ("SMC_NCNT", _ffi.SMC_NCNT), ("SMC_NLOC", _ffi.SMC_NLOC),
                       ("SMC_C_STRONG", _ffi.C_STRONG), ("SMC_L_STATUS", _ffi.L_STATUS), ("SMC_A_G", _ffi.A_G)):
        assert int(re.search(r"#define\s+%s\s+(\d+)" % cname, hdr).group(1)) == val, cname


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created and the Python binding raises; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is exercised on the CPU box")
    from smcounter_b200 import _ffi
    from smcounter_b200.caller import GpuCaller, VcParams
    p = _ffi.smc_params(20, 30, 100, 0, 0, 2, 4.0, 6.0)
    h = C.c_void_p()
    rc = lib.smc_ctx_create(0, C.byref(p), C.byref(h))
    assert rc == -1 and not h.value
    assert b"no CPU fallback" in lib.smc_last_error(None)
    with pytest.raises(RuntimeError, match="smc_ctx_create failed"):
        GpuCaller(VcParams(mtDepth=100, rpb=4.0))
    assert lib.smc_ctx_create(0, None, C.byref(h)) == -2


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    from smcounter_b200 import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setenv("SMC_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _ffi.load()


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under smcounter_b200/ may import it."""
    pkg = os.path.join(ROOT, "smcounter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    code = "import sys; import smcounter_b200.smCounter, smcounter_b200.caller, smcounter_b200.rows; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_bamio_library_exports_and_layout(tmp_path):
    """include/smc_bamio.h (host-side BAM decoder): every declared symbol is exported, struct layout matches ctypes."""
    from smcounter_b200 import _bamio, build
    build.build_bamio()
    lib = _bamio.load()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "smc_bamio.h")).read(), flags=re.S)
    src += re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "smc_rows.h")).read(), flags=re.S)      # same library
    src += re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "smc_soa.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(smc_(?:bam|rows|soa)_[a-z_]+)\s*\(", src)))
    assert set(names) == set(_bamio.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    cls = _bamio.smc_bam_reads
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smc_bamio.h"', 'int main(void){',
             'printf("size %zu\\n", sizeof(smc_bam_reads));']
    for f, _ in cls._fields_:
        lines.append('printf("%s %%zu\\n", offsetof(smc_bam_reads, %s));' % (f, f))
    lines.append('return 0;}')
    (tmp_path / "l.c").write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "l"), str(tmp_path / "l.c")])
    got = dict(l.split() for l in subprocess.check_output([str(tmp_path / "l")], text=True).splitlines())
    assert int(got["size"]) == C.sizeof(cls)
    for f, _ in cls._fields_:
        assert int(got[f]) == getattr(cls, f).offset, f
    for cname, cls2 in (("smc_rows_in", _bamio.smc_rows_in), ("smc_rows_out", _bamio.smc_rows_out),                 # include/smc_rows.h
                        ("smc_soa_view", _bamio.smc_soa_view), ("smc_soa_pack_opts", _bamio.smc_soa_pack_opts),     # include/smc_soa.h
                        ("smc_soa_pack_sizes", _bamio.smc_soa_pack_sizes), ("smc_soa_pack_bufs", _bamio.smc_soa_pack_bufs),
                        ("smc_soa_pack_exc", _bamio.smc_soa_pack_exc)):
        lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smc_rows.h"', '#include "smc_soa.h"', 'int main(void){',
                 'printf("size %%zu\\n", sizeof(%s));' % cname]
        for f, _ in cls2._fields_:
            lines.append('printf("%s %%zu\\n", offsetof(%s, %s));' % (f, cname, f))
        lines.append('return 0;}')
        (tmp_path / "r.c").write_text("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "r"), str(tmp_path / "r.c")])
        got = dict(l.split() for l in subprocess.check_output([str(tmp_path / "r")], text=True).splitlines())
        assert int(got["size"]) == C.sizeof(cls2), cname
        for f, _ in cls2._fields_:
            assert int(got[f]) == getattr(cls2, f).offset, (cname, f)
    h = C.c_void_p()
    assert lib.smc_bam_open(str(tmp_path / "missing.bam").encode(), 1, C.byref(h)) != 0
    assert b"cannot open" in lib.smc_bam_last_error(None)
