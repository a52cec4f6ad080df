"""CPU-side checks of the C-ABI shared library: it loads, exports every symbol that
include/pic_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pypic_b200 import build, _lib
    build.build()
    return _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pic_[a-zA-Z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    l = lib.load()
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(l, s), "libpic_b200.so does not export %s" % s
    # and the binding table covers the header
    assert set(syms) == set(lib.EXPORTS)


def test_binding_table_matches_the_prototypes(lib):
    """Every ctypes signature in pypic_b200/_lib.py has as many arguments as the prototype in
    include/pic_b200.h (a wrong count only shows up as a TypeError on the GPU box otherwise)."""
    src = open(os.path.join(ROOT, "include", "pic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|const char\*|void)\s+(pic_[a-zA-Z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) > 40
    for name, args in protos:
        if name in ("pic_last_error", "pic_version"):
            continue
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert len(lib._SIGS[name]) == n, "%s: header has %d arguments, binding table %d" % (name, n, len(lib._SIGS[name]))


def test_struct_layouts_match_the_header(lib, tmp_path):
    """The ctypes mirrors of the parameter structs (pypic_b200/_lib.py) against the C compiler's view of
    include/pic_b200.h: same size, same offset for every field (a mismatch would silently shift every
    argument behind it)."""
    import ctypes as C
    import subprocess
    pairs = [("pic_dd_params", lib.DDParams), ("pic_dd_prologue", lib.DDPrologue), ("pic_pypic_params", lib.PypicParams),
             ("pic_l_params", lib.LParams), ("pic_gc_params", lib.GCParams)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pic_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_version_and_error_string(lib):
    l = lib.load()
    assert l.pic_version() >= 100
    assert isinstance(l.pic_last_error(), bytes)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pypic_b200 import device
    with pytest.raises(lib.PicError):
        device.require_cuda()
    import numpy as np
    from pypic_b200 import ops
    with pytest.raises(lib.PicError):
        ops.smooth(np.zeros(8), 0)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (test infrastructure)."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "pypic_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                if re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(d, f)).read(), flags=re.M):
                    bad.append(f)
    for f in ("pypic.py", "PIC_L.py", "PIC_L_DD.py", "pygcpic.py", "convert.py"):
        p = os.path.join(ROOT, f)
        if os.path.isfile(p) and re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), flags=re.M):
            bad.append(f)
    assert not bad, bad
