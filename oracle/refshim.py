"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified reference sources* from
/root/reference through mechanical import shims.

Nothing in the product path (pypic_b200/, the drop-in modules, bench.py's GPU
arm) may import this file.  It exists so that tests and the golden-vector
generator (oracle/make_golden.py) can execute the reference's own arithmetic
in the authoring container.  /root/reference does not exist on the GPU box,
so every user must call ``available()`` first and skip when it is False.

Shims applied (none changes arithmetic):
  * MagicMock stand-ins for matplotlib / imageio / vpython (absent here);
  * ``scipy.diag`` aliased to ``numpy.diag`` (removed from SciPy);
  * ``np.trapz`` aliased to ``np.trapezoid`` when missing (NumPy >= 2.0 keeps it,
    guarded anyway);
  * pypic.py: the decorator on line 216 (``particle_push_p``) is commented out
    because numba 0.65's parfor pass crashes on it; the six other JIT kernels
    compile as written, so gather/deposit are the reference's real numerics;
  * PIC_L.py / PIC_L_DD.py (Python 2): integer-division idioms used as
    indices/slices are rewritten ``N/2 -> N//2`` etc.
"""
import os
import re
import sys
import types
from unittest import mock

REF_DIR = os.environ.get("PYPIC_REFERENCE_DIR", "/root/reference")
_cache = {}


def available():
    return os.path.isfile(os.path.join(REF_DIR, "pypic.py"))


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.colors", "imageio", "vpython"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    import numpy as np
    import scipy
    if not hasattr(scipy, "diag"):
        scipy.diag = np.diag
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid


def _read(fname):
    with open(os.path.join(REF_DIR, fname), "r") as f:
        return f.read()


def _exec_module(name, src, fname):
    mod = types.ModuleType("ref_" + name)
    mod.__file__ = os.path.join(REF_DIR, fname)
    code = compile(src, mod.__file__, "exec")
    exec(code, mod.__dict__)
    return mod


def _py2_fix(src):
    src = re.sub(r"\bN/2\b", "N//2", src)
    src = re.sub(r"\bNg/2\b", "Ng//2", src)
    src = src.replace("(Ng)/2", "(Ng)//2").replace("(Ng+1)/2", "(Ng+1)//2")
    src = src.replace("N*2/6", "N*2//6").replace("N*4/6", "N*4//6")
    # the __main__ guard calls main()/main_i() without arguments
    return src


def load(name):
    """name in {'pypic','PIC_L','PIC_L_DD','pygcpic'} -> module object."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_DIR)
    _install_stubs()
    if name == "pypic":
        lines = _read("pypic.py").split("\n")
        assert lines[215].startswith("@nb.jit(nb.types.UniTuple"), lines[215][:40]
        lines[215] = "#" + lines[215]
        mod = _exec_module(name, "\n".join(lines), "pypic.py")
    elif name in ("PIC_L", "PIC_L_DD"):
        mod = _exec_module(name, _py2_fix(_read(name + ".py")), name + ".py")
    elif name == "pygcpic":
        if "convert" not in sys.modules:
            sys.modules["convert"] = mock.MagicMock(name="convert")
        mod = _exec_module(name, _read("pygcpic.py"), "pygcpic.py")
    else:
        raise KeyError(name)
    _cache[name] = mod
    return mod
