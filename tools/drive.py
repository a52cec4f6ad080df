#!/usr/bin/env python
"""Runs one of the reference's launch scripts UNCHANGED against the drop-in modules of this
repository.

    python tools/drive.py /path/to/reference/run_pypic_dd.py [--seed 1]

`python run_pypic.py` would put the script's own directory first on sys.path and import the
reference's pypic / PIC_L_DD / convert; runpy.run_path does not touch sys.path, so with this
repository's root inserted first the launcher's `import pypic as p`, `import PIC_L_DD as p`
and `import convert as c` bind to the B200-backed modules.  The reference never seeds the
global NumPy stream in these drivers; --seed does it before the launcher starts so that runs
are reproducible (and comparable with a seeded reference run).  The launchers end with a
PNG->GIF step through imageio; when imageio (or matplotlib, which writes the PNGs) is not
installed that step is skipped with a note -- the numerical outputs are already on disk.

--set NAME=VALUE (repeatable) overrides a hard-coded literal of the driver function the launcher
calls (the drop-in's main / main_i take the reference's literals as keyword defaults), e.g.
    python tools/drive.py /path/to/run_pypic_dd.py --seed 1 --set N=20000000 --set Ng=4097 --steps 200
runs the unchanged launcher on a 2e7-particle sheath; --steps replaces the launcher's own step count
(its `stop` literal) the same way.  A timing line (particle-steps/s) is printed at the end.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("launcher")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--quiet", action="store_true", help="drop the per-step prints of the driver")
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    if a.seed is not None:
        import numpy as np
        np.random.seed(a.seed)
    os.makedirs("plots", exist_ok=True)
    timing = {}
    if a.set or a.steps is not None or a.quiet:
        import ast
        import contextlib
        import functools
        import io
        import time
        kw = {}
        for item in a.set:
            name, val = item.split("=", 1)
            kw[name] = ast.literal_eval(val)
        for modname, fn in (("PIC_L_DD", "main_i"), ("pypic", "main"), ("PIC_L", "main")):
            if os.path.isfile(os.path.join(ROOT, modname + ".py")) and modname.lower() in open(a.launcher).read().lower():
                mod = __import__(modname)
                orig = getattr(mod, fn)

                def patched(T, nplot, _orig=orig, **more):
                    if a.steps is not None:
                        T = a.steps
                    t0 = time.perf_counter()
                    with (contextlib.redirect_stdout(io.StringIO()) if a.quiet else contextlib.nullcontext()):
                        out = _orig(T, nplot, **dict(kw, **more))
                    timing.update(seconds=time.perf_counter() - t0, steps=T + 1, N=kw.get("N"))
                    return out
                setattr(mod, fn, functools.wraps(orig)(patched))
    try:
        runpy.run_path(a.launcher, run_name="__main__")
    except (ImportError, FileNotFoundError) as ex:
        if "imageio" in str(ex) or ".png" in str(ex):
            print("drive.py: GIF assembly skipped (%s)" % ex)
        else:
            raise
    if timing:
        n = timing["N"]
        print("drive.py: %d steps in %.3f s" % (timing["steps"], timing["seconds"]) +
              (" = %.3e particle-steps/s (whole driver: initialisation, upload, steps, outputs)"
               % (n * timing["steps"] / timing["seconds"]) if n else ""))


if __name__ == "__main__":
    main()
