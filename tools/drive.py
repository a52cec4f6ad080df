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
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    if a.seed is not None:
        import numpy as np
        np.random.seed(a.seed)
    os.makedirs("plots", exist_ok=True)
    try:
        runpy.run_path(a.launcher, run_name="__main__")
    except (ImportError, FileNotFoundError) as ex:
        if "imageio" in str(ex) or ".png" in str(ex):
            print("drive.py: GIF assembly skipped (%s)" % ex)
        else:
            raise


if __name__ == "__main__":
    main()
