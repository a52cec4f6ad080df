"""Optional matplotlib access for the drop-in drivers.  The reference imports
matplotlib unconditionally and plots inside its time loops; the GPU box has no
matplotlib, so plotting is skipped there (numerical outputs are unaffected)."""


def get_plt():
    try:
        import matplotlib
        import matplotlib.pyplot as plt
        return matplotlib, plt
    except Exception:
        return None, None
