"""Shim for `matplotlib` (absent in this image).  The reference only draws the trajectory picture at the end of a run
(system/modules/recoder.py `draw_trajectory`, system/modules/utils.py): every pyplot call becomes a no-op and
`savefig` writes a one-line placeholder so the run's file list stays the same."""
__version__ = "0.0-shim"


def use(*a, **k):
    pass
