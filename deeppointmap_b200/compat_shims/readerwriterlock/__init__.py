"""Shim for `readerwriterlock` 1.0.9 (absent in this image): the reference takes
`rwlock.RWLockFair()` and uses `with self.locker.gen_rlock():` / `gen_wlock()`
(system/modules/pose_graph.py:9, 171 ff.).  A fair readers-writer lock over one condition variable."""
from . import rwlock  # noqa: F401
