import threading


class _Guard:
    def __init__(self, acquire, release):
        self._a, self._r = acquire, release
        self._held = False

    def acquire(self, blocking=True, timeout=-1):
        self._a()
        self._held = True
        return True

    def release(self):
        self._held = False
        self._r()

    def locked(self):
        return self._held

    def __enter__(self):
        self.acquire()
        return self

    def __exit__(self, *a):
        self.release()
        return False


class RWLockFair:
    """readers share, writers are exclusive, arrival order decides (tickets), re-entrant per thread for readers
    inside a writer section (the pose graph calls read helpers while it holds the write lock)."""

    def __init__(self):
        self._cv = threading.Condition()
        self._readers = 0
        self._writer = None
        self._wdepth = 0
        self._next, self._serving = 0, 0

    def _r_acquire(self):
        me = threading.get_ident()
        with self._cv:
            if self._writer == me:
                self._wdepth += 1
                return
            t = self._next
            self._next += 1
            while self._serving != t or self._writer is not None:
                self._cv.wait()
            self._readers += 1
            self._serving += 1
            self._cv.notify_all()

    def _r_release(self):
        me = threading.get_ident()
        with self._cv:
            if self._writer == me:
                self._wdepth -= 1
                return
            self._readers -= 1
            self._cv.notify_all()

    def _w_acquire(self):
        me = threading.get_ident()
        with self._cv:
            if self._writer == me:
                self._wdepth += 1
                return
            t = self._next
            self._next += 1
            while self._serving != t or self._writer is not None or self._readers > 0:
                self._cv.wait()
            self._writer, self._wdepth = me, 1
            self._serving += 1

    def _w_release(self):
        with self._cv:
            self._wdepth -= 1
            if self._wdepth == 0:
                self._writer = None
                self._cv.notify_all()

    def gen_rlock(self):
        return _Guard(self._r_acquire, self._r_release)

    def gen_wlock(self):
        return _Guard(self._w_acquire, self._w_release)


RWLockFairD = RWLockRead = RWLockWrite = RWLockFair
