class _Anything:
    """absorbs any attribute access / call / iteration the plotting code does"""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def savefig(fname, *a, **k):
    try:
        with open(fname, "w") as f:
            f.write("matplotlib is not installed in this image: the figure was not drawn\n")
    except (OSError, TypeError):
        pass


def subplots(*a, **k):
    return _Anything(), _Anything()


def __getattr__(name):  # figure, plot, close, axis, legend, ...
    return _Anything()
