"""Import shim for `easydict` (absent in this image); the reference wraps the
YAML config in EasyDict (pipeline/parameters.py:18-34) and reads it with both
attribute and item access plus `.get` (network/encoder/encoder.py:14-22)."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {}, **kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setattr__(self, name, value):
        value = self._wrap(value)
        super().__setitem__(name, value)
        super().__setattr__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {}, **f)
        for k, v in d.items():
            setattr(self, k, v)

    def pop(self, k, *a):
        if hasattr(self, k):
            try:
                delattr(self, k)
            except AttributeError:
                pass
        return super().pop(k, *a)
