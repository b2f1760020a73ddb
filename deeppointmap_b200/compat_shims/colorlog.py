"""Import shim for `colorlog` (absent in this image): the reference does
`import colorlog as logging` (network/encoder/utils.py:7, system/modules/*.py).
Everything it touches exists on the stdlib `logging` module."""
from logging import *  # noqa: F401,F403
import logging as _logging


class ColoredFormatter(_logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, style='%', log_colors=None, reset=True,
                 secondary_log_colors=None, **kw):
        if fmt is not None:
            for tok in ('%(log_color)s', '%(reset)s', '%(bold)s'):
                fmt = fmt.replace(tok, '')
        super().__init__(fmt, datefmt, style)


getLogger = _logging.getLogger
basicConfig = _logging.basicConfig
StreamHandler = _logging.StreamHandler
