"""Debug timing of the two plugin calls of ``latents2mesh``.

Same switch and log line as the reference's ``synchronize_timer`` (hy3dgen/shapegen/utils.py:38-86): with
``HY3DGEN_DEBUG=1`` the elapsed device time of the wrapped region is logged as ``"<name> takes <ms> ms"`` on the
``hy3dgen.shapgen`` logger (the reference's spelling).  Here the two events are recorded on the *current stream* and only
the closing event is waited for — the whole device is not synchronised — and the helper is a generator-based context
manager (usable as a decorator through ``contextlib``)."""
from __future__ import annotations

import contextlib
import logging
import os

import torch

_LOG = logging.getLogger("hy3dgen.shapgen")


def _debug_timing_enabled() -> bool:
    return os.environ.get("HY3DGEN_DEBUG", "0") == "1" and torch.cuda.is_available()


@contextlib.contextmanager
def synchronize_timer(name=None):
    """``with synchronize_timer("Volume decoding") as elapsed: ...`` — ``elapsed()`` returns the milliseconds after the
    block has closed (``None`` when timing is off)."""
    if not _debug_timing_enabled():
        yield None
        return
    stream = torch.cuda.current_stream()
    opened, closed = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    result = {}
    opened.record(stream)
    try:
        yield lambda: result.get("ms")
    finally:
        closed.record(stream)
        closed.synchronize()
        result["ms"] = opened.elapsed_time(closed)
        if name is not None:
            _LOG.info("%s takes %s ms", name, result["ms"])
