"""Timer mirror of the reference ``synchronize_timer`` (hy3dgen/shapegen/utils.py:38-86):
CUDA-event timing printed only when ``HY3DGEN_DEBUG=1``, same log wording."""
from __future__ import annotations

import logging
import os

import torch

logger = logging.getLogger('hy3dgen.shapgen')      # the reference's (misspelt) logger name, utils.py:22-35


class synchronize_timer:
    def __init__(self, name=None):
        self.name = name

    def __enter__(self):
        if os.environ.get('HY3DGEN_DEBUG', '0') == '1' and torch.cuda.is_available():
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()
            return lambda: self.time

    def __exit__(self, exc_type, exc_value, exc_tb):
        if os.environ.get('HY3DGEN_DEBUG', '0') == '1' and torch.cuda.is_available():
            self.end.record()
            torch.cuda.synchronize()
            self.time = self.start.elapsed_time(self.end)
            if self.name is not None:
                logger.info(f'{self.name} takes {self.time} ms')
