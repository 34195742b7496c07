"""gims_b200 — B200-native (sm_100a) implementation of the GIMS matcher forward path.

Public API mirrors the reference: `Matching(config)(data)` / `GMatcher(config)(data)`
(models/matching.py, models/gmatcher.py of songxf1024/GIMS).  All arithmetic runs in
libgims_b200.so (hand-written CUDA, C ABI in include/gims_b200.h).
"""
import os as _os

# A forward enqueues ~125 dependent launches per pair on the caller's stream plus a forked side stream; callers that keep
# several pairs in flight use 8-24 streams.  The CUDA default of 8 hardware work queues makes streams share queues (false
# dependencies): 32 gave +3.6 % pairs/s at 2048 keypoints with 12 streams.  The variable is read when the CUDA context is
# created, i.e. it takes effect if this package is imported before the first CUDA call; an explicit setting wins.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

from .config import DEFAULT_CONFIG  # noqa: F401,E402
from .gmatcher import GMatcher  # noqa: F401,E402
from .matching import Matching  # noqa: F401,E402
