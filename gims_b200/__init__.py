"""gims_b200 — B200-native (sm_100a) implementation of the GIMS matcher forward path.

Public API mirrors the reference: `Matching(config)(data)` / `GMatcher(config)(data)`
(models/matching.py, models/gmatcher.py of songxf1024/GIMS).  All arithmetic runs in
libgims_b200.so (hand-written CUDA, C ABI in include/gims_b200.h).
"""
from .config import DEFAULT_CONFIG  # noqa: F401
from .gmatcher import GMatcher  # noqa: F401
from .matching import Matching  # noqa: F401
