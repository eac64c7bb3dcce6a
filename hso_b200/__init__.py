"""hso_b200 — B200-native (sm_100a CUDA) implementation of HSO's per-frame tracking hot path behind a C-ABI.

Product code only: nothing here imports oracle/. The CUDA library is loaded lazily by hso_b200._capi.load() and there is
no CPU fallback."""
from . import _capi  # noqa: F401
from .api import Context, CoarseTracker, HsoError, make_cam, PINHOLE, FOV, EQUIDISTANT  # noqa: F401
