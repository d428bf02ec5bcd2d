"""webaudio-modem_b200 — B200-native FSK physical layer behind the reference's FSKCore API.

Import with importlib (the directory name carries a hyphen):
    wam = importlib.import_module("webaudio-modem_b200")
"""
import os as _os

# The fast path checks its doubtful decisions on a dozen CUDA streams at once (window classes x configuration groups).
# With the default of 8 hardware work queues several of those streams share a queue and a chain that waits for its own
# kernel holds up the chain behind it (measured: +0.5 ms per call whenever an end-of-data check runs).  Takes effect
# when set before the process creates its CUDA context, hence here, at import.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402
from ._lib import WamError, lib  # noqa: F401,E402
from .fsk import DEFAULT_FSK_CONFIG, ChunkedModulator, FSKBatch, FSKCore, FSKSessionMux, normalize_config  # noqa: F401
from .filters import FilterDesign, FilterFactory, FIRFilter, IIRFilter  # noqa: F401
from .xmodem import CRC16, XModemPacket, XModemBatchReceiver, xmodem_batch_check, crc16_batch, PKT_STATUS  # noqa: F401
