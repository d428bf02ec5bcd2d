"""webaudio-modem_b200 — B200-native FSK physical layer behind the reference's FSKCore API.

Import with importlib (the directory name carries a hyphen):
    wam = importlib.import_module("webaudio-modem_b200")
"""
from . import _lib  # noqa: F401
from ._lib import WamError, lib  # noqa: F401
from .fsk import DEFAULT_FSK_CONFIG, ChunkedModulator, FSKBatch, FSKCore, FSKSessionMux, normalize_config  # noqa: F401
from .filters import FilterDesign, FilterFactory, FIRFilter, IIRFilter  # noqa: F401
from .xmodem import CRC16, XModemPacket, XModemBatchReceiver, xmodem_batch_check, crc16_batch, PKT_STATUS  # noqa: F401
