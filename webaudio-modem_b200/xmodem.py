"""Host-side mirror of src/utils/crc16.ts and src/transports/xmodem/packet.ts, plus the batched
GPU frame check (wam_xmodem_batch_check: one warp per stream, warp-level CRC-16)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

PKT_STATUS = ["OK", "DUPLICATE", "NO_SOH", "EOT", "INCOMPLETE", "BAD_COMPLEMENT", "BAD_CRC", "UNEXPECTED_SEQ"]
_u8p = C.POINTER(C.c_uint8)


class CRC16:
    """src/utils/crc16.ts:11-50"""

    @staticmethod
    def calculate(data) -> int:
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(L.lib().wam_crc16(a.ctypes.data_as(_u8p) if len(a) else None, len(a)))

    @staticmethod
    def verify(data, expectedCrc: int) -> bool:
        return CRC16.calculate(data) == expectedCrc


class XModemPacket:
    """src/transports/xmodem/packet.ts:17-65"""

    SOH = 0x01

    @staticmethod
    def createData(sequence: int, payload) -> dict:
        payload = bytes(payload)
        if sequence < 1 or sequence > 255:
            raise ValueError(f"Invalid sequence: {sequence}. Must be 1-255.")
        if len(payload) > 255:
            raise ValueError(f"Payload too large: {len(payload)}. Max 255 bytes.")
        return {"soh": 0x01, "sequence": sequence, "invSequence": (~sequence) & 0xFF, "length": len(payload),
                "payload": payload, "checksum": CRC16.calculate(payload)}

    @staticmethod
    def serialize(packet: dict) -> bytes:
        p = packet["payload"]
        out = np.zeros(len(p) + 6, dtype=np.uint8)
        pa = np.frombuffer(p, dtype=np.uint8)
        n = L.lib().wam_xmodem_serialize(packet["sequence"], pa.ctypes.data_as(_u8p) if len(pa) else None, len(pa),
                                         out.ctypes.data_as(_u8p), len(out))
        L.check(n)
        res = bytearray(out[:n].tobytes())
        # serialize() writes the packet's own checksum field (packet.ts:50-51), which tests may corrupt
        res[4 + len(p)] = (packet["checksum"] >> 8) & 0xFF
        res[5 + len(p)] = packet["checksum"] & 0xFF
        return bytes(res)

    @staticmethod
    def verify(packet: dict) -> bool:
        return CRC16.calculate(packet["payload"]) == packet["checksum"]

    @staticmethod
    def serializeControl(controlType: int) -> bytes:
        return bytes([controlType])


def xmodem_batch_check(byte_rows: np.ndarray, lengths, expected_seq=None, device: int = 0) -> list[dict]:
    """byte_rows uint8 [n_streams, stride]; lengths[s] valid bytes.  One warp per stream on the GPU."""
    rows = np.ascontiguousarray(byte_rows, dtype=np.uint8)
    assert rows.ndim == 2
    n = rows.shape[0]
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    es = np.ascontiguousarray(expected_seq, dtype=np.int32) if expected_seq is not None else None
    res = (L.PktResult * max(n, 1))()
    L.check(L.lib().wam_xmodem_batch_check(device, rows.ctypes.data if rows.size else None, rows.shape[1], ln.ctypes.data,
                                           es.ctypes.data if es is not None else None, n, res))
    return [{k: getattr(res[i], k) for k, _ in L.PktResult._fields_} for i in range(n)]


def crc16_batch(byte_rows: np.ndarray, lengths, device: int = 0) -> np.ndarray:
    rows = np.ascontiguousarray(byte_rows, dtype=np.uint8)
    n = rows.shape[0]
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    out = np.zeros(n, dtype=np.uint16)
    L.check(L.lib().wam_crc16_batch(device, rows.ctypes.data if rows.size else None, rows.shape[1], ln.ctypes.data, n,
                                    out.ctypes.data))
    return out
