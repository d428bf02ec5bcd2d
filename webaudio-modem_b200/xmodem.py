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


RX_STATE_DTYPE = np.dtype([("expectedSequence", "<i4"), ("retries", "<i4"), ("done", "<i4"), ("dataLen", "<i4"),
                           ("packetsReceived", "<i4"), ("packetsDropped", "<i4")])
ACK, NAK, SOH, EOT = 0x06, 0x15, 0x01, 0x04  # ControlType, src/transports/xmodem/types.ts:29-34


class XModemBatchReceiver:
    """Receive side of XModemTransport (src/transports/xmodem/xmodem.ts:232-321) for n independent sessions:
    feed every session's burst of demodulated bytes, get back the ACK / NAK bytes the transport would send,
    the reassembled payloads and the per-session statistics.  Bytes of an unfinished packet are kept and
    prepended to the next burst (the reference's receive.buffer)."""

    def __init__(self, n_sessions: int, max_retries: int = 10, device: int = 0, data_capacity: int = 1 << 16,
                 reply_cap: int = 16):
        self.n = int(n_sessions)
        self.max_retries = int(max_retries)  # XModemConfig.maxRetries default 10 (xmodem.ts:47)
        self.device = device
        self.reply_cap = reply_cap
        self.state = np.zeros(self.n, dtype=RX_STATE_DTYPE)
        self.state["expectedSequence"] = 1
        self.data = np.zeros((self.n, data_capacity), dtype=np.uint8)
        self.pending: list[bytes] = [b""] * self.n

    def feed(self, bursts) -> list[bytes]:
        """bursts: one bytes-like per session (may be empty).  Returns the reply bytes per session."""
        assert len(bursts) == self.n
        joined = [self.pending[i] + bytes(bursts[i]) for i in range(self.n)]
        stride = max(1, max(len(j) for j in joined))
        rows = np.zeros((self.n, stride), dtype=np.uint8)
        ln = np.zeros(self.n, dtype=np.int32)
        for i, j in enumerate(joined):
            rows[i, : len(j)] = np.frombuffer(j, dtype=np.uint8)
            ln[i] = len(j)
        replies = np.zeros((self.n, self.reply_cap), dtype=np.uint8)
        nrep = np.zeros(self.n, dtype=np.int32)
        consumed = np.zeros(self.n, dtype=np.int32)
        L.check(L.lib().wam_xmodem_batch_receive(self.device, rows.ctypes.data, stride, ln.ctypes.data, self.n,
                                                 self.max_retries, self.state.ctypes.data, replies.ctypes.data,
                                                 self.reply_cap, nrep.ctypes.data, consumed.ctypes.data,
                                                 self.data.ctypes.data, self.data.shape[1]))
        self.pending = [joined[i][consumed[i]:] for i in range(self.n)]
        return [bytes(replies[i, : min(nrep[i], self.reply_cap)]) for i in range(self.n)]

    def received(self, i: int) -> bytes:
        """assembleData(receive.data) of session i so far (xmodem.ts:323-334)"""
        return bytes(self.data[i, : min(int(self.state["dataLen"][i]), self.data.shape[1])])


def crc16_batch(byte_rows: np.ndarray, lengths, device: int = 0) -> np.ndarray:
    rows = np.ascontiguousarray(byte_rows, dtype=np.uint8)
    n = rows.shape[0]
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    out = np.zeros(n, dtype=np.uint16)
    L.check(L.lib().wam_crc16_batch(device, rows.ctypes.data if rows.size else None, rows.shape[1], ln.ctypes.data, n,
                                    out.ctypes.data))
    return out
