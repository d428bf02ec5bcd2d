"""Pins the CPU oracle against every deterministic expectation the reference's own tests hold for
the FSKCore / CRC / packet path.  The reference cannot be executed here (no
Node); these known answers are what anchors the oracle.  Each test cites the reference test
file:line it restates.  CPU only."""
import numpy as np
import pytest


def new_core(O, cfg=None):
    m = O.FSKCore()
    m.configure(cfg or {})
    return m


# ---- tests/utils/crc16.node.test.ts ---------------------------------------------------------
@pytest.mark.parametrize("data,crc", [
    (b"", 0xFFFF),                       # :12-16
    (b"A", 0xB915),                      # :18-22
    (b"123456789", 0x29B1),              # :24-28
    (b"\x00", 0xE1F0),                   # :33-37
    (b"\xff", 0xFF00),                   # :39-43
    (b"\xaa\xaa", 0xFB1A),               # :45-49
    (bytes(range(256)), 0x3FBD),         # :51-61
    (b"\x00\x00", 0x1D0F),               # :121-135
    (b"\x01\x02", 0x0E7C),
    (b"\xab\xcd", 0xD46A),
])
def test_crc16_golden_vectors(oracle, data, crc):
    assert oracle.crc16(data) == crc


def test_crc16_single_bit_error_detected(oracle):  # :79-87
    assert oracle.crc16(b"\x31\x32\x33\x34\x35") != oracle.crc16(b"\x30\x32\x33\x34\x35")


# ---- tests/transports/xmodem/xmodem.node.test.ts (packet format) -----------------------------
def test_packet_header_bytes_and_bad_crc(oracle):
    pkt = oracle.xmodem_serialize(1, b"BCD")
    assert pkt[:4] == bytes([0x01, 0x01, 0xFE, 0x03])            # :437-457 SOH SEQ ~SEQ LEN
    assert pkt[4:7] == b"BCD" and len(pkt) == 9
    crc = oracle.crc16(b"BCD")
    assert pkt[7] == crc >> 8 and pkt[8] == crc & 0xFF
    assert oracle.xmodem_check(pkt, 1)["status"] == 0
    bad = bytes([0x01, 0x01, 0xFE, 0x03, 0x42, 0x43, 0x44, 0xFF, 0xFF])  # :1034-1058
    assert oracle.PKT_STATUS[oracle.xmodem_check(bad, 1)["status"]] == "BAD_CRC"
    with pytest.raises(ValueError, match="Invalid sequence"):      # packet.ts:22-24
        oracle.xmodem_serialize(0, b"x")
    with pytest.raises(ValueError, match="Payload too large"):     # packet.ts:25-27
        oracle.xmodem_serialize(1, bytes(256))


def test_packet_receive_rules(oracle):
    """src/transports/xmodem/xmodem.ts:232-321"""
    S = oracle.PKT_STATUS
    pkt2 = oracle.xmodem_serialize(2, b"hello")
    assert S[oracle.xmodem_check(b"\x55\xaa" + pkt2, 2)["status"]] == "OK"          # junk before SOH ignored
    assert oracle.xmodem_check(b"\x55\xaa" + pkt2, 2)["payloadOffset"] == 6
    assert S[oracle.xmodem_check(pkt2, 3)["status"]] == "DUPLICATE"                 # previous sequence
    assert S[oracle.xmodem_check(oracle.xmodem_serialize(255, b"z"), 1)["status"]] == "DUPLICATE"  # wraps 1 -> 255
    assert S[oracle.xmodem_check(pkt2, 7)["status"]] == "UNEXPECTED_SEQ"
    assert S[oracle.xmodem_check(b"\x04" + pkt2, 2)["status"]] == "EOT"
    assert S[oracle.xmodem_check(b"\x10\x20", 1)["status"]] == "NO_SOH"
    assert S[oracle.xmodem_check(pkt2[:-1], 2)["status"]] == "INCOMPLETE"
    bad = bytearray(pkt2); bad[2] ^= 1
    assert S[oracle.xmodem_check(bytes(bad), 2)["status"]] == "BAD_COMPLEMENT"


# ---- tests/modems/fsk-modulation.node.test.ts ------------------------------------------------
def test_default_config_and_lengths(oracle):
    m = new_core(oracle)
    p = m.params()
    assert (p["samplesPerBit"], p["downsampledSamplesPerBit"], p["bitsPerByte"], p["nbits"]) == (40, 20, 10, 30)
    assert p["syncRingCapacity"] == 1364 and p["ampRingCapacity"] == 160 and p["samplesForEOD"] == 140
    # :75-90 single byte; :191-207 start/stop framing; :92-109 per-byte increment
    assert len(m.modulateData(b"H")) == 4 * 400 + 80 + 400
    assert len(m.modulateData(b"\x00")) == 2080
    l1, l2, l3 = (len(m.modulateData(b"Hel"[:k])) for k in (1, 2, 3))
    assert l2 - l1 == 400 and l3 - l2 == 400
    assert len(m.modulateData(b"AB")) == 2480
    assert len(m.modulateData(b"")) == 3 * 400 + 80 + 400           # fsk-sfd :96-113 style
    m300 = new_core(oracle, dict(baudRate=300))
    assert len(m300.modulateData(b"Hello, World!")) == 27520
    sig = m.modulateData(b"\x55")                                    # :111-124 amplitude bounds
    assert 0.8 < sig.max() <= 1.1 and -1.1 <= sig.min() < -0.8
    sig = m.modulateData(b"\x3c")                                    # :126-135 phase continuity
    assert np.max(np.abs(np.diff(sig))) < 0.5


def test_not_configured_errors(oracle):  # fsk-modulation :211-216, fsk-demodulation :30-36
    m = oracle.FSKCore()
    with pytest.raises(RuntimeError, match="not configured"):
        m.modulateData(b"H")
    with pytest.raises(RuntimeError, match="not configured"):
        m.demodulateData(np.zeros(3, dtype=np.float32))


# ---- tests/modems/fsk-demodulation.node.test.ts ----------------------------------------------
def test_roundtrip_ab_exact_one_sync(oracle):  # :81-106
    m = new_core(oracle)
    out = m.demodulateData(m.modulateData(b"AB"))
    assert out == b"AB"
    assert m.getStatus()["syncDetections"] == 1


def test_empty_and_short_signals(oracle):  # :13-28
    m = new_core(oracle)
    assert m.demodulateData(np.zeros(0, dtype=np.float32)) == b""
    assert m.demodulateData(np.zeros(100, dtype=np.float32)) == b""


def test_chunked_128_equals_whole(oracle):  # :363-398
    m = new_core(oracle)
    sig = m.modulateData(b"AB")
    out = b"".join(m.demodulateData(sig[i:i + 128].copy()) for i in range(0, len(sig), 128))
    assert out == b"AB" and m.getStatus()["syncDetections"] == 1


def test_leading_silence_2000(oracle):  # :400-437
    m = new_core(oracle)
    sig = np.concatenate([np.zeros(2000, dtype=np.float32), m.modulateData(b"AB")])
    out = b"".join(m.demodulateData(sig[i:i + 128].copy()) for i in range(0, len(sig), 128))
    assert out == b"AB"


def test_low_amplitude_agc(oracle):  # :493-521
    m = new_core(oracle)
    sig = (m.modulateData(b"AB") * np.float32(0.1)).astype(np.float32)
    out = b"".join(m.demodulateData(sig[i:i + 128].copy()) for i in range(0, len(sig), 128))
    assert out == b"AB"


def test_all_128_chunk_offsets(oracle):  # :668-716
    sig = new_core(oracle).modulateData(b"AB")
    for offset in range(128):
        m = new_core(oracle)
        out = b"".join(m.demodulateData(sig[i:i + 128].copy()) for i in range(offset, len(sig), 128))
        assert len(out) >= 2, offset


@pytest.mark.parametrize("chunk", [32, 64, 128, 256])
def test_chunk_sizes(oracle, chunk):  # :718-753
    m = new_core(oracle)
    sig = m.modulateData(b"AB")
    out = b"".join(m.demodulateData(sig[i:i + chunk].copy()) for i in range(0, len(sig), chunk))
    assert out == b"AB" and m.getStatus()["syncDetections"] > 0.95


def test_three_messages_with_gaps(oracle):  # :854-925
    m = new_core(oracle)
    got = b""
    for p in (b"AB", b"Hel", b"lo"):
        sig = np.concatenate([m.modulateData(p), np.zeros(500, dtype=np.float32)])
        got += b"".join(m.demodulateData(sig[i:i + 128].copy()) for i in range(0, len(sig), 128))
    assert got == b"ABHello"


@pytest.mark.parametrize("byte", [0x00, 0xFF, 0x55, 0xAA, 0x7E, 0x0F, 0xF0, 0x33, 0xCC])
def test_single_byte_patterns(oracle, byte):  # :1110-1131
    m = new_core(oracle)
    assert m.demodulateData(m.modulateData(bytes([byte]))) == bytes([byte])


@pytest.mark.parametrize("byte", [0xFF, 0x00, 0x55, 0x7E])
def test_identical_runs_one_eod(oracle, byte):  # :1133-1161
    m = new_core(oracle)
    assert m.demodulateData(m.modulateData(bytes([byte] * 3))) == bytes([byte] * 3)
    assert m.getStatus()["eodEvents"] == 1


@pytest.mark.parametrize("cfg", [dict(baudRate=300), dict(baudRate=1200),
                                 dict(markFrequency=1650, spaceFrequency=1850),
                                 dict(markFrequency=2125, spaceFrequency=2295)])
def test_bauds_and_tone_pairs(oracle, cfg):  # :301-345
    m = new_core(oracle, cfg)
    out = m.demodulateData(m.modulateData(b"\x48"))
    assert b"\x48" in out


# ---- tests/modems/fsk-sfd.node.test.ts -------------------------------------------------------
@pytest.mark.parametrize("payload", [bytes([0x48, 0x55, 0x65, 0x55, 0x6C]), bytes([0x55, 0x7E, 0x48, 0x55, 0x7E]),
                                     bytes([0x55, 0x55, 0x48]), b""])
def test_sfd_payloads(oracle, payload):  # :36-93, :163-171
    m = new_core(oracle)
    assert m.demodulateData(m.modulateData(payload)) == payload


def test_two_frames_two_eods(oracle):  # :139-159
    m = new_core(oracle)
    for p in (b"A", b"B"):
        assert m.demodulateData(m.modulateData(p)) == p
    assert m.getStatus()["eodEvents"] == 2


# ---- tests/modems/fsk-simplesync.node.test.ts (48 kHz / 300 Bd / 1650 / 1850) ----------------
SIMPLESYNC = dict(sampleRate=48000, baudRate=300, markFrequency=1650, spaceFrequency=1850, syncThreshold=0.85)


@pytest.mark.parametrize("payload", [b"H", b"Hello", bytes([0x55, 0x55, 0x7E, 0x48]), bytes([0x54, 0x55, 0x56])])
def test_simplesync_payloads(oracle, payload):  # :24-63, :105-116
    m = new_core(oracle, SIMPLESYNC)
    assert m.demodulateData(m.modulateData(payload)) == payload


def test_simplesync_reconfigure_on_mutated_buffer(oracle):  # :83-102 (AGC mutates the caller's buffer)
    m = new_core(oracle, SIMPLESYNC)
    sig = m.modulateData(b"Hi")
    before = sig.copy()
    for thr in (0.7, 0.8, 0.9):
        m.configure({**SIMPLESYNC, "syncThreshold": thr})
        out = m.demodulateData(sig)  # same buffer re-used, like the reference test
        assert isinstance(out, bytes)
    assert not np.array_equal(sig, before)


# ---- tests/modems/fsk-false-positive.node.test.ts ---------------------------------------------
def test_false_positives(oracle):
    m = new_core(oracle)
    assert m.demodulateData(np.zeros(4000, dtype=np.float32)) == b""                        # :14-24
    assert new_core(oracle).demodulateData(np.full(4000, 0.5, dtype=np.float32)) == b""     # :26-36
    assert new_core(oracle).demodulateData(np.full(4000, -0.3, dtype=np.float32)) == b""    # :38-48
    assert new_core(oracle).demodulateData(np.zeros(12000, dtype=np.float32)) == b""        # :50-60
    t = np.arange(8000)
    tone = np.sin(2 * np.pi * 2000 / 48000 * t).astype(np.float32)                           # :64-82
    assert new_core(oracle).demodulateData(tone) == b""
    alt = np.where(t % 2 == 0, 1.0, -1.0).astype(np.float32)                                 # :102-118
    assert new_core(oracle).demodulateData(alt) == b""


def _raw_bits_signal(bits, cfg):
    spb = int(cfg["sampleRate"] // cfg["baudRate"])
    out, phase = [], 0.0
    for b in bits:
        f = cfg["markFrequency"] if b else cfg["spaceFrequency"]
        for _ in range(spb):
            out.append(np.sin(phase))
            phase += 2 * np.pi * f / cfg["sampleRate"]
    return np.array(out, dtype=np.float32)


def test_partial_and_wrong_preamble(oracle):  # :134-206
    cfg = oracle.DEFAULT_FSK_CONFIG
    assert new_core(oracle).demodulateData(_raw_bits_signal([(0x55 >> b) & 1 for b in range(7, 3, -1)], cfg)) == b""
    assert new_core(oracle).demodulateData(_raw_bits_signal([(0xAA >> b) & 1 for b in range(7, -1, -1)], cfg)) == b""


def test_zero_then_valid(oracle):  # :209-242
    m = new_core(oracle)
    for _ in range(3):
        assert m.demodulateData(np.zeros(4000, dtype=np.float32)) == b""
    assert b"\x48" in m.demodulateData(m.modulateData(b"\x48"))


# ---- tests/modems/fsk-preamble-robustness.node.test.ts ----------------------------------------
def test_preamble_truncation(oracle):
    m = new_core(oracle)
    full = m.modulateData(b"\x48")
    sync_len = 3 * 10 * 40
    assert new_core(oracle).demodulateData(full[int(sync_len * 0.75):].copy()) == b""       # :65-84
    data = bytes([0x55, 0x48, 0x65])                                                          # :86-121
    assert new_core(oracle).demodulateData(new_core(oracle).modulateData(data)) == data


# ---- SURVEY section 0 facts the oracle must reproduce -----------------------------------------
def test_polarity_r9(oracle):
    """R9: the lower tone decodes as 1.  Bell 103 as named (1270/1070) decodes nothing."""
    m = new_core(oracle, dict(baudRate=300, markFrequency=1270, spaceFrequency=1070))
    assert m.demodulateData(m.modulateData(b"Hello, World!")) == b""
    m = new_core(oracle, dict(baudRate=300, markFrequency=1070, spaceFrequency=1270))
    assert m.demodulateData(m.modulateData(b"Hello, World!")) == b"Hello, World!"


def test_fractional_ring_capacity_r10(oracle):
    """R10: at 44.1 kHz the sync ring capacity is fractional (1227.6): only the first frame of a stream
    decodes; with parity 'even' the capacity is integral (1287) and every frame decodes."""
    assert abs(new_core(oracle, dict(sampleRate=44100)).params()["syncRingCapacity"] - 1227.6) < 1e-9
    assert new_core(oracle, dict(sampleRate=44100, parity="even")).params()["syncRingCapacity"] == 1287.0
    for cfg, want in ((dict(sampleRate=44100), [b"one"]), (dict(sampleRate=44100, parity="even"), [b"one", b"two", b"3", b"4"])):
        m = new_core(oracle, cfg)
        sig = np.concatenate([m.modulateData(p) for p in (b"one", b"two", b"3", b"4")])
        assert m.demodulateData(sig) == b"".join(want), cfg


def test_ring_buffer_semantics(oracle):
    """tests/utils.test.ts: overwrite-oldest, negative index, bounds errors; plus utils.ts:14-47 with a
    fractional size (typed-array length truncates, indices go fractional after the first wrap)."""
    r = oracle.RingBuffer("u8", 3)
    r.put(1, 2, 3, 4)
    assert [r.get(0), r.get(1), r.get(2), r.get(-1)] == [2, 3, 4, 4] and r.length == 3
    with pytest.raises(IndexError):
        r.get(3)
    r.clear()
    assert r.length == 0
    f = oracle.RingBuffer("u8", 2.5)
    f.put(1, 1, 1)
    assert f.length == 3                       # _length grows past the integer capacity
    assert f.get(0) == 1 and f.get(2) is None  # index 2 is outside the 2-element typed array
    f.put(1)
    assert f.get(0) == 1                       # readIndex 1 (still integral): stale element
    f.put(1)
    assert f.get(0) is None                    # readIndex 2: outside the typed array
    f.put(1)
    assert f.get(0) is None and f.get(1) is None  # readIndex 0.5: fractional for good
