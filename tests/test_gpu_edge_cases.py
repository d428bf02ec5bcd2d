"""Edge cases of the batched C ABI on the GPU: empty and one-sample calls, a too-small output buffer (flagged, not
silently truncated), maximum XModem payloads, zero-length payload frames, argument errors, a stream count that is not
a multiple of the warp size, and a size-independent round-trip property at a larger batch."""
import ctypes as C

import numpy as np
import pytest

import siggen

pytestmark = pytest.mark.gpu


def test_empty_and_tiny_calls_match_oracle(gpu_wam, oracle):
    cfg = {}
    x = siggen.multi_frame_stream(cfg, 30000, 5, 25.0, seed=1)[0]
    core = oracle.FSKCore()
    core.configure(cfg)
    b = gpu_wam.FSKBatch(3, cfg)
    want = b""
    got = [b""] * 3
    pos = 0
    for n in (0, 1, 1, 2, 31, 32, 33, 0, 255, 256, 257, 5000, 1, 30000):  # empty, one sample, tile edges, pipeline threshold
        n = min(n, len(x) - pos)
        seg = x[pos:pos + n]
        want += core.demodulateData(seg.copy())
        part = b.demodulate_bytes(np.ascontiguousarray(np.stack([seg] * 3)).reshape(3, n))
        got = [g + p for g, p in zip(got, part)]
        pos += n
    assert got == [want] * 3 and len(want) >= 5
    st, o = b.status()[1], core.getStatus()
    for k in ("demodulationCalls", "totalSamplesProcessed", "syncDetections", "eodEvents", "globalSampleCounter"):
        assert float(st[k]) == float(o[k]), k


def test_output_overflow_is_flagged(gpu_wam):
    """A caller-provided output row that is too small: the bytes that fit are written, out_len is clamped to the row,
    and the stream's errorEvents carries WAM_ERR_OUT_OVERFLOW — never a write past the row."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    payload = bytes(range(40))
    sig = siggen.modulate({}, payload)
    n_streams = 5
    x = torch.from_numpy(np.stack([sig] * n_streams)).to(dev)
    guard = 64
    out = torch.full((n_streams, 8 + guard), 0xEE, dtype=torch.uint8, device=dev)
    ln = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    for flags in (0, gpu_wam._lib.WAM_BATCH_NO_PIPELINE):
        b = gpu_wam.FSKBatch(n_streams, {})
        out.fill_(0xEE)
        # out_stride = row pitch of the tensor, but only 8 bytes of capacity are announced through a view
        view = out[:, :8]
        rows = torch.empty((n_streams, 8), dtype=torch.uint8, device=dev)
        b.demodulate_device(x.data_ptr(), x.shape[1], x.shape[1], rows.data_ptr(), 8, ln.data_ptr(), flags=flags)
        torch.cuda.synchronize()
        assert ln.cpu().tolist() == [8] * n_streams
        assert bytes(rows[2].cpu().numpy()) == payload[:8]
        assert all(s["errorEvents"] & 1 for s in b.status())
        assert view.shape == (n_streams, 8)
        b.close()


def test_maximum_and_empty_xmodem_payloads(gpu_wam, oracle):
    rng = np.random.default_rng(3)
    cases = [b"", bytes([7]), rng.integers(0, 256, 255, dtype=np.uint8).tobytes()]  # PacketConstants.MAX_PAYLOAD_SIZE
    rows = np.zeros((len(cases), 261), dtype=np.uint8)  # MAX_PACKET_SIZE
    lens = []
    for i, p in enumerate(cases):
        pk = gpu_wam.XModemPacket.serialize(gpu_wam.XModemPacket.createData(255, p))
        assert pk == oracle.xmodem_serialize(255, p)
        rows[i, :len(pk)] = np.frombuffer(pk, dtype=np.uint8)
        lens.append(len(pk))
    res = gpu_wam.xmodem_batch_check(rows, lens, [255] * len(cases))
    assert [r["status"] for r in res] == [0, 0, 0] and [r["length"] for r in res] == [0, 1, 255]
    with pytest.raises(ValueError):
        gpu_wam.XModemPacket.createData(1, bytes(256))
    with pytest.raises(ValueError):
        gpu_wam.XModemPacket.createData(0, b"x")
    # maximum packet through the modem and a receive session that starts at sequence 255
    m = gpu_wam.FSKCore()
    m.configure({})
    pk = gpu_wam.XModemPacket.serialize(gpu_wam.XModemPacket.createData(255, cases[2]))
    rx = bytes(m.demodulateData(m.modulateData(pk)))
    assert rx == pk
    sess = gpu_wam.XModemBatchReceiver(1)
    sess.state["expectedSequence"] = 255
    assert sess.feed([rx]) == [b"\x06"] and sess.received(0) == cases[2] and int(sess.state["expectedSequence"][0]) == 1


def test_zero_byte_frame_and_argument_errors(gpu_wam, oracle):
    b = gpu_wam.FSKBatch(4, {})
    sig, out_len = b.modulate(np.zeros((4, 0), dtype=np.uint8))  # preamble + SFD only (fsk.ts:391-394)
    want = siggen.modulate({}, b"")
    assert out_len.tolist() == [len(want)] * 4
    np.testing.assert_allclose(sig[3], want, atol=1e-4, rtol=0)
    assert b.demodulate_bytes(sig) == [b""] * 4
    L = gpu_wam._lib
    lib = gpu_wam.lib()
    assert lib.wam_fsk_batch_demodulate(b._h, None, 10, 10, None, 0, None, 0) == L.WAM_E_INVALID  # out_len missing
    nv = np.array([11, 0, 0, 0], dtype=np.int32)
    x = np.zeros((4, 10), dtype=np.float32)
    ol = np.zeros(4, dtype=np.int32)
    rc = lib.wam_fsk_batch_demodulate_ragged(b._h, x.ctypes.data, 10, 10, nv.ctypes.data, None, 0, ol.ctypes.data, 0)
    assert rc == L.WAM_E_INVALID and b"n_valid" in lib.wam_last_error()
    h = C.c_void_p()
    assert lib.wam_fsk_mux_create(0, 0, None, 0, None, 128, C.byref(h)) == L.WAM_E_INVALID
    with pytest.raises(gpu_wam.WamError):
        gpu_wam.FSKBatch(0, {})


def test_large_ragged_round_trip_property(gpu_wam):
    """Size-independent property at a larger batch (20,000 streams, fused kernel with TMA): every stream carries its
    own payload and its own length; modulate -> demodulate returns exactly that payload."""
    rng = np.random.default_rng(12)
    n_streams = 20000
    data = rng.integers(0, 256, (n_streams, 6), dtype=np.uint8)
    dl = rng.integers(0, 7, n_streams).astype(np.int32)
    b = gpu_wam.FSKBatch(n_streams, {})
    sig, out_len = b.modulate(data, dl)
    nv = out_len.copy()
    nv[::97] = -1  # some streams are not called at all
    got = b.demodulate_ragged(sig, nv)
    for s in range(0, n_streams, 13):
        want = b"" if nv[s] < 0 else data[s, :dl[s]].tobytes()
        assert got[s] == want, s
    assert sum(len(g) for g in got) == int(dl[nv >= 0].sum())
