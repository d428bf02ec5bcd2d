"""CPU-only checks of the product side: libwam.so builds, loads and exports every symbol that
include/wam.h declares; host-only entry points (designs, CRC, packet serialisation, config
defaults) match the oracle; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wam.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wam_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(wam):
    lib = C.CDLL(os.path.join(ROOT, "webaudio-modem_b200", "libwam.so"))
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libwam.so does not export {n}"
    bound = set(wam._lib.SYMBOLS)
    assert bound == set(names), (bound ^ set(names))
    assert wam.lib().wam_version() == 100


def test_no_cpu_fallback(wam):
    """Without a CUDA device every compute entry point must fail with WAM_E_CUDA."""
    n = C.c_int(0)
    rc = wam.lib().wam_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(wam.WamError) as e:
        wam.FSKCore().configure({})
    assert e.value.code == wam._lib.WAM_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(wam.WamError):
        wam.FSKBatch(4, {})
    with pytest.raises(wam.WamError):
        wam.IIRFilter([1.0], [1.0]).processBuffer(np.ones(4, dtype=np.float32))
    with pytest.raises(wam.WamError):
        wam.xmodem_batch_check(np.zeros((1, 8), dtype=np.uint8), [8])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "webaudio-modem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "fsk.py" and "import oracle" not in src, f
                assert "import oracle" not in src and "wam_oracle" not in src, f


def test_host_side_reference_surface(wam):
    m = wam.FSKCore()
    assert m.name == "FSK" and m.type == "FSK" and not m.isReady()
    with pytest.raises(RuntimeError, match="FSK modulator not configured"):      # fsk.ts:378-380
        m.modulateData(b"x")
    with pytest.raises(RuntimeError, match="FSK demodulator not configured"):    # fsk.ts:191-193
        m.demodulateData(np.zeros(4, dtype=np.float32))
    assert m.getSignalQuality() == {"snr": 0, "ber": 0, "eyeOpening": 0, "phaseJitter": 0, "frequencyOffset": 0}
    assert m.getStatus()["ready"] is False and m.getStatus()["silenceThreshold"] == 0.01
    seen = []
    cb = lambda e: seen.append(e)
    m.on("x", cb); m.emit("x", 1); m.off("x", cb); m.emit("x", 2)
    assert seen == [1]
    cfg = wam.normalize_config(dict(baud=300, markFreq=980, spaceFreq=1180))      # README aliases
    assert cfg["baudRate"] == 300 and cfg["markFrequency"] == 980 and cfg["spaceFrequency"] == 1180
    assert cfg["preamblePattern"] == [0x55, 0x55] and cfg["sfdPattern"] == [0x7E] and cfg["syncThreshold"] == 0.85
    assert wam.DEFAULT_FSK_CONFIG == dict(sampleRate=48000, baudRate=1200, markFrequency=1650, spaceFrequency=1850,
                                          preamblePattern=[0x55, 0x55], sfdPattern=[0x7E], startBits=1, stopBits=1,
                                          parity="none", syncThreshold=0.85, agcEnabled=True, preFilterBandwidth=800,
                                          adaptiveThreshold=True)                  # fsk.ts:19-33


def test_host_math_equals_oracle(wam, oracle):
    G, O = wam.FilterDesign, oracle.FilterDesign
    for a, b in ((G.butterworthLowpass(300, 48000), O.butterworthLowpass(300, 48000)),
                 (G.butterworthHighpass(700, 44100), O.butterworthHighpass(700, 44100)),
                 (G.butterworthBandpass(1080, 800, 48000), O.butterworthBandpass(1080, 800, 48000))):
        np.testing.assert_array_equal(a["b"], b["b"]); np.testing.assert_array_equal(a["a"], b["a"])
    for n in (51, 50):
        np.testing.assert_array_equal(G.sincLowpass(1000, 44100, n), O.sincLowpass(1000, 44100, n))
        np.testing.assert_array_equal(G.sincHighpass(1000, 44100, n), O.sincHighpass(1000, 44100, n))
        np.testing.assert_array_equal(G.sincBandpass(1500, 400, 44100, n), O.sincBandpass(1500, 400, 44100, n))
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 9, 128, 255, 1024):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert wam.CRC16.calculate(d) == oracle.crc16(d)
    assert wam.CRC16.calculate(b"123456789") == 0x29B1 and wam.CRC16.verify(b"123456789", 0x29B1)
    p = wam.XModemPacket.createData(7, b"payload")
    assert wam.XModemPacket.serialize(p) == oracle.xmodem_serialize(7, b"payload") and wam.XModemPacket.verify(p)
    with pytest.raises(ValueError, match="Invalid sequence: 0"):
        wam.XModemPacket.createData(0, b"")
    with pytest.raises(ValueError, match="Payload too large: 256"):
        wam.XModemPacket.createData(1, bytes(256))
    with pytest.raises(ValueError, match="cannot be empty"):
        wam.IIRFilter([], [1])


# ---- the boundary seen from C and from the Node side --------------------------------------------------------------
def test_c_abi_layouts_match_the_ctypes_mirror(tmp_path):
    """tests/c_abi_smoke.c (plain C over include/wam.h) prints every boundary struct's size and field offsets; the
    ctypes structs in _lib.py must agree field by field."""
    import ctypes as C
    import importlib
    import json
    import subprocess

    exe = tmp_path / "c_abi_smoke"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-o", str(exe), os.path.join(ROOT, "tests", "c_abi_smoke.c")],
                   check=True)
    layout = json.loads(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    L = importlib.import_module("webaudio-modem_b200._lib")
    mirror = {"wam_fsk_config": L.FSKConfigStruct, "wam_fsk_status": L.StatusStruct, "wam_pkt_result": L.PktResult,
              "wam_xmodem_rx_state": L.XmodemRxState, "wam_chunk_result": L.ChunkResult, "wam_fast_stats": L.FastStats}
    assert set(layout) == set(mirror)
    for name, st in mirror.items():
        want = layout[name]
        assert C.sizeof(st) == want["size"], name
        fields = {f[0]: getattr(st, f[0]).offset for f in st._fields_}
        listed = {k: v for k, v in want.items() if k != "size"}
        if listed:
            assert fields == listed, name


def test_napi_addon_is_valid_c_against_the_header_and_binds_what_the_ts_host_calls():
    """host/wam_napi.c compiles (syntax + types, -Wall -Wextra -Werror) against include/wam.h and a stub node_api.h, and
    registers every native that host/fsk_core_gpu.ts declares or calls."""
    import re
    import subprocess

    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"),
                    os.path.join(ROOT, "host", "wam_napi.c")], check=True)
    c_src = open(os.path.join(ROOT, "host", "wam_napi.c")).read()
    ts_src = open(os.path.join(ROOT, "host", "fsk_core_gpu.ts")).read()
    registered = set(re.findall(r'\{"(\w+)", NULL, \w+, NULL', c_src))
    called = set(re.findall(r"native\.(\w+)\(", ts_src))
    block = ts_src[ts_src.index("const native = require"):ts_src.index("};", ts_src.index("const native = require"))]
    declared = set(re.findall(r"^\s{2}(\w+)\(", block, flags=re.M))
    assert called and called <= declared, called - declared
    assert declared <= registered, declared - registered
    assert registered <= declared, registered - declared
    # every libwam function the addon calls is declared in the header
    header = open(os.path.join(ROOT, "include", "wam.h")).read()
    for fn in set(re.findall(r"\b(wam_\w+)\(", c_src)):
        assert re.search(r"\b%s\(" % fn, header), fn
    assert os.path.exists(os.path.join(ROOT, "host", "alias-hook.mjs")) and os.path.exists(os.path.join(ROOT, "host", "alias-hook-impl.mjs"))


def test_traffic_record_belongs_to_the_current_kernel_sources():
    """bench.py reports roofline.traffic from profiles/r02_demod_traffic.json only while it was measured on the sources in
    the tree (it says "stale" otherwise): the committed record must be the current one."""
    import json
    import sys

    sys.path.insert(0, ROOT)
    import bench

    rec = json.load(open(os.path.join(ROOT, "profiles", "r02_demod_traffic.json")))
    assert rec["source_hash"] == bench.kernel_source_hash()
    assert rec["algorithmic_bytes_per_step"] == 65536 * 48000 * 4 and rec["bytes_per_step"] > rec["algorithmic_bytes_per_step"]
