"""Oracle pins for the batched XModem receive session (wamo_xmodem_receive): the deterministic expectations of
the reference's own 'Data Reception' tests (tests/transports/xmodem/xmodem.node.test.ts:765-983), restated
without the transport's timers.  The reference counts the initial NAK of receiveData() in sentData; the session
walk starts after it, so `sentData.length == k` there is `k - 1` replies here."""
import oracle

ACK, NAK, EOT = 0x06, 0x15, 0x04


def run(bursts, max_retries=10):
    st, replies, data = None, b"", b""
    pending = b""
    for b in bursts:
        buf = pending + bytes(b)
        st, rep, nrep, consumed, payload = oracle.xmodem_receive(buf, st, max_retries)
        assert nrep == len(rep)
        replies += rep
        data += payload
        pending = buf[consumed:]
    return st, replies, data


def pkt(seq, payload):
    return oracle.xmodem_serialize(seq, bytes(payload))


def test_receive_single_packet():  # xmodem.node.test.ts:766-781
    st, rep, data = run([pkt(1, [0x48, 0x65, 0x6C, 0x6C, 0x6F]) + bytes([EOT])])
    assert data == bytes([0x48, 0x65, 0x6C, 0x6C, 0x6F])
    assert st["packetsReceived"] == 1 and st["done"] == 1
    assert rep == bytes([ACK, ACK])  # 3 sent in the reference including the initial NAK


def test_receive_multiple_packets_reassembly():  # :783-803
    st, rep, data = run([pkt(1, [1, 2, 3]), pkt(2, [4, 5, 6]), pkt(3, [7, 8]), bytes([EOT])])
    assert data == bytes([1, 2, 3, 4, 5, 6, 7, 8])
    assert st["packetsReceived"] == 3
    assert rep == bytes([ACK] * 4)  # 5 in the reference


def test_out_of_sequence_fails_after_max_retries():  # :805-826 (maxRetries: 1)
    st, rep, data = run([pkt(2, [4, 5, 6])], max_retries=1)
    assert rep == bytes([NAK]) and st["done"] == 0 and st["packetsDropped"] == 1 and st["retries"] == 1
    st2, rep2, nrep2, consumed, payload = oracle.xmodem_receive(pkt(2, [4, 5, 6]), st, 1)
    assert st2["done"] == 2 and rep2 == b"" and st2["packetsDropped"] == 2  # 'Receive failed after max retries'


def test_duplicate_basic():  # :828-852
    st, rep, data = run([pkt(1, [0x42, 0x43]), pkt(1, [0x42, 0x43]), bytes([EOT])])
    assert data == bytes([0x42, 0x43])
    assert rep == bytes([ACK] * 3)  # 4 in the reference
    assert st["packetsReceived"] == 1 and st["packetsDropped"] == 1


def test_duplicate_in_multi_packet_transfer():  # :854-884
    st, rep, data = run([pkt(1, [0x41]) + pkt(2, [0x42]) + pkt(2, [0x42]) + pkt(3, [0x43]) + bytes([EOT])])
    assert data == bytes([0x41, 0x42, 0x43])
    assert rep == bytes([ACK] * 5)  # 6 in the reference
    assert st["packetsReceived"] == 3 and st["packetsDropped"] == 1


def test_byte_by_byte():  # :908-962
    bursts = [bytes([b]) for p in (pkt(1, [1, 2, 3]), pkt(2, [4, 5, 6]), pkt(3, [7, 8])) for b in p] + [bytes([EOT])]
    st, rep, data = run(bursts)
    assert data == bytes([1, 2, 3, 4, 5, 6, 7, 8]) and rep == bytes([ACK] * 4) and st["packetsReceived"] == 3


def test_bad_crc_packet_naks():  # bad-CRC packet of :1046 -> 'Invalid CRC' -> NAK (xmodem.ts:286-290, 251-260)
    st, rep, data = run([bytes([0x01, 0x01, 0xFE, 0x03, 0x42, 0x43, 0x44, 0xFF, 0xFF])])
    assert rep == bytes([NAK]) and data == b"" and st["packetsReceived"] == 1 and st["packetsDropped"] == 1
    st, rep2, nrep, consumed, payload = oracle.xmodem_receive(pkt(1, [0x42, 0x43, 0x44]), st)
    assert rep2 == bytes([ACK]) and payload == bytes([0x42, 0x43, 0x44]) and st["retries"] == 0


def test_sequence_wraps_255_to_1():  # xmodem.ts:298
    st = {"expectedSequence": 255, "retries": 0, "done": 0, "dataLen": 0, "packetsReceived": 0, "packetsDropped": 0}
    st, rep, nrep, consumed, payload = oracle.xmodem_receive(pkt(255, [9]) + pkt(1, [8]) + pkt(255, [9]), st)
    assert payload == bytes([9, 8]) and st["expectedSequence"] == 2
    assert rep == bytes([ACK, ACK, NAK])  # 255 is no longer "previous" once 1 has been received


def test_garbage_before_soh_is_ignored_and_bad_complement():  # xmodem.ts:248-250, 270-274
    st, rep, data = run([bytes([0x00, 0xFF, 0x33]) + pkt(1, [5]) + bytes([0x01, 0x02, 0x02, 0x00])])
    assert data == bytes([5]) and rep == bytes([ACK, NAK]) and st["packetsDropped"] == 1
