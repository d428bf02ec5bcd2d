"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: contiguous stream shards, no
data-path collective, results gathered in global order.  The CPU oracle stands in for the GPU
demodulator here (it is test infrastructure; the sharding code under test is the product's)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_streams, tmp):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("webaudio-modem_b200.shard")
    import oracle as O
    import siggen

    x = np.load(os.path.join(tmp, "x.npy"))
    lo, hi = shard.shard_range(n_streams, rank, world)
    local, st = O.batch_demodulate([siggen.V21_CH2], None, np.ascontiguousarray(x[lo:hi]), n_threads=2)
    dist.barrier()
    allres = shard.gather_stream_results(local, n_streams, rank, world, dist)
    counts = shard.reduce_counters(np.array([sum(len(b) for b in local), hi - lo]), dist)
    if rank == 0:
        np.save(os.path.join(tmp, "lens.npy"), np.array([len(b) for b in allres]))
        open(os.path.join(tmp, "bytes.bin"), "wb").write(b"".join(allres))
        np.save(os.path.join(tmp, "counts.npy"), counts)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    shard = importlib.import_module("webaudio-modem_b200.shard")
    for n in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_matches_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp

    import siggen

    n_streams = 11  # odd on purpose: ragged shards
    x, _ = siggen.noisy_streams(siggen.V21_CH2, n_streams, 24000, 8, 12.0, seed=5, max_offset=400)
    np.save(tmp_path / "x.npy", x)
    want, _ = oracle.batch_demodulate([siggen.V21_CH2], None, x.copy(), n_threads=2)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_streams, str(tmp_path)), nprocs=2, join=True)
    lens = np.load(tmp_path / "lens.npy")
    blob = (tmp_path / "bytes.bin").read_bytes()
    got, p = [], 0
    for ln in lens:
        got.append(blob[p:p + ln]); p += ln
    assert got == want
    counts = np.load(tmp_path / "counts.npy")
    assert counts[0] == sum(len(b) for b in want) and counts[1] == n_streams
