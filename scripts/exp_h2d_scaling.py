"""Host->device copy ceiling per rank, alone and with all ranks copying at once (pinned memory, 1-D and the
65,536-row 2-D copy of the host-buffer demodulate call).  Launch: python -m torch.distributed.run --nnodes=1
--nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/exp_h2d_scaling.py [--no-affinity]
One JSON line per rank on stdout."""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--no-affinity", action="store_true")
ap.add_argument("--gib", type=float, default=3.0)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("gloo")
wam = importlib.import_module("webaudio-modem_b200")
lib = wam.lib()
near = 0 if a.no_affinity else int(lib.wam_host_bind_near_device(lr))
info = {"rank": rank, "near_cpus": near, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."], "n_aff": len(os.sched_getaffinity(0))}
try:
    bus = torch.cuda.get_device_properties(lr).pci_bus_id
    info["pci_bus"] = bus
except Exception:
    bus = None
try:
    info["numa_online"] = open("/sys/devices/system/node/online").read().strip()
except Exception:
    pass
n = int(a.gib * (1 << 30)) // 4
rows = 65536; cols = n // rows
h = torch.empty(rows * cols, dtype=torch.float32, pin_memory=True)
h.fill_(1.0)
d = torch.empty(rows * cols, dtype=torch.float32, device="cuda")
h2 = h.view(rows, cols); d2 = d.view(rows, cols)
sub = cols // 4  # a quarter of every row: strided on both sides like one time slab
st = torch.cuda.Stream()

def barrier():
    if world > 1:
        dist.barrier()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt / reps

def copy1d():
    with torch.cuda.stream(st):
        d.copy_(h, non_blocking=True)

from cuda.bindings import runtime as rt
H2D = rt.cudaMemcpyKind.cudaMemcpyHostToDevice

def make2d(width_floats):
    """the host-buffer demodulate call's copy: time slabs of `width_floats` samples of all rows, row pitch = full row"""
    nsl = cols // width_floats
    def f():
        for q in range(nsl):
            err, = rt.cudaMemcpy2DAsync(d.data_ptr(), width_floats * 4, h.data_ptr() + q * width_floats * 4, cols * 4,
                                        width_floats * 4, rows, H2D, st.cuda_stream)
            assert int(err) == 0, err
    return f, nsl * width_floats * 4 * rows

copy2d, _b2 = make2d(3072)

bytes1 = rows * cols * 4; bytes2 = _b2
# alone: ranks take turns
def alone(fn, nbytes):
    out = None
    for r in range(world):
        if r == rank:
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize()
            out = 2 * nbytes / (time.perf_counter() - t0) / 1e9
        barrier()
    return out
info["alone_1d_GBps"] = alone(copy1d, bytes1)
info["alone_2d_GBps"] = alone(copy2d, bytes2)
info["all_1d_GBps"] = bytes1 / timed(copy1d) / 1e9
info["all_2d_GBps"] = bytes2 / timed(copy2d) / 1e9
for w in (768, 1536, 6144):
    f, nb = make2d(w)
    info[f"all_2d_w{w}_GBps"] = nb / timed(f) / 1e9
# host memory read bandwidth of this rank's threads (memcpy pinned -> pageable), all ranks at once
import numpy as np
src = h.numpy(); dst = np.empty_like(src)
def hostcopy():
    np.copyto(dst, src)
info["all_host_memcpy_GBps_1thread"] = bytes1 / timed(hostcopy, 2) / 1e9
print(json.dumps(info), flush=True)
if world > 1:
    dist.destroy_process_group()
