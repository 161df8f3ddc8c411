"""Micro-benchmark (developer tool): plain pinned host-to-device copy bandwidth of the box, 1 / 2 / 4 / 8 GPUs at once.

This is the ceiling of bench.py's end-to-end number (every 4 MP pair is 8.4 MB of frames that have to cross PCIe).
One thread per GPU issues cudaMemcpyAsync from its own pinned buffer; variants: one or two copy streams per GPU
(two halves of the buffer in flight), default pinned vs write-combined host memory, 16 MB chunks vs one copy.

    python tests/micro/h2d_bw.py [max_gpus] > profiles/rNN_h2d_bw.txt
"""
import ctypes
import sys
import threading
import time

import torch

def _load_cudart():
    import glob
    import os
    cands = ["libcudart.so.12", "libcudart.so"] + \
        glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
        glob.glob("/usr/local/cuda/lib64/libcudart.so*")
    for name in cands:
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    raise OSError("libcudart not found")


rt = _load_cudart()
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
H2D = 1
SIZE = 256 << 20
REPS = 12


def worker(dev, flags, n_streams, chunk, barrier, out):
    torch.cuda.set_device(dev)
    host = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(host), SIZE, flags) == 0
    ctypes.memset(host, 1, SIZE)
    dst = torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{dev}")
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]

    def issue():
        part = SIZE // n_streams
        for si, st in enumerate(streams):
            off = si * part
            step = chunk if chunk else part
            for o in range(off, off + part, step):
                n = min(step, off + part - o)
                rt.cudaMemcpyAsync(dst.data_ptr() + o, host.value + o, n, H2D, ctypes.c_void_p(st.cuda_stream))
    issue()
    torch.cuda.synchronize(dev)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(REPS):
        issue()
    torch.cuda.synchronize(dev)
    out[dev] = SIZE * REPS / (time.perf_counter() - t0) / 1e9
    barrier.wait()
    rt.cudaFreeHost(host)


def run(n_gpus, flags, n_streams, chunk):
    barrier = threading.Barrier(n_gpus)
    out = {}
    ts = [threading.Thread(target=worker, args=(d, flags, n_streams, chunk, barrier, out)) for d in range(n_gpus)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return sum(out.values()), out


def main():
    have = torch.cuda.device_count()
    want = int(sys.argv[1]) if len(sys.argv) > 1 else have
    print(f"# {have} GPU(s) visible: {torch.cuda.get_device_name(0)}; buffer {SIZE >> 20} MiB x {REPS} copies per GPU")
    for n in (1, 2, 4, 8):
        if n > min(have, want):
            break
        for label, flags, ns, chunk in (("pinned, 1 stream", 1, 1, 0), ("pinned, 2 streams", 1, 2, 0),
                                        ("pinned, 1 stream, 16 MiB chunks", 1, 1, 16 << 20),
                                        ("write-combined, 1 stream", 1 | 4, 1, 0), ("write-combined, 2 streams", 1 | 4, 2, 0)):
            tot, per = run(n, flags, ns, chunk)
            print(f"{n} GPU(s)  {label:34s} aggregate {tot:7.1f} GB/s   per GPU " + " ".join(f"{per[d]:5.1f}" for d in sorted(per)))


if __name__ == "__main__":
    main()
