#!/usr/bin/env python
"""Development aid: wall-clock split of one end-to-end batch call (host parse / upload / decode / download)."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpegdec_b200 as J
from tests import synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    jp = [synth.synth_jpeg(1920, 1080, s, 75) for s in range(8)]
    sizes = [len(jp[i % 8]) for i in range(n)]
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += (s + 15) & ~15
    L = J.lib()
    in_ptr = L.JPEGB200_hostAlloc(o + 64)
    arr = np.ctypeslib.as_array(C.cast(in_ptr, C.POINTER(C.c_ubyte)), shape=(o + 64,))
    for i in range(n):
        arr[offs[i]:offs[i] + sizes[i]] = np.frombuffer(jp[i % 8], dtype=np.uint8)
    ptrs = [in_ptr + x for x in offs]
    ob = 1920 * 1080 * 4
    stride = (ob + 255) & ~255
    out_ptr = L.JPEGB200_hostAlloc(stride * n + 256)
    outs = [out_ptr + i * stride for i in range(n)]
    ctx = J.Context(0, J.JPEG_ARITH_SSE2)
    for it in range(4):
        t0 = time.time(); b = J.Batch(ctx, ptrs, sizes, J.RGB8888, 0)
        t1 = time.time()
        for i in range(n): b.set_output(i, outs[i], 0)
        t2 = time.time(); b.upload(); b.decode(0)
        t3 = time.time(); b.download()
        t4 = time.time(); st = b.wait()
        t5 = time.time(); tm = b.timings(); b.close()
        t6 = time.time()
        print("iter %d: create %.1f set_output %.1f upload+decode(enqueue) %.1f download(enqueue) %.1f wait %.1f close %.1f total %.1f ms | dev: %s" % (
            it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5), 1e3 * (t6 - t0),
            {k: round(v, 2) for k, v in tm.items()}))
main()
