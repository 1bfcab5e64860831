#!/usr/bin/env python
"""Development aid: throughput with 1 vs 2 device-resident batches in flight (entropy of batch k+1 overlapping IDCT of batch k)."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpegdec_b200 as J
from tests import synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    K = 12
    jp = synth.synth_set(8, 1920, 1080, quality=75)
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
    ctx = J.Context(0, J.JPEG_ARITH_SSE2)
    for nb in (1, 2, 3):
        bs = []
        for _ in range(nb):
            b = J.Batch(ctx, ptrs, sizes, J.RGB8888, 0); b.alloc_device_output(); b.upload(); b.decode(J.JPEGB200_OUT_DEVICE); b.download(); assert not any(b.wait())
            bs.append(b)
        for rep in range(2):
            t0 = time.time()
            for k in range(K):
                b = bs[k % nb]
                if k >= nb: b.wait()
                b.decode(J.JPEGB200_OUT_DEVICE); b.download()
            for b in bs: b.wait()
            t1 = time.time()
        print("batches in flight %d: %.3f ms per step (wall, %d steps)" % (nb, 1e3 * (t1 - t0) / K, K))
        for b in bs: b.close()
main()
