#!/usr/bin/env python
"""Development aid: ONE JPEGB200_decodeBatch(OUT_DEVICE) call over a slice of UHD files, swept over the pipeline depth
(JPEGB200_setPipelineDepth) -- run once per JPEGDEC_B200_JOB_MB value (the job size is read once per process).
Prints wall ms per call and the summed per-job device timings."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpegdec_b200 as J
from tests import synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 625
    jp = synth.synth_set(16, 3840, 2160, quality=85, seed0=0, restart_rows=1, workers=16)
    sizes = [len(jp[i % 16]) for i in range(n)]
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += (s + 15) & ~15
    L = J.lib()
    ctx = J.Context(0, J.JPEG_ARITH_SSE2)
    ctx.bind_host_to_device()
    in_ptr = L.JPEGB200_hostAlloc(o + 64)
    arr = np.ctypeslib.as_array(C.cast(in_ptr, C.POINTER(C.c_ubyte)), shape=(o + 64,))
    for i in range(n):
        arr[offs[i]:offs[i] + sizes[i]] = np.frombuffer(jp[i % 16], dtype=np.uint8)
    ptrs = [in_ptr + x for x in offs]
    per = 3840 * 2160 * 2
    stride = (per + 255) & ~255
    dev = ctx.device_alloc(stride * n)
    douts = [dev + i * stride for i in range(n)]
    ctx.set_shared_tables(ctx.export_tables(jp[0]))
    for depth in [int(x) for x in os.environ.get("ONECALL_DEPTHS", "2,3,4,6,8").split(",")]:
        ctx.set_pipeline_depth(depth)
        best = None
        for it in range(4):
            t0 = time.time()
            rc, st, cnt = J.decode_batch(ctx, ptrs, sizes, J.RGB565_LITTLE_ENDIAN, 0, douts, None, J.JPEGB200_OUT_DEVICE)
            t1 = time.time()
            assert rc == 1
            ms = 1e3 * (t1 - t0)
            best = ms if best is None else min(best, ms)
        tm, jobs = ctx.last_call_timings()
        print("job_mb %s depth %d: %.1f ms per call (%d images, %d jobs, h2d %.0f MB) dev sums: %s" % (
            os.environ.get("JPEGDEC_B200_JOB_MB", "192"), depth, best, n, jobs, cnt["h2d_bytes"] / 1e6, {k: round(v, 1) for k, v in tm.items()}), flush=True)
main()
