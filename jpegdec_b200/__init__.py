"""jpegdec_b200 -- B200-native baseline JPEG decoder behind JPEGDEC's API.

This package is a thin ctypes mirror of the C ABI in include/JPEGDEC.h (the drop-in
API of bitbank2/JPEGDEC: reference src/JPEGDEC.h:249-309) and include/jpegdec_b200.h
(the batch / device-resident extension).  All decode work runs in the hand-written
sm_100a kernels inside libjpegdec_b200.so; there is no CPU fallback -- if the shared
library or a CUDA device is missing, calls fail loudly.

    from jpegdec_b200 import JPEGDEC, BatchDecoder, RGB565_LITTLE_ENDIAN
    j = JPEGDEC(); j.openRAM(data, draw_cb); j.decode(0, 0, 0); j.close()
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JPEGDEC_B200_LIB") or os.path.join(HERE, "libjpegdec_b200.so")  # env override: A/B of kernel variants

# --- constants (reference src/JPEGDEC.h:68-75, :102-111, :119-126) ---
JPEG_AUTO_ROTATE, JPEG_SCALE_HALF, JPEG_SCALE_QUARTER, JPEG_SCALE_EIGHTH = 1, 2, 4, 8
JPEG_LE_PIXELS, JPEG_EXIF_THUMBNAIL, JPEG_LUMA_ONLY, JPEG_USES_DMA = 16, 32, 64, 128
(RGB565_LITTLE_ENDIAN, RGB565_BIG_ENDIAN, RGB8888, EIGHT_BIT_GRAYSCALE, FOUR_BIT_DITHERED,
 TWO_BIT_DITHERED, ONE_BIT_DITHERED, INVALID_PIXEL_TYPE) = range(8)
(JPEG_SUCCESS, JPEG_INVALID_PARAMETER, JPEG_DECODE_ERROR, JPEG_UNSUPPORTED_FEATURE,
 JPEG_INVALID_FILE, JPEG_ERROR_MEMORY) = range(6)
JPEG_ARITH_SSE2, JPEG_ARITH_SCALAR = 0, 1
JPEGB200_OUT_DEVICE = 1
TIMING_NAMES = ["h2d", "prescan", "entropy", "stitch", "idct", "dither", "d2h", "total"]
COUNTER_NAMES = ["launches", "segments", "blocks", "events", "compressed_bytes", "output_bytes",
                 "record_bytes", "h2d_bytes", "d2h_bytes", "event_candidates"]
TABLE_BLOB_BYTES = 10496 * 2 + 3 * 64 * 2 + 16


class JPEGDRAW(C.Structure):
    """reference src/JPEGDEC.h:143-151"""
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("iWidth", C.c_int), ("iHeight", C.c_int),
                ("iWidthUsed", C.c_int), ("iBpp", C.c_int), ("pPixels", C.c_void_p),
                ("pUser", C.c_void_p)]


DRAW_CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(JPEGDRAW))

_lib = None


def lib():
    """Load libjpegdec_b200.so (build it with `python -m jpegdec_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libjpegdec_b200.so not built (run `python -m jpegdec_b200.build`); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ip, i32p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int32)
    L.JPEG_sizeofImage.restype = C.c_int
    L.JPEG_openRAM.argtypes = [vp, vp, C.c_int, vp]
    L.JPEG_openFile.argtypes = [vp, C.c_char_p, vp]
    for name in ("JPEG_getWidth", "JPEG_getHeight", "JPEG_getLastError", "JPEG_getOrientation",
                 "JPEG_getBpp", "JPEG_getSubSample", "JPEG_getJPEGType", "JPEG_hasThumb",
                 "JPEG_getThumbWidth", "JPEG_getThumbHeight", "JPEG_getPixelType"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.JPEG_close.argtypes = [vp]
    L.JPEG_close.restype = None
    L.JPEG_decode.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.JPEG_decodeDither.argtypes = [vp, vp, C.c_int]
    L.JPEG_setFramebuffer.argtypes = [vp, vp]
    L.JPEG_setCropArea.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.JPEG_getCropArea.argtypes = [vp, ip, ip, ip, ip]
    for name in ("JPEG_setPixelType", "JPEG_setMaxOutputSize", "JPEG_setArithMode", "JPEG_setDevice"):
        getattr(L, name).argtypes = [vp, C.c_int]
        getattr(L, name).restype = None
    L.JPEG_setUserPointer.argtypes = [vp, vp]
    L.JPEG_setUserPointer.restype = None
    L.JPEG_setFramebuffer.restype = None
    L.JPEG_setCropArea.restype = None
    L.JPEG_getCropArea.restype = None
    # batch API
    L.JPEGB200_create.argtypes = [C.c_int, C.c_int]
    L.JPEGB200_create.restype = vp
    L.JPEGB200_destroy.argtypes = [vp]
    L.JPEGB200_destroy.restype = None
    L.JPEGB200_lastErrorString.argtypes = [vp]
    L.JPEGB200_lastErrorString.restype = C.c_char_p
    L.JPEGB200_deviceCount.restype = C.c_int
    L.JPEGB200_hostAlloc.argtypes = [C.c_size_t]
    L.JPEGB200_hostAlloc.restype = vp
    L.JPEGB200_hostFree.argtypes = [vp]
    L.JPEGB200_hostFree.restype = None
    L.JPEGB200_batchCreate.argtypes = [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, C.c_int]
    L.JPEGB200_batchCreate.restype = vp
    L.JPEGB200_batchDestroy.argtypes = [vp]
    L.JPEGB200_batchDestroy.restype = None
    L.JPEGB200_batchCount.argtypes = [vp]
    L.JPEGB200_batchImageInfo.argtypes = [vp, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p]
    L.JPEGB200_batchOutputBytes.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.JPEGB200_batchOutputBytes.restype = C.c_int64
    L.JPEGB200_batchSetOutput.argtypes = [vp, C.c_int, vp, C.c_int64]
    L.JPEGB200_batchAllocDeviceOutput.argtypes = [vp]
    L.JPEGB200_batchGetDeviceOutput.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.JPEGB200_batchReadOutput.argtypes = [vp, C.c_int, vp]
    L.JPEGB200_batchUpload.argtypes = [vp]
    L.JPEGB200_batchDecode.argtypes = [vp, C.c_int]
    L.JPEGB200_batchDownload.argtypes = [vp]
    L.JPEGB200_batchWait.argtypes = [vp, i32p]
    L.JPEGB200_batchGetTimings.argtypes = [vp, C.POINTER(C.c_float)]
    L.JPEGB200_batchGetCounters.argtypes = [vp, C.POINTER(C.c_int64)]
    L.JPEGB200_batchStream.argtypes = [vp]
    L.JPEGB200_batchStream.restype = vp
    L.JPEGB200_decodeBatch.argtypes = [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(vp), C.POINTER(C.c_int64), C.c_int, i32p]
    L.JPEGB200_lastCallCounters.argtypes = [vp, C.POINTER(C.c_int64)]
    L.JPEGB200_lastCallTimings.argtypes = [vp, C.POINTER(C.c_float), ip]
    L.JPEGB200_setPipelineDepth.argtypes = [vp, C.c_int]
    L.JPEGB200_numaNode.argtypes = [vp]
    L.JPEGB200_bindHostToDevice.argtypes = [vp]
    L.JPEGB200_deviceAlloc.argtypes = [vp, C.c_size_t]
    L.JPEGB200_deviceAlloc.restype = vp
    L.JPEGB200_deviceFree.argtypes = [vp, vp]
    L.JPEGB200_deviceFree.restype = None
    L.JPEGB200_deviceRead.argtypes = [vp, vp, vp, C.c_size_t]
    L.JPEGB200_digestDevice.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_uint64)]
    L.JPEGB200_exportTables.argtypes = [C.c_char_p, C.c_int, vp]
    L.JPEGB200_setSharedTables.argtypes = [vp, vp]
    L.JPEGB200_sharedTableHits.argtypes = [vp]
    _lib = L
    return L


def bits_per_pixel(pixel_type):
    return {RGB565_LITTLE_ENDIAN: 16, RGB565_BIG_ENDIAN: 16, RGB8888: 32, EIGHT_BIT_GRAYSCALE: 8,
            FOUR_BIT_DITHERED: 4, TWO_BIT_DITHERED: 2, ONE_BIT_DITHERED: 1}[pixel_type]


class JPEGDEC:
    """Mirror of the reference's C++ class (src/JPEGDEC.h:249-287): same method names, argument
    meaning and 1/0 return convention; getLastError() gives the reference's error codes."""

    def __init__(self):
        L = lib()
        self._img = C.create_string_buffer(L.JPEG_sizeofImage())
        self._p = C.cast(self._img, C.c_void_p)
        self._data = None
        self._cb = None
        self._keep = []

    def _wrap_cb(self, draw):
        if draw is None:
            return None
        self._cb = DRAW_CALLBACK(lambda pd: int(draw(pd.contents)))
        return C.cast(self._cb, C.c_void_p)

    def openRAM(self, data, draw=None):
        self._data = (C.c_ubyte * len(data)).from_buffer_copy(bytes(data))
        return lib().JPEG_openRAM(self._p, C.cast(self._data, C.c_void_p), len(data), self._wrap_cb(draw))

    openFLASH = openRAM

    def open(self, filename, draw=None):
        return lib().JPEG_openFile(self._p, filename.encode(), self._wrap_cb(draw))

    def close(self):
        lib().JPEG_close(self._p)

    def setFramebuffer(self, buf):
        """buf: numpy array (host) that must cover whole MCU rows, as in the reference."""
        self._keep.append(buf)
        lib().JPEG_setFramebuffer(self._p, buf.ctypes.data if buf is not None else None)

    def setCropArea(self, x, y, w, h):
        lib().JPEG_setCropArea(self._p, x, y, w, h)

    def getCropArea(self):
        v = [C.c_int() for _ in range(4)]
        lib().JPEG_getCropArea(self._p, *[C.byref(i) for i in v])
        return tuple(i.value for i in v)

    def decode(self, x, y, options):
        return lib().JPEG_decode(self._p, x, y, options)

    def decodeDither(self, dither_buf, options):
        self._keep.append(dither_buf)
        return lib().JPEG_decodeDither(self._p, dither_buf.ctypes.data, options)

    def getOrientation(self): return lib().JPEG_getOrientation(self._p)
    def getWidth(self): return lib().JPEG_getWidth(self._p)
    def getHeight(self): return lib().JPEG_getHeight(self._p)
    def getBpp(self): return lib().JPEG_getBpp(self._p)
    def getSubSample(self): return lib().JPEG_getSubSample(self._p)
    def getJPEGType(self): return lib().JPEG_getJPEGType(self._p)
    def hasThumb(self): return lib().JPEG_hasThumb(self._p)
    def getThumbWidth(self): return lib().JPEG_getThumbWidth(self._p)
    def getThumbHeight(self): return lib().JPEG_getThumbHeight(self._p)
    def getLastError(self): return lib().JPEG_getLastError(self._p)
    def setPixelType(self, t): lib().JPEG_setPixelType(self._p, t)
    def getPixelType(self): return lib().JPEG_getPixelType(self._p)
    def setMaxOutputSize(self, n): lib().JPEG_setMaxOutputSize(self._p, n)
    def setArithMode(self, m): lib().JPEG_setArithMode(self._p, m)
    def setDevice(self, d): lib().JPEG_setDevice(self._p, d)


class Context:
    """One per (process, GPU)."""

    def __init__(self, device=-1, arith=JPEG_ARITH_SSE2):
        self.h = lib().JPEGB200_create(device, arith)
        if not self.h:
            raise RuntimeError("JPEGB200_create failed: " + lib().JPEGB200_lastErrorString(None).decode())

    def close(self):
        if self.h:
            lib().JPEGB200_destroy(self.h)
            self.h = None

    def export_tables(self, jpeg):
        blob = np.zeros(TABLE_BLOB_BYTES, dtype=np.uint8)
        if not lib().JPEGB200_exportTables(bytes(jpeg), len(jpeg), blob.ctypes.data):
            raise RuntimeError("exportTables: header parse failed")
        return blob

    def set_shared_tables(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        return lib().JPEGB200_setSharedTables(self.h, blob.ctypes.data)

    def shared_table_hits(self):
        return lib().JPEGB200_sharedTableHits(self.h)

    def numa_node(self):
        return lib().JPEGB200_numaNode(self.h)

    def bind_host_to_device(self):
        """Pin the calling thread to the CPUs next to this context's GPU (see include/jpegdec_b200.h)."""
        return lib().JPEGB200_bindHostToDevice(self.h)

    def set_pipeline_depth(self, jobs):
        return lib().JPEGB200_setPipelineDepth(self.h, jobs)

    def last_call_timings(self):
        ms = (C.c_float * len(TIMING_NAMES))()
        jobs = C.c_int()
        lib().JPEGB200_lastCallTimings(self.h, ms, C.byref(jobs))
        return dict(zip(TIMING_NAMES, list(ms))), jobs.value

    def device_alloc(self, nbytes):
        p = lib().JPEGB200_deviceAlloc(self.h, nbytes)
        if not p:
            raise RuntimeError("deviceAlloc(%d) failed" % nbytes)
        return p

    def device_free(self, p):
        lib().JPEGB200_deviceFree(self.h, p)

    def device_read(self, dev_ptr, nbytes):
        o = np.empty(nbytes, dtype=np.uint8)
        if not lib().JPEGB200_deviceRead(self.h, o.ctypes.data, dev_ptr, nbytes):
            raise RuntimeError("deviceRead failed: " + lib().JPEGB200_lastErrorString(self.h).decode())
        return o

    def digest_device(self, ptrs, lengths):
        """64-bit digests of device byte ranges, computed on the GPU (same function as digest_host)."""
        n = len(ptrs)
        pa = (C.c_void_p * n)(*ptrs)
        la = (C.c_int64 * n)(*lengths)
        out = (C.c_uint64 * n)()
        if not lib().JPEGB200_digestDevice(self.h, pa, la, n, out):
            raise RuntimeError("digestDevice failed: " + lib().JPEGB200_lastErrorString(self.h).decode())
        return list(out)


def digest_host(a):
    """The digest JPEGB200_digestDevice computes, for a host array (include/jpegdec_b200.h)."""
    b = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    if b.size % 8:
        b = np.concatenate([b, np.zeros(8 - b.size % 8, dtype=np.uint8)])
    w = b.view("<u8")
    with np.errstate(over="ignore"):
        z = w ^ (np.arange(w.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))


class Batch:
    """A decode job over n JPEG files that live in host memory at (ptr, size) pairs."""

    def __init__(self, ctx, ptrs, sizes, pixel_type, options=0):
        n = len(ptrs)
        self.n = n
        self._ptrs = (C.c_void_p * n)(*ptrs)
        self._sizes = (C.c_int32 * n)(*sizes)
        self.ctx = ctx
        self.h = lib().JPEGB200_batchCreate(ctx.h, self._ptrs, self._sizes, n, pixel_type, options)
        if not self.h:
            raise RuntimeError("batchCreate failed: " + lib().JPEGB200_lastErrorString(ctx.h).decode())

    def _ck(self, rc, what):
        if not rc:
            raise RuntimeError(what + " failed: " + lib().JPEGB200_lastErrorString(self.ctx.h).decode())
        return rc

    def info(self, i):
        v = [C.c_int32() for _ in range(6)]
        lib().JPEGB200_batchImageInfo(self.h, i, *[C.byref(x) for x in v])
        return dict(zip(("width", "height", "subsample", "out_w", "out_h", "status"), [x.value for x in v]))

    def output_bytes(self, i):
        p = C.c_int64()
        b = lib().JPEGB200_batchOutputBytes(self.h, i, C.byref(p))
        return b, p.value

    def set_output(self, i, ptr, pitch=0):
        lib().JPEGB200_batchSetOutput(self.h, i, ptr, pitch)

    def alloc_device_output(self):
        self._ck(lib().JPEGB200_batchAllocDeviceOutput(self.h), "batchAllocDeviceOutput")

    def device_output(self, i):
        p, pitch = C.c_void_p(), C.c_int64()
        self._ck(lib().JPEGB200_batchGetDeviceOutput(self.h, i, C.byref(p), C.byref(pitch)), "batchGetDeviceOutput")
        return p.value, pitch.value

    def read_output(self, i):
        nbytes, pitch = self.output_bytes(i)
        o = np.empty(nbytes, dtype=np.uint8)
        self._ck(lib().JPEGB200_batchReadOutput(self.h, i, o.ctypes.data), "batchReadOutput")
        return o.reshape(-1, pitch)

    def upload(self): self._ck(lib().JPEGB200_batchUpload(self.h), "batchUpload")
    def decode(self, flags=0): self._ck(lib().JPEGB200_batchDecode(self.h, flags), "batchDecode")
    def download(self): self._ck(lib().JPEGB200_batchDownload(self.h), "batchDownload")

    def wait(self):
        st = (C.c_int32 * self.n)()
        self._ck(lib().JPEGB200_batchWait(self.h, st), "batchWait")
        return list(st)

    def timings(self):
        ms = (C.c_float * len(TIMING_NAMES))()
        lib().JPEGB200_batchGetTimings(self.h, ms)
        return dict(zip(TIMING_NAMES, list(ms)))

    def counters(self):
        c = (C.c_int64 * len(COUNTER_NAMES))()
        lib().JPEGB200_batchGetCounters(self.h, c)
        return dict(zip(COUNTER_NAMES, list(c)))

    def close(self):
        if self.h:
            lib().JPEGB200_batchDestroy(self.h)
            self.h = None


def decode_batch(ctx, ptrs, sizes, pixel_type, options, outs, pitches=None, flags=0):
    """JPEGB200_decodeBatch: one call for n files (host pointers) -> n outputs (host pointers, or device pointers with
    JPEGB200_OUT_DEVICE).  Returns (rc, per-image status list, counters summed over the internal jobs)."""
    n = len(ptrs)
    pa = (C.c_void_p * n)(*ptrs)
    sa = (C.c_int32 * n)(*sizes)
    oa = (C.c_void_p * n)(*outs)
    pi = (C.c_int64 * n)(*pitches) if pitches is not None else None
    st = (C.c_int32 * n)()
    rc = lib().JPEGB200_decodeBatch(ctx.h, pa, sa, n, pixel_type, options, oa, pi, flags, st)
    cnt = (C.c_int64 * len(COUNTER_NAMES))()
    lib().JPEGB200_lastCallCounters(ctx.h, cnt)
    return rc, list(st), dict(zip(COUNTER_NAMES, list(cnt)))


def decode_batch_to_host(ctx, jpegs, pixel_type, options=0):
    """Convenience: list of bytes -> list of numpy arrays [out_h, pitch_bytes] (uint8).
    One public-API call per batch with HOST buffers on both sides."""
    bufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
    b = Batch(ctx, [x.ctypes.data for x in bufs], [len(x) for x in bufs], pixel_type, options)
    try:
        outs = []
        for i in range(b.n):
            nbytes, pitch = b.output_bytes(i)
            inf = b.info(i)
            if inf["status"] != JPEG_SUCCESS:
                outs.append(None)
                continue
            o = np.zeros((inf["out_h"], pitch), dtype=np.uint8)
            b.set_output(i, o.ctypes.data, pitch)
            outs.append(o)
        b.upload(); b.decode(0); b.download()
        status = b.wait()
        return outs, status, b.timings(), b.counters()
    finally:
        b.close()
