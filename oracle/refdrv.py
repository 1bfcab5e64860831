"""ctypes driver for the compiled, unmodified reference (oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by tests/golden/make_golden.py,
by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
arm.  Never imported by the jpegdec_b200 package.

The shared objects are built by oracle/Makefile from the read-only reference
tree (ref_shim.c includes src/jpeg.inl the way linux/examples/c_cmdline/main.c:10-11
does).  Two builds exist because the reference has two arithmetic paths on
x86-64 (src/jpeg.inl:49-55): 'sse' (default flags) and 'scalar' (-DNO_SIMD).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# pixel types (reference src/JPEGDEC.h:102-111)
RGB565_LITTLE_ENDIAN, RGB565_BIG_ENDIAN, RGB8888, EIGHT_BIT_GRAYSCALE, \
    FOUR_BIT_DITHERED, TWO_BIT_DITHERED, ONE_BIT_DITHERED = range(7)
# options (reference src/JPEGDEC.h:68-75)
JPEG_SCALE_HALF, JPEG_SCALE_QUARTER, JPEG_SCALE_EIGHTH = 2, 4, 8
JPEG_EXIF_THUMBNAIL, JPEG_LUMA_ONLY, JPEG_USES_DMA = 32, 64, 128


class RefInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "subsample", "bpp", "orientation", "has_thumb",
        "thumb_w", "thumb_h", "mode", "res_interval", "error")]


class RefDrawRec(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("x", "y", "w", "h", "wused", "bpp", "buf_toggle")]


def bytes_per_pixel_num_den(pixel_type):
    """(bits per pixel)"""
    return {RGB565_LITTLE_ENDIAN: 16, RGB565_BIG_ENDIAN: 16, RGB8888: 32,
            EIGHT_BIT_GRAYSCALE: 8, FOUR_BIT_DITHERED: 4, TWO_BIT_DITHERED: 2,
            ONE_BIT_DITHERED: 1}[pixel_type]


def scale_shift(options):
    if options & JPEG_SCALE_HALF:
        return 1
    if options & JPEG_SCALE_QUARTER:
        return 2
    if options & JPEG_SCALE_EIGHTH:
        return 3
    return 0


def available(mode="sse"):
    return os.path.exists(os.path.join(HERE, "_ref", "libjpegdec_ref_%s.so" % mode))


class Ref:
    """One compiled build of the reference ('sse' or 'scalar')."""

    def __init__(self, mode="sse"):
        path = os.path.join(HERE, "_ref", "libjpegdec_ref_%s.so" % mode)
        if not os.path.exists(path):
            raise FileNotFoundError(
                path + " missing: run `make -C oracle ref` where /root/reference exists")
        self.mode = mode
        self.lib = L = C.CDLL(path)
        L.ref_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(RefInfo)]
        L.ref_decode_fb.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.POINTER(C.c_int)]
        L.ref_decode_cb.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(RefDrawRec), C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]
        L.ref_decode_dither.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_int,
                                        C.POINTER(RefDrawRec), C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]
        L.ref_decode_batch.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                       C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(C.c_double)]
        assert bool(L.ref_is_simd()) == (mode == "sse")

    def info(self, data):
        inf = RefInfo()
        rc = self.lib.ref_info(data, len(data), C.byref(inf))
        return rc, inf

    def decode_cb(self, data, pixel_type=0, options=0, xoff=0, yoff=0, crop=None,
                  max_mcus=0, abort_after=0, want_log=True):
        """Decode through the draw callback.  Returns (rc, err, image ndarray (tight,
        uint8 [h, pitch_bytes]), log list of tuples)."""
        rc0, inf = self.info(data)
        if not rc0:
            return -1, inf.error, None, []
        s = scale_shift(options)
        if inf.mode == 0xC2:
            s = 3
        pt = pixel_type
        if (options & JPEG_LUMA_ONLY) and pt < EIGHT_BIT_GRAYSCALE:
            pt = EIGHT_BIT_GRAYSCALE
        w, h = inf.width, inf.height
        if options & JPEG_EXIF_THUMBNAIL:
            w, h = inf.thumb_w, inf.thumb_h
        ow, oh = (w + (1 << s) - 1) >> s, (h + (1 << s) - 1) >> s
        if crop is not None:
            ow, oh = crop[2] + 32, crop[3] + 32  # snapped-up crop; caller slices
        bpp = bytes_per_pixel_num_den(pt)
        pitch = (ow * bpp + 7) // 8
        out = np.zeros((oh, pitch), dtype=np.uint8)
        cap = 1 << 16
        log = (RefDrawRec * cap)() if want_log else None
        n = C.c_int(0)
        err = C.c_int(0)
        cx, cy, cw, ch = crop if crop is not None else (0, 0, 0, 0)
        rc = self.lib.ref_decode_cb(data, len(data), pixel_type, options, xoff, yoff,
                                    cx, cy, cw, ch, max_mcus, abort_after,
                                    out.ctypes.data, pitch, oh,
                                    log, cap if want_log else 0, C.byref(n), C.byref(err))
        recs = []
        if want_log:
            recs = [(r.x, r.y, r.w, r.h, r.wused, r.bpp, r.buf_toggle)
                    for r in log[:min(n.value, cap)]]
        return rc, err.value, out, recs

    def decode_fb(self, data, pixel_type=0, options=0, crop=None):
        """Framebuffer-mode decode.  Returns (rc, err, ndarray [rows, pitch_bytes]) where
        the buffer covers whole MCU rows (rows = ceil to 16 of the height + slack)."""
        rc0, inf = self.info(data)
        if not rc0:
            return -1, inf.error, None
        bpp = bytes_per_pixel_num_den(pixel_type if not (options & JPEG_LUMA_ONLY) else 3)
        w = inf.width if crop is None else crop[2] + 32
        rows = ((inf.height + 15) // 16) * 16 + 32
        pitch_px = inf.width if crop is None else None
        # pitch = iCropCX pixels (src/jpeg.inl:5116): whole width when not cropped
        buf = np.zeros(((rows + 2) * max(w, inf.width) * bpp // 8 + 4096,), dtype=np.uint8)
        err = C.c_int(0)
        cx, cy, cw, ch = crop if crop is not None else (0, 0, 0, 0)
        rc = self.lib.ref_decode_fb(data, len(data), pixel_type, options, cx, cy, cw, ch,
                                    buf.ctypes.data, C.byref(err))
        return rc, err.value, buf

    def decode_dither(self, data, pixel_type=ONE_BIT_DITHERED, options=0):
        rc0, inf = self.info(data)
        if not rc0:
            return -1, inf.error, None, []
        s = scale_shift(options)
        ow, oh = (inf.width + (1 << s) - 1) >> s, (inf.height + (1 << s) - 1) >> s
        bpp = bytes_per_pixel_num_den(pixel_type)
        pitch = ((ow + 31) * bpp + 7) // 8
        out = np.zeros((oh, pitch), dtype=np.uint8)
        cap = 1 << 14
        log = (RefDrawRec * cap)()
        n = C.c_int(0)
        err = C.c_int(0)
        rc = self.lib.ref_decode_dither(data, len(data), pixel_type, options,
                                        out.ctypes.data, pitch, oh, log, cap, C.byref(n),
                                        C.byref(err))
        recs = [(r.x, r.y, r.w, r.h, r.wused, r.bpp, r.buf_toggle) for r in log[:n.value]]
        return rc, err.value, out, recs

    def decode_batch(self, datas, pixel_type, options, nthreads, framebuffers=None):
        """Multi-threaded CPU decode of a batch.  Returns (fails, seconds)."""
        n = len(datas)
        arr = (C.c_char_p * n)(*datas)
        lens = (C.c_int * n)(*[len(d) for d in datas])
        fbs = None
        if framebuffers is not None:
            fbs = (C.c_void_p * n)(*[fb.ctypes.data for fb in framebuffers])
        secs = C.c_double(0)
        fails = self.lib.ref_decode_batch(arr, lens, fbs, n, pixel_type, options, nthreads,
                                          C.byref(secs))
        return fails, secs.value
