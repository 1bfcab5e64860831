"""Seeded synthetic baseline JPEGs: image = 0.55*F64 + 0.30*F16 + 0.15*F4 + N(0,4),
Fs = uniform random RGB field of (H/s+2)x(W/s+2) bicubically upsampled to WxH; standard Huffman
tables, libjpeg quality scaling, DRI = one MCU row.  Used by tests and bench.py (test/bench
infrastructure, not part of the product)."""
import io
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
from PIL import Image


def synth_pixels(w, h, seed):
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w, 3), np.float32)
    for s, wt in ((64, 0.55), (16, 0.30), (4, 0.15)):
        f = rng.integers(0, 256, size=(h // s + 2, w // s + 2, 3), dtype=np.uint8)
        acc += wt * np.asarray(Image.fromarray(f).resize((w, h), Image.BICUBIC), dtype=np.float32)
    acc += rng.normal(0, 4, size=acc.shape).astype(np.float32)
    return np.clip(acc, 0, 255).astype(np.uint8)


def synth_jpeg(w, h, seed, quality=75, subsampling="4:2:0", gray=False, restart_rows=1, optimize=False, progressive=False):
    img = Image.fromarray(synth_pixels(w, h, seed))
    if gray:
        img = img.convert("L")
    b = io.BytesIO()
    kw = dict(quality=quality, optimize=optimize)
    if not gray:
        kw["subsampling"] = subsampling
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    if progressive:
        kw["progressive"] = True   # libjpeg's default script: first scan = interleaved DC of all components, Al = 1
    img.save(b, "JPEG", **kw)
    return b.getvalue()


def _job(args):
    return synth_jpeg(*args[0], **args[1])


def synth_set(n, w, h, quality=75, subsampling="4:2:0", gray=False, seed0=0, workers=None, restart_rows=1):
    """n unique images (seed = index), generated on all host cores."""
    jobs = [((w, h, seed0 + i), dict(quality=quality, subsampling=subsampling, gray=gray, restart_rows=restart_rows))
            for i in range(n)]
    workers = workers or min(os.cpu_count() or 1, 64)
    if n <= 2 or workers <= 1:
        return [_job(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_job, jobs, chunksize=max(1, n // (workers * 4))))
