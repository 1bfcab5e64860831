"""The drop-in boundary from C and C++: programs written like the reference's own examples
(linux/examples/c_cmdline/main.c, linux/examples/jpeg_perf_test/main.cpp) compile against include/JPEGDEC.h and
link against libjpegdec_b200.so.  CPU tier: they build, open() works, decode fails loudly without a GPU.
GPU tier: their output is bit-exact with the compiled reference."""
import os
import subprocess

import numpy as np
import pytest

import jpegdec_b200 as J
from tests import common as T

HERE = os.path.join(T.ROOT, "tests", "c_api")
OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(T.ROOT, "jpegdec_b200")


def _build():
    os.makedirs(OUT, exist_ok=True)
    common = ["-O2", "-Wall", "-I" + os.path.join(T.ROOT, "include")]
    link = ["-L" + LIBDIR, "-ljpegdec_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(["gcc"] + common + [os.path.join(HERE, "accept.c"), "-o", os.path.join(OUT, "accept")] + link, check=True)
    subprocess.run(["g++"] + common + [os.path.join(HERE, "accept_cpp.cpp"), "-o", os.path.join(OUT, "accept_cpp")] + link, check=True)


def test_c_and_cpp_callers_build_and_fail_loudly_without_gpu():
    _build()
    img = os.path.join(T.GOLD, "images", "tulips.jpg")
    r = subprocess.run([os.path.join(OUT, "accept"), img, "0", "0"], stdout=subprocess.PIPE, text=True)
    assert "w=640 h=480 sub=0x22" in r.stdout
    if J.lib().JPEGB200_deviceCount() == 0:
        assert r.returncode == 3 and "err=5" in r.stdout          # JPEG_ERROR_MEMORY, nothing decoded on the CPU
    r = subprocess.run([os.path.join(OUT, "accept"), os.path.join(T.GOLD, "images", "corrupt1.jpg"), "0", "0"],
                       stdout=subprocess.PIPE, text=True)
    assert r.returncode == 2                                       # open fails like the reference (JPEG_DECODE_ERROR)


@pytest.mark.gpu
def test_c_and_cpp_callers_match_reference(tmp_path):
    from oracle import refdrv
    _build()
    ref = refdrv.Ref("sse") if refdrv.available("sse") else None
    img = os.path.join(T.GOLD, "images", "tulips.jpg")
    for pt, opt in ((0, 0), (2, 0), (3, 0)):
        out = str(tmp_path / ("o_%d_%d.raw" % (pt, opt)))
        r = subprocess.run([os.path.join(OUT, "accept"), img, str(pt), str(opt), out], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout
        got = np.fromfile(out, dtype=np.uint8)
        if ref is not None:
            rc, err, want, _ = ref.decode_cb(T.image("tulips"), pt, opt, want_log=False)
            assert np.array_equal(got, want.reshape(-1))
        else:
            assert T.sha(got.reshape(480, -1)) == T.digests()["tulips"]["sse/%s/full" % dict(T.PTS)[pt]]["sha"]
    r = subprocess.run([os.path.join(OUT, "accept"), img, "cb", "0"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "callbacks=150 pixels=307200" in r.stdout, r.stdout
    r = subprocess.run([os.path.join(OUT, "accept_cpp"), img], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.count("rc=1") == 4, r.stdout
