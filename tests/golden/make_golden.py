#!/usr/bin/env python
"""Generate tests/golden/ from the reference tree (run HERE, where /root/reference exists).

1. Extracts the JPEG byte streams the reference ships as C arrays (test_images/*.h,
   examples/**.h, MacOS/JPEGDEC_Test/**/corrupt*.h) into tests/golden/images/*.jpg --
   they are *data fixtures* (the reference's own test inputs), not source code.
2. Runs the compiled reference (oracle/_ref, both arithmetic builds) on every image x
   pixel type x scale and records SHA-256 digests of the tight callback-assembled image
   plus the draw-callback log digest into tests/golden/digests.json.
3. Stores a few small raw golden frames (tests/golden/frames/*.bin) for pixel-level
   diffs on the GPU box, where neither /root/reference nor (necessarily) the _ref
   build exist.

The GPU box never runs this script; it only reads the committed outputs.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refdrv  # noqa: E402

REF = "/root/reference"
SOURCES = {
    "tulips": "test_images/tulips.h",
    "st_peters": "test_images/st_peters.h",
    "sciopero": "test_images/sciopero.h",
    "zebra": "test_images/zebra.h",
    "thumb_test": "test_images/thumb_test.h",
    "croptest": "examples/crop_area/croptest.h",
    "octocat_small": "examples/jpegdisplay_demo/octocat_small.h",
    "batman": "examples/M5Stack/M5Stack/batman.h",
    "ncc1701": "examples/M5Stack/M5Stack/ncc1701.h",
    "lange": "examples/epd_demo/lange.h",
    "corrupt1": "MacOS/JPEGDEC_Test/JPEGDEC_Test/corrupt1.h",
    "corrupt2": "MacOS/JPEGDEC_Test/JPEGDEC_Test/corrupt2.h",
    "corrupt3": "MacOS/JPEGDEC_Test/JPEGDEC_Test/corrupt3.h",
    "corrupt4": "MacOS/JPEGDEC_Test/JPEGDEC_Test/corrupt4.h",
    "corrupt5": "MacOS/JPEGDEC_Test/JPEGDEC_Test/corrupt5.h",
}
VALID = ["tulips", "st_peters", "sciopero", "zebra", "croptest", "octocat_small",
         "batman", "ncc1701", "lange"]


def extract(path):
    txt = open(path, "r", errors="replace").read()
    a = txt.index("{")
    body = txt[a + 1: txt.index("}", a)]  # first array only (lange.h holds several)
    body = re.sub(r"//[^\n]*", "", body)
    vals = re.findall(r"0[xX][0-9a-fA-F]+|\d+", body)
    return bytes(int(v, 0) & 0xFF for v in vals)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    os.makedirs(os.path.join(HERE, "images"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "frames"), exist_ok=True)
    blobs = {}
    for name, rel in SOURCES.items():
        data = extract(os.path.join(REF, rel))
        assert data[:2] == b"\xff\xd8", name
        blobs[name] = data
        with open(os.path.join(HERE, "images", name + ".jpg"), "wb") as f:
            f.write(data)
        print("%-14s %7d bytes" % (name, len(data)))

    refs = {m: refdrv.Ref(m) for m in ("sse", "scalar")}
    digests = {}
    P = refdrv
    ptypes = [(P.RGB565_LITTLE_ENDIAN, "rgb565le"), (P.RGB565_BIG_ENDIAN, "rgb565be"),
              (P.RGB8888, "rgb8888"), (P.EIGHT_BIT_GRAYSCALE, "gray8")]
    scales = [(0, "full"), (P.JPEG_SCALE_HALF, "half"), (P.JPEG_SCALE_QUARTER, "quarter"),
              (P.JPEG_SCALE_EIGHTH, "eighth")]
    for name in VALID:
        data = blobs[name]
        rc, inf = refs["sse"].info(data)
        digests[name] = {"info": {k: getattr(inf, k) for k, _ in inf._fields_}, "open": rc}
        if not rc:
            continue
        for mode, ref in refs.items():
            for pt, ptn in ptypes:
                for opt, sn in scales:
                    rc, err, img, log = ref.decode_cb(data, pt, opt)
                    key = "%s/%s/%s" % (mode, ptn, sn)
                    digests[name][key] = {
                        "rc": rc, "err": err, "sha": sha(img), "shape": list(img.shape),
                        "ncb": len(log), "logsha": hashlib.sha256(
                            json.dumps(log).encode()).hexdigest()[:16]}
            # dithered (decodeDither) full size
            for pt, ptn in [(P.ONE_BIT_DITHERED, "dither1"), (P.TWO_BIT_DITHERED, "dither2"),
                            (P.FOUR_BIT_DITHERED, "dither4")]:
                rc, err, img, log = ref.decode_dither(data, pt, 0)
                digests[name]["%s/%s/full" % (mode, ptn)] = {
                    "rc": rc, "err": err, "sha": sha(img), "shape": list(img.shape),
                    "ncb": len(log)}
            # luma only
            rc, err, img, log = ref.decode_cb(data, P.RGB565_LITTLE_ENDIAN, P.JPEG_LUMA_ONLY)
            digests[name]["%s/lumaonly/full" % mode] = {"rc": rc, "err": err, "sha": sha(img),
                                                        "shape": list(img.shape)}
    # crop known-answer (reference test 2: MacOS/JPEGDEC_Test/JPEGDEC_Test/main.cpp:106-137)
    for mode, ref in refs.items():
        rc, err, img, log = ref.decode_cb(blobs["tulips"], 0, 0, crop=(50, 50, 125, 170))
        img = img[:176, :128 * 2]
        digests["tulips"]["%s/rgb565le/crop_50_50_125_170" % mode] = {
            "rc": rc, "sha": sha(img), "shape": list(img.shape), "ncb": len(log)}
    # corrupt files: what the reference returns
    for name in ["corrupt1", "corrupt2", "corrupt3", "corrupt4", "corrupt5", "thumb_test"]:
        rc, inf = refs["sse"].info(blobs[name])
        digests[name] = {"open": rc, "info": {k: getattr(inf, k) for k, _ in inf._fields_}}
    # EXIF thumbnail known-answer (reference test 10, main.cpp:236-260)
    for mode, ref in refs.items():
        rc, err, img, log = ref.decode_cb(blobs["thumb_test"], 0, P.JPEG_EXIF_THUMBNAIL)
        digests["thumb_test"]["%s/rgb565le/exif_thumb" % mode] = {
            "rc": rc, "err": err, "sha": sha(img), "shape": list(img.shape), "ncb": len(log)}

    # raw frames for pixel-level diffs (small ones only)
    for name in ["tulips", "ncc1701"]:
        for mode, ref in refs.items():
            for pt, ptn in [(P.RGB565_LITTLE_ENDIAN, "rgb565le"), (P.RGB8888, "rgb8888")]:
                if name == "tulips" and (mode, ptn) != ("sse", "rgb565le"):
                    continue  # keep the committed fixtures small
                rc, err, img, log = ref.decode_cb(blobs[name], pt, 0)
                img.tofile(os.path.join(HERE, "frames", "%s_%s_%s_full.bin" % (name, mode, ptn)))

    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    # cross-check against the survey-time digests
    expect = {
        ("tulips", "sse/rgb565le/full"): "ac3b5ca6e3c8b405",
        ("tulips", "scalar/rgb565le/full"): "772ff2897a08d2c7",
        ("tulips", "sse/rgb565le/half"): "ec11c03d434bb77b",
        ("tulips", "sse/rgb565le/quarter"): "7f4ff7880b4fab1f",
        ("tulips", "sse/rgb565le/eighth"): "4fc77b98c8054fe6",
        ("tulips", "sse/rgb8888/full"): "ce60e5a67dc70170",
        ("tulips", "scalar/rgb8888/full"): "ddc23693636b7a4b",
        ("tulips", "sse/gray8/full"): "dce584b7c35f9173",
        ("sciopero", "sse/rgb565le/full"): "212ef1e85f7f7d6f",
        ("sciopero", "scalar/rgb565le/full"): "fce422abca5925c3",
        ("zebra", "sse/rgb565le/full"): "0df95f2f5aa46e24",
        ("zebra", "scalar/rgb8888/full"): "19a4d43e665e7e49",
    }
    bad = 0
    for (n, k), v in expect.items():
        got = digests[n][k]["sha"]
        ok = got == v
        bad += not ok
        print("%-10s %-24s %s %s" % (n, k, got, "OK" if ok else "MISMATCH (survey %s)" % v))
    print("survey digest cross-check: %d mismatches" % bad)


if __name__ == "__main__":
    main()
