"""CPU tier: the host C code (header parser, table builders, JPEG_open*/getters/crop/close) under ASan + UBSan on mutated
and truncated files -- the host-side half of the reference's fuzz tests (MacOS/JPEGDEC_Test/JPEGDEC_Test/main.cpp:262-300;
the GPU half runs in tests/test_c_api.py through the reference's own harness)."""
import os
import subprocess

import pytest

from tests import common as T

HERE = os.path.join(T.ROOT, "tests", "fuzz")
OUT = os.path.join(HERE, "_build")
CS = os.path.join(T.ROOT, "jpegdec_b200", "csrc")
FILES = ["tulips", "lange", "thumb_test", "prog_420", "corrupt1", "corrupt2", "sciopero", "croptest"]


def _build(name, sources, extra):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    cmd = ["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I" + os.path.join(T.ROOT, "include"), "-I" + CS] + sources + ["-o", exe, "-lm", "-lpthread"] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available here: " + r.stdout[-300:])
    return exe


@pytest.mark.parametrize("which", ["parse", "api"])
def test_host_code_survives_mutated_files_under_sanitizers(which):
    imgs = [os.path.join(T.GOLD, "images", n + ".jpg") for n in FILES]
    if which == "parse":
        exe = _build("fuzz_parse", [os.path.join(HERE, "fuzz_parse.c"), os.path.join(CS, "jd_host.c")], [])
        iters = "1500"
    else:
        exe = _build("fuzz_api", [os.path.join(HERE, "fuzz_api.c"), os.path.join(CS, "jd_host.c"), os.path.join(CS, "jd_api.c")],
                     ["-Wl,--unresolved-symbols=ignore-all"])
        iters = "800"
    r = subprocess.run([exe, iters] + imgs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "cases" in r.stdout and "ERROR" not in r.stdout and "runtime error" not in r.stdout, r.stdout[-2000:]


def test_device_decode_code_survives_corrupt_scans_under_sanitizers():
    """The kernels' per-thread decode code (stepped on the CPU by tests/hostsim) on corrupted and truncated scans."""
    os.makedirs(OUT, exist_ok=True)
    flags = ["-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
             "-I" + os.path.join(T.ROOT, "include"), "-I" + CS]
    objs = []
    for cc, src, extra in (("g++", os.path.join(T.ROOT, "tests", "hostsim", "hostsim.cpp"), ["-std=c++17", "-w"]),
                           ("gcc", os.path.join(CS, "jd_host.c"), []), ("gcc", os.path.join(HERE, "fuzz_sim.c"), [])):
        o = os.path.join(OUT, os.path.basename(src) + ".o")
        r = subprocess.run([cc] + flags + extra + ["-c", src, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            pytest.skip("sanitizer build not available here: " + r.stdout[-300:])
        objs.append(o)
    exe = os.path.join(OUT, "fuzz_sim")
    r = subprocess.run(["g++", "-fsanitize=address,undefined"] + objs + ["-o", exe, "-lm"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer link not available here: " + r.stdout[-300:])
    imgs = [os.path.join(T.GOLD, "images", n + ".jpg") for n in ("tulips", "sciopero", "croptest", "zebra", "lange", "ncc1701")]
    r = subprocess.run([exe, "120"] + imgs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "cases 720" in r.stdout and "ERROR" not in r.stdout and "runtime error" not in r.stdout, r.stdout[-2000:]
