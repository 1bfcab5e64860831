#!/usr/bin/env python
"""Progressive fixtures (the reference bundles none): seeded synthetic images written by Pillow with progressive=True,
decoded by the compiled reference (oracle/_ref, both builds) -> tests/golden/images/prog_*.jpg + tests/golden/progressive.json.
The reference decodes only the DC coefficients of the first scan and returns a 1/8-size image (src/jpeg.inl:4964-4966).
Run where /root/reference exists:  python tests/golden/make_progressive_golden.py"""
import hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refdrv
from tests import synth

CASES = {
    "prog_420": dict(w=320, h=240, seed=11, quality=75, subsampling="4:2:0", restart_rows=0),
    "prog_420_dri": dict(w=333, h=251, seed=12, quality=85, subsampling="4:2:0", restart_rows=1),
    "prog_444": dict(w=301, h=203, seed=13, quality=90, subsampling="4:4:4", restart_rows=0),
    "prog_422": dict(w=640, h=360, seed=14, quality=60, subsampling="4:2:2", restart_rows=2),
    "prog_gray": dict(w=257, h=129, seed=15, quality=80, gray=True, restart_rows=0),
}


def main():
    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        w, h, seed = kw.pop("w"), kw.pop("h"), kw.pop("seed")
        data = synth.synth_jpeg(w, h, seed, progressive=True, **kw)
        open(os.path.join(HERE, "images", name + ".jpg"), "wb").write(data)
        ent = {"w": w, "h": h, "bytes": len(data)}
        for mode in ("sse", "scalar"):
            ref = refdrv.Ref(mode)
            for pt, ptn in ((0, "565le"), (1, "565be"), (2, "8888")):
                if kw.get("gray") and pt == 2:
                    continue
                for opt in (0, 8):
                    rc, err, img, _ = ref.decode_cb(data, pt, opt, want_log=False)
                    assert rc == 1 and img.shape[0] == (h + 7) // 8, (name, mode, pt, opt, rc, err)
                    ent["%s/%s/opt%d" % (mode, ptn, opt)] = {"shape": list(img.shape), "sha": hashlib.sha256(img.tobytes()).hexdigest()[:16]}
        out[name] = ent
    json.dump(out, open(os.path.join(HERE, "progressive.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "fixtures")


main()
