#!/usr/bin/env python
"""Detailed GPU parity report (development aid; the pass/fail tests are tests/test_gpu_*.py).
Compares libjpegdec_b200.so with the compiled reference (oracle/_ref) on the fixtures and on
synthetic images; prints the first mismatching pixels."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpegdec_b200 as J
from oracle import refdrv as P
from tests import synth

def main():
    names = ['tulips', 'sciopero', 'st_peters', 'zebra', 'croptest', 'octocat_small', 'batman', 'ncc1701', 'lange']
    blobs = {n: open(os.path.join(ROOT, 'tests/golden/images/%s.jpg' % n), 'rb').read() for n in names}
    blobs['hd_q75'] = synth.synth_jpeg(1920, 1080, 0, 75)
    blobs['gray_q75'] = synth.synth_jpeg(640, 360, 1, 75, gray=True)
    blobs['s444'] = synth.synth_jpeg(333, 251, 2, 80, subsampling='4:4:4')
    blobs['s422'] = synth.synth_jpeg(333, 251, 3, 80, subsampling='4:2:2')
    blobs['odd420'] = synth.synth_jpeg(301, 203, 4, 90, restart_rows=0)
    names = list(blobs)
    pts = [(0, '565le'), (1, '565be'), (2, '8888'), (3, 'gray8')]
    scs = [(0, 'full'), (2, 'half'), (4, 'quarter'), (8, 'eighth')]
    tot = bad = 0
    for mode, arith in (('sse', 0), ('scalar', 1)):
        ref = P.Ref(mode)
        ctx = J.Context(-1, arith)
        for pt, ptn in pts:
            for opt, sn in scs:
                use = [n for n in names if not (pt == 2 and n == 'gray_q75')]
                t0 = time.time()
                outs, status, tim, cnt = J.decode_batch_to_host(ctx, [blobs[n] for n in use], pt, opt)
                for n, o, st in zip(use, outs, status):
                    rc, err, img, _ = ref.decode_cb(blobs[n], pt, opt, want_log=False)
                    tot += 1
                    if o is None or st != 0 or o.shape != img.shape or (o != img).any():
                        bad += 1
                        if o is not None and o.shape == img.shape:
                            ys, xs = np.nonzero(o != img)
                            print('MISMATCH', mode, ptn, sn, n, 'status', st, 'diff bytes', len(ys), 'first', ys[0], xs[0],
                                  'got', o[ys[0], xs[0]:xs[0]+8], 'want', img[ys[0], xs[0]:xs[0]+8], flush=True)
                        else:
                            print('MISMATCH', mode, ptn, sn, n, 'status', st, 'shape', None if o is None else o.shape, img.shape, flush=True)
                print(mode, ptn, sn, 'done %.2fs' % (time.time() - t0), 'events', cnt['events'], 'launches', cnt['launches'], flush=True)
        ctx.close()
    print('TOTAL', tot, 'BAD', bad)
    return bad

if __name__ == '__main__':
    sys.exit(1 if main() else 0)
