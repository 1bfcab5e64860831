/*
 * tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE: a sequential stepper for the per-thread
 * code of the CUDA kernels (jpegdec_b200/csrc/jd_core.h) plus the host parser/table code
 * (jd_host.c).  It lets the CPU-only test tier prove, against the compiled reference
 * (oracle/_ref), that the exact arithmetic / bit-reader / window-phase logic the kernels run
 * is bit-exact -- where no GPU exists.  It is NOT part of libjpegdec_b200.so and nothing in
 * the product calls it.
 *
 * Stages mirror the GPU pipeline (jd_device.cu): prescan -> per-segment entropy decode ->
 * phase stitch + patch -> per-block dequant/IDCT -> pixel assembly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <stdio.h>

#include "../../jpegdec_b200/csrc/jd_core.h"
#include "../../jpegdec_b200/csrc/jd_chunk.h"
#include "jd_flat_walk.h"
#include "../../jpegdec_b200/csrc/jd_internal.h"

static uint32_t g_ring[64];
static uint16_t g_stage[8];   /* record staging chunk of the walk */   /* stream ring of the CLEAN reader (one walker at a time here) */

struct VecSink {
    std::vector<JDEvent> ev;
    void push(const JDEvent &e) { ev.push_back(e); }
};

static const uint8_t kTpos[64] = JD_TPOS_INIT;
static uint32_t kTposW[64];
static int g_chunk_iters = 0, g_chunk_dc_mismatch = 0;
extern "C" int hostsim_last_chunk_dc_mismatch(void) { return g_chunk_dc_mismatch; }
extern "C" int hostsim_last_chunk_iters(void) { return g_chunk_iters; }

struct Planes {
    int sshift;          /* 0 full / half (block bytes are full 8x8), 2 quarter, 3 eighth */
    int bs;              /* block-byte edge: 8, 2 or 1 */
    int yw, yh, cw, ch;  /* plane sizes in block-bytes */
    std::vector<uint8_t> Y, Cb, Cr;
};

/* serial statement of what jdk_unstuff_segs writes for one restart segment: FF00 -> FF, the data ends at the first other
 * FFxx (or at `end`); zero padded; the un-stuffed length is stored in the vector's last word */
static void hostsim_unstuff(const uint8_t *data, uint32_t start, uint32_t end, std::vector<uint32_t> &out)
{
    out.assign((end - start) / 4 + 16, 0);
    uint8_t *o = (uint8_t *)out.data();
    uint32_t n = 0;
    for (uint32_t p = start; p < end; p++) {
        if (data[p] == 0xFF) { if (p + 1 < end && data[p + 1] == 0) { o[n++] = 0xFF; p++; } else break; }
        else o[n++] = data[p];
    }
    out[out.size() - 1] = n;
}

extern "C" int hostsim_decode(const uint8_t *data, int size, int pixel_type, int options, int arith,
                              uint8_t *out, int out_pitch, int *out_w, int *out_h, int *n_events,
                              int *err)
{
    JDInfo info;
    int start = 0;
    if (!jd_parse_header(data, size, 0, &info)) { *err = info.error; return -1; }
    if (options & JPEG_EXIF_THUMBNAIL) {
        if (info.thumb_data == 0 || info.thumb_w == 0) { *err = JPEG_INVALID_PARAMETER; return 0; }
        start = info.thumb_data;
        if (!jd_parse_header(data, size, start, &info)) { *err = info.error; return 0; }
    }
    if ((info.mode != 0xC0 && info.mode != 0xC2) || !info.tables_ok) { *err = JPEG_UNSUPPORTED_FEATURE; return 0; }
    *err = 0;
    std::vector<uint16_t> lut(JD_LUT_ENTRIES);
    jd_build_lut(&info, lut.data());
    int16_t quant[3 * 64];
    jd_build_quant(&info, quant);

    int sshift = 0;
    if (options & JPEG_SCALE_HALF) sshift = 1;
    else if (options & JPEG_SCALE_QUARTER) sshift = 2;
    else if (options & JPEG_SCALE_EIGHTH) sshift = 3;
    const bool prog = info.mode == 0xC2;   /* progressive: DC of the first scan -> 1/8 image (as batchCreate / JPEG_decode) */
    if (prog) {
        if (!(options & (JPEG_SCALE_HALF | JPEG_SCALE_QUARTER))) sshift = 3;
        if (sshift != 3 || info.p.ncomp_in_scan != info.ncomp || info.p.scan_start != 0 || info.p.scan_end != 0 || (info.approx >> 4) != 0) { *err = JPEG_UNSUPPORTED_FEATURE; return 0; }
    }
    if ((options & JPEG_LUMA_ONLY) && pixel_type < EIGHT_BIT_GRAYSCALE) pixel_type = EIGHT_BIT_GRAYSCALE;

    /* ---- prescan: restart segments ---- */
    const int total_mcus = info.mcus_x * info.mcus_y;
    const int mps = info.restart_interval ? info.restart_interval : total_mcus;
    const int nseg = (total_mcus + mps - 1) / mps;
    std::vector<uint32_t> seg_start(nseg, 0xFFFFFFFFu);
    seg_start[0] = (uint32_t)info.scan_offset;
    {
        int k = 1;
        for (int i = info.scan_offset; i + 1 < size && k < nseg; i++) {
            if (data[i] == 0xFF && data[i + 1] >= 0xD0 && data[i + 1] <= 0xD7) { seg_start[k++] = (uint32_t)(i + 2); i++; }
        }
    }
    /* compressed buffer padded + aligned like the device buffer */
    std::vector<uint32_t> padded((size + 64) / 4 + 16, 0);
    memcpy(padded.data(), data, (size_t)size);
    const uint8_t *cdata = (const uint8_t *)padded.data();

    const int nblk = total_mcus * info.bpm;
    std::vector<jd_u64> hdr(nblk, 0);
    /* records: JD_REC_INDEX layout, byte offsets relative to the file start, one slot per segment and per chunk */
    std::vector<uint16_t> rec((size_t)size * JD_REC_PER_BYTE + (size_t)JD_REC_SLOT_SLACK * ((size_t)nseg + (size_t)size / JD_CHUNK_BYTES + 4) + 1024, 0);
    std::vector<uint32_t> jmap(nseg, 0);
    const bool clean = (options & 0x40000) != 0;   /* test hook: un-stuff first, CLEAN bit reader (the GPU default) */
    std::vector<uint32_t> cbuf;
    VecSink sink;
    int bad = 0;
    for (int i = 0; i < 64; i++) kTposW[i] = jd_tposw(kTpos[i]);
    const bool chunked = (options & 0x20000) && nseg == 1;   /* test hook: restart-free parallel path (jd_chunk.h) */
    std::vector<uint32_t> phase_slot;                          /* chunked: true phase per chunk */
    int chunk_iters = 0;
    if (chunked) {
        /* jdk_unstuff: FF00 -> FF, stop at the first marker; zero tail */
        std::vector<uint32_t> fbuf((size + 64) / 4 + 64, 0);
        uint8_t *filt = (uint8_t *)fbuf.data();
        uint32_t flen = 0;
        for (int i = info.scan_offset; i < size; i++) {
            if (data[i] == 0xFF) { if (i + 1 < size && data[i + 1] == 0) { filt[info.scan_offset + flen++] = 0xFF; i++; } else break; }
            else filt[info.scan_offset + flen++] = data[i];
        }
        JDScanIn sc;
        sc.filt = filt; sc.f0 = (uint32_t)info.scan_offset; sc.flen = flen;
        sc.bpm = (uint32_t)info.bpm; sc.ncomp = (uint32_t)info.ncomp; sc.tsel = (uint32_t)info.tsel; sc.total_blocks = (uint32_t)nblk;
        const uint32_t nch = ((uint32_t)(size - info.scan_offset) + JD_CHUNK_BYTES - 1) / JD_CHUNK_BYTES + 1;
        std::vector<uint32_t> E(nch, JD_CS_PACK(0, 0, 0)), E2(nch), nst(nch, 0), pre(nch, 0);
        std::vector<int32_t> dcs(nch * 3, 0);
        std::vector<uint32_t> first(nch, 0);
        for (;;) {   /* fix point of the entry states */
            bool changed = false;
            E2[0] = E[0];
            for (uint32_t c = 0; c < nch; c++) {
                uint32_t badc;
                uint32_t ex = jd_chunk_parse(sc, lut.data(), c, E[c], &nst[c], &badc, &dcs[3 * c], &first[c]);
                first[c] |= badc << 31;
                if (c + 1 < nch) { E2[c + 1] = ex; if (ex != E[c + 1]) changed = true; }
            }
            E.swap(E2);
            chunk_iters++;
            if (getenv("HOSTSIM_TRACE")) { int nchg = 0, firstchg = -1; for (uint32_t c = 0; c < nch; c++) if (E[c] != E2[c]) { nchg++; if (firstchg < 0) firstchg = (int)c; } fprintf(stderr, "iter %d changed %d first %d\n", chunk_iters, nchg, firstchg); }
            if (!changed || chunk_iters > (int)nch + 2) break;
        }
        g_chunk_iters = chunk_iters;
        { uint32_t run = 0; for (uint32_t c = 0; c < nch; c++) { pre[c] = run; run += nst[c]; } }
        /* DC predictors at each chunk's first block: prefix sums of the parse pass's per-chunk DC sums (jdk_chunk_prefix) */
        std::vector<int32_t> pe(nch * 3, 0);
        { int run[3] = {0, 0, 0}; for (uint32_t c = 0; c < nch; c++) for (int q = 0; q < 3; q++) { pe[c * 3 + q] = run[q]; run[q] += dcs[c * 3 + q]; } }
        /* emit: the entropy walk itself (CLEAN reader), started at each chunk's first block (jdk_chunk_emit) */
        std::vector<uint32_t> cjmap(nch, JD_JW_INIT);
        g_chunk_dc_mismatch = 0;
        for (uint32_t c = 0; c < nch; c++) {
            uint32_t n = nst[c];
            n = (pre[c] >= (uint32_t)nblk) ? 0u : ((n < (uint32_t)nblk - pre[c]) ? n : (uint32_t)nblk - pre[c]);
            uint32_t status = JD_SEG_OK;
            if (n) {
                const uint32_t P0 = c * JD_CHUNK_BYTES * 8u + (first[c] & 0xFFFFu);
                const uint32_t byte0 = (uint32_t)info.scan_offset + (P0 >> 3);
                JDSegIn in;
                in.data = filt; in.start = byte0 & ~15u; in.end = (uint32_t)info.scan_offset + flen;
                in.nmcu = 0; in.bpm = (uint32_t)info.bpm; in.ncomp = (uint32_t)info.ncomp; in.tsel = (uint32_t)info.tsel;
                in.skip_bits = (byte0 - in.start) * 8u + (P0 & 7u);
                in.blk_first = (first[c] >> 16) & 0xFu; in.nblk = n; in.midstream = 1;
                in.pred[0] = pe[3 * c]; in.pred[1] = pe[3 * c + 1]; in.pred[2] = pe[3 * c + 2];
                in.rec_index0 = JD_REC_INDEX((uint32_t)info.scan_offset + c * JD_CHUNK_BYTES, 1u + c);
                in.rec_cap = JD_REC_CAP(JD_CHUNK_BYTES);
                in.seg = c; in.img = 0; in.blk0 = pre[c]; in.al = 0; in.ring = g_ring; in.stage = g_stage;
                JDSegOut so;
                jd_decode_segment<VecSink, JD_MODE_BASELINE, true>(in, lut.data(), kTposW, hdr.data() + pre[c], rec.data() + in.rec_index0, sink, so);
                cjmap[c] = so.jmap; status = so.status;
            }
            if (status == JD_SEG_OK && (first[c] >> 31)) status = JD_SEG_BADCODE;
            if (status != JD_SEG_OK) bad = 1;
        }
        /* stitch over chunks: true phase per chunk */
        phase_slot.assign(nch, 0);
        { uint32_t cur = 0;
          for (uint32_t c = 0; c < nch; c++) {
              phase_slot[c] = cur;
              uint32_t j = (cjmap[c] >> (4 * cur)) & 15u; cur = (j >= 6) ? 0 : j;
          } }
        jmap[0] = JD_JW_INIT;
    } else
    for (int sgi = 0; sgi < nseg; sgi++) {
        JDSegIn in;
        jd_segin_whole_interval(&in);
        in.data = cdata;
        in.start = seg_start[sgi];
        in.end = (uint32_t)size;
        int m0 = sgi * mps;
        in.nmcu = (uint32_t)((m0 + mps <= total_mcus) ? mps : total_mcus - m0);
        in.bpm = (uint32_t)info.bpm;
        in.ncomp = (uint32_t)info.ncomp;
        in.tsel = (uint32_t)info.tsel;
        if (in.start == 0xFFFFFFFFu) { bad = 1; break; }
        const uint32_t seg_end = (sgi + 1 < nseg && seg_start[sgi + 1] != 0xFFFFFFFFu) ? seg_start[sgi + 1] : (uint32_t)size;
        in.rec_index0 = JD_REC_INDEX(in.start, sgi);
        in.rec_cap = JD_REC_CAP(seg_end - in.start);
        in.seg = (uint32_t)sgi;
        in.img = 0;
        in.ring = g_ring; in.stage = g_stage;
        in.blk0 = (uint32_t)(m0 * info.bpm);
        JDSegOut so;
        in.al = prog ? (uint32_t)(info.approx & 15) : 0u;
        jd_u64 *hp = hdr.data() + (size_t)m0 * info.bpm;
        uint16_t *rp0 = rec.data() + in.rec_index0;
        if (clean) {
            hostsim_unstuff(cdata, in.start, (sgi + 1 < nseg && seg_start[sgi + 1] != 0xFFFFFFFFu) ? seg_start[sgi + 1] - 2u : (uint32_t)size, cbuf);
            in.data = (const uint8_t *)cbuf.data(); in.end = cbuf[cbuf.size() - 1]; in.start = 0;
            if (prog) jd_decode_segment<VecSink, JD_MODE_DC_SCAN, true>(in, lut.data(), kTposW, hp, rp0, sink, so);
            else if (sshift == 3) jd_decode_segment<VecSink, JD_MODE_PARSE_AC, true>(in, lut.data(), kTposW, hp, rp0, sink, so);
            else if (sshift == 2) jd_decode_segment<VecSink, JD_MODE_STORE_LOW, true>(in, lut.data(), kTposW, hp, rp0, sink, so);
            else jd_decode_segment<VecSink, JD_MODE_BASELINE, true>(in, lut.data(), kTposW, hp, rp0, sink, so);
        }
        else if (prog) jd_decode_segment<VecSink, JD_MODE_DC_SCAN>(in, lut.data(), kTposW, hp, rp0, sink, so);
        else if (sshift == 3) jd_decode_segment<VecSink, JD_MODE_PARSE_AC>(in, lut.data(), kTposW, hp, rp0, sink, so);
        else if (sshift == 2) jd_decode_segment<VecSink, JD_MODE_STORE_LOW>(in, lut.data(), kTposW, hp, rp0, sink, so);
        else jd_decode_segment(in, lut.data(), kTposW, hp, rp0, sink, so);
        jmap[sgi] = so.jmap;
        if (so.err_mcu >= 0) { bad = 1; break; }
    }
    if (bad) { *err = JPEG_DECODE_ERROR; }
    /* ---- stitch: true window phase at each segment start, then patch ---- */
    std::vector<uint32_t> phase(nseg, 0);
    {
        uint32_t c = 0;
        for (int sgi = 0; sgi < nseg; sgi++) {
            phase[sgi] = c;
            uint32_t j = (jmap[sgi] >> (4 * c)) & 15u;
            c = (j >= 6) ? 0 : j;
        }
    }
    int nev = 0;
    for (size_t i = 0; i < sink.ev.size(); i++) {
        const JDEvent &e = sink.ev[i];
        uint32_t jc = (e.j1 >> (4 * (chunked ? phase_slot[e.seg] : phase[e.seg]))) & 15u;
        if (8 * (int)jc + e.p7 + e.s > 64) {
            int v = jd_event_value(&e, jc);
            jd_patch_record(rec.data(), hdr[e.blk], e.ord, v);
            nev++;
        }
    }
    if (n_events) *n_events = nev;

    /* ---- per-block dequant + IDCT into block-byte planes ---- */
    Planes pl;
    pl.sshift = sshift;
    pl.bs = (sshift >= 2) ? (8 >> sshift) : 8;
    const int hs = info.mcu_w / 8, vs = info.mcu_h / 8;
    pl.yw = info.mcus_x * hs * pl.bs; pl.yh = info.mcus_y * vs * pl.bs;
    pl.cw = info.mcus_x * pl.bs; pl.ch = info.mcus_y * pl.bs;
    pl.Y.assign((size_t)pl.yw * pl.yh, 0);
    if (info.ncomp == 3) { pl.Cb.assign((size_t)pl.cw * pl.ch, 0); pl.Cr.assign((size_t)pl.cw * pl.ch, 0); }
    const int nluma = (info.ncomp == 3) ? info.bpm - 2 : info.bpm;
    for (int m = 0; m < total_mcus; m++) {
        int mx = m % info.mcus_x, my = m / info.mcus_x;
        for (int b = 0; b < info.bpm; b++) {
            int comp = (b < nluma) ? 0 : (b - nluma + 1);
            const int16_t *q = quant + comp * 64;
            jd_u64 h = hdr[(size_t)m * info.bpm + b];
            const uint32_t ri = JD_HDR_REC(h);
            const int dc = JD_HDR_DC(h);
            const int ncoef = (int)JD_HDR_NCOEF(h);
            const bool bigb = JD_HDR_BIG(h) != 0;
            int16_t tile[64]; /* natural order here */
            memset(tile, 0, sizeof(tile));
            tile[0] = (int16_t)dc;
            uint32_t flags = 0;
            for (int i = 0; i < ncoef; i++) {
                uint32_t t; int v;
                if (bigb) { t = rec[ri + 2 * i] & 63u; v = (int16_t)rec[ri + 2 * i + 1]; }
                else { const uint32_t r = rec[ri + i]; t = r >> 10; v = (int)(r << 22) >> 22; }
                const int n = (int)((t & 7u) * 8u + (t >> 3));
                /* 1/4 and 1/8 scale keep only zigzag 1..4 = natural 1, 8, 16, 9 (jpeg.inl:2117-2119) */
                if (sshift >= 2 && !(n == 1 || n == 8 || n == 16 || n == 9)) continue;
                tile[n] = (int16_t)v;
                flags |= 1u << (n & 7);
                flags |= (uint32_t)n << 8;
            }
            if (sshift < 2) {
                /* the header carries the same flags for the full-size path */
                uint32_t hf = JD_HDR_COLMASK(h) | (JD_HDR_HI(h) ? 0x2000u : 0u);
                if ((flags & 0x20FFu) != hf) { *err = 99; }
            }
            uint8_t px[64];
            if (sshift == 3 || flags == 0) {
                uint8_t c = (uint8_t)jd_range(dc * (int)q[0]);
                memset(px, c, 64);
            } else if (sshift == 2) {
                int t4 = tile[0] * q[0], t5 = tile[8] * q[8];
                int t0 = t4 + t5, t2 = t4 - t5;
                t4 = tile[1] * q[1]; t5 = tile[9] * q[9];
                int t1 = t4 + t5, t3 = t4 - t5;
                px[0] = (uint8_t)jd_range(t0 + t1); px[1] = (uint8_t)jd_range(t0 - t1);
                px[2] = (uint8_t)jd_range(t2 + t3); px[3] = (uint8_t)jd_range(t2 - t3);
            } else {
                int16_t col[64];
                const bool r47 = (flags & 0x2000u) == 0;
                for (int c = 0; c < 8; c++) {
                    int o[8];
                    if (arith == JPEG_ARITH_SSE2) {
                        int d[8];
                        for (int r = 0; r < 8; r++) d[r] = tile[r * 8 + c] * q[r * 8 + c];
                        jd_col_sse16(d, r47, o);
                    } else {
                        int mm[8], qq[8];
                        for (int r = 0; r < 8; r++) { mm[r] = tile[r * 8 + c]; qq[r] = q[r * 8 + c]; }
                        jd_col_scalar(mm, qq, r47, o);
                    }
                    for (int r = 0; r < 8; r++) col[r * 8 + c] = (int16_t)o[r];
                }
                for (int r = 0; r < 8; r++) {
                    int p[8];
                    uint32_t o[8];
                    for (int c = 0; c < 8; c++) p[c] = col[r * 8 + c];
                    jd_row(p, flags & 0xFFu, o);
                    for (int c = 0; c < 8; c++) px[r * 8 + c] = (uint8_t)o[c];
                }
            }
            /* store block bytes: full 8x8, or 2x2 (bytes 0..3 row-major) / 1 */
            uint8_t *plane; int pw, bx, by;
            if (comp == 0) {
                plane = pl.Y.data(); pw = pl.yw;
                int lx = (hs == 2) ? (b & 1) : 0;
                int ly = (hs == 2 && vs == 2) ? (b >> 1) : ((vs == 2 && hs == 1) ? b : 0);
                bx = mx * hs + lx; by = my * vs + ly;
            } else {
                plane = (comp == 1) ? pl.Cb.data() : pl.Cr.data(); pw = pl.cw;
                bx = mx; by = my;
            }
            for (int y = 0; y < pl.bs; y++)
                for (int x = 0; x < pl.bs; x++)
                    plane[(size_t)(by * pl.bs + y) * pw + bx * pl.bs + x] = px[y * pl.bs + x];
        }
    }

    /* ---- pixel assembly ---- */
    const int ow = (info.width + (1 << sshift) - 1) >> sshift, oh = (info.height + (1 << sshift) - 1) >> sshift;
    *out_w = ow; *out_h = oh;
    const bool sse_full = (arith == JPEG_ARITH_SSE2) && sshift == 0 && (info.subsample == 0x22 || info.subsample == 0x11);
    for (int oy = 0; oy < oh; oy++) {
        uint8_t *row = out + (size_t)oy * out_pitch;
        for (int ox = 0; ox < ow; ox++) {
            int Y, Y12, Cb = 128, Cr = 128;
            if (sshift == 1) {
                const uint8_t *yp = &pl.Y[(size_t)(2 * oy) * pl.yw + 2 * ox];
                int sum = yp[0] + yp[1] + yp[pl.yw] + yp[pl.yw + 1];
                Y = (sum + 2) >> 2;   /* gray paths */
                Y12 = sum << 10;      /* colour paths: no rounding */
                if (info.ncomp == 3) {
                    if (hs == 2 && vs == 2) { Cb = pl.Cb[(size_t)oy * pl.cw + ox]; Cr = pl.Cr[(size_t)oy * pl.cw + ox]; }
                    else if (hs == 1 && vs == 1) {
                        const uint8_t *a = &pl.Cb[(size_t)(2 * oy) * pl.cw + 2 * ox], *b2 = &pl.Cr[(size_t)(2 * oy) * pl.cw + 2 * ox];
                        Cb = (a[0] + a[1] + a[pl.cw] + a[pl.cw + 1] + 2) >> 2;
                        Cr = (b2[0] + b2[1] + b2[pl.cw] + b2[pl.cw + 1] + 2) >> 2;
                    } else if (hs == 2) { /* 4:2:2: chroma full height, half width: average 2 rows */
                        Cb = (pl.Cb[(size_t)(2 * oy) * pl.cw + ox] + pl.Cb[(size_t)(2 * oy + 1) * pl.cw + ox] + 1) >> 1;
                        Cr = (pl.Cr[(size_t)(2 * oy) * pl.cw + ox] + pl.Cr[(size_t)(2 * oy + 1) * pl.cw + ox] + 1) >> 1;
                    } else { /* 4:4:0 */
                        Cb = (pl.Cb[(size_t)oy * pl.cw + 2 * ox] + pl.Cb[(size_t)oy * pl.cw + 2 * ox + 1] + 1) >> 1;
                        Cr = (pl.Cr[(size_t)oy * pl.cw + 2 * ox] + pl.Cr[(size_t)oy * pl.cw + 2 * ox + 1] + 1) >> 1;
                    }
                }
            } else {
                Y = pl.Y[(size_t)oy * pl.yw + ox];
                Y12 = Y << 12;
                if (info.ncomp == 3) {
                    Cb = pl.Cb[(size_t)(oy / vs) * pl.cw + ox / hs];
                    Cr = pl.Cr[(size_t)(oy / vs) * pl.cw + ox / hs];
                }
            }
            if (pixel_type >= EIGHT_BIT_GRAYSCALE) {
                row[ox] = (uint8_t)Y;
            } else if (info.ncomp == 1) {
                uint32_t v = jd_gray565((uint32_t)Y);
                if (pixel_type != RGB565_LITTLE_ENDIAN) v = jd_bswap16(v);
                ((uint16_t *)row)[ox] = (uint16_t)v;
            } else if (sse_full) {
                int tr, tg, tb, R, G, B;
                jd_chroma_sse(Cb, Cr, &tr, &tg, &tb);
                jd_rgb_sse(Y, tr, tg, tb, &R, &G, &B);
                if (pixel_type == RGB8888) ((uint32_t *)row)[ox] = 0xFF000000u | ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B;
                else ((uint16_t *)row)[ox] = (uint16_t)(((R >> 3) << 11) | ((G >> 2) << 5) | (B >> 3));
            } else {
                if (pixel_type == RGB8888) ((uint32_t *)row)[ox] = jd_rgb8888_scalar(Y12, Cb, Cr);
                else {
                    uint32_t v = jd_rgb565_scalar(Y12, Cb, Cr);
                    if (pixel_type == RGB565_BIG_ENDIAN) v = jd_bswap16(v);
                    ((uint16_t *)row)[ox] = (uint16_t)v;
                }
            }
        }
    }
    return bad ? 0 : 1;
}

/* open()-level behaviour for conformance tests */
extern "C" int hostsim_open(const uint8_t *data, int size, int *w, int *h, int *subsample, int *err,
                            int *has_thumb, int *tw, int *th, int *orientation, int *bpp)
{
    JDInfo info;
    int rc = jd_parse_header(data, size, 0, &info);
    *w = info.width; *h = info.height; *subsample = info.subsample; *err = info.error;
    *has_thumb = info.has_thumb; *tw = info.thumb_w; *th = info.thumb_h; *orientation = info.orientation;
    *bpp = info.bpp;
    return rc;
}

/* one block through the kernels' per-thread IDCT code (column pass per lane, (short) store, row pass per lane) */
extern "C" void hostsim_idct(const int16_t *coef, const int16_t *quant, unsigned flags, int arith, uint8_t *out)
{
    int16_t col[64];
    const bool r47 = (flags & 0x2000u) == 0;
    for (int c = 0; c < 8; c++) {
        int o[8];
        if (arith == JPEG_ARITH_SSE2) {
            int d[8];
            for (int r = 0; r < 8; r++) d[r] = coef[r * 8 + c] * quant[r * 8 + c];
            jd_col_sse16(d, r47, o);
        } else {
            int mm[8], qq[8];
            for (int r = 0; r < 8; r++) { mm[r] = coef[r * 8 + c]; qq[r] = quant[r * 8 + c]; }
            jd_col_scalar(mm, qq, r47, o);
        }
        for (int r = 0; r < 8; r++) col[r * 8 + c] = (int16_t)o[r];
    }
    for (int r = 0; r < 8; r++) {
        int p[8];
        uint32_t o[8];
        for (int c = 0; c < 8; c++) p[c] = col[r * 8 + c];
        jd_row(p, flags & 0xFFu, o);
        for (int c = 0; c < 8; c++) out[r * 8 + c] = (uint8_t)o[c];
    }
}

/* one block through the packed thread-per-block code of jdk_idct_p (jd_core.h jd_idct_block_packed, SSE2-build arithmetic) */
extern "C" void hostsim_idct_packed(const int16_t *coef, const int16_t *quant, unsigned flags, uint8_t *out)
{
    const bool hi = (flags & 0x2000u) != 0;
    const uint32_t colmask = flags & 0xFFu;
    uint16_t d[64];
    for (int n = 0; n < 64; n++) d[n] = (uint16_t)(coef[n] * quant[n]);
    d[0] = (uint16_t)(d[0] + JD_ROW_BIAS);
    uint32_t px[16];
    if ((colmask & 0xF0u) == 0u) {
        uint32_t x[8][2];
        for (int r = 0; r < 8; r++) for (int q = 0; q < 2; q++) x[r][q] = (uint32_t)d[r * 8 + 2 * q] | ((uint32_t)d[r * 8 + 2 * q + 1] << 16);
        jd_idct_block_packed<2>(x, hi, colmask, (uint8_t *)px, 8);
    } else {
        uint32_t x[8][4];
        for (int r = 0; r < 8; r++) for (int q = 0; q < 4; q++) x[r][q] = (uint32_t)d[r * 8 + 2 * q] | ((uint32_t)d[r * 8 + 2 * q + 1] << 16);
        jd_idct_block_packed<4>(x, hi, colmask, (uint8_t *)px, 8);
    }
    memcpy(out, px, 64);
}

/* the same block with the 4-column lanes forced through the general (8-column) instantiation, as happens in a warp that
 * mixes both kinds */
extern "C" void hostsim_idct_packed_general(const int16_t *coef, const int16_t *quant, unsigned flags, uint8_t *out)
{
    const bool hi = (flags & 0x2000u) != 0;
    const uint32_t colmask = flags & 0xFFu;
    uint16_t d[64];
    for (int n = 0; n < 64; n++) d[n] = (uint16_t)(coef[n] * quant[n]);
    d[0] = (uint16_t)(d[0] + JD_ROW_BIAS);
    uint32_t px[16], x[8][4];
    for (int r = 0; r < 8; r++) for (int q = 0; q < 4; q++) x[r][q] = (uint32_t)d[r * 8 + 2 * q] | ((uint32_t)d[r * 8 + 2 * q + 1] << 16);
    jd_idct_block_packed<4>(x, hi, colmask, (uint8_t *)px, 8);
    memcpy(out, px, 64);
}

/* statistics of the block classes the IDCT kernel branches on (development aid) */
extern "C" int hostsim_block_stats(const uint8_t *data, int size, double *out /* 8 */)
{
    JDInfo info;
    if (!jd_parse_header(data, size, 0, &info)) return 0;
    std::vector<uint16_t> lut(JD_LUT_ENTRIES);
    jd_build_lut(&info, lut.data());
    for (int i = 0; i < 64; i++) kTposW[i] = jd_tposw(kTpos[i]);
    const int total_mcus = info.mcus_x * info.mcus_y;
    const int mps = info.restart_interval ? info.restart_interval : total_mcus;
    const int nseg = (total_mcus + mps - 1) / mps;
    std::vector<uint32_t> seg_start(nseg, 0xFFFFFFFFu);
    seg_start[0] = (uint32_t)info.scan_offset;
    { int k = 1; for (int i = info.scan_offset; i + 1 < size && k < nseg; i++) if (data[i] == 0xFF && data[i + 1] >= 0xD0 && data[i + 1] <= 0xD7) { seg_start[k++] = (uint32_t)(i + 2); i++; } }
    std::vector<uint32_t> padded((size + 64) / 4 + 16, 0);
    memcpy(padded.data(), data, (size_t)size);
    const int nblk = total_mcus * info.bpm;
    std::vector<jd_u64> hdr(nblk, 0);
    std::vector<uint16_t> rec((size_t)size * JD_REC_PER_BYTE + (size_t)JD_REC_SLOT_SLACK * (nseg + 1) + 1024, 0);
    VecSink sink;
    for (int sgi = 0; sgi < nseg; sgi++) {
        JDSegIn in; jd_segin_whole_interval(&in); in.data = (const uint8_t *)padded.data(); in.start = seg_start[sgi]; in.end = (uint32_t)size;
        int m0 = sgi * mps; in.nmcu = (uint32_t)((m0 + mps <= total_mcus) ? mps : total_mcus - m0);
        in.bpm = (uint32_t)info.bpm; in.ncomp = (uint32_t)info.ncomp; in.tsel = (uint32_t)info.tsel; in.img = 0; in.al = 0; in.ring = g_ring; in.stage = g_stage;
        in.rec_index0 = JD_REC_INDEX(in.start, sgi); in.rec_cap = JD_REC_CAP((uint32_t)size - in.start); in.seg = (uint32_t)sgi; in.blk0 = (uint32_t)(m0 * info.bpm);
        JDSegOut so;
        jd_decode_segment(in, lut.data(), kTposW, hdr.data() + (size_t)m0 * info.bpm, rec.data() + in.rec_index0, sink, so);
    }
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (int b = 0; b < nblk; b++) {
        jd_u64 h = hdr[b];
        uint32_t n = JD_HDR_NCOEF(h), cm = JD_HDR_COLMASK(h), hi = JD_HDR_HI(h);
        out[0] += (n == 0);
        out[1] += (n != 0 && (cm & 0xfc) == 0);
        out[2] += (n != 0 && (cm & 0xfc) != 0 && (cm & 0xf0) == 0);
        out[3] += ((cm & 0xf0) != 0);
        out[4] += (n != 0 && hi == 0);
        out[5] += (hi != 0);
        out[6] += n;
        out[7] += JD_HDR_BIG(h);
    }
    return nblk;
}

/* ---- the block-synchronous walk the kernels run (jd_decode_segment, raw and CLEAN reader) against the flat state-machine
 * form of the same walk (jd_decode_segment_flat): headers, records, window-phase map, events, status and failing MCU must be
 * identical.  Returns the number of mismatches. ---- */
extern "C" int hostsim_walk_check(const uint8_t *data, int size, int *n_segments, int *n_records, int *n_events, int *n_bad_segments)
{
    JDInfo info;
    if (!jd_parse_header(data, size, 0, &info) || info.mode != 0xC0 || !info.tables_ok) return -1;
    std::vector<uint16_t> lut(JD_LUT_ENTRIES);
    jd_build_lut(&info, lut.data());
    for (int i = 0; i < 64; i++) kTposW[i] = jd_tposw(kTpos[i]);
    const int total_mcus = info.mcus_x * info.mcus_y;
    const int mps = info.restart_interval ? info.restart_interval : total_mcus;
    const int nseg = (total_mcus + mps - 1) / mps;
    std::vector<uint32_t> seg_start(nseg, 0xFFFFFFFFu);
    seg_start[0] = (uint32_t)info.scan_offset;
    { int k = 1; for (int i = info.scan_offset; i + 1 < size && k < nseg; i++) if (data[i] == 0xFF && data[i + 1] >= 0xD0 && data[i + 1] <= 0xD7) { seg_start[k++] = (uint32_t)(i + 2); i++; } }
    std::vector<uint32_t> padded((size + 64) / 4 + 16, 0), cbuf;
    memcpy(padded.data(), data, (size_t)size);
    const int nblk = total_mcus * info.bpm;
    const size_t nrec_cap = (size_t)size * JD_REC_PER_BYTE + (size_t)JD_REC_SLOT_SLACK * (nseg + 1) + 4096;
    std::vector<jd_u64> hdrA(nblk, 0), hdrB(nblk, 0), hdrC(nblk, 0);
    std::vector<uint16_t> recA(nrec_cap, 0), recB(nrec_cap, 0), recC(nrec_cap, 0);
    int bad = 0, nrec = 0, nev = 0, nbadseg = 0;
    for (int sgi = 0; sgi < nseg; sgi++) {
        if (seg_start[sgi] == 0xFFFFFFFFu) break;
        JDSegIn in;
        jd_segin_whole_interval(&in);
        in.data = (const uint8_t *)padded.data(); in.start = seg_start[sgi]; in.end = (uint32_t)size;
        const int m0 = sgi * mps;
        in.nmcu = (uint32_t)((m0 + mps <= total_mcus) ? mps : total_mcus - m0);
        in.bpm = (uint32_t)info.bpm; in.ncomp = (uint32_t)info.ncomp; in.tsel = (uint32_t)info.tsel;
        const uint32_t seg_end = (sgi + 1 < nseg && seg_start[sgi + 1] != 0xFFFFFFFFu) ? seg_start[sgi + 1] : (uint32_t)size;
        in.rec_index0 = JD_REC_INDEX(in.start, sgi); in.seg = (uint32_t)sgi; in.img = 0; in.ring = g_ring; in.stage = g_stage;
        in.rec_cap = JD_REC_CAP(seg_end - in.start);
        in.blk0 = (uint32_t)(m0 * info.bpm); in.al = 0;
        const uint32_t nb = in.nmcu * in.bpm;
        VecSink sA, sB, sC;
        JDSegOut oA, oB, oC;
        jd_decode_segment_flat(in, lut.data(), kTposW, hdrA.data() + in.blk0, recA.data() + in.rec_index0, sA, oA);
        jd_decode_segment(in, lut.data(), kTposW, hdrB.data() + in.blk0, recB.data() + in.rec_index0, sB, oB);
        JDSegIn ic = in;
        hostsim_unstuff((const uint8_t *)padded.data(), in.start, (seg_end == (uint32_t)size) ? seg_end : seg_end - 2u, cbuf);
        ic.data = (const uint8_t *)cbuf.data(); ic.start = 0; ic.end = cbuf[cbuf.size() - 1];
        jd_decode_segment<VecSink, JD_MODE_BASELINE, true>(ic, lut.data(), kTposW, hdrC.data() + in.blk0, recC.data() + in.rec_index0, sC, oC);
        nrec += (int)oA.nrec; nev += (int)sA.ev.size();
        if (oA.status != JD_SEG_OK) nbadseg++;
        const JDSegOut *os[2] = {&oB, &oC};
        const VecSink *ss[2] = {&sB, &sC};
        const std::vector<jd_u64> *hs[2] = {&hdrB, &hdrC};
        const std::vector<uint16_t> *rs[2] = {&recB, &recC};
        for (int v = 0; v < 2; v++) {
            /* the flat form tests the record capacity per coefficient, the block form per block: an overflow may be reported
             * one block earlier, everything else must agree */
            if (oA.status == JD_SEG_OVERFLOW || os[v]->status == JD_SEG_OVERFLOW) { if (oA.status != os[v]->status) bad++; continue; }
            if (oA.status != os[v]->status || oA.err_mcu != os[v]->err_mcu) { bad++; continue; }
            if (oA.status == JD_SEG_OK && (oA.jmap != os[v]->jmap || oA.nrec != os[v]->nrec)) bad++;
            if (sA.ev.size() != ss[v]->ev.size()) bad++;
            else for (size_t e = 0; e < sA.ev.size(); e++) if (memcmp(&sA.ev[e], &ss[v]->ev[e], sizeof(JDEvent)) != 0) { bad++; break; }
            const uint32_t ngood = (oA.status == JD_SEG_OK) ? nb : (uint32_t)oA.err_mcu * in.bpm;   /* whole MCUs before the failing one */
            for (uint32_t b = 0; b < nb; b++) {
                const jd_u64 ha = hdrA[in.blk0 + b], hb = (*hs[v])[in.blk0 + b];
                if (b >= ngood && b < ngood + in.bpm) continue;   /* blocks of the failing MCU before the error: both forms finish the same ones, checked by err_mcu */
                if (ha != hb) { bad++; continue; }
                const uint32_t n = JD_HDR_NCOEF(ha) * (JD_HDR_BIG(ha) ? 2u : 1u);
                if (n && memcmp(recA.data() + JD_HDR_REC(ha), rs[v]->data() + JD_HDR_REC(hb), n * 2u) != 0) bad++;
            }
        }
    }
    if (n_segments) *n_segments = nseg;
    if (n_records) *n_records = nrec;
    if (n_events) *n_events = nev;
    if (n_bad_segments) *n_bad_segments = nbadseg;
    return bad;
}
