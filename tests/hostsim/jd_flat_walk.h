/*
 * jd_flat_walk.h -- TEST INFRASTRUCTURE: the entropy walk as one flat state machine (one symbol per loop iteration), the form
 * the first kernels ran.  tests/hostsim checks that jd_decode_segment (jd_core.h, what the kernels run now) produces the same
 * headers, records, window-phase maps, truncation events, status and failing MCU (tests/test_oracle.py).  Not part of the
 * library.
 */
#ifndef JD_FLAT_WALK_H
#define JD_FLAT_WALK_H
#include "../../jpegdec_b200/csrc/jd_core.h"

template <typename EventSink, int MODE = JD_MODE_BASELINE>
JD_HD void jd_decode_segment_flat(const JDSegIn &in, const uint16_t *lut /* JD_LUT_ENTRIES, shared/global */,
                             const uint32_t *tposw /* 64 words: jd_tposw(JD_TPOS[k]), shared/global */,
                             jd_u64 *blk_hdr /* nmcu*bpm headers */, uint16_t *rec /* this segment's records */,
                             EventSink &sink, JDSegOut &out)
{
    /* ---- bit reader: aligned 32-bit words, one word prefetched ahead of use ---- */
    const uint32_t *words = (const uint32_t *)in.data;
    const uint32_t endw = (in.end + 3u) >> 2;    /* first word index past this file */
    uint32_t wi = in.start >> 2;                 /* index of the next word to consume */
    uint32_t wnext = (wi < endw) ? words[wi] : 0u;
    uint32_t skip = in.start & 3u;               /* bytes of the first word that precede the segment */
    uint32_t ffp = 0;                            /* previous byte was 0xFF (stuffing / marker undecided) */
    uint32_t eos = 0;                            /* marker or end of data reached: zeros from here on */
    jd_u64 bb = 0;                               /* bit buffer, MSB first */
    int nb = 0;                                  /* valid bits in bb */

    int pred0 = 0, pred1 = 0, pred2 = 0;
    uint32_t jw = JD_JW_INIT;
    int P = 0, Pb = 0;                           /* bits consumed in this segment, and P >> 3 */
    uint16_t *rp = rec;                          /* next record slot */
    uint16_t *const rend = rec + in.rec_cap;
    int err = -1;
    bool last_was_eob = true;

    const uint32_t nluma = (in.ncomp == 3) ? in.bpm - 2 : in.bpm;
    const uint32_t nblk_total = in.nblk ? in.nblk : in.nmcu * in.bpm;
    /* per-MCU block schedule, one nibble per block: component (2 bits) | DC table << 2 | AC table << 3 */
    uint32_t sched = 0;
    for (uint32_t i = 0; i < in.bpm && i < 8u; i++) {
        const uint32_t c = (i < nluma) ? 0u : (i - nluma + 1u);
        sched |= (c | (((in.tsel >> (2 * c)) & 1u) << 2) | (((in.tsel >> (2 * c + 1)) & 1u) << 3)) << (4 * i);
    }
    const uint32_t bsh_end = 4u * in.bpm;
    uint32_t bsh = 4u * in.blk_first;            /* 4 * (block index inside the MCU) */
    uint32_t cur = sched & 15u;                  /* schedule nibble of the current block */
    uint32_t nleft = nblk_total;                 /* blocks still to finish */
    jd_u64 *hp = blk_hdr;

    /* per-block state */
    uint32_t k = 0;                              /* zigzag index; 0 = DC pending */
    uint32_t cnt = 0, bflags = 0;                /* cnt: stored coefficients << 16 | BIG << 22 (header layout); bflags: OR of tposw words */
    uint32_t ridx0 = in.rec_index0;              /* global index of this block's first record */
    const uint32_t rec_lo = (uint32_t)(uintptr_t)rec;
    int dcval = 0;
    /* current table geometry (DC at block start) */
    const uint16_t *tb = lut + JD_LUT_DC((cur >> 2) & 1u);
    uint32_t thr = 0xF800u, sh = 4u, msk = 0x7Fu;

    if (nblk_total == 0) { out.status = JD_SEG_OK; out.err_mcu = -1; out.jmap = jw; out.nrec = 0; return; }

    for (;;) {
        /* ---- refill: keep >= 32 valid bits ---- */
        while (nb <= 32) {
            const uint32_t w = wnext;
            wi++;
            wnext = (wi < endw) ? words[wi] : 0u;
            if ((((((~w) - 0x01010101u) & w & 0x80808080u)) | skip | ffp | eos) == 0u) {
#ifdef __CUDA_ARCH__
                const uint32_t be = __byte_perm(w, 0, 0x0123);
#else
                const uint32_t be = __builtin_bswap32(w);
#endif
                bb |= (jd_u64)be << (32 - nb);
                nb += 32;
            } else if (eos) {
                nb = 64;                          /* bb's low bits are zero: the stream continues as zeros */
            } else {
                /* byte path: FF00 -> FF; FFxx (xx != 0) = marker: this segment's data ends (JPEGFilter :1519-1538) */
                for (int i = 0; i < 4; i++) {
                    const uint32_t c = (w >> (8 * i)) & 0xFFu;
                    if (skip) { skip--; continue; }
                    if (eos) break;
                    if (wi - 1u == (in.end >> 2) && (uint32_t)i >= (in.end & 3u)) { eos = 1; break; } /* past the file */
                    if (ffp) {
                        ffp = 0;
                        if (c != 0u) { eos = 1; break; }
                        bb |= (jd_u64)0xFFu << (56 - nb);
                        nb += 8;
                        continue;
                    }
                    if (c == 0xFFu) { ffp = 1; continue; }
                    bb |= (jd_u64)c << (56 - nb);
                    nb += 8;
                }
                if (wi >= endw && !eos && nb <= 32) eos = 1;
            }
        }
        /* ---- window checkpoint (R1 at block entry / R3 at AC loop top; also the previous R4) ---- */
        jw = jd_jw_ckpt(jw);
        /* ---- code lookup ---- */
        const uint32_t w16 = (uint32_t)(bb >> 48);
        const uint32_t idx = (w16 >= thr) ? (1024u + ((w16 >> sh) & msk)) : (w16 >> 6);
        const uint32_t e = tb[idx];
        if (e == 0u) { err = JD_SEG_BADCODE; break; }
        const int len = (int)(e >> 8);
        const uint32_t rs = e & 0xFFu;
        const int s = (int)(rs & 15u);
        bb <<= len;
        const uint32_t hi32 = (uint32_t)(bb >> 32);
        const uint32_t field = s ? (hi32 >> (32 - s)) : 0u;
        const uint32_t half = s ? (1u << (s - 1)) : 1u;
        const int v = (field < half) ? (int)field - ((1 << s) - 1) : (int)field;
        bb <<= s;
        nb -= len + s;
        if (k == 0u) {
            /* DC: jpeg.inl:2128-2165.  Window reload R2 (:2149) only when the LUT has no
             * precomputed difference, i.e. not (SSSS != 0 && len + SSSS <= 6) (:1132). */
            P += len;
            { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            if (s != 0 && len + s > 6) jw = jd_jw_ckpt(jw);
            P += s;
            { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            const uint32_t comp = cur & 3u;
            const int pv = ((comp == 0u) ? pred0 : ((comp == 1u) ? pred1 : pred2)) + ((MODE == JD_MODE_DC_SCAN) ? (int)((uint32_t)v << in.al) : v);
            pred0 = (comp == 0u) ? pv : pred0;
            pred1 = (comp == 1u) ? pv : pred1;
            pred2 = (comp >= 2u) ? pv : pred2;
            dcval = pv;
            if (MODE != JD_MODE_DC_SCAN) {
                k = 1;
                tb = lut + JD_LUT_AC(cur >> 3);
                thr = 0xFC00u; sh = 0u; msk = 0x3FFu;
                continue;
            }
            k = 64;
        } else {
        last_was_eob = (rs == 0u);
        if (rs == 0u) {
            /* EOB (:2241-2244): leaves without the trailing window check */
            k = 64;
        } else {
            k += rs >> 4;
            if (MODE != JD_MODE_PARSE_AC && s && k < ((MODE == JD_MODE_STORE_LOW) ? 5u : 64u)) {
                /* stored coefficient (jpeg.inl:2247-2256) */
                if (s > 11) { err = JD_SEG_BADSIZE; break; }
                if (len + s >= 18) {
                    /* possible truncated read for some start phases */
                    const int P1 = P + len;
                    const uint32_t j1 = jw + (uint32_t)((P1 >> 3) - Pb) * JD_JW_ONES;
                    const int p7 = P1 & 7;
                    if (((j1 + 0x222222u) & 0x888888u) != 0u) {
                        bool any = false;
                        for (int c = 0; c < 6; c++) {
                            const int jc = (int)((j1 >> (4 * c)) & 15u);
                            if (8 * jc + p7 + s > 64) any = true;
                        }
                        if (any) {
                            JDEvent ev;
                            ev.blk = in.blk0 + (nblk_total - nleft);
                            ev.seg = in.seg;
                            ev.j1 = j1;
                            ev.field = (uint16_t)field;
                            ev.s = (uint8_t)s;
                            ev.p7 = (uint8_t)p7;
                            ev.ord = (cnt >> 16) & 63u;
                            ev.img = in.img;
                            sink.push(ev);
                        }
                    }
                }
                const uint32_t tw = tposw[k];
                bflags |= tw;
                if (s >= 10 && !(cnt & (1u << 22))) {
                    /* first >= 10-bit magnitude of this block: switch its records to (t, value) pairs */
                    const uint32_t ncoef = (cnt >> 16) & 63u;
                    uint16_t *rec0 = rec + (ridx0 - in.rec_index0);
                    if (rp + ncoef + 2 > rend) { err = JD_SEG_OVERFLOW; break; }
                    for (uint32_t i = ncoef; i-- > 0u;) {
                        const uint32_t r = rec0[i];
                        rec0[2u * i] = (uint16_t)(r >> 10);
                        rec0[2u * i + 1u] = (uint16_t)(int16_t)((int)(r << 22) >> 22);
                    }
                    rp += ncoef;
                    cnt |= 1u << 22;
                }
                if (cnt & (1u << 22)) {
                    if (rp + 2 > rend) { err = JD_SEG_OVERFLOW; break; }
                    rp[0] = (uint16_t)(tw & 63u);
                    rp[1] = (uint16_t)(int16_t)v;
                    rp += 2;
                } else {
                    if (rp >= rend) { err = JD_SEG_OVERFLOW; break; }
                    *rp++ = (uint16_t)((tw << 10) | ((uint32_t)v & 0x3FFu));
                }
                cnt += 1u << 16;
            }
            k++;
        }
        P += len + s;
        { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
        }
        if (k >= 64u) {
            /* ---- block finished: header = first record | dc << 32 | count << 48 | BIG << 54 | rows-4..7 << 55 | columns << 56 ---- */
            *hp++ = (jd_u64)ridx0 | ((jd_u64)((bflags & JD_BF_MASK) | cnt | ((uint32_t)dcval & 0xFFFFu)) << 32);
            if (--nleft == 0u) break;
            /* next block of the MCU: luma blocks first, then Cb, Cr (jpeg.inl:5138-5275) */
            bsh += 4u;
            if (bsh == bsh_end) bsh = 0u;
            cur = (sched >> bsh) & 15u;
            tb = lut + JD_LUT_DC((cur >> 2) & 1u);
            thr = 0xF800u; sh = 4u; msk = 0x7Fu;
            k = 0; cnt = 0; bflags = 0;
            ridx0 = in.rec_index0 + (((uint32_t)(uintptr_t)rp - rec_lo) >> 1);
        }
    }
    const uint32_t b = nblk_total - nleft;       /* blocks finished */
    if (err >= 0) {
        /* undecodable from here: later stages must still find well-formed (empty) headers */
        out.err_mcu = (int32_t)((b + in.blk_first) / in.bpm);
        for (uint32_t bb2 = b; bb2 < nblk_total; bb2++) blk_hdr[bb2] = jd_pack_hdr(in.rec_index0, 0, 0, 0, 0, 0);
    }
    out.status = (err < 0) ? (uint32_t)JD_SEG_OK : (uint32_t)err;
    if (err < 0) {
        out.err_mcu = -1;
        /* end of restart interval (jpeg.inl:5337-5347): R4 already happened unless the last
         * block ended with EOB; then the bit offset is rounded up to a byte without a reload. */
        if (!last_was_eob) jw = jd_jw_ckpt(jw);
        if (P & 7) jw += JD_JW_ONES;
    }
    out.jmap = jw;
    out.nrec = (uint32_t)(rp - rec);
}
#endif
