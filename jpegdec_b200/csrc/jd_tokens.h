/*
 * jd_tokens.h -- two-phase form of the per-restart-interval entropy decode (PROTOTYPE).
 *
 * jd_decode_segment (jd_core.h) does everything in one sequential walk: bit reader, code lookup, DC prediction, record
 * packing, block headers.  On the GPU each warp then executes the union of the DC / AC / block-end paths of its 32 lanes
 * (~220 instructions per 32 symbols).  The split below keeps only what is inherently sequential in the walk:
 *
 *   phase 1  jd_parse_segment        bit reader + code lookup + zigzag index + table switch + window-phase nibbles;
 *                                    emits one 32-bit token per stored value and one token cursor per block
 *   phase 2  jd_materialize_segment  per block, independent of the bit stream: DC prediction (a prefix sum of the DC
 *                                    differences per component), record packing (normal / BIG), column mask and
 *                                    rows-4..7 flag, record index (a prefix sum of record counts), block header
 *
 * Both produce exactly the headers, records, window-phase map and truncation events of jd_decode_segment: that equality is
 * what tests/test_oracle.py::test_two_phase_entropy_equals_the_single_walk checks on the CPU for every fixture.  Not yet
 * wired into a kernel (phase 2 becomes one thread per block with two segmented scans); baseline mode only.
 *
 * Reference semantics: JPEGDecodeMCU src/jpeg.inl:2090-2274 (see jd_core.h for the window-phase bookkeeping).
 */
#ifndef JD_TOKENS_H
#define JD_TOKENS_H

#include "jd_core.h"

/* token: bit 31 = DC difference, bits 16..21 = zigzag index k of an AC coefficient, bits 0..15 = value (int16) */
#define JD_TOK_DC 0x80000000u
#define JD_TOK_K(t) (((t) >> 16) & 63u)
#define JD_TOK_VAL(t) ((int)(int16_t)((t) & 0xFFFFu))

typedef struct {
    uint32_t jmap;      /* as JDSegOut.jmap */
    uint32_t status;    /* JD_SEG_* */
    uint32_t err_blk;   /* first block without a complete token set (= number of blocks when status is JD_SEG_OK) */
    uint32_t ntok;
} JDParseOut;

/* phase 1: tok[] receives the tokens of the segment in stream order, blk_tok[b] = number of tokens emitted up to and
 * including block b.  tok_cap bounds tok[]; a segment that would overflow it ends with JD_SEG_OVERFLOW. */
template <typename EventSink>
JD_HD void jd_parse_segment(const JDSegIn &in, const uint16_t *lut, uint32_t *tok, uint32_t tok_cap, uint32_t *blk_tok,
                            EventSink &sink, JDParseOut &out)
{
    const uint32_t *words = (const uint32_t *)in.data;
    const uint32_t endw = (in.end + 3u) >> 2;
    uint32_t wi = in.start >> 2;
    uint32_t wnext = (wi < endw) ? words[wi] : 0u;
    uint32_t skip = in.start & 3u, ffp = 0, eos = 0;
    jd_u64 bb = 0;
    int nb = 0;

    uint32_t jw = JD_JW_INIT;
    int P = 0, Pb = 0;
    int err = -1;
    bool last_was_eob = true;

    const uint32_t nluma = (in.ncomp == 3) ? in.bpm - 2 : in.bpm;
    const uint32_t nblk_total = in.nmcu * in.bpm;
    uint32_t blk_in_mcu = 0, b = 0;
    uint32_t k = 0;
    uint32_t cursor = 0, blk_start = 0;   /* tokens emitted; cursor at the start of the current block */
    uint32_t comp = 0;
    const uint16_t *tb = lut + JD_LUT_DC(in.tsel & 1u);
    uint32_t thr = 0xF800u, sh = 4u, msk = 0x7Fu;

    if (nblk_total == 0) { out.status = JD_SEG_OK; out.err_blk = 0; out.jmap = jw; out.ntok = 0; return; }

    for (;;) {
        while (nb <= 32) {   /* same reader as jd_decode_segment */
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
                nb = 64;
            } else {
                for (int i = 0; i < 4; i++) {
                    const uint32_t c = (w >> (8 * i)) & 0xFFu;
                    if (skip) { skip--; continue; }
                    if (eos) break;
                    if (wi - 1u == (in.end >> 2) && (uint32_t)i >= (in.end & 3u)) { eos = 1; break; }
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
        jw = jd_jw_ckpt(jw);
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
            P += len;
            { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            if (s != 0 && len + s > 6) jw = jd_jw_ckpt(jw);
            P += s;
            { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            if (cursor >= tok_cap) { err = JD_SEG_OVERFLOW; break; }
            tok[cursor++] = JD_TOK_DC | ((uint32_t)v & 0xFFFFu);
            k = 1;
            tb = lut + JD_LUT_AC((in.tsel >> (2 * comp + 1)) & 1u); thr = 0xFC00u; sh = 0u; msk = 0x3FFu;
            continue;
        }
        last_was_eob = (rs == 0u);
        if (rs == 0u) {
            k = 64;
        } else {
            k += rs >> 4;
            if (s && k < 64u) {
                if (s > 11) { err = JD_SEG_BADSIZE; break; }
                if (len + s >= 18) {   /* possible truncated read for some start phases (as jd_decode_segment) */
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
                            ev.blk = in.blk0 + b;
                            ev.seg = in.seg;
                            ev.j1 = j1;
                            ev.field = (uint16_t)field;
                            ev.s = (uint8_t)s;
                            ev.p7 = (uint8_t)p7;
                            ev.ord = cursor - blk_start - 1u;   /* AC ordinal inside the block */
                            sink.push(ev);
                        }
                    }
                }
                if (cursor >= tok_cap) { err = JD_SEG_OVERFLOW; break; }
                tok[cursor++] = (k << 16) | ((uint32_t)v & 0xFFFFu);
            }
            k++;
        }
        P += len + s;
        { const int nPb = P >> 3; jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
        if (k >= 64u) {
            blk_tok[b] = cursor;
            blk_start = cursor;
            if (++b == nblk_total) break;
            if (++blk_in_mcu == in.bpm) blk_in_mcu = 0;
            comp = (blk_in_mcu < nluma) ? 0u : (blk_in_mcu - nluma + 1u);
            tb = lut + JD_LUT_DC((in.tsel >> (2 * comp)) & 1u); thr = 0xF800u; sh = 4u; msk = 0x7Fu;
            k = 0;
        }
    }
    out.err_blk = b;
    if (err >= 0) {
        for (uint32_t bb2 = b; bb2 < nblk_total; bb2++) blk_tok[bb2] = blk_start;   /* the unfinished block's tokens are dropped */
        cursor = blk_start;
    } else {
        if (!last_was_eob) jw = jd_jw_ckpt(jw);
        if (P & 7) jw += JD_JW_ONES;
    }
    out.status = (err < 0) ? (uint32_t)JD_SEG_OK : (uint32_t)err;
    out.jmap = jw;
    out.ntok = cursor;
}

/* phase 1 again, written the way the GPU wants it: ONE instruction stream per symbol.  DC / AC / EOB / block end are
 * selects and predicated stores, so the lanes of a warp never run different paths; only the rare events keep a branch
 * (a 0xFF byte in the refill word, an invalid code, a possibly truncated read).  Same outputs as jd_parse_segment. */
template <typename EventSink>
JD_HD void jd_parse_segment_uniform(const JDSegIn &in, const uint16_t *lut, uint32_t *tok, uint32_t tok_cap, uint32_t *blk_tok,
                                    EventSink &sink, JDParseOut &out)
{
    const uint32_t *words = (const uint32_t *)in.data;
    const uint32_t endw = (in.end + 3u) >> 2;
    uint32_t wi = in.start >> 2;
    uint32_t wnext = (wi < endw) ? words[wi] : 0u;
    uint32_t skip = in.start & 3u, ffp = 0, eos = 0;
    jd_u64 bb = 0;
    int nb = 0;
    uint32_t jw = JD_JW_INIT;
    int P = 0, Pb = 0;
    int err = -1;
    uint32_t last_eob = 1;

    const uint32_t nluma = (in.ncomp == 3) ? in.bpm - 2 : in.bpm;
    const uint32_t nblk_total = in.nmcu * in.bpm;
    /* per-MCU schedule, one byte per block: DC table offset / 64 in the low nibble... kept simple: nibble = comp | dc << 2 | ac << 3 */
    uint32_t sched = 0;
    for (uint32_t i = 0; i < in.bpm && i < 8u; i++) {
        const uint32_t c = (i < nluma) ? 0u : (i - nluma + 1u);
        sched |= (c | (((in.tsel >> (2 * c)) & 1u) << 2) | (((in.tsel >> (2 * c + 1)) & 1u) << 3)) << (4 * i);
    }
    const uint32_t bsh_end = 4u * in.bpm;
    uint32_t bsh = 0, cur = sched & 15u;
    uint32_t b = 0, k = 0, cursor = 0, blk_start = 0;
    uint32_t off_dc = JD_LUT_DC((cur >> 2) & 1u), off_ac = JD_LUT_AC(cur >> 3);   /* tables of the current block */
    uint32_t *tp = tok;                                                           /* next token slot */
    uint32_t *bp = blk_tok;                                                       /* next block's cursor slot */

    if (nblk_total == 0) { out.status = JD_SEG_OK; out.err_blk = 0; out.jmap = jw; out.ntok = 0; return; }
    if (tok_cap < 64u) { out.status = JD_SEG_OVERFLOW; out.err_blk = 0; out.jmap = jw; out.ntok = 0; for (uint32_t i = 0; i < nblk_total; i++) blk_tok[i] = 0; return; }

    for (;;) {
        /* ---- refill: the common case (no 0xFF in the next word) as selects ---- */
        {
            const uint32_t w = wnext;
            const bool need = nb <= 32;
            /* fast path: no 0xFF in the word, nothing pending, and the word after it still inside the file */
            const bool clean = (((((~w) - 0x01010101u) & w & 0x80808080u) | skip | ffp | eos) == 0u) && (wi + 2u < endw);
            if (need && !clean) {
                /* rare: byte path / end of data, as in jd_decode_segment */
                while (nb <= 32) {
                    const uint32_t w2 = wnext;
                    wi++;
                    wnext = (wi < endw) ? words[wi] : 0u;
                    if ((((((~w2) - 0x01010101u) & w2 & 0x80808080u)) | skip | ffp | eos) == 0u) {
#ifdef __CUDA_ARCH__
                        const uint32_t be2 = __byte_perm(w2, 0, 0x0123);
#else
                        const uint32_t be2 = __builtin_bswap32(w2);
#endif
                        bb |= (jd_u64)be2 << (32 - nb);
                        nb += 32;
                    } else if (eos) {
                        nb = 64;
                    } else {
                        for (int i = 0; i < 4; i++) {
                            const uint32_t c = (w2 >> (8 * i)) & 0xFFu;
                            if (skip) { skip--; continue; }
                            if (eos) break;
                            if (wi - 1u == (in.end >> 2) && (uint32_t)i >= (in.end & 3u)) { eos = 1; break; }
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
            } else {
#ifdef __CUDA_ARCH__
                const uint32_t be = __byte_perm(w, 0, 0x0123);
#else
                const uint32_t be = __builtin_bswap32(w);
#endif
                const int shl = need ? (32 - nb) : 0;
                bb |= need ? ((jd_u64)be << shl) : 0ull;
                nb += need ? 32 : 0;
                wi += need ? 1u : 0u;
                if (need) wnext = words[wi];
            }
        }
        jw = jd_jw_ckpt(jw);
        /* ---- code lookup: the table and its geometry follow from (k == 0) and the block's schedule nibble ---- */
        const bool is_dc = (k == 0u);
        const uint32_t toff = is_dc ? off_dc : off_ac;
        const uint32_t thr = is_dc ? 0xF800u : 0xFC00u, sh = is_dc ? 4u : 0u, msk = is_dc ? 0x7Fu : 0x3FFu;
        const uint32_t w16 = (uint32_t)(bb >> 48);
        const uint32_t idx = (w16 >= thr) ? (1024u + ((w16 >> sh) & msk)) : (w16 >> 6);
        const uint32_t e = lut[toff + idx];
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
        /* ---- what the symbol is ---- */
        const bool is_eob = !is_dc && rs == 0u;
        const uint32_t kpos = k + (rs >> 4);                              /* zigzag position of an AC coefficient */
        const bool is_coef = !is_dc && s != 0 && kpos < 64u;
        if (is_coef && s > 11) { err = JD_SEG_BADSIZE; break; }
        const bool store = is_dc || is_coef;
        /* ---- window phases: length, optional reload (DC only, jpeg.inl:2149), extra bits ---- */
        const int P1 = P + len;
        const uint32_t j1 = jw + (uint32_t)((P1 >> 3) - Pb) * JD_JW_ONES;
        if (is_coef && len + s >= 18 && ((j1 + 0x222222u) & 0x888888u) != 0u) {   /* rare: possibly truncated read */
            const int p7 = P1 & 7;
            bool any = false;
            for (int c = 0; c < 6; c++) {
                const int jc = (int)((j1 >> (4 * c)) & 15u);
                if (8 * jc + p7 + s > 64) any = true;
            }
            if (any) {
                JDEvent ev;
                ev.blk = in.blk0 + b; ev.seg = in.seg; ev.j1 = j1; ev.field = (uint16_t)field; ev.s = (uint8_t)s; ev.p7 = (uint8_t)p7;
                ev.ord = cursor - blk_start - 1u;
                sink.push(ev);
            }
        }
        const uint32_t j1c = (is_dc && s != 0 && len + s > 6) ? jd_jw_ckpt(j1) : j1;
        P = P1 + s;
        { const int nPb = P >> 3; jw = j1c + (uint32_t)(nPb - (P1 >> 3)) * JD_JW_ONES; Pb = nPb; }
        /* ---- token ---- */
        if (is_dc && cursor + 64u > tok_cap) { err = JD_SEG_OVERFLOW; break; }   /* a block emits at most 64 tokens */
        if (store) *tp = ((is_dc ? 0x8000u : kpos) << 16) | ((uint32_t)v & 0xFFFFu);
        tp += store ? 1 : 0;
        cursor += store ? 1u : 0u;
        last_eob = is_dc ? last_eob : (is_eob ? 1u : 0u);
        /* ---- zigzag index, block end ---- */
        k = is_dc ? 1u : (is_eob ? 64u : kpos + 1u);
        const bool done = k >= 64u;
        if (done) *bp = cursor;
        bp += done ? 1 : 0;
        blk_start = done ? cursor : blk_start;
        b += done ? 1u : 0u;
        if (b == nblk_total) break;
        {
            uint32_t nb2 = bsh + 4u;
            nb2 = (nb2 == bsh_end) ? 0u : nb2;
            bsh = done ? nb2 : bsh;
            cur = (sched >> bsh) & 15u;
            off_dc = JD_LUT_DC((cur >> 2) & 1u); off_ac = JD_LUT_AC(cur >> 3);
            k = done ? 0u : k;
        }
    }
    out.err_blk = b;
    if (err >= 0) {
        for (uint32_t bb2 = b; bb2 < nblk_total; bb2++) blk_tok[bb2] = blk_start;
        cursor = blk_start;
    } else {
        if (!last_eob) jw = jd_jw_ckpt(jw);
        if (P & 7) jw += JD_JW_ONES;
    }
    out.status = (err < 0) ? (uint32_t)JD_SEG_OK : (uint32_t)err;
    out.jmap = jw;
    out.ntok = cursor;
}

/* phase 2, written as the sequential loop a CPU runs; on the GPU the two running sums (DC predictor per component, record
 * cursor) become segmented scans and every block is independent.  Writes the same headers and records as jd_decode_segment. */
/* Returns JD_SEG_OK, or JD_SEG_OVERFLOW when the records do not fit in.rec_cap (blocks from there on get empty headers). */
JD_HD uint32_t jd_materialize_segment(const JDSegIn &in, const uint32_t *tposw, const uint32_t *tok, const uint32_t *blk_tok,
                                      uint32_t err_blk, jd_u64 *blk_hdr, uint16_t *rec, uint32_t *nrec)
{
    uint32_t status = JD_SEG_OK;
    const uint32_t nluma = (in.ncomp == 3) ? in.bpm - 2 : in.bpm;
    const uint32_t nblk_total = in.nmcu * in.bpm;
    int pred[3] = {0, 0, 0};
    uint32_t ri = 0;   /* u16 records written so far in this segment */
    for (uint32_t b = 0; b < nblk_total; b++) {
        if (b >= err_blk) { blk_hdr[b] = jd_pack_hdr(in.rec_index0, 0, 0, 0, 0, 0); continue; }
        const uint32_t t0 = b ? blk_tok[b - 1] : 0u, t1 = blk_tok[b];
        const uint32_t bim = b % in.bpm, comp = (bim < nluma) ? 0u : (bim - nluma + 1u);
        pred[comp] += JD_TOK_VAL(tok[t0]);                       /* the block's first token is its DC difference */
        const uint32_t ncoef = t1 - t0 - 1u;
        uint32_t big = 0, bflags = 0;
        for (uint32_t i = 0; i < ncoef; i++) { const int v = JD_TOK_VAL(tok[t0 + 1 + i]); if (v > 511 || v < -511) big = 1; }
        if (ri + (big ? 2 * ncoef : ncoef) > in.rec_cap) { status = JD_SEG_OVERFLOW; err_blk = b; blk_hdr[b] = jd_pack_hdr(in.rec_index0, 0, 0, 0, 0, 0); continue; }
        for (uint32_t i = 0; i < ncoef; i++) {
            const uint32_t t = tok[t0 + 1 + i], tw = tposw[JD_TOK_K(t)];
            const int v = JD_TOK_VAL(t);
            bflags |= tw;
            if (big) { rec[ri + 2 * i] = (uint16_t)(tw & 63u); rec[ri + 2 * i + 1] = (uint16_t)(int16_t)v; }
            else rec[ri + i] = (uint16_t)((tw << 10) | ((uint32_t)v & 0x3FFu));
        }
        blk_hdr[b] = jd_pack_hdr(in.rec_index0 + ri, pred[comp], ncoef, big, JD_BF_HI(bflags), JD_BF_COLMASK(bflags));
        ri += big ? 2 * ncoef : ncoef;
    }
    *nrec = ri;
    return status;
}

#endif /* JD_TOKENS_H */
