/*
 * jd_chunk.h -- per-thread code of the parallel entropy decode for scans WITHOUT restart markers
 * (SURVEY.md section 8(f) item 2).  One restart-free scan is one long dependent bit stream; the reference
 * (JPEGDecodeMCU, src/jpeg.inl:2090-2274, driven by DecodeJPEG :5128-5275) walks it on one core.  Here the
 * un-stuffed stream (jdk_unstuff = JPEGFilter :1431-1540) is cut into fixed-size chunks:
 *
 *   1. jd_chunk_parse  every chunk is parsed from a guessed entry state (bit offset, zigzag index, block-in-MCU
 *                      index); Huffman streams self-synchronise, so after a few passes in which each chunk hands
 *                      its exit state to its right neighbour the entry states are exact (fix point).
 *   2. prefix sums     blocks started before each chunk.
 *   3. jd_chunk_emit   every chunk decodes the blocks that START inside it (running past its end to finish the
 *                      last one) and writes the same headers/records jd_decode_segment writes, tracking the
 *                      reference's six bit-window phase candidates exactly as jd_decode_segment does.  The DC
 *                      predictors at its first block come from the parse pass (sum of the DC differences per chunk
 *                      and component, prefix-summed over the chunks with the block counts).
 *
 * `__host__ __device__` like jd_core.h: tests/hostsim steps it on the CPU against the compiled reference.
 */
#ifndef JD_CHUNK_H
#define JD_CHUNK_H
#include "jd_core.h"

#define JD_CHUNK_BYTES 512u

/* entry/exit state of a chunk: bit offset past the chunk start (symbols straddle by < 32 bits), zigzag index k
 * (0 = next symbol is a DC), block-in-MCU index.  JD_CS_NONE = nothing decodable here (past the end). */
#define JD_CS_PACK(bit, k, bim) (((uint32_t)(bit) & 0xFFu) | (((uint32_t)(k) & 0x7Fu) << 8) | (((uint32_t)(bim) & 0xFu) << 16))
#define JD_CS_BIT(s) ((s) & 0xFFu)
#define JD_CS_K(s) (((s) >> 8) & 0x7Fu)
#define JD_CS_BIM(s) (((s) >> 16) & 0xFu)
#define JD_CS_NONE 0xFFFFFFFFu

typedef struct {
    const uint8_t *filt;   /* un-stuffed stream buffer (same offsets as the raw buffer) */
    uint32_t f0;           /* byte offset of the scan's first un-stuffed byte */
    uint32_t flen;         /* un-stuffed length in bytes (zeros follow up to the raw length + 8) */
    uint32_t bpm, ncomp, tsel;
    uint32_t total_blocks; /* blocks in the scan */
} JDScanIn;

/* Bit window over the un-stuffed scan: 64-bit register buffer, MSB first, one aligned 32-bit word fetched per 32 bits
 * consumed (the first version fetched two words per SYMBOL, which made the L1 the bound of every pass). */
struct JDBitWin {
    const uint32_t *words;
    uint32_t wi;
    jd_u64 bb;
    int nb;
    JD_HDM void init(const JDScanIn &sc, uint32_t rel)
    {
        const uint32_t ap = sc.f0 * 8u + rel;
        words = (const uint32_t *)sc.filt;
        wi = ap >> 5;
        const uint32_t sft = ap & 31u;
        bb = (jd_u64)jd_bswap32(words[wi++]) << (32u + sft);
        nb = 32 - (int)sft;
    }
    JD_HDM void refill() { if (nb <= 32) { bb |= (jd_u64)jd_bswap32(words[wi++]) << (32 - nb); nb += 32; } }
    JD_HDM uint32_t hi() const { return (uint32_t)(bb >> 32); }       /* the next 32 bits (>= 33 valid after refill) */
    JD_HDM void drop(uint32_t n) { bb <<= n; nb -= (int)n; }
};

/* per-MCU block schedule, one nibble per block: component (2 bits) | DC table << 2 | AC table << 3 */
JD_HD uint32_t jd_block_schedule(uint32_t tsel, uint32_t bpm, uint32_t ncomp)
{
    const uint32_t nluma = (ncomp == 3) ? bpm - 2 : bpm;
    uint32_t sched = 0;
    for (uint32_t i = 0; i < bpm && i < 8u; i++) {
        const uint32_t c = (i < nluma) ? 0u : (i - nluma + 1u);
        sched |= (c | (((tsel >> (2 * c)) & 1u) << 2) | (((tsel >> (2 * c + 1)) & 1u) << 3)) << (4 * i);
    }
    return sched;
}

/* one AC symbol's fast-table entry (JD_LUT_ACF layout, jd_core.h), 0 = invalid code */
template <typename T16, typename T32>
JD_HD uint32_t jd_ac_entry(const T16 &T, const T32 &TF, uint32_t actab, uint32_t hi)
{
    uint32_t e = TF.at((JD_LUT_ACF(actab) >> 1) + (hi >> 22));
    if (e == 0u) {
        const uint32_t e16 = (hi >= 0xFC000000u) ? T.at(JD_LUT_AC(actab) + 1024u + ((hi >> 16) & 0x3FFu)) : 0u;
        if (e16 != 0u) e = JD_ACF_PACK(e16 >> 8, e16 & 0xFFu);
    }
    return e;
}

/* Pass 1: parse chunk `ci` from `entry`; returns the state at which the first symbol of chunk ci+1 starts
 * (JD_CS_NONE if the stream ends before) and counts the DC symbols (= block starts) inside this chunk.
 * `lut` = the image's table set (on the device: in the CTA's shared memory). */
JD_HD uint32_t jd_chunk_parse(const JDScanIn &sc, const uint16_t *lut, uint32_t ci, uint32_t entry, uint32_t *nstart, uint32_t *bad,
                               int32_t *dcs /* [3]: per component, sum of the DC differences of the blocks that start here */)
{
    *nstart = 0; *bad = 0;
    dcs[0] = dcs[1] = dcs[2] = 0;
    if (entry == JD_CS_NONE) return JD_CS_NONE;
    const uint32_t c0 = ci * JD_CHUNK_BYTES * 8u, c1 = c0 + JD_CHUNK_BYTES * 8u, endbits = sc.flen * 8u;
    uint32_t rel = c0 + JD_CS_BIT(entry), k = JD_CS_K(entry);
    if (c0 >= endbits) return JD_CS_NONE;
    const JDTab16 T(lut);
    const JDTab32 TF((const uint32_t *)lut);
    const uint32_t sched = jd_block_schedule(sc.tsel, sc.bpm, sc.ncomp), bsh_end = 4u * sc.bpm;
    uint32_t bsh = 4u * JD_CS_BIM(entry);
    JDBitWin w;
    w.init(sc, rel);
    uint32_t n = 0;
    int d0 = 0, d1 = 0, d2 = 0;
    while (rel < c1) {
        if (rel >= endbits) { *nstart = n; dcs[0] = d0; dcs[1] = d1; dcs[2] = d2; return JD_CS_NONE; }
        w.refill();
        const uint32_t hi = w.hi(), cur = (sched >> bsh) & 15u;
        uint32_t adv;
        if (k == 0) {
            const uint32_t w16 = hi >> 16;
            const uint32_t e = T.at(JD_LUT_DC((cur >> 2) & 1u) + ((w16 >= 0xF800u) ? (1024u + ((w16 >> 4) & 0x7Fu)) : (w16 >> 6)));
            if (e == 0u) { *bad = 1; *nstart = n; return JD_CS_PACK(0, 0, 0); }
            const uint32_t len = e >> 8, s = e & 15u;
            const int v = jd_extend_top(hi << len, s);
            const uint32_t comp = cur & 3u;
            d0 += (comp == 0u) ? v : 0; d1 += (comp == 1u) ? v : 0; d2 += (comp >= 2u) ? v : 0;
            adv = len + s;
            n++; k = 1;
        } else {
            const uint32_t e = jd_ac_entry(T, TF, cur >> 3, hi);
            /* an invalid code under a guessed entry state only says the guess was wrong: let the right neighbour keep
             * speculating from its own first bit (a truly corrupt stream is reported by jd_chunk_emit) */
            if (e == 0u) { *bad = 1; *nstart = n; return JD_CS_PACK(0, 0, 0); }
            adv = e & 0x1Fu;
            k += e >> 24;                        /* run + 1; 128 for EOB */
        }
        w.drop(adv);
        rel += adv;
        if (k >= 64u) {
            k = 0;
            bsh += 4u;
            if (bsh == bsh_end) bsh = 0;
        }
    }
    *nstart = n;
    dcs[0] = d0; dcs[1] = d1; dcs[2] = d2;
    return JD_CS_PACK(rel - c1, k, bsh >> 2);
}

typedef struct {
    uint32_t jmap;       /* window-phase candidates at the point where the next chunk's first symbol starts */
    int32_t dcsum[3];    /* per component: sum of the DC differences of the blocks owned by this chunk */
    uint32_t status;     /* JD_SEG_* */
    uint32_t nown;       /* blocks owned (started here and inside the scan) */
} JDChunkOut;

/* Pass 3: decode and emit the blocks that start in chunk `ci`.
 *   blk_first : index (within the scan) of the first block that starts in this chunk
 *   next_entry: entry state of chunk ci+1 (where to snapshot the phase map), JD_CS_NONE for the last chunk
 *   blk_hdr   : headers of the scan (indexed by block index within the scan)
 *   rec/rec_index0/rec_cap: this chunk's record area (rec_index0 = image-relative index of rec[0])
 *   slot      : phase slot id written into events (the stitch resolves the true phase per slot)
 *   blk0      : global index of the scan's first block (events) */
template <typename EventSink>
JD_HD void jd_chunk_emit(const JDScanIn &sc, const uint16_t *lut, const uint32_t *tposw, uint32_t ci, uint32_t entry,
                         uint32_t next_entry, uint32_t blk_first, jd_u64 *blk_hdr, uint16_t *rec, uint32_t rec_index0,
                         uint32_t rec_cap, uint32_t slot, uint32_t blk0, uint32_t img, const int32_t *pred_in /* [3] DC predictors at the chunk's first block */,
                         EventSink &sink, JDChunkOut &out)
{
    out.jmap = JD_JW_INIT; out.dcsum[0] = out.dcsum[1] = out.dcsum[2] = 0; out.status = JD_SEG_OK; out.nown = 0;
    if (entry == JD_CS_NONE) return;
    const uint32_t c0 = ci * JD_CHUNK_BYTES * 8u, c1 = c0 + JD_CHUNK_BYTES * 8u, endbits = sc.flen * 8u;
    if (c0 >= endbits) return;
    const uint32_t snap_at = (next_entry == JD_CS_NONE) ? 0xFFFFFFFFu : c1 + JD_CS_BIT(next_entry);
    uint32_t rel = c0 + JD_CS_BIT(entry), k = JD_CS_K(entry);
    const JDTab16 T(lut);
    const JDTab32 TF((const uint32_t *)lut);
    const uint32_t sched = jd_block_schedule(sc.tsel, sc.bpm, sc.ncomp), bsh_end = 4u * sc.bpm;
    uint32_t bsh = 4u * JD_CS_BIM(entry);
    JDBitWin w;
    w.init(sc, rel);
    uint32_t jw = JD_JW_INIT;
    int Pb = (int)(rel >> 3);            /* P = rel: bit position in the un-stuffed scan (no restart: segment == scan) */
    bool snapped = false, last_was_eob = true;
    bool own = false;                    /* the block being parsed is owned by this chunk */
    uint32_t bi = blk_first;             /* index of the next block to start */
    uint16_t *rp = rec, *const rend = rec + rec_cap, *rec0 = rec;
    uint32_t ncoef = 0, big = 0, bflags = 0;
    int pred[3] = {pred_in[0], pred_in[1], pred_in[2]}, dcval = 0;
    for (;;) {
        if (!snapped && rel >= snap_at) { out.jmap = jw; snapped = true; }   /* before the checkpoint, like a segment end */
        if (k == 0) {
            /* a block starts here: ours only if it starts inside the chunk and inside the scan */
            if (rel >= c1 || bi >= sc.total_blocks || rel >= endbits) break;
            own = true;
            ncoef = 0; big = 0; bflags = 0; rec0 = rp;
        } else if (rel >= endbits) break;
        jw = jd_jw_ckpt(jw);
        w.refill();
        const uint32_t hi = w.hi(), cur = (sched >> bsh) & 15u;
        if (k == 0) {
            const uint32_t w16 = hi >> 16;
            const uint32_t e = T.at(JD_LUT_DC((cur >> 2) & 1u) + ((w16 >= 0xF800u) ? (1024u + ((w16 >> 4) & 0x7Fu)) : (w16 >> 6)));
            if (e == 0u) { out.status = JD_SEG_BADCODE; break; }
            const uint32_t len = e >> 8, s = e & 15u;
            const int v = jd_extend_top(hi << len, s);
            w.drop(len + s);
            rel += len;
            { const int nPb = (int)(rel >> 3); jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            if (s != 0u && len + s > 6u) jw = jd_jw_ckpt(jw);
            rel += s;
            { const int nPb = (int)(rel >> 3); jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
            const uint32_t comp = cur & 3u;
            pred[comp] += v;
            dcval = pred[comp];
            k = 1;
            last_was_eob = false;
            continue;
        }
        const uint32_t e = jd_ac_entry(T, TF, cur >> 3, hi);
        if (e == 0u) { out.status = JD_SEG_BADCODE; break; }
        const uint32_t tot = e & 0x1Fu, len = (e >> 8) & 0xFFu, s = (e >> 16) & 0xFFu, adv = e >> 24;
        if (adv == 128u) {
            k = 64;
            last_was_eob = true;
        } else {
            const uint32_t kz = k + adv - 1u;    /* zigzag index of this symbol's coefficient */
            if (s && kz < 64u && own) {
                const uint32_t x = hi << len;    /* the S extra bits at the top */
                const int v = jd_extend_top(x, s);
                if (s > 11u) { out.status = JD_SEG_BADSIZE; break; }
                if (len + s >= 18u) {
                    const uint32_t P1 = rel + len;
                    const uint32_t j1 = jw + (uint32_t)((int)(P1 >> 3) - Pb) * JD_JW_ONES;
                    const int p7 = (int)(P1 & 7u);
                    if (((j1 + 0x222222u) & 0x888888u) != 0u) {
                        bool any = false;
                        for (int c = 0; c < 6; c++) if (8 * (int)((j1 >> (4 * c)) & 15u) + p7 + (int)s > 64) any = true;
                        if (any) {
                            JDEvent ev;
                            ev.blk = blk0 + bi; ev.seg = slot; ev.j1 = j1; ev.field = (uint16_t)(x >> (32u - s));
                            ev.s = (uint8_t)s; ev.p7 = (uint8_t)p7; ev.ord = ncoef; ev.img = img;
                            sink.push(ev);
                        }
                    }
                }
                const uint32_t tw = tposw[kz];
                bflags |= tw;
                if (s >= 10u && !big) {
                    if (rp + ncoef + 2 > rend) { out.status = JD_SEG_OVERFLOW; break; }
                    for (uint32_t i = ncoef; i-- > 0u;) {
                        const uint32_t r = rec0[i];
                        rec0[2u * i] = (uint16_t)(r >> 10);
                        rec0[2u * i + 1u] = (uint16_t)(int16_t)((int)(r << 22) >> 22);
                    }
                    rp += ncoef;
                    big = 1;
                }
                if (big) {
                    if (rp + 2 > rend) { out.status = JD_SEG_OVERFLOW; break; }
                    rp[0] = (uint16_t)(tw & 63u); rp[1] = (uint16_t)(int16_t)v; rp += 2;
                } else {
                    if (rp >= rend) { out.status = JD_SEG_OVERFLOW; break; }
                    *rp++ = (uint16_t)(((tw & 63u) << 10) | ((uint32_t)v & 0x3FFu));
                }
                ncoef++;
            }
            k = kz + 1u;
            last_was_eob = false;
        }
        w.drop(tot);
        rel += tot;
        { const int nPb = (int)(rel >> 3); jw += (uint32_t)(nPb - Pb) * JD_JW_ONES; Pb = nPb; }
        if (k >= 64u) {
            if (own) {
                blk_hdr[bi] = jd_pack_hdr(rec_index0 + (uint32_t)(rec0 - rec), dcval, ncoef, big, JD_BF_HI(bflags), JD_BF_COLMASK(bflags));
                out.nown++;
                bi++;
                own = false;
            }
            k = 0;
            bsh += 4u;
            if (bsh == bsh_end) bsh = 0;
        }
    }
    if (!snapped) {
        /* last chunk: the scan ends here; state as jd_decode_segment leaves it (only the image-end matters to nobody) */
        if (!last_was_eob) jw = jd_jw_ckpt(jw);
        out.jmap = jw;
    }
    out.dcsum[0] = pred[0] - pred_in[0]; out.dcsum[1] = pred[1] - pred_in[1]; out.dcsum[2] = pred[2] - pred_in[2];
}
#endif
