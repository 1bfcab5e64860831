/*
 * oracle/jpegdec_oracle.c -- TEST INFRASTRUCTURE: a plain sequential C restatement of the
 * reference's hot path (bitbank2/JPEGDEC src/jpeg.inl), used ONLY as a checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product never calls it.
 *
 * Parity pinning: the reference's own tests pin no pixel values, so this
 * restatement is pinned against outputs of the reference itself, compiled unmodified into
 * oracle/_ref (see oracle/Makefile, tests/test_oracle.py): every bundled image x pixel type x
 * scale x both arithmetic builds, plus the digests committed in tests/golden/digests.json.
 *
 * It is written as the reference *behaves* -- one image, MCU by MCU, a literal 64-bit bit
 * window -- deliberately unlike the GPU formulation (per-segment decode + phase stitching), so
 * that the two are independent derivations of the same semantics.
 *
 * Function -> reference lines:
 *   or_filter            JPEGFilter                       src/jpeg.inl:1431-1540
 *   or_build_tables      JPEGMakeHuffTables (canonical)   src/jpeg.inl:1066-1275
 *   or_fix_quant         JPEGFixQuantD                    src/jpeg.inl:1789-1811
 *   or_decode_block      JPEGDecodeMCU                    src/jpeg.inl:2090-2274
 *   or_idct              JPEGIDCT (+DC-only :5146-5154)   src/jpeg.inl:2278-2798
 *   or_pixel_*           JPEGPixelLE/BE/RGB + SSE2 paths  src/jpeg.inl:3101-3278, :3409-3517, :4006-4308
 *   or_emit_mcu          JPEGPutMCU22/11/12/21/Gray/8BitGray  src/jpeg.inl:2799-4868
 *   or_dither_rows       JPEGDither                       src/jpeg.inl:4871-4940
 *   oracle_decode        DecodeJPEG                       src/jpeg.inl:4946-5357
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "jpegdec_oracle.h"

static const uint8_t ZZ_NAT[64] = { /* zigzag position -> natural index */
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

typedef struct {
    int w, h, ncomp, sub, dri, scan;
    int prog, scan_nc, ss, se, ahal;   /* SOF2: first-scan parameters (jpeg.inl:1406-1418) */
    int cq[4], cdc[4], cac[4];
    uint16_t q[4][64];
    uint8_t bits[8][16], vals[8][256];
    int defined[8];
    uint8_t scratch[4352]; /* DHT bytes as the reference parks them in usPixels (jpeg.inl:843) */
} OrHdr;

static int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

/* minimal marker walk: SOF0, DQT, DHT, DRI, SOS (jpeg.inl:1611-1763) */
static int or_parse(const uint8_t *d, int n, OrHdr *H)
{
    memset(H, 0, sizeof(*H));
    if (n < 256 || d[0] != 0xFF || d[1] != 0xD8) return 0;
    int off = 2;
    while (off + 4 <= n) {
        int m = be16(d + off), len = be16(d + off + 2);
        off += 2;
        if (m < 0xFFC0 || m == 0xFFFF) { off++; continue; }
        if (m == 0xFFC0 || m == 0xFFC2) {
            H->prog = (m == 0xFFC2);
            H->h = be16(d + off + 3); H->w = be16(d + off + 5); H->ncomp = d[off + 7];
            if (H->ncomp > 4) return 0;
            for (int i = 0; i < H->ncomp; i++) {
                if (i == 0) H->sub = d[off + 9];
                H->cq[i] = d[off + 10 + 3 * i];
                H->cdc[i] = d[off + 8 + 3 * i]; /* component id, matched at SOS */
            }
            if (H->ncomp == 1) H->sub = 0;
        } else if (m == 0xFFC1 || m == 0xFFC3) {
            return 0;
        } else if (m == 0xFFDD) {
            H->dri = be16(d + off + 2);
        } else if (m == 0xFFDB) {
            int p = off + 2, end = off + len;
            while (p < end) {
                int t = d[p++];
                for (int i = 0; i < 64; i++) {
                    if (t & 0xF0) { H->q[t & 3][i] = (uint16_t)be16(d + p); p += 2; }
                    else H->q[t & 3][i] = d[p++];
                }
            }
        } else if (m == 0xFFC4) {
            int p = off + 2, end = off + len;
            while (p + 17 <= end) {
                int t = d[p++];
                if (t & 0x10) t ^= 0x14;
                t &= 7;
                int tot = 0;
                for (int i = 0; i < 16; i++) { H->bits[t][i] = d[p]; H->scratch[t * 273 + i] = d[p]; tot += d[p++]; }
                if (tot > 256) return 0;
                for (int i = 0; i < tot; i++) { H->vals[t][i] = d[p]; H->scratch[t * 273 + 16 + i] = d[p++]; }
                H->defined[t] = 1;
            }
        } else if (m == 0xFFDA) {
            int nc = d[off + 2];
            int ids[4];
            for (int i = 0; i < H->ncomp; i++) ids[i] = H->cdc[i];
            for (int i = 0; i < nc; i++) {
                int id = d[off + 3 + 2 * i], tb = d[off + 4 + 2 * i];
                for (int j = 0; j < H->ncomp; j++) if (ids[j] == id) { H->cdc[j] = tb >> 4; H->cac[j] = tb & 15; }
            }
            H->scan_nc = nc; H->ss = d[off + 3 + 2 * nc]; H->se = d[off + 4 + 2 * nc]; H->ahal = d[off + 5 + 2 * nc];
            H->scan = off + len;
            return 1;
        }
        off += len;
    }
    return 0;
}

/* JPEGFilter: FF00 -> FF, every other FFxx pair removed */
static uint8_t *or_filter(const uint8_t *d, int from, int n, int *outn)
{
    uint8_t *f = (uint8_t *)calloc((size_t)(n - from) + 64, 1);
    int k = 0;
    for (int i = from; i < n; i++) {
        if (d[i] == 0xFF) {
            if (i + 1 < n && d[i + 1] == 0) f[k++] = 0xFF;
            i++;
        } else f[k++] = d[i];
    }
    *outn = k;
    return f;
}

/* canonical Huffman decode tables: per code length the first code and value index */
typedef struct { int mincode[17], maxcode[18], valptr[17]; const uint8_t *vals; } OrHuff;
static void or_build_tables(const OrHdr *H, int t, OrHuff *T)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        T->valptr[l] = k; T->mincode[l] = code;
        code += H->bits[t][l - 1]; k += H->bits[t][l - 1];
        T->maxcode[l] = H->bits[t][l - 1] ? code - 1 : -1;
        code <<= 1;
    }
    T->maxcode[17] = 0x7FFFFFFF;
    T->vals = H->vals[t];
}

/* the reference's bit window, literally: 64 bits loaded at a byte position, a bit offset that
 * is only rebased at fixed points (src/JPEGDEC.h:128-133, jpeg.inl:2110-2114) */
typedef struct { const uint8_t *f; int pos; uint64_t bits; int off; } OrWin;
static uint64_t or_load(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; return v; }
static void or_rebase(OrWin *W) { if (W->off > 47) { W->pos += W->off >> 3; W->off &= 7; W->bits = or_load(W->f + W->pos); } }

static int or_code(OrWin *W, const OrHuff *T, int maxlen, int *len)
{
    uint32_t peek = (uint32_t)((W->bits >> (64 - 16 - W->off)) & 0xFFFF);
    for (int l = 1; l <= maxlen; l++) {
        int c = (int)(peek >> (16 - l));
        if (T->maxcode[l] >= 0 && c <= T->maxcode[l] && c >= T->mincode[l]) { *len = l; return T->vals[T->valptr[l] + c - T->mincode[l]]; }
    }
    return -1;
}

/* JPEGDecodeMCU.  store_limit: 64 full, 5 quarter/eighth, 1 skipped block.  Returns 0 ok. */
static int or_decode_block(OrWin *W, const OrHuff *dc, const OrHuff *ac, int *pred, int16_t *blk, int store_limit, unsigned *flags)
{
    int len;
    or_rebase(W);                                    /* :2110 */
    memset(blk, 0, 64 * sizeof(int16_t));
    *flags = 0;
    int s = or_code(W, dc, 12, &len);
    if (s < 0) return -1;
    s &= 15;
    W->off += len;
    if (s) {
        int pre = (len + s <= 6);                    /* magnitude folded into the LUT (:1132): no rebase */
        if (!pre) or_rebase(W);                      /* :2149 */
        uint64_t c = W->bits << W->off;
        int v = (int)(c >> (64 - s));
        if (!(c >> 63)) v -= (1 << s) - 1;
        W->off += s;
        *pred += v;
    }
    blk[0] = (int16_t)*pred;
    int k = 1;
    while (k < 64) {
        or_rebase(W);                                /* :2225 */
        int rs = or_code(W, ac, 16, &len);
        if (rs < 0) return -1;
        W->off += len;
        if (rs == 0) break;                          /* EOB: no trailing rebase (:2241) */
        k += rs >> 4;
        s = rs & 15;
        if (k < store_limit && s) {
            uint64_t c = W->bits << W->off;          /* no rebase here: bits past the window read 0 (:2249) */
            int v = (int)(c >> (64 - s));
            if (!(c >> 63)) v -= (1 << s) - 1;
            int n = ZZ_NAT[k];
            blk[n] = (int16_t)v;
            *flags |= 1u << (n & 7);
            *flags |= (unsigned)n << 8;
        }
        W->off += s;
        k++;
        or_rebase(W);                                /* :2259 */
    }
    return 0;
}

/* JPEGDecodeMCU_P for the only case the reference can finish: first scan = DC scan (iScanStart = iScanEnd = 0,
 * cApproxBitsHigh = 0), jpeg.inl:1837-1884: window reload at bit offset > 47 before the code and before the extra bits
 * (so no truncated reads), difference << cApproxBitsLow added to the predictor. */
static int or_decode_dc_prog(OrWin *W, const OrHuff *dc, int *pred, int al)
{
    int len;
    or_rebase(W);
    int s = or_code(W, dc, 12, &len);
    if (s < 0) return -1;
    s &= 15;
    W->off += len;
    if (s) {
        or_rebase(W);
        uint64_t c = W->bits << W->off;
        int v = (int)(c >> (64 - s));
        if (!(c >> 63)) v -= (1 << s) - 1;
        W->off += s;
        *pred += (int)((unsigned)v << al);
    }
    return 0;
}

static int or_scale(int r, int c)
{
    /* AAN prescale 16384 * s_r * s_c, s_0 = 1, s_k = sqrt(2) cos(k pi/16): the standard IFAST table
     * (the reference's iScaleBits, jpeg.inl:146-153); tests/test_host.py re-derives it from the formula */
    static const int T[64] = {
        16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520, 22725, 31521, 29692, 26722, 22725, 17855, 12299, 6270,
        21407, 29692, 27969, 25172, 21407, 16819, 11585, 5906, 19266, 26722, 25172, 22654, 19266, 15137, 10426, 5315,
        16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520, 12873, 17855, 16819, 15137, 12873, 10114, 6967, 3552,
        8867, 12299, 11585, 10426, 8867, 6967, 4799, 2446, 4520, 6270, 5906, 5315, 4520, 3552, 2446, 1247};
    return T[r * 8 + c];
}

/* JPEGFixQuantD: tables with index < ncomp are de-zigzagged and prescaled; others left raw */
static void or_fix_quant(const OrHdr *H, int16_t out[4][64])
{
    for (int t = 0; t < 4; t++)
        for (int n = 0; n < 64; n++) {
            if (t < H->ncomp) {
                int z = 0;
                for (int k = 0; k < 64; k++) if (ZZ_NAT[k] == n) z = k;
                out[t][n] = (int16_t)(uint16_t)(((unsigned)H->q[t][z] * (unsigned)or_scale(n >> 3, n & 7)) >> 12);
            } else out[t][n] = (int16_t)H->q[t][n];
        }
}

static uint8_t or_range(int v)
{
    int i = (v >> 5) & 0x3ff; /* ucRangeTable (:159-222) */
    if (i < 128) return (uint8_t)(i + 128);
    if (i < 512) return 255;
    if (i < 896) return 0;
    return (uint8_t)(i - 896);
}

#define S16(x) ((int16_t)(x))
static int16_t mh(int16_t a, int k) { return (int16_t)(((int)a * k) >> 16); }
static int16_t sl2(int16_t a) { return (int16_t)((uint16_t)a << 2); }

/* JPEGIDCT for one block; out = 64 pixel bytes.  mode: 0 full/half, 2 quarter. */
static void or_idct(const int16_t *m, const int16_t *q, unsigned flags, int arith, int mode, uint8_t *out)
{
    int16_t w[64];
    if (mode == 2) { /* :2305-2326 */
        int a = m[0] * q[0], b = m[8] * q[8], t0 = a + b, t2 = a - b;
        a = m[1] * q[1]; b = m[9] * q[9];
        int t1 = a + b, t3 = a - b;
        out[0] = or_range(t0 + t1); out[1] = or_range(t0 - t1); out[2] = or_range(t2 + t3); out[3] = or_range(t2 - t3);
        return;
    }
    const int low_only = (flags & 0x2000) == 0;
    if (arith == 0) { /* SSE2 build: int16 lanes, all 8 columns (:2327-2440) */
        for (int c = 0; c < 8; c++) {
            int16_t d[8], T0, T1, T2, T3, T4, T5, T6, T7;
            for (int r = 0; r < 8; r++) d[r] = S16(m[r * 8 + c] * q[r * 8 + c]);
            if (low_only) {
                int16_t t12 = mh(sl2(d[2]), 1697 * 4);
                T0 = S16(d[0] + d[2]); T3 = S16(d[0] - d[2]); T1 = S16(d[0] + t12); T2 = S16(d[0] - t12);
                T7 = S16(d[1] + d[3]);
                int16_t df = S16(d[1] - d[3]);
                int16_t t11 = mh(sl2(df), 5793 * 4), z5 = mh(sl2(df), 7568 * 4);
                t12 = mh(sl2(d[3]), 10703 * 2); t12 = S16(t12 + t12); t12 = S16(t12 + z5);
                T6 = S16(t12 - T7); T5 = S16(t11 - T6);
                T4 = S16(S16(mh(sl2(d[1]), 4433 * 4) - z5) + T5);
            } else {
                int16_t t10 = S16(d[0] + d[4]), t11 = S16(d[0] - d[4]), t13 = S16(d[2] + d[6]);
                int16_t t12 = S16(mh(sl2(S16(d[2] - d[6])), 5793 * 4) - t13);
                T0 = S16(t10 + t13); T3 = S16(t10 - t13); T1 = S16(t11 + t12); T2 = S16(t11 - t12);
                int16_t z13 = S16(d[5] + d[3]), z10 = S16(d[5] - d[3]), z11 = S16(d[1] + d[7]), z12 = S16(d[1] - d[7]);
                T7 = S16(z11 + z13);
                t11 = mh(sl2(S16(z11 - z13)), 5793 * 4);
                int16_t z5 = mh(sl2(S16(z10 + z12)), 7568 * 4);
                t12 = mh(sl2(z10), -10703 * 2); t12 = S16(t12 + t12); t12 = S16(t12 + z5);
                T6 = S16(t12 - T7); T5 = S16(t11 - T6);
                T4 = S16(S16(mh(sl2(z12), 4433 * 4) - z5) + T5);
            }
            w[c] = S16(T0 + T7); w[8 + c] = S16(T1 + T6); w[16 + c] = S16(T2 + T5); w[24 + c] = S16(T3 - T4);
            w[32 + c] = S16(T3 + T4); w[40 + c] = S16(T2 - T5); w[48 + c] = S16(T1 - T6); w[56 + c] = S16(T0 - T7);
        }
    } else { /* -DNO_SIMD build (:2555-2678): only flagged columns (+ column 0); others stay as they are */
        memcpy(w, m, sizeof(w));
        unsigned f = flags | 1;
        for (int c = 0; c < 8; c++) {
            if (!(f & (1u << c))) continue;
            int t0, t1, t2, t3, t4, t5, t6, t7, t10, t11, t12, t13, z5, z10, z11, z12, z13;
            if (low_only) {
                t10 = m[c] * q[c]; t1 = m[c + 16] * q[c + 16]; t12 = (t1 * 106) >> 8;
                t0 = t10 + t1; t3 = t10 - t1; t1 = t10 + t12; t2 = t10 - t12;
                t4 = m[c + 8] * q[c + 8];
                t5 = m[c + 24];
                if (t5) {
                    t5 *= q[c + 24]; t7 = t4 + t5; t11 = ((t4 - t5) * 362) >> 8; z5 = ((t4 - t5) * 473) >> 8;
                    t12 = ((-t5 * -669) >> 8) + z5; t6 = t12 - t7; t5 = t11 - t6; t10 = ((t4 * 277) >> 8) - z5; t4 = t10 + t5;
                } else { t7 = t4; t5 = (145 * t4) >> 8; t6 = (217 * t4) >> 8; t4 = (-51 * t4) >> 8; }
            } else {
                t0 = m[c] * q[c]; t2 = m[c + 32] * q[c + 32]; t10 = t0 + t2; t11 = t0 - t2;
                t1 = m[c + 16] * q[c + 16]; t3 = m[c + 48] * q[c + 48]; t13 = t1 + t3; t12 = (((t1 - t3) * 362) >> 8) - t13;
                t0 = t10 + t13; t3 = t10 - t13; t1 = t11 + t12; t2 = t11 - t12;
                t5 = m[c + 24] * q[c + 24]; t6 = m[c + 40] * q[c + 40]; z13 = t6 + t5; z10 = t6 - t5;
                t4 = m[c + 8] * q[c + 8]; t7 = m[c + 56] * q[c + 56]; z11 = t4 + t7; z12 = t4 - t7;
                t7 = z11 + z13; t11 = ((z11 - z13) * 362) >> 8; z5 = ((z10 + z12) * 473) >> 8;
                t12 = ((z10 * -669) >> 8) + z5; t6 = t12 - t7; t5 = t11 - t6; t10 = ((z12 * 277) >> 8) - z5; t4 = t10 + t5;
            }
            w[c] = S16(t0 + t7); w[c + 8] = S16(t1 + t6); w[c + 16] = S16(t2 + t5); w[c + 24] = S16(t3 - t4);
            w[c + 32] = S16(t3 + t4); w[c + 40] = S16(t2 - t5); w[c + 48] = S16(t1 - t6); w[c + 56] = S16(t0 - t7);
        }
    }
    for (int r = 0; r < 64; r += 8) { /* rows (:2681-2797) */
        int t0, t1, t2, t3, t4, t5, t6, t7;
        const int16_t *p = w + r;
        if ((flags & 0xf0) == 0) {
            if ((flags & 0xfc) == 0) {
                t0 = t1 = t2 = t3 = p[0]; t7 = p[1]; t6 = (t7 * 217) >> 8; t5 = (t7 * 145) >> 8; t4 = -((t7 * 51) >> 8);
            } else {
                int t10 = p[0], t13 = p[2], t12 = (t13 * 106) >> 8;
                t0 = t10 + t13; t3 = t10 - t13; t1 = t10 + t12; t2 = t10 - t12;
                int z13 = p[3], z11 = p[1];
                t7 = z11 + z13;
                int t11 = ((z11 - z13) * 362) >> 8, z5 = ((z11 - z13) * 473) >> 8;
                t10 = ((z11 * 277) >> 8) - z5; t12 = ((z13 * 669) >> 8) + z5;
                t6 = t12 - t7; t5 = t11 - t6; t4 = t10 + t5;
            }
        } else {
            int t10 = p[0] + p[4], t11 = p[0] - p[4], t13 = p[2] + p[6], t12 = (((p[2] - p[6]) * 362) >> 8) - t13;
            t0 = t10 + t13; t3 = t10 - t13; t1 = t11 + t12; t2 = t11 - t12;
            int z13 = p[5] + p[3], z10 = p[5] - p[3], z11 = p[1] + p[7], z12 = p[1] - p[7];
            t7 = z11 + z13; t11 = ((z11 - z13) * 362) >> 8;
            int z5 = ((z10 + z12) * 473) >> 8;
            t10 = ((z12 * 277) >> 8) - z5; t12 = ((z10 * -669) >> 8) + z5;
            t6 = t12 - t7; t5 = t11 - t6; t4 = t10 + t5;
        }
        out[r + 0] = or_range(t0 + t7); out[r + 1] = or_range(t1 + t6); out[r + 2] = or_range(t2 + t5); out[r + 3] = or_range(t3 - t4);
        out[r + 4] = or_range(t3 + t4); out[r + 5] = or_range(t2 - t5); out[r + 6] = or_range(t1 - t6); out[r + 7] = or_range(t0 - t7);
    }
}

/* ---- colour ---- */
static int clamp8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static int rt(int v) { v &= 0x3ff; return v < 256 ? v : (v < 512 ? 255 : 0); } /* usRangeTableR/G/B (:262-555) */
static void or_ycc(int Y12, int Cb, int Cr, int *R, int *G, int *B)
{
    Cb -= 128; Cr -= 128;
    *B = (7258 * Cb + Y12) >> 12; *G = (-1409 * Cb - 2925 * Cr + Y12) >> 12; *R = (5742 * Cr + Y12) >> 12;
}
static void or_put_scalar(uint8_t *dst, int pt, int Y12, int Cb, int Cr)
{
    int R, G, B;
    or_ycc(Y12, Cb, Cr, &R, &G, &B);
    if (pt == OR_RGB8888) { dst[0] = (uint8_t)clamp8(R); dst[1] = (uint8_t)clamp8(G); dst[2] = (uint8_t)clamp8(B); dst[3] = 0xFF; }
    else {
        unsigned v = (unsigned)((rt(R) >> 3) << 11) | (unsigned)((rt(G) >> 2) << 5) | (unsigned)(rt(B) >> 3);
        if (pt == OR_RGB565_BE) { dst[0] = (uint8_t)(v >> 8); dst[1] = (uint8_t)v; } else { dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); }
    }
}
static void or_put_sse(uint8_t *dst, int pt, int Y, int Cb, int Cr)
{
    int16_t cb = (int16_t)((Cb - 128) << 8), cr = (int16_t)((Cr - 128) << 8), y4 = (int16_t)(Y << 4);
    int R = clamp8(S16(y4 + mh(cr, 5742)) >> 4);
    int G = clamp8(S16(S16(y4 + mh(cr, -2925)) + mh(cb, -1409)) >> 4);
    int B = clamp8(S16(y4 + mh(cb, 7258)) >> 4);
    if (pt == OR_RGB8888) { dst[0] = (uint8_t)B; dst[1] = (uint8_t)G; dst[2] = (uint8_t)R; dst[3] = 0xFF; }
    else { unsigned v = (unsigned)((R >> 3) << 11) | (unsigned)((G >> 2) << 5) | (unsigned)(B >> 3); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); }
}

/* Floyd-Steinberg rows of one MCU row (:4871-4940); errors persists between calls */
static void or_dither_rows(const uint8_t *src, int W, int rows, int bits, uint8_t *errors, uint8_t *dst, int dpitch)
{
    const int mask = bits == 4 ? 0xF0 : (bits == 2 ? 0xC0 : 0x80), xmask = bits == 4 ? 1 : (bits == 2 ? 3 : 7);
    errors[0] = errors[1] = errors[2] = 0;
    for (int y = 0; y < rows; y++) {
        const uint8_t *p = src + (size_t)y * W;
        uint8_t *d = dst + (size_t)y * dpitch, *pe = errors + 1;
        int fwd = 0;
        unsigned acc = 0;
        for (int x = 0; x < W; x++) {
            int c = p[x] + fwd;
            if (c > 255) c = 255;
            acc = ((acc << bits) | (unsigned)(c >> (8 - bits))) & 0xFF;
            if ((x & xmask) == xmask) { *d++ = (uint8_t)acc; acc = 0; }
            int v = c - (c & mask), h = v >> 1, e1 = (7 * h) >> 3, e2 = h - e1, e3 = (5 * h) >> 3, e4 = h - e3;
            fwd = e1 + pe[1];
            pe[1] = (uint8_t)e2; pe[0] = (uint8_t)(pe[0] + e3); pe[-1] = (uint8_t)(pe[-1] + e4);
            pe++;
        }
    }
}

int oracle_decode(const uint8_t *jpeg, int len, int pixel_type, int options, int arith,
                  uint8_t *out, int out_pitch, int *out_w, int *out_h)
{
    OrHdr *H = (OrHdr *)malloc(sizeof(OrHdr));
    if (!or_parse(jpeg, len, H)) { free(H); return 0; }
    int hs = 1, vs = 1;
    switch (H->sub) { case 0x00: case 0x11: break; case 0x21: hs = 2; break; case 0x12: vs = 2; break; case 0x22: hs = vs = 2; break; default: free(H); return 0; }
    if (H->ncomp != 1 && H->ncomp != 3) { free(H); return 0; }
    if (H->prog) options |= OR_SCALE_EIGHTH;         /* progressive: DC of the first scan only -> 1/8 image (:4964-4966) */
    const int sh = (options & OR_SCALE_HALF) ? 1 : (options & OR_SCALE_QUARTER) ? 2 : (options & OR_SCALE_EIGHTH) ? 3 : 0;
    if (H->prog && (sh != 3 || H->scan_nc != H->ncomp || H->ss != 0 || H->se != 0 || (H->ahal >> 4) != 0)) { free(H); return 0; }
    if ((options & OR_LUMA_ONLY) && pixel_type < OR_GRAY8) pixel_type = OR_GRAY8;
    const int dbits = pixel_type == OR_DITHER4 ? 4 : pixel_type == OR_DITHER2 ? 2 : pixel_type == OR_DITHER1 ? 1 : 0;
    const int gray_out = pixel_type >= OR_GRAY8;
    const int cx = (H->w + hs * 8 - 1) / (hs * 8), cy = (H->h + vs * 8 - 1) / (vs * 8);
    const int nl = hs * vs, bpm = nl + (H->ncomp == 3 ? 2 : 0);
    int16_t Q[4][64];
    or_fix_quant(H, Q);
    OrHuff dct[2], act[2];
    for (int t = 0; t < 2; t++) { or_build_tables(H, t, &dct[t]); or_build_tables(H, 4 + t, &act[t]); }
    int fn;
    uint8_t *F = or_filter(jpeg, H->scan, len, &fn);
    OrWin W = {F, 0, or_load(F), 0};
    const int bs = sh >= 2 ? (8 >> sh) : 8;          /* bytes per block edge kept by the IDCT stage */
    const int pw = cx * hs * bs, ph = cy * vs * bs;  /* luma plane, MCU aligned */
    const int cw = cx * bs, chh = cy * bs;
    uint8_t *PY = (uint8_t *)calloc((size_t)pw * ph + 64, 1), *PB = (uint8_t *)calloc((size_t)cw * chh + 64, 1), *PR = (uint8_t *)calloc((size_t)cw * chh + 64, 1);
    int pred[3] = {0, 0, 0}, rc = 1, rescount = H->dri;
    for (int my = 0; my < cy && rc; my++) {
        for (int mx = 0; mx < cx && rc; mx++) {
            for (int b = 0; b < bpm; b++) {
                const int comp = b < nl ? 0 : b - nl + 1;
                int16_t blk[64];
                uint8_t px[64];
                unsigned flags;
                const int skip = comp > 0 && gray_out;  /* chroma parsed with MCU_SKIP (:5225-5233) */
                const int limit = skip ? 1 : (sh >= 2 ? 5 : 64);
                if (H->prog) {
                    flags = 0;
                    if (H->cdc[comp] > 1 || or_decode_dc_prog(&W, &dct[H->cdc[comp]], &pred[comp], H->ahal & 15)) { rc = 0; break; }
                } else
                if (H->cdc[comp] > 1 || H->cac[comp] > 1 || or_decode_block(&W, &dct[H->cdc[comp]], &act[H->cac[comp]], &pred[comp], blk, limit, &flags)) { rc = 0; break; }
                if (skip) continue;
                const int16_t *q = Q[H->cq[comp] & 3];
                if (flags == 0 || sh == 3) memset(px, or_range(pred[comp] * q[0]), 64);   /* :5146-5154 */
                else or_idct(blk, q, flags, arith, sh == 2 ? 2 : 0, px);
                uint8_t *pl; int plw, bx, by;
                if (comp == 0) { pl = PY; plw = pw; bx = mx * hs + (hs == 2 ? (b & 1) : 0); by = my * vs + (hs == 2 && vs == 2 ? (b >> 1) : (vs == 2 ? b : 0)); }
                else { pl = comp == 1 ? PB : PR; plw = cw; bx = mx; by = my; }
                for (int y = 0; y < bs; y++) memcpy(pl + (size_t)(by * bs + y) * plw + bx * bs, px + y * bs, (size_t)bs);
            }
            if (H->dri && --rescount == 0) {           /* :5337-5347 */
                rescount = H->dri; pred[0] = pred[1] = pred[2] = 0;
                if (W.off & 7) W.off += 8 - (W.off & 7);
            }
        }
    }
    /* ---- pixels: what JPEGPutMCU* deliver, expressed per output pixel ---- */
    const int ow = (H->w + (1 << sh) - 1) >> sh, oh = (H->h + (1 << sh) - 1) >> sh;
    *out_w = ow; *out_h = oh;
    const int sse_full = arith == 0 && sh == 0 && (H->sub == 0x22 || H->sub == 0x11);
    const int gw = dbits ? cx * ((hs * 8) >> sh) : ow;   /* dither works on whole MCU rows */
    uint8_t *gray = dbits ? (uint8_t *)calloc((size_t)gw * (size_t)(cy * ((vs * 8) >> sh)) + 64, 1) : NULL;
    const int rows_total = dbits ? cy * ((vs * 8) >> sh) : oh;
    for (int oy = 0; oy < rows_total; oy++) {
        for (int ox = 0; ox < gw; ox++) {
            int Y, Y12, Cb = 128, Cr = 128;
            if (sh == 1) {
                const uint8_t *yp = PY + (size_t)(2 * oy) * pw + 2 * ox;
                int sum = yp[0] + yp[1] + yp[pw] + yp[pw + 1];
                Y = (sum + 2) >> 2; Y12 = sum << 10;
                if (H->ncomp == 3 && !gray_out) {
                    if (hs == 2 && vs == 2) { Cb = PB[(size_t)oy * cw + ox]; Cr = PR[(size_t)oy * cw + ox]; }
                    else if (hs == 1 && vs == 1) {
                        const uint8_t *a = PB + (size_t)(2 * oy) * cw + 2 * ox, *c = PR + (size_t)(2 * oy) * cw + 2 * ox;
                        Cb = (a[0] + a[1] + a[cw] + a[cw + 1] + 2) >> 2; Cr = (c[0] + c[1] + c[cw] + c[cw + 1] + 2) >> 2;
                    } else if (hs == 2) {
                        Cb = (PB[(size_t)(2 * oy) * cw + ox] + PB[(size_t)(2 * oy + 1) * cw + ox] + 1) >> 1;
                        Cr = (PR[(size_t)(2 * oy) * cw + ox] + PR[(size_t)(2 * oy + 1) * cw + ox] + 1) >> 1;
                    } else {
                        Cb = (PB[(size_t)oy * cw + 2 * ox] + PB[(size_t)oy * cw + 2 * ox + 1] + 1) >> 1;
                        Cr = (PR[(size_t)oy * cw + 2 * ox] + PR[(size_t)oy * cw + 2 * ox + 1] + 1) >> 1;
                    }
                }
            } else {
                Y = PY[(size_t)oy * pw + ox]; Y12 = Y << 12;
                if (H->ncomp == 3 && !gray_out) { Cb = PB[(size_t)(oy / vs) * cw + ox / hs]; Cr = PR[(size_t)(oy / vs) * cw + ox / hs]; }
            }
            if (dbits) { gray[(size_t)oy * gw + ox] = (uint8_t)Y; continue; }
            uint8_t *dst = out + (size_t)oy * out_pitch;
            if (gray_out) dst[ox] = (uint8_t)Y;
            else if (H->ncomp == 1) {
                unsigned v = (unsigned)((Y >> 3) << 11) | (unsigned)((Y >> 2) << 5) | (unsigned)(Y >> 3); /* usGrayTo565 (:227-258) */
                if (pixel_type == OR_RGB565_LE) { dst[2 * ox] = (uint8_t)v; dst[2 * ox + 1] = (uint8_t)(v >> 8); }
                else { dst[2 * ox] = (uint8_t)(v >> 8); dst[2 * ox + 1] = (uint8_t)v; }
            } else if (sse_full) or_put_sse(dst + (size_t)ox * (pixel_type == OR_RGB8888 ? 4 : 2), pixel_type, Y, Cb, Cr);
            else or_put_scalar(dst + (size_t)ox * (pixel_type == OR_RGB8888 ? 4 : 2), pixel_type, Y12, Cb, Cr);
        }
    }
    if (dbits) {
        const int mrows = (vs * 8) >> sh, dpitch = (gw * dbits + 7) / 8;
        uint8_t *errors = (uint8_t *)calloc((size_t)gw + 4400, 1);
        memcpy(errors, H->scratch, sizeof(H->scratch)); /* the error line aliases the DHT scratch area (:4881) */
        uint8_t *tmp = (uint8_t *)calloc((size_t)dpitch * mrows + 16, 1);
        for (int my = 0; my < cy; my++) {
            or_dither_rows(gray + (size_t)my * mrows * gw, gw, mrows, dbits, errors, tmp, dpitch);
            for (int r = 0; r < mrows && my * mrows + r < oh; r++) memcpy(out + (size_t)(my * mrows + r) * out_pitch, tmp + (size_t)r * dpitch, (size_t)((ow * dbits + 7) / 8));
        }
        free(errors); free(tmp); free(gray);
    }
    free(PY); free(PB); free(PR); free(F); free(H);
    return rc;
}

/* test hook: one block through the restated IDCT (flags as JPEGDecodeMCU leaves them; mode 0 full, 2 quarter) */
void oracle_idct(const int16_t *coef, const int16_t *quant, unsigned flags, int arith, int mode, uint8_t *out)
{
    or_idct(coef, quant, flags, arith, mode, out);
}
