/*
 * jd_host.c -- host-side header parsing and table construction (plain C).
 *
 * Replaces, for the GPU pipeline, the reference's JPEGParseInfo (src/jpeg.inl:1572-1785),
 * JPEGGetHuffTables (:837-873), JPEGGetSOS (:1378-1425), the acceptance rules of
 * JPEGMakeHuffTables (:1066-1275) and JPEGFixQuantD (:1789-1811).  Written fresh against a
 * whole-file buffer (the reference walks a 2 KB window); behaviour that decides
 * open()'s return value / error code is kept, reads are bounds-checked.
 */
#include <string.h>
#include <stdlib.h>
#include "jd_internal.h"

#define HUFF_TABLEN 273 /* reference src/JPEGDEC.h:58: stride of one DHT table in the scratch area */

static unsigned be16(const uint8_t *p) { return ((unsigned)p[0] << 8) | p[1]; }

static unsigned tiff16(const uint8_t *p, int mot) { return mot ? ((unsigned)p[0] << 8) | p[1] : ((unsigned)p[1] << 8) | p[0]; }
static unsigned tiff32(const uint8_t *p, int mot)
{
    return mot ? ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]
               : ((unsigned)p[3] << 24) | ((unsigned)p[2] << 16) | ((unsigned)p[1] << 8) | p[0];
}

/* value of one 12-byte TIFF tag (reference TIFFVALUE, src/jpeg.inl:1313-1345) */
static int tiff_value(const uint8_t *p, int mot)
{
    int type = (int)tiff16(p + 2, mot);
    if (tiff16(p + 4, mot) > 1) type = 4;
    switch (type) {
        case 3: return (int)tiff16(p + 8, mot);
        case 6: return (signed char)p[8];
        case 2: case 4: case 5: case 7: case 10: return (int)tiff32(p + 8, mot);
        default: return 0;
    }
}

/* one IFD: orientation / thumbnail size / thumbnail offset (reference GetTIFFInfo, :1346-1376) */
static void tiff_ifd(const uint8_t *data, int size, int off, int mot, JDInfo *info)
{
    if (off < 0 || off + 2 > size) return;
    int n = (int)tiff16(data + off, mot);
    if (n < 1 || n > 256) return;
    for (int i = 0; i < n; i++) {
        int t = off + 2 + i * 12;
        if (t + 12 > size) return;
        int tag = (int)tiff16(data + t, mot);
        if (tag == 274) info->orientation = tiff_value(data + t, mot) & 0xFF;
        else if (tag == 256) info->thumb_w = tiff_value(data + t, mot);
        else if (tag == 257) info->thumb_h = tiff_value(data + t, mot);
        else if (tag == 513) info->thumb_data = tiff_value(data + t, mot);
    }
}

/* DHT payload (reference JPEGGetHuffTables, :837-873).  Returns 0 ok, -1 bad. */
static int parse_dht(const uint8_t *p, int len, int avail, JDInfo *info)
{
    int off = 0;
    uint8_t *hv = info->p.huffvals;
    while (len > 17) {
        if (off + 17 > avail) return -1;
        unsigned t = p[off++];
        if (t & 0x10) t ^= 0x14; /* AC class -> tables 4..7 */
        if (t <= 7) {
            info->p.huff_defined |= (uint8_t)(1u << t);
            int base = (int)t * HUFF_TABLEN, total = 0;
            for (int i = 0; i < 16; i++) { total += p[off]; hv[base + i] = p[off++]; }
            len -= 17;
            if (total == 0 || total > 256 || total > len) return -1;
            if (off + total > avail) return -1;
            memcpy(hv + base + 16, p + off, (size_t)total);
            off += total;
            len -= total;
        }
    }
    return 0;
}

static int fail(JDInfo *info, int code) { info->error = code; return 0; }

int jd_parse_header(const uint8_t *data, int size, int start, JDInfo *info)
{
    const uint8_t *s = data;
    if (start == 0) {
        memset(info, 0, sizeof(*info));
    } else {
        /* thumbnail re-parse (:4967-4976): table state and EXIF facts persist like in the reference */
        JDInfo keep = *info;
        memset(info, 0, sizeof(*info));
        info->p = keep.p;
        info->orientation = keep.orientation; info->has_thumb = keep.has_thumb;
        info->thumb_w = keep.thumb_w; info->thumb_h = keep.thumb_h;
        info->thumb_data = keep.thumb_data; info->exif = keep.exif;
    }
    if (start < 0 || start > size) return fail(info, JPEG_INVALID_FILE);
    /* the reference reads up to 2048 bytes and rejects < 256 (:1597-1602) */
    if (size - start < 256) return fail(info, JPEG_INVALID_FILE);
    if (be16(s + start) != 0xFFD8) return fail(info, JPEG_INVALID_FILE);
    int off = start + 2;
    unsigned marker = 0, len = 0;
    while (marker != 0xFFDA && off < size) {
        if (off + 4 > size) return fail(info, JPEG_DECODE_ERROR);
        marker = be16(s + off);
        off += 2;
        len = be16(s + off);
        if (marker < 0xFFC0 || marker == 0xFFFF) { off++; continue; } /* resync (:1642-1646) */
        switch (marker) {
            case 0xFFC1: case 0xFFC3:
                return fail(info, JPEG_UNSUPPORTED_FEATURE);
            case 0xFFE1: /* APP1 / EXIF (:1654-1678) */
                if (off + 20 <= size && s[off + 2] == 'E' && s[off + 3] == 'x' && (s[off + 8] == 'M' || s[off + 8] == 'I')) {
                    int mot = (s[off + 8] == 'M');
                    int tiff = off + 8;
                    info->exif = tiff;
                    int ifd = (int)tiff32(s + off + 12, mot);
                    int ntags = (int)tiff16(s + off + 16, mot);
                    tiff_ifd(s, size, ifd + tiff, mot, info);
                    if (ntags >= 1 && ntags < 32) {
                        ifd += 12 * ntags + 2;
                        if (ifd >= 0 && ifd + tiff + 4 <= size) {
                            ifd = (int)tiff32(s + ifd + tiff, mot);
                            if (ifd != 0 && ifd + (tiff - start) < 2048) { /* window bound of the reference (:1671) */
                                info->has_thumb = 1;
                                tiff_ifd(s, size, ifd + tiff, mot, info);
                                info->thumb_data += tiff;
                            }
                        }
                    }
                }
                break;
            case 0xFFC0: case 0xFFC2: { /* SOF (:1679-1714) */
                if (off + 8 > size) return fail(info, JPEG_DECODE_ERROR);
                info->mode = (int)(marker & 0xFF);
                int bits = s[off + 2];
                info->height = (int)be16(s + off + 3);
                info->width = (int)be16(s + off + 5);
                info->ncomp = s[off + 7];
                info->bpp = (bits * info->ncomp) & 0xFF;
                if (info->ncomp > 4 || off + 8 + 3 * info->ncomp > size) return fail(info, JPEG_DECODE_ERROR);
                len -= 8; off += 8;
                for (int i = 0; i < info->ncomp; i++) {
                    info->p.comp_id[i] = s[off++];
                    unsigned samp = s[off++];
                    if (i == 0) info->subsample = (int)samp;
                    info->p.comp_quant[i] = s[off++];
                    if (info->p.comp_quant[i] > 3) return fail(info, JPEG_DECODE_ERROR);
                    len -= 3;
                }
                if (info->ncomp == 1) info->subsample = 0;
                len &= 0xFFFF;
                break;
            }
            case 0xFFDD: /* DRI (:1715-1718) */
                if (len == 4 && off + 4 <= size) info->restart_interval = (int)be16(s + off + 2);
                break;
            case 0xFFC4: /* DHT (:1719-1727) */
                off += 2; len = (len - 2) & 0xFFFF;
                if (parse_dht(s + off, (int)len, size - off, info) != 0) return fail(info, JPEG_DECODE_ERROR);
                break;
            case 0xFFDB: { /* DQT (:1728-1760) */
                off += 2;
                int rem = (int)len - 2;
                while (rem > 0) {
                    if (off >= size) return fail(info, JPEG_DECODE_ERROR);
                    unsigned t = s[off++];
                    if ((t & 0xF) > 3) return fail(info, JPEG_DECODE_ERROR);
                    uint16_t *q = info->p.quant_raw[t & 0xF];
                    if (t & 0xF0) {
                        if (off + 128 > size) return fail(info, JPEG_DECODE_ERROR);
                        for (int i = 0; i < 64; i++) { q[i] = (uint16_t)be16(s + off); off += 2; }
                        rem -= 129;
                    } else {
                        if (off + 64 > size) return fail(info, JPEG_DECODE_ERROR);
                        for (int i = 0; i < 64; i++) q[i] = s[off++];
                        rem -= 65;
                    }
                }
                len = 0; /* off already points past the payload the way the reference's bookkeeping ends up */
                break;
            }
            default:
                break;
        }
        off += (int)len;
    }
    if (marker != 0xFFDA) return fail(info, JPEG_DECODE_ERROR);
    /* SOS (reference JPEGGetSOS :1378-1425; its error return is ignored at :1769) */
    off -= (int)len;
    if (off + 3 > size) return fail(info, JPEG_DECODE_ERROR);
    int slen = (int)be16(s + off);
    off += 2;
    int nc = s[off++];
    info->p.ncomp_in_scan = (uint8_t)nc;
    slen -= 3;
    if (nc >= 1 && nc <= 4 && slen == nc * 2 + 3 && off + nc * 2 + 3 <= size) {
        int bad = 0;
        for (int i = 0; i < nc && !bad; i++) {
            unsigned cc = s[off++], c = s[off++];
            int j;
            for (j = 0; j < 4; j++) if (info->p.comp_id[j] == cc) break;
            if (j == 4) { bad = 1; break; }
            if ((c & 0xF) > 3 || (c & 0xF0) > 0x30) { bad = 1; break; }
            info->p.comp_dc[j] = (uint8_t)(c >> 4);
            info->p.comp_ac[j] = (uint8_t)(c & 0xF);
        }
        if (!bad) {
            info->p.scan_start = s[off++];
            info->p.scan_end = s[off++];
            info->approx = s[off++]; /* successive approximation: Ah << 4 | Al (:1417) */
        }
    }
    info->scan_offset = off;
    info->p.scan_offset = off;
    if (!jd_check_huffman(info)) return fail(info, JPEG_UNSUPPORTED_FEATURE);

    /* geometry (reference DecodeJPEG :5008-5049).  Deviation: sampling factors the reference
     * does not know end in a division by zero there (:5062); we refuse them at open. */
    if (info->ncomp != 1 && info->ncomp != 3) return fail(info, JPEG_UNSUPPORTED_FEATURE);
    switch (info->subsample) {
        case 0x00: case 0x11: info->mcu_w = 8; info->mcu_h = 8; info->bpm = (info->ncomp == 3) ? 3 : 1; break;
        case 0x21: info->mcu_w = 16; info->mcu_h = 8; info->bpm = 4; break;
        case 0x12: info->mcu_w = 8; info->mcu_h = 16; info->bpm = 4; break;
        case 0x22: info->mcu_w = 16; info->mcu_h = 16; info->bpm = 6; break;
        default: return fail(info, JPEG_UNSUPPORTED_FEATURE);
    }
    if (info->width <= 0 || info->height <= 0) return fail(info, JPEG_DECODE_ERROR);
    info->mcus_x = (info->width + info->mcu_w - 1) / info->mcu_w;
    info->mcus_y = (info->height + info->mcu_h - 1) / info->mcu_h;
    info->tsel = 0;
    info->tables_ok = 1;
    for (int c = 0; c < info->ncomp; c++) {
        if (info->p.comp_dc[c] > 1 || info->p.comp_ac[c] > 1) info->tables_ok = 0; /* :2166 / ucHuffDC holds 2 */
        info->tsel |= (info->p.comp_dc[c] & 1) << (2 * c);
        info->tsel |= (info->p.comp_ac[c] & 1) << (2 * c + 1);
    }
    info->error = JPEG_SUCCESS;
    return 1;
}

/* Code-length classes the reference's two-level LUTs can hold (JPEGMakeHuffTables :1066-1275). */
int jd_check_huffman(const JDInfo *info)
{
    const uint8_t *hv = info->p.huffvals;
    for (int t = 0; t < 4; t++) {
        if (!(info->p.huff_defined & (1u << t))) continue;
        const uint8_t *bits = hv + t * HUFF_TABLEN;
        unsigned cc = 0;
        for (int n = 1; n <= 16; n++) {
            int cnt = bits[n - 1];
            if (n > 12 && cnt > 0) return 0;
            while (cnt--) {
                int is_long = (n >= 5) && ((cc >> (n - 5)) == 0x1F);
                if (!is_long && n > 6) return 0;
                cc++;
            }
            cc <<= 1;
        }
    }
    if (info->mode == 0xC2) return 1;
    for (int t = 0; t < 4; t++) {
        if (!(info->p.huff_defined & (1u << (t + 4)))) continue;
        if (t >= 2) return 0; /* usHuffAC holds two tables (:1189-1190) */
        const uint8_t *bits = hv + (t + 4) * HUFF_TABLEN;
        unsigned cc = 0;
        for (int n = 1; n <= 16; n++) {
            int cnt = bits[n - 1];
            while (cnt--) {
                int is_long = (n >= 6) && ((cc >> (n - 6)) == 0x3F);
                if (!is_long && n > 10) return 0;
                cc++;
            }
            cc <<= 1;
        }
    }
    return 1;
}

/* Device LUT set: layout documented in jd_core.h. */
void jd_build_lut(const JDInfo *info, uint16_t *lut)
{
    const uint8_t *hv = info->p.huffvals;
    memset(lut, 0, JD_LUT_ENTRIES_H * sizeof(uint16_t));
    for (int t = 0; t < 2; t++) { /* DC */
        if (!(info->p.huff_defined & (1u << t))) continue;
        const uint8_t *bits = hv + t * HUFF_TABLEN, *vals = bits + 16;
        uint16_t *L = lut + t * 1152;
        unsigned cc = 0;
        for (int n = 1; n <= 16; n++) {
            int cnt = bits[n - 1];
            while (cnt--) {
                unsigned sym = *vals++;
                uint16_t e = (uint16_t)((n << 8) | (sym & 0xF));
                if (n >= 5 && n <= 12 && (cc >> (n - 5)) == 0x1F) {
                    unsigned first = (cc << (12 - n)) & 0x7F, rep = 1u << (12 - n);
                    for (unsigned i = 0; i < rep && first + i < 128; i++) L[1024 + first + i] = e;
                } else if (n <= 10) {
                    unsigned first = cc << (10 - n), rep = 1u << (10 - n);
                    for (unsigned i = 0; i < rep && first + i < 1024; i++) L[first + i] = e;
                }
                cc++;
            }
            cc <<= 1;
        }
    }
    for (int t = 0; t < 2; t++) { /* AC */
        if (!(info->p.huff_defined & (1u << (t + 4)))) continue;
        const uint8_t *bits = hv + (t + 4) * HUFF_TABLEN, *vals = bits + 16;
        uint16_t *L = lut + 2 * 1152 + t * 2048;
        unsigned cc = 0;
        for (int n = 1; n <= 16; n++) {
            int cnt = bits[n - 1];
            while (cnt--) {
                unsigned sym = *vals++;
                uint16_t e = (uint16_t)((n << 8) | sym);
                if (n >= 6 && (cc >> (n - 6)) == 0x3F) {
                    unsigned first = (cc << (16 - n)) & 0x3FF, rep = 1u << (16 - n);
                    for (unsigned i = 0; i < rep && first + i < 1024; i++) L[1024 + first + i] = e;
                } else if (n <= 10) {
                    unsigned first = cc << (10 - n), rep = 1u << (10 - n);
                    for (unsigned i = 0; i < rep && first + i < 1024; i++) L[first + i] = e;
                }
                cc++;
            }
            cc <<= 1;
        }
    }
    /* fast AC tables (jd_core.h JD_LUT_ACF): the next 10 bits alone decide.  Prefixes 111111xxxx belong to the long-code
     * half of the table above; such a prefix gets a direct entry when all of its 64 extensions are one code of <= 10 bits,
     * else len = 0 = "look in the long-code half".  Invalid prefixes are len = 0 too (the decoder tells them apart). */
    for (int t = 0; t < 2; t++) {
        const uint16_t *L = lut + 2 * 1152 + t * 2048;
        uint16_t *F = lut + 2 * 1152 + 2 * 2048 + t * 2048;      /* 1024 32-bit entries (little endian halves) */
        for (unsigned idx = 0; idx < 1024; idx++) {
            uint16_t e = 0;
            if ((idx >> 4) != 0x3F) e = L[idx];
            else {
                const uint16_t *x = L + 1024 + ((idx & 15u) << 6);
                e = x[0];
                for (int j = 1; j < 64; j++) if (x[j] != e) e = 0;
                if ((e >> 8) > 10) e = 0;
            }
            uint32_t f = 0;
            if (e) {
                const unsigned len = e >> 8, rs = e & 0xFFu, s = rs & 15u;
                f = (len + s) | (len << 8) | (s << 16) | ((rs == 0 ? 128u : (rs >> 4) + 1u) << 24);
                if (s >= 10 || len + s >= 18) f |= 0x80u;
            }
            F[2 * idx] = (uint16_t)(f & 0xFFFFu);
            F[2 * idx + 1] = (uint16_t)(f >> 16);
        }
    }
}

/* AAN prescale factors 16384 * s[r] * s[c], s[0] = 1, s[k] = cos(k*pi/16) * sqrt(2)
 * (the IFAST scaling every AAN integer IDCT uses; the reference's copy is iScaleBits,
 * src/jpeg.inl:146-153; tests/test_host.py re-derives the numbers from the formula). */
static const int jd_aan_scale[64] = {
    16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520,
    22725, 31521, 29692, 26722, 22725, 17855, 12299, 6270,
    21407, 29692, 27969, 25172, 21407, 16819, 11585, 5906,
    19266, 26722, 25172, 22654, 19266, 15137, 10426, 5315,
    16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520,
    12873, 17855, 16819, 15137, 12873, 10114, 6967, 3552,
    8867, 12299, 11585, 10426, 8867, 6967, 4799, 2446,
    4520, 6270, 5906, 5315, 4520, 3552, 2446, 1247};

const int *jd_aan_table(void) { return jd_aan_scale; }

/* natural index -> zigzag index */
static const uint8_t jd_zigzag_of_natural[64] = {
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

/* Prescaled quant per component, natural order, read back as signed 16-bit like the reference does
 * (JPEGFixQuantD :1789-1811 + pQuant as signed short :2304).  Quirk kept: only tables with index
 * < number of components are reordered/prescaled (:1796). */
void jd_build_quant(const JDInfo *info, int16_t *q)
{
    for (int c = 0; c < 3; c++) {
        int t = (c < info->ncomp) ? info->p.comp_quant[c] : 0;
        const uint16_t *raw = info->p.quant_raw[t & 3];
        for (int n = 0; n < 64; n++) {
            uint16_t v;
            if (t < info->ncomp) v = (uint16_t)(((unsigned)raw[jd_zigzag_of_natural[n]] * (unsigned)jd_aan_scale[n]) >> 12);
            else v = raw[n];
            q[c * 64 + n] = (int16_t)v;
        }
    }
}

/* 1 when both headers define the same Huffman tables (same table ids, same counts, same symbols) */
int jd_tables_equal(const JDInfo *x, const JDInfo *y)
{
    if (x->p.huff_defined != y->p.huff_defined) return 0;
    for (int t = 0; t < 8; t++) {
        if (!(x->p.huff_defined & (1u << t))) continue;
        const uint8_t *bx = x->p.huffvals + t * HUFF_TABLEN, *by = y->p.huffvals + t * HUFF_TABLEN;
        int total = 0;
        for (int i = 0; i < 16; i++) total += bx[i];
        if (total > 256) total = 256;
        if (memcmp(bx, by, (size_t)(16 + total)) != 0) return 0;
    }
    return 1;
}

/* second, independent 64-bit digest of the same bytes (keys the shared-table blob together with jd_tables_hash) */
uint64_t jd_tables_hash2(const JDInfo *info)
{
    uint64_t h = 0x9E3779B97F4A7C15ull;
    const uint8_t *hv = info->p.huffvals;
    for (int t = 0; t < 8; t++) {
        if (!(info->p.huff_defined & (1u << t))) continue;
        const uint8_t *b = hv + t * HUFF_TABLEN;
        int total = 0;
        for (int i = 0; i < 16; i++) total += b[i];
        if (total > 256) total = 256;
        h = (h + (uint64_t)(t + 1)) * 0xD6E8FEB86659FD93ull; h ^= h >> 32;
        for (int i = 0; i < 16 + total; i++) { h = (h + b[i]) * 0xD6E8FEB86659FD93ull; h ^= h >> 29; }
    }
    return h;
}

uint64_t jd_tables_hash(const JDInfo *info)
{
    /* FNV-1a over the defined DHT tables */
    uint64_t h = 1469598103934665603ull;
    const uint8_t *hv = info->p.huffvals;
    for (int t = 0; t < 8; t++) {
        if (!(info->p.huff_defined & (1u << t))) continue;
        const uint8_t *b = hv + t * HUFF_TABLEN;
        int total = 0;
        for (int i = 0; i < 16; i++) total += b[i];
        if (total > 256) total = 256;
        h = (h ^ (uint64_t)(t + 1)) * 1099511628211ull;
        for (int i = 0; i < 16 + total; i++) h = (h ^ b[i]) * 1099511628211ull;
    }
    return h;
}
