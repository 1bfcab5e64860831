/*
 * jd_internal.h -- structures shared between the host C code (jd_host.c, jd_api.c),
 * the device pipeline (jd_device.cu) and the kernels (jd_kernels.cuh).
 */
#ifndef JD_INTERNAL_H
#define JD_INTERNAL_H

#include <stdint.h>
#include "../../include/JPEGDEC.h"
#include "../../include/jpegdec_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define JD_LUT_ENTRIES_H 10496 /* == JD_LUT_ENTRIES in jd_core.h */

/* Host-side result of parsing one JPEG header (all the per-image facts the GPU needs). */
typedef struct {
    int width, height;
    int subsample;          /* 0x00 gray, 0x11, 0x21, 0x12, 0x22 */
    int ncomp;              /* 1 or 3 */
    int mode;               /* 0xC0 baseline, 0xC2 progressive */
    int bpp;
    int restart_interval;   /* MCUs, 0 = none */
    int scan_offset;        /* byte offset of entropy-coded data */
    int orientation, has_thumb, thumb_w, thumb_h, thumb_data, exif;
    int mcu_w, mcu_h;       /* MCU size in pixels */
    int mcus_x, mcus_y;
    int bpm;                /* blocks per MCU */
    int tsel;               /* per comp c: bit 2c = DC table, bit 2c+1 = AC table */
    int tables_ok;          /* scan uses only tables 0/1 (DC and AC) */
    int error;              /* JPEG_* error code when parse fails */
    int approx;             /* first scan's successive-approximation byte (Ah << 4 | Al); progressive files */
    JDPARSED p;
} JDInfo;

/* jd_host.c */
int jd_parse_header(const uint8_t *data, int size, int start_offset, JDInfo *info);
int jd_check_huffman(const JDInfo *info);                      /* 1 ok, 0 -> JPEG_UNSUPPORTED_FEATURE */
void jd_build_lut(const JDInfo *info, uint16_t *lut /* JD_LUT_ENTRIES_H */);
void jd_build_quant(const JDInfo *info, int16_t *q /* [3][64] natural order, per component */);
uint64_t jd_tables_hash(const JDInfo *info);
uint64_t jd_tables_hash2(const JDInfo *info);
int jd_tables_equal(const JDInfo *x, const JDInfo *y);
const int *jd_aan_table(void);

/* Device-visible per-image descriptor (96 B). */
typedef struct {
    uint32_t scan_off;      /* absolute offset of first entropy byte in the batch buffer */
    uint32_t scan_end;      /* absolute end of this file's bytes */
    uint16_t width, height;
    uint16_t mcus_x, mcus_y;
    uint8_t subsample, ncomp, bpm, tsel;
    uint32_t mcus_per_seg;
    uint32_t nseg;
    uint32_t seg_base;      /* first global segment index */
    uint32_t blk_base;      /* first global block index */
    uint32_t lutset;        /* index of the Huffman LUT set */
    uint32_t out_pitch;     /* bytes */
    uint32_t prog;          /* bit 0: progressive file, only the DC coefficients of its first scan are decoded (reference
                             * JPEGDecodeMCU_P, src/jpeg.inl:1819-2084, 1/8-scale output); bits 8..11: point transform Al */
    uint64_t out_off;       /* byte offset from the output base pointer */
    uint32_t out_w, out_h;  /* output size in pixels after scaling */
    uint32_t status;        /* written by kernels: 0 ok */
    uint32_t err_mcu;
    uint32_t chunk_base;    /* restart-free scans decoded in parallel chunks: first global chunk index ... */
    uint32_t nch;           /* ... and number of chunks (0 = the scan is decoded per restart segment) */
    uint32_t comp_off;      /* offset of this file's first byte in the batch buffer */
    uint32_t pad_;
    uint64_t rec_base;      /* index of this image's first coefficient record: block headers hold record indices relative
                             * to it (jd_core.h JD_REC_INDEX with byte offsets relative to comp_off and image-local slots) */
} JDImageDesc;

#ifdef __cplusplus
}
#endif
#endif
