/*
 * jd_api.c -- the JPEGDEC C API (include/JPEGDEC.h) on top of the batch pipeline.
 *
 * Boundary being replaced (reference file:line):
 *   JPEG_openRAM / JPEG_openFile / getters / setters ... src/jpeg.inl:564-738
 *   JPEG_decode / JPEG_decodeDither -> DecodeJPEG ....... src/jpeg.inl:655-668, :4946-5357
 *
 * decode() = one-image batch on the GPU (whole MCU-aligned frame into a pinned staging
 * buffer) followed by a host replay of the reference's *delivery* rules: MCU skipping for
 * crop, draw-callback batching (iMCUCount), iWidth / iWidthUsed / iHeight trimming, the DMA
 * ping-pong pointer, framebuffer pitch = crop width.  Pixels are never computed on the host.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "jd_internal.h"

#define JPEGB200_OPT_PADDED 0x10000 /* internal option bit: write the whole MCU-aligned frame */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static JPEGB200_CTX *g_ctx[64][2];
static uint8_t *g_stage;
static size_t g_stage_bytes;

static JPEGB200_CTX *get_ctx(int device, int arith)
{
    int d = device;
    if (d < 0) d = 0; /* slot for "current device" */
    if (d >= 64) return NULL;
    if (!g_ctx[d][arith ? 1 : 0]) g_ctx[d][arith ? 1 : 0] = JPEGB200_create(device, arith);
    return g_ctx[d][arith ? 1 : 0];
}

static void fill_from_info(JPEGIMAGE *p, const JDInfo *inf)
{
    p->iWidth = p->iCropCX = inf->width;
    p->iHeight = p->iCropCY = inf->height;
    p->iCropX = p->iCropY = 0;
    p->ucBpp = (uint8_t)inf->bpp;
    p->ucSubSample = (uint8_t)inf->subsample;
    p->ucMode = (uint8_t)inf->mode;
    p->ucNumComponents = (uint8_t)inf->ncomp;
    p->ucComponentsInScan = inf->p.ncomp_in_scan;
    p->ucHuffTableUsed = inf->p.huff_defined;
    p->iResInterval = inf->restart_interval;
}

static int init_common(JPEGIMAGE *p)
{
    JDInfo *inf = (JDInfo *)malloc(sizeof(JDInfo));
    if (!inf) { p->iError = JPEG_ERROR_MEMORY; return 0; }
    int rc = jd_parse_header(p->pFileData, p->iFileSize, 0, inf);
    /* fields the reference fills while it walks the markers, even when open fails later */
    fill_from_info(p, inf);
    p->ucOrientation = (uint8_t)inf->orientation;
    p->ucHasThumb = (uint8_t)inf->has_thumb;
    p->iThumbWidth = inf->thumb_w;
    p->iThumbHeight = inf->thumb_h;
    p->iThumbData = inf->thumb_data;
    p->iEXIF = inf->exif;
    p->iError = rc ? JPEG_SUCCESS : inf->error;
    p->parsed.scan_offset = inf->scan_offset;
    free(inf);
    p->pFramebuffer = NULL; /* must be set after open (jpeg.inl:1580) */
    return rc;
}

int JPEG_openRAM(JPEGIMAGE *p, uint8_t *pData, int iDataSize, JPEG_DRAW_CALLBACK *pfnDraw)
{
    memset(p, 0, sizeof(JPEGIMAGE));
    p->ucMemType = JPEG_MEM_RAM;
    p->pfnDraw = pfnDraw;
    p->JPEGFile.iSize = iDataSize;
    p->JPEGFile.pData = pData;
    p->iMaxMCUs = 1000;
    p->iDevice = -1;
    p->pFileData = pData;
    p->iFileSize = iDataSize;
    if (!pData || iDataSize <= 0) { p->iError = JPEG_INVALID_FILE; return 0; }
    return init_common(p);
}

/* whole file through the user's (or stdio) callbacks */
static int slurp(JPEGIMAGE *p)
{
    int size = p->JPEGFile.iSize;
    if (size <= 0) { p->iError = JPEG_INVALID_FILE; return 0; }
    uint8_t *buf = (uint8_t *)malloc((size_t)size + 16);
    if (!buf) { p->iError = JPEG_ERROR_MEMORY; return 0; }
    int got = 0;
    if (p->pfnSeek) (*p->pfnSeek)(&p->JPEGFile, 0);
    while (got < size) {
        int want = size - got; if (want > 65536) want = 65536;
        int r = (*p->pfnRead)(&p->JPEGFile, buf + got, want);
        if (r <= 0) break;
        got += r;
    }
    memset(buf + got, 0, (size_t)(size - got) + 16);
    p->pFileData = buf;
    p->iFileSize = got;
    p->bOwnsFileData = 1;
    return 1;
}

static int32_t file_read(JPEGFILE *f, uint8_t *buf, int32_t len)
{
    if (len > f->iSize - f->iPos) len = f->iSize - f->iPos;
    if (len <= 0) return 0;
    int32_t r = (int32_t)fread(buf, 1, (size_t)len, (FILE *)f->fHandle);
    f->iPos += r;
    return r;
}
static int32_t file_seek(JPEGFILE *f, int32_t pos)
{
    if (pos < 0) pos = 0; else if (pos >= f->iSize) pos = f->iSize - 1;
    f->iPos = pos;
    fseek((FILE *)f->fHandle, pos, SEEK_SET);
    return pos;
}
static void file_close(void *h) { if (h) fclose((FILE *)h); }

int JPEG_openFile(JPEGIMAGE *p, const char *szFilename, JPEG_DRAW_CALLBACK *pfnDraw)
{
    memset(p, 0, sizeof(JPEGIMAGE));
    p->ucMemType = JPEG_MEM_RAM;
    p->pfnRead = file_read;
    p->pfnSeek = file_seek;
    p->pfnDraw = pfnDraw;
    p->pfnClose = file_close;
    p->iMaxMCUs = 1000;
    p->iDevice = -1;
    FILE *f = fopen(szFilename, "rb");
    if (!f) return 0;
    p->JPEGFile.fHandle = f;
    fseek(f, 0, SEEK_END);
    p->JPEGFile.iSize = (int32_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    if (!slurp(p)) return 0;
    return init_common(p);
}

int JPEG_openCallbacks(JPEGIMAGE *p, const char *szFilename, void *fHandle, int iDataSize, JPEG_OPEN_CALLBACK *pfnOpen,
                       JPEG_CLOSE_CALLBACK *pfnClose, JPEG_READ_CALLBACK *pfnRead, JPEG_SEEK_CALLBACK *pfnSeek,
                       JPEG_DRAW_CALLBACK *pfnDraw)
{
    memset(p, 0, sizeof(JPEGIMAGE));
    p->pfnRead = pfnRead; p->pfnSeek = pfnSeek; p->pfnDraw = pfnDraw; p->pfnOpen = pfnOpen; p->pfnClose = pfnClose;
    p->iMaxMCUs = 1000;
    p->iDevice = -1;
    if (!pfnRead) { p->iError = JPEG_INVALID_PARAMETER; return 0; }
    if (pfnOpen) {
        int32_t sz = 0;
        p->JPEGFile.fHandle = (*pfnOpen)(szFilename, &sz);
        p->JPEGFile.iSize = sz;
        if (p->JPEGFile.fHandle == NULL) return 0;
    } else {
        p->JPEGFile.fHandle = fHandle;
        p->JPEGFile.iSize = iDataSize;
    }
    if (!slurp(p)) return 0;
    return init_common(p);
}

void JPEG_close(JPEGIMAGE *p)
{
    if (p->pfnClose) (*p->pfnClose)(p->JPEGFile.fHandle);
    if (p->bOwnsFileData && p->pFileData) free(p->pFileData);
    p->pFileData = NULL;
    p->bOwnsFileData = 0;
    p->pfnClose = NULL;
}

int JPEG_getLastError(JPEGIMAGE *p) { return p->iError; }
int JPEG_getWidth(JPEGIMAGE *p) { return p->iWidth; }
int JPEG_getHeight(JPEGIMAGE *p) { return p->iHeight; }
int JPEG_getOrientation(JPEGIMAGE *p) { return (int)p->ucOrientation; }
int JPEG_getBpp(JPEGIMAGE *p) { return (int)p->ucBpp; }
int JPEG_getSubSample(JPEGIMAGE *p) { return (int)p->ucSubSample; }
int JPEG_getJPEGType(JPEGIMAGE *p) { return (p->ucMode == 0xc2) ? JPEG_MODE_PROGRESSIVE : JPEG_MODE_BASELINE; }
int JPEG_hasThumb(JPEGIMAGE *p) { return (int)p->ucHasThumb; }
int JPEG_getThumbWidth(JPEGIMAGE *p) { return p->iThumbWidth; }
int JPEG_getThumbHeight(JPEGIMAGE *p) { return p->iThumbHeight; }
void JPEG_setPixelType(JPEGIMAGE *p, int iType) { p->ucPixelType = (uint8_t)iType; }
int JPEG_getPixelType(JPEGIMAGE *p) { return (int)p->ucPixelType; }
void JPEG_setMaxOutputSize(JPEGIMAGE *p, int iMaxMCUs) { if (iMaxMCUs < 1) iMaxMCUs = 1; p->iMaxMCUs = iMaxMCUs; }
void JPEG_setUserPointer(JPEGIMAGE *p, void *u) { p->pUser = u; }
void JPEG_setFramebuffer(JPEGIMAGE *p, void *fb) { p->pFramebuffer = fb; }
void JPEG_setArithMode(JPEGIMAGE *p, int iMode) { p->ucArithMode = (uint8_t)(iMode ? JPEG_ARITH_SCALAR : JPEG_ARITH_SSE2); }
void JPEG_setDevice(JPEGIMAGE *p, int iDevice) { p->iDevice = iDevice; }
int JPEG_sizeofImage(void) { return (int)sizeof(JPEGIMAGE); }

static void mcu_size(int subsample, int *w, int *h)
{
    switch (subsample) {
        case 0x12: *w = 8; *h = 16; break;
        case 0x21: *w = 16; *h = 8; break;
        case 0x22: *w = 16; *h = 16; break;
        default: *w = 8; *h = 8; break;
    }
}

/* snap the crop to MCU boundaries, grow to cover the request (semantics of jpeg.inl:682-727) */
void JPEG_setCropArea(JPEGIMAGE *p, int x, int y, int w, int h)
{
    int mw, mh;
    mcu_size(p->ucSubSample, &mw, &mh);
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    w = (w + mw - 1) / mw * mw;
    h = (h + mh - 1) / mh * mh;
    if (x > p->iWidth - mw) x = p->iWidth - mw;
    if (y > p->iHeight - mh) y = p->iHeight - mh;
    if (x + w > p->iWidth) w = p->iWidth - mw;
    if (y + h > p->iHeight) h = p->iHeight - mh;
    x &= ~(mw - 1);
    y &= ~(mh - 1);
    p->iCropX = x; p->iCropY = y; p->iCropCX = w; p->iCropCY = h;
}

void JPEG_getCropArea(JPEGIMAGE *p, int *x, int *y, int *w, int *h)
{
    *x = p->iCropX; *y = p->iCropY; *w = p->iCropCX; *h = p->iCropCY;
}

/* ---- decode ---- */
static int bits_per_pixel(int pt)
{
    switch (pt) {
        case RGB8888: return 32;
        case EIGHT_BIT_GRAYSCALE: return 8;
        case FOUR_BIT_DITHERED: return 4;
        case TWO_BIT_DITHERED: return 2;
        case ONE_BIT_DITHERED: return 1;
        default: return 16;
    }
}

/* copies one scaled MCU (mw x mh pixels of `bypp` bytes) out of the decoded frame */
static void copy_mcu(uint8_t *dst, int dst_pitch_bytes, const uint8_t *frame, int frame_pitch, int mx, int my, int mw, int mh,
                     int bypp, int max_cols)
{
    const uint8_t *src = frame + (size_t)my * mh * frame_pitch + (size_t)mx * mw * bypp;
    int cols = mw;
    if (max_cols < cols) cols = max_cols;
    if (cols <= 0) return;
    for (int r = 0; r < mh; r++) memcpy(dst + (size_t)r * dst_pitch_bytes, src + (size_t)r * frame_pitch, (size_t)cols * bypp);
}

static int decode_common(JPEGIMAGE *p)
{
    int options = p->iOptions;
    if (!p->pFileData) { p->iError = JPEG_INVALID_PARAMETER; return 0; }
    if (p->ucMode == 0xc2) options = (p->iOptions |= JPEG_SCALE_EIGHTH); /* progressive: DC-only 1/8 image (jpeg.inl:4964-4966) */
    if (options & JPEG_EXIF_THUMBNAIL) {
        if (p->iThumbData == 0 || p->iThumbWidth == 0) { p->iError = JPEG_INVALID_PARAMETER; return 0; } /* jpeg.inl:4969 */
    }
    int shift = (options & JPEG_SCALE_HALF) ? 1 : (options & JPEG_SCALE_QUARTER) ? 2 : (options & JPEG_SCALE_EIGHTH) ? 3 : 0;
    if ((options & JPEG_LUMA_ONLY) && p->ucPixelType < EIGHT_BIT_GRAYSCALE) p->ucPixelType = EIGHT_BIT_GRAYSCALE; /* :4991 */
    if (p->ucPixelType >= INVALID_PIXEL_TYPE) { p->iError = JPEG_INVALID_PARAMETER; return 0; }
    const int pt = p->ucPixelType;
    const int dither = pt > EIGHT_BIT_GRAYSCALE;
    if (dither && !p->pDitherBuffer) { p->iError = JPEG_INVALID_PARAMETER; return 0; }
    if (!p->pFramebuffer && !p->pfnDraw) { p->iError = JPEG_INVALID_PARAMETER; return 0; }

    pthread_mutex_lock(&g_lock);
    JPEGB200_CTX *ctx = get_ctx(p->iDevice, p->ucArithMode);
    if (!ctx) { pthread_mutex_unlock(&g_lock); p->iError = JPEG_ERROR_MEMORY; return 0; }
    const uint8_t *datas[1] = {p->pFileData};
    int32_t sizes[1] = {p->iFileSize};
    JPEGB200_BATCH *b = JPEGB200_batchCreate(ctx, datas, sizes, 1, pt, (options & 0xFF) | JPEGB200_OPT_PADDED);
    if (!b) { pthread_mutex_unlock(&g_lock); p->iError = JPEG_ERROR_MEMORY; return 0; }
    int32_t w = 0, h = 0, sub = 0, fw = 0, fh = 0, st = 0;
    JPEGB200_batchImageInfo(b, 0, &w, &h, &sub, &fw, &fh, &st);
    if (st != JPEG_SUCCESS) { JPEGB200_batchDestroy(b); pthread_mutex_unlock(&g_lock); p->iError = st; return 0; }
    if (options & JPEG_EXIF_THUMBNAIL) { /* the reference re-parses into the same state (:4975) */
        p->iWidth = p->iCropCX = w; p->iHeight = p->iCropCY = h; p->iCropX = p->iCropY = 0; p->ucSubSample = (uint8_t)sub;
    }
    int64_t fpitch = 0;
    int64_t fbytes = JPEGB200_batchOutputBytes(b, 0, &fpitch);
    if ((size_t)fbytes + 64 > g_stage_bytes) {
        if (g_stage) JPEGB200_hostFree(g_stage);
        g_stage_bytes = (size_t)fbytes + 64 + (1u << 20);
        g_stage = (uint8_t *)JPEGB200_hostAlloc(g_stage_bytes);
        if (!g_stage) { g_stage_bytes = 0; JPEGB200_batchDestroy(b); pthread_mutex_unlock(&g_lock); p->iError = JPEG_ERROR_MEMORY; return 0; }
    }
    JPEGB200_batchSetOutput(b, 0, g_stage, 0);
    int32_t dst_status = 0;
    int ok = JPEGB200_batchUpload(b) && JPEGB200_batchDecode(b, 0) && JPEGB200_batchDownload(b);
    int wrc = ok ? JPEGB200_batchWait(b, &dst_status) : 0;
    extern int JPEGB200_batchErrMcu(JPEGB200_BATCH * b, int i);
    int err_mcu = wrc ? JPEGB200_batchErrMcu(b, 0) : -1;
    JPEGB200_batchDestroy(b);
    if (!wrc) { pthread_mutex_unlock(&g_lock); p->iError = JPEG_ERROR_MEMORY; return 0; }
    const int decode_failed = (dst_status != JPEG_SUCCESS);

    /* ---- delivery: geometry of DecodeJPEG (jpeg.inl:5008-5127, :5300-5336) ---- */
    int mw, mh;
    mcu_size(p->ucSubSample, &mw, &mh);
    const int cx = (p->iWidth + mw - 1) / mw;
    const int cy = (p->iCropY + p->iCropCY + mh - 1) / mh;
    mw >>= shift; mh >>= shift;
    const int bpp = bits_per_pixel(pt);
    const int bypp = bpp >= 8 ? bpp / 8 : 1;
    const uint8_t *frame = g_stage;
    const int frame_pitch = (int)fpitch;
    int per_cb = MAX_BUFFERED_PIXELS / (mw * mh);
    if (pt == RGB8888) per_cb /= 2;
    int dma_size = 0, dma_off = 0;
    if (pt == EIGHT_BIT_GRAYSCALE) per_cb *= 2;
    if (per_cb > cx) per_cb = cx;
    if (per_cb > p->iMaxMCUs) per_cb = p->iMaxMCUs;
    else if (options & JPEG_USES_DMA) { per_cb /= 2; dma_size = MAX_BUFFERED_PIXELS / 2; }
    if (dither) per_cb = cx;
    if (p->iCropCX != p->iWidth && per_cb * mw > p->iCropCX) per_cb = p->iCropCX / mw;
    if (per_cb < 1) per_cb = 1;
    const int adj = (1 << shift) - 1;
    const int cur_w = (p->iWidth + adj) >> shift, cur_h = (p->iHeight + adj) >> shift;
    /* pixel staging the callbacks see: same size as the reference's usPixels (2048 px + slack), 16-byte aligned */
    uint16_t pixbuf_raw[MAX_BUFFERED_PIXELS + 64];
    uint16_t *pixbuf = (uint16_t *)(((uintptr_t)pixbuf_raw + 15) & ~(uintptr_t)15);
    JPEGDRAW jd;
    memset(&jd, 0, sizeof(jd));
    jd.iBpp = bpp;
    jd.iHeight = mh;
    int keep_going = 1;
    const int sse_mcu_writes = p->pFramebuffer && p->ucArithMode == JPEG_ARITH_SSE2 && shift == 0 && p->ucNumComponents == 3 &&
                               pt <= RGB8888 && (p->ucSubSample == 0x11 || p->ucSubSample == 0x22);
    const int dpitch = dither ? (cx * mw * bpp + 7) / 8 : 0;
    for (int y = 0; y < cy && keep_going; y++) {
        const int skip_row = (y * mh < p->iCropY);
        int pitch_px, xoff = 0;
        uint8_t *rowbase = NULL;
        jd.x = p->iXOffset;
        if (p->pFramebuffer) {
            pitch_px = p->iCropCX;
            const int ty = y * mh - p->iCropY;
            rowbase = (uint8_t *)p->pFramebuffer + (ptrdiff_t)ty * pitch_px * bypp;
        } else {
            pitch_px = per_cb * mw;
        }
        for (int x = 0; x < cx && keep_going; x++) {
            const int skip = skip_row || x * mw < p->iCropX || x * mw > p->iCropX + p->iCropCX;
            const int mcu_index = y * cx + x;
            if (decode_failed && err_mcu >= 0 && mcu_index > err_mcu) { keep_going = 0; break; } /* loops stop after the failing MCU (:5128) */
            if (!skip) {
                if (dither) {
                    /* packed rows come from the dither kernel; nothing to gather per MCU */
                } else if (p->pFramebuffer) {
                    int rows = mh;
                    const uint8_t *src = frame + (size_t)y * mh * frame_pitch + (size_t)x * mw * bypp;
                    if (sse_mcu_writes) {
                        /* the reference's SSE2 colour paths store whole MCUs with no edge clipping (jpeg.inl:3409, :4006):
                         * the right-edge MCU runs on into the start of the next line (and is partly overwritten later, in
                         * MCU order), the bottom MCU row continues below the image -- which is why the caller's buffer
                         * must cover whole MCU rows (c_cmdline/main.c:180).  Same writes, same order; clipped only at the
                         * end of that MCU-row-aligned buffer. */
                        const size_t fb_px = (size_t)pitch_px * (size_t)(cy * mh - p->iCropY);
                        for (int r = 0; r < rows; r++) {
                            const size_t at = (size_t)(y * mh - p->iCropY + r) * pitch_px + xoff;
                            size_t n = (size_t)mw;
                            if (at >= fb_px) break;
                            if (at + n > fb_px) n = fb_px - at;
                            memcpy((uint8_t *)p->pFramebuffer + at * bypp, src + (size_t)r * frame_pitch, n * bypp);
                        }
                    } else {
                        if (y * mh + rows > cur_h) rows = cur_h - y * mh; /* scalar paths clip at the image edges (:3520-3524, :4311-4332) */
                        if (rows > 0) {
                            int cols = pitch_px - xoff; if (cols > mw) cols = mw;
                            for (int r = 0; r < rows && cols > 0; r++)
                                memcpy(rowbase + ((size_t)r * pitch_px + xoff) * bypp, src + (size_t)r * frame_pitch, (size_t)cols * bypp);
                        }
                    }
                } else {
                    uint8_t *dst = (uint8_t *)(pixbuf + dma_off) + (size_t)xoff * bypp;
                    copy_mcu(dst, pitch_px * bypp, frame, frame_pitch, x, y, mw, mh, bypp, pitch_px > 0 ? mw : 0);
                }
                xoff += mw;
            }
            if (!p->pFramebuffer && (xoff == pitch_px || x == cx - 1) && !skip) {
                jd.iWidth = jd.iWidthUsed = pitch_px;
                jd.pUser = p->pUser;
                if ((jd.x - p->iXOffset) + pitch_px > cur_w) jd.iWidthUsed = cur_w - (jd.x - p->iXOffset);
                else if ((jd.x - p->iXOffset) + pitch_px > p->iCropCX) jd.iWidthUsed = p->iCropCX - (jd.x - p->iXOffset);
                jd.y = p->iYOffset + y * mh - p->iCropY;
                if ((jd.y - p->iYOffset + mh) > cur_h) jd.iHeight = cur_h - (jd.y - p->iYOffset);
                if (dither) {
                    memcpy(p->pDitherBuffer, frame + (size_t)y * mh * frame_pitch, (size_t)dpitch * mh);
                    jd.pPixels = (uint16_t *)p->pDitherBuffer;
                } else jd.pPixels = pixbuf + dma_off;
                keep_going = (*p->pfnDraw)(&jd);
                dma_off ^= dma_size;
                jd.x += pitch_px;
                if (p->iCropCX != cx * mw && (pitch_px + jd.x) > (p->iCropX + p->iCropCX)) pitch_px = p->iCropCX - (jd.x - p->iXOffset);
                else if ((cx - 1 - x) < per_cb) pitch_px = (cx - 1 - x) * mw;
                xoff = 0;
                if (pitch_px & (mw - 1)) pitch_px = (pitch_px + (mw - 1)) & ~(mw - 1);
                if (pitch_px < 0) pitch_px = 0;
            }
        }
    }
    pthread_mutex_unlock(&g_lock);
    if (decode_failed) { p->iError = JPEG_DECODE_ERROR; return 0; }
    return 1;
}

int JPEG_decode(JPEGIMAGE *p, int x, int y, int iOptions)
{
    p->iXOffset = x;
    p->iYOffset = y;
    p->iOptions = iOptions;
    return decode_common(p);
}

int JPEG_decodeDither(JPEGIMAGE *p, uint8_t *pDither, int iOptions)
{
    p->iOptions = iOptions;
    p->pDitherBuffer = pDither;
    return decode_common(p);
}
