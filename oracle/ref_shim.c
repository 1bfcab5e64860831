/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C driver around the UNMODIFIED reference decoder.  The reference is one
 * include file; this TU pulls it in from its read-only location (the Makefile
 * passes -I/root/reference/src) exactly the way the reference's own pure-C
 * caller does (linux/examples/c_cmdline/main.c:10-11), so the JPEG_* C API
 * (src/jpeg.inl:556-738) is compiled with -D__LINUX__.  Nothing from the
 * reference is copied into this repository; only the resulting .so lands in the
 * git-ignored oracle/_ref/ directory.
 *
 * Built twice (see oracle/Makefile):
 *   libjpegdec_ref_sse.so     default flags  -> HAS_SSE   (src/jpeg.inl:49-55)
 *   libjpegdec_ref_scalar.so  -DNO_SIMD      -> 32-bit scalar IDCT / colour
 *
 * Exposed helpers (plain C ABI, driven from Python ctypes in oracle/refdrv.py and
 * from bench.py's reference arm):
 *   ref_decode_fb      openRAM -> setPixelType -> setFramebuffer -> decode -> close
 *   ref_decode_cb      same but through the JPEGDRAW callback; assembles the
 *                      tight ceil(w/s) x ceil(h/s) image using iWidthUsed/iHeight
 *                      and records the callback sequence
 *   ref_decode_dither  JPEG_decodeDither path (1/2/4 bpp)
 *   ref_decode_batch   pthread pool, one JPEGIMAGE per worker, framebuffer mode
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <time.h>

#include "JPEGDEC.h"
#include "jpeg.inl"

int ref_sizeof_image(void) { return (int)sizeof(JPEGIMAGE); }

int ref_is_simd(void)
{
#ifdef HAS_SSE
    return 1;
#else
    return 0;
#endif
}

/* ---- open + query ------------------------------------------------------- */
typedef struct {
    int width, height, subsample, bpp, orientation, has_thumb, thumb_w, thumb_h;
    int mode, res_interval, error;
} RefInfo;

int ref_info(const uint8_t *data, int len, RefInfo *out)
{
    JPEGIMAGE *img = (JPEGIMAGE *)malloc(sizeof(JPEGIMAGE));
    int rc = JPEG_openRAM(img, (uint8_t *)data, len, NULL);
    out->error = JPEG_getLastError(img);
    out->width = JPEG_getWidth(img);
    out->height = JPEG_getHeight(img);
    out->subsample = JPEG_getSubSample(img);
    out->bpp = JPEG_getBpp(img);
    out->orientation = JPEG_getOrientation(img);
    out->has_thumb = JPEG_hasThumb(img);
    out->thumb_w = JPEG_getThumbWidth(img);
    out->thumb_h = JPEG_getThumbHeight(img);
    out->mode = img->ucMode;
    out->res_interval = img->iResInterval;
    if (rc) JPEG_close(img);
    free(img);
    return rc;
}

/* ---- framebuffer decode -------------------------------------------------
 * fb must hold whole MCU rows (the reference writes them all).  crop: pass
 * cw<=0 for "no crop".  Returns decode()'s return value, -1 if open failed. */
int ref_decode_fb(const uint8_t *data, int len, int pixel_type, int options,
                  int cx, int cy, int cw, int ch,
                  uint8_t *fb, int *err)
{
    JPEGIMAGE *img = (JPEGIMAGE *)malloc(sizeof(JPEGIMAGE));
    int rc = JPEG_openRAM(img, (uint8_t *)data, len, NULL);
    if (!rc) { *err = JPEG_getLastError(img); free(img); return -1; }
    JPEG_setPixelType(img, pixel_type);
    if (cw > 0) JPEG_setCropArea(img, cx, cy, cw, ch);
    JPEG_setFramebuffer(img, fb);
    rc = JPEG_decode(img, 0, 0, options);
    *err = JPEG_getLastError(img);
    JPEG_close(img);
    free(img);
    return rc;
}

/* ---- callback decode ----------------------------------------------------- */
typedef struct {
    int x, y, w, h, wused, bpp;
    int buf_toggle; /* 0/1: which half of the internal pixel buffer pPixels pointed at */
} RefDrawRec;

typedef struct {
    uint8_t *out;      /* tight output image */
    int out_pitch;     /* bytes */
    int out_h;
    int xoff, yoff;
    RefDrawRec *log;
    int log_cap, log_n;
    int abort_after;   /* return 0 from the callback after this many calls (<=0: never) */
    const void *first_ptr;
} RefCbCtx;

static int ref_draw_cb(JPEGDRAW *d)
{
    RefCbCtx *c = (RefCbCtx *)d->pUser;
    if (c->log_n == 0) c->first_ptr = d->pPixels;
    if (c->log && c->log_n < c->log_cap) {
        RefDrawRec *r = &c->log[c->log_n];
        r->x = d->x; r->y = d->y; r->w = d->iWidth; r->h = d->iHeight;
        r->wused = d->iWidthUsed; r->bpp = d->iBpp;
        r->buf_toggle = ((const void *)d->pPixels != c->first_ptr);
    }
    c->log_n++;
    if (c->out) {
        int bpp = d->iBpp;
        const uint8_t *src = (const uint8_t *)d->pPixels;
        int src_pitch = (d->iWidth * bpp + 7) / 8;       /* bytes per source line */
        int x0 = d->x - c->xoff, y0 = d->y - c->yoff;
        for (int row = 0; row < d->iHeight; row++) {
            int oy = y0 + row;
            if (oy < 0 || oy >= c->out_h) continue;
            if (bpp >= 8) {
                int bytespp = bpp / 8;
                memcpy(c->out + (size_t)oy * c->out_pitch + (size_t)x0 * bytespp,
                       src + (size_t)row * src_pitch, (size_t)d->iWidthUsed * bytespp);
            } else { /* packed: x0 is always 0 for dithered output (whole row per call) */
                memcpy(c->out + (size_t)oy * c->out_pitch, src + (size_t)row * src_pitch,
                       (size_t)((d->iWidthUsed * bpp + 7) / 8));
            }
        }
    }
    if (c->abort_after > 0 && c->log_n >= c->abort_after) return 0;
    return 1;
}

/* out: tight image, out_pitch bytes per line, out_h lines (caller computes from
 * scale/crop).  log may be NULL.  Returns decode() rc, -1 on open failure. */
int ref_decode_cb(const uint8_t *data, int len, int pixel_type, int options,
                  int xoff, int yoff, int cx, int cy, int cw, int ch,
                  int max_mcus, int abort_after,
                  uint8_t *out, int out_pitch, int out_h,
                  RefDrawRec *log, int log_cap, int *log_n, int *err)
{
    JPEGIMAGE *img = (JPEGIMAGE *)malloc(sizeof(JPEGIMAGE));
    RefCbCtx ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.out = out; ctx.out_pitch = out_pitch; ctx.out_h = out_h;
    ctx.xoff = xoff; ctx.yoff = yoff; ctx.log = log; ctx.log_cap = log_cap;
    ctx.abort_after = abort_after;
    int rc = JPEG_openRAM(img, (uint8_t *)data, len, ref_draw_cb);
    if (!rc) { *err = JPEG_getLastError(img); free(img); if (log_n) *log_n = 0; return -1; }
    img->pUser = &ctx;
    JPEG_setPixelType(img, pixel_type);
    if (max_mcus > 0) JPEG_setMaxOutputSize(img, max_mcus);
    if (cw > 0) JPEG_setCropArea(img, cx, cy, cw, ch);
    rc = JPEG_decode(img, xoff, yoff, options);
    *err = JPEG_getLastError(img);
    if (log_n) *log_n = ctx.log_n;
    JPEG_close(img);
    free(img);
    return rc;
}

/* Dithered decode (JPEG_decodeDither, src/jpeg.inl:663-668). */
int ref_decode_dither(const uint8_t *data, int len, int pixel_type, int options,
                      uint8_t *out, int out_pitch, int out_h,
                      RefDrawRec *log, int log_cap, int *log_n, int *err)
{
    JPEGIMAGE *img = (JPEGIMAGE *)malloc(sizeof(JPEGIMAGE));
    RefCbCtx ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.out = out; ctx.out_pitch = out_pitch; ctx.out_h = out_h;
    ctx.log = log; ctx.log_cap = log_cap;
    int rc = JPEG_openRAM(img, (uint8_t *)data, len, ref_draw_cb);
    if (!rc) { *err = JPEG_getLastError(img); free(img); if (log_n) *log_n = 0; return -1; }
    img->pUser = &ctx;
    JPEG_setPixelType(img, pixel_type);
    /* examples/dithering/dithering.ino:77 allocates width*16 bytes */
    uint8_t *dither = (uint8_t *)calloc((size_t)(JPEG_getWidth(img) + 32) * 16, 1);
    rc = JPEG_decodeDither(img, dither, options);
    *err = JPEG_getLastError(img);
    if (log_n) *log_n = ctx.log_n;
    JPEG_close(img);
    free(dither);
    free(img);
    return rc;
}

/* ---- one 8x8 block through the reference's JPEGIDCT (static, but in this TU) ----
 * coef: 64 int16 natural order; quant: 64 prescaled int16 natural order; flags: u16MCUFlags as JPEGDecodeMCU
 * would have left them; options: 0 or JPEG_SCALE_QUARTER.  out: the 64 bytes JPEGIDCT writes over the slot. */
void ref_idct(const int16_t *coef, const int16_t *quant, int flags, int options, uint8_t *out)
{
    JPEGIMAGE *img = (JPEGIMAGE *)calloc(1, sizeof(JPEGIMAGE));
    img->sMCUs = img->sUnalignedMCUs;
    memcpy(img->sMCUs, coef, 64 * sizeof(int16_t));
    memcpy(img->sQuantTable, quant, 64 * sizeof(int16_t));
    img->u16MCUFlags = (uint16_t)flags;
    img->iOptions = options;
    JPEGIDCT(img, 0, 0);
    memcpy(out, img->sMCUs, 64);
    free(img);
}

/* ---- multi-threaded batch (CPU baseline) --------------------------------- */
typedef struct {
    const uint8_t **datas; const int *lens; uint8_t **fbs;
    int n, pixel_type, options, nthreads, tid;
    int fails;
    int nocb; /* unused */
} RefBatchArg;

static int ref_null_draw(JPEGDRAW *d) { (void)d; return 1; }

static void *ref_batch_worker(void *p)
{
    RefBatchArg *a = (RefBatchArg *)p;
    JPEGIMAGE *img = (JPEGIMAGE *)malloc(sizeof(JPEGIMAGE));
    for (int i = a->tid; i < a->n; i += a->nthreads) {
        int rc = JPEG_openRAM(img, (uint8_t *)a->datas[i], a->lens[i], ref_null_draw);
        if (!rc) { a->fails++; continue; }
        JPEG_setPixelType(img, a->pixel_type);
        if (a->fbs) JPEG_setFramebuffer(img, a->fbs[i]);
        rc = JPEG_decode(img, 0, 0, a->options);
        if (!rc) a->fails++;
        JPEG_close(img);
    }
    free(img);
    return NULL;
}

/* Decodes n images with nthreads workers (image i -> worker i % nthreads).
 * fbs[i] = per-image framebuffer (whole MCU rows) or fbs==NULL for callback mode
 * with a no-op draw callback.  Returns number of failed images; *seconds = wall time. */
int ref_decode_batch(const uint8_t **datas, const int *lens, uint8_t **fbs, int n,
                     int pixel_type, int options, int nthreads, double *seconds)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    RefBatchArg *args = (RefBatchArg *)calloc(nthreads, sizeof(RefBatchArg));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        args[t].datas = datas; args[t].lens = lens; args[t].fbs = fbs; args[t].n = n;
        args[t].pixel_type = pixel_type; args[t].options = options;
        args[t].nthreads = nthreads; args[t].tid = t;
        pthread_create(&th[t], NULL, ref_batch_worker, &args[t]);
    }
    int fails = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); fails += args[t].fails; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    free(th); free(args);
    return fails;
}
