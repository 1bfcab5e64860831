/*
 * tests/c_api/accept.c -- a pure-C caller written the way the reference's own C example is
 * (linux/examples/c_cmdline/main.c: JPEG_openFile/openRAM -> poke jpg.ucPixelType -> JPEG_setFramebuffer ->
 * JPEG_decode -> JPEG_getLastError -> JPEG_close), compiled against include/JPEGDEC.h and linked to
 * libjpegdec_b200.so.  Proves C linkage / struct-field source compatibility of the drop-in boundary.
 *
 *   accept <file.jpg> <pixel_type> <options> <out.raw>      (framebuffer mode)
 *   accept <file.jpg> cb <options>                          (callback mode: prints the draw sequence)
 * exit code: 0 ok, 2 open failed, 3 decode failed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "JPEGDEC.h"

static JPEGIMAGE jpg;
static int ncb;
static long cb_pixels;

static int draw(JPEGDRAW *d)
{
    ncb++;
    cb_pixels += (long)d->iWidthUsed * d->iHeight;
    if (d->pUser != (void *)&ncb) return 0;
    return 1;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage\n"); return 1; }
    int options = atoi(argv[3]);
    if (strcmp(argv[2], "cb") == 0) {
        if (!JPEG_openFile(&jpg, argv[1], draw)) { printf("open failed err=%d\n", JPEG_getLastError(&jpg)); return 2; }
        jpg.pUser = (void *)&ncb;                 /* users poke the struct directly (c_cmdline/main.c:170-187) */
        jpg.ucPixelType = RGB565_LITTLE_ENDIAN;
        int rc = JPEG_decode(&jpg, 0, 0, options);
        printf("rc=%d err=%d w=%d h=%d callbacks=%d pixels=%ld\n", rc, JPEG_getLastError(&jpg), jpg.iWidth, jpg.iHeight, ncb, cb_pixels);
        JPEG_close(&jpg);
        return rc ? 0 : 3;
    }
    int pt = atoi(argv[2]);
    if (!JPEG_openFile(&jpg, argv[1], NULL)) { printf("open failed err=%d\n", JPEG_getLastError(&jpg)); return 2; }
    int bpp = pt == RGB8888 ? 4 : (pt >= EIGHT_BIT_GRAYSCALE ? 1 : 2);
    size_t bytes = (size_t)jpg.iWidth * (jpg.iHeight + 15) * bpp;
    unsigned char *fb = (unsigned char *)calloc(bytes, 1);
    JPEG_setFramebuffer(&jpg, fb);
    jpg.ucPixelType = (uint8_t)pt;
    int rc = JPEG_decode(&jpg, 0, 0, options);
    printf("rc=%d err=%d w=%d h=%d sub=0x%02x\n", rc, JPEG_getLastError(&jpg), JPEG_getWidth(&jpg), JPEG_getHeight(&jpg), JPEG_getSubSample(&jpg));
    if (rc && argc > 4) {
        FILE *f = fopen(argv[4], "wb");
        fwrite(fb, 1, (size_t)jpg.iWidth * jpg.iHeight * bpp, f);
        fclose(f);
    }
    JPEG_close(&jpg);
    free(fb);
    return rc ? 0 : 3;
}
