/*
 * JPEGDEC.h -- drop-in public API of the B200-native baseline JPEG decoder.
 *
 * Written from scratch.  It keeps the *surface* of bitbank2/JPEGDEC's src/JPEGDEC.h
 * (reference file:line cited per item) so that a program written against the
 * reference -- linux/examples/c_cmdline/main.c, linux/examples/jpeg_perf_test/main.cpp,
 * MacOS/JPEGDEC_Test/JPEGDEC_Test/main.cpp -- recompiles against libjpegdec_b200.so
 * without source changes:
 *
 *   - option bits, pixel types, error codes ........ reference src/JPEGDEC.h:68-75, :102-111, :119-126
 *   - JPEGFILE / JPEGDRAW / callback typedefs ...... reference src/JPEGDEC.h:135-158
 *   - JPEGIMAGE (caller-owned state; the fields user programs poke directly --
 *     iWidth, iHeight, ucPixelType, pUser, pDitherBuffer, iError -- keep their names;
 *     c_cmdline/main.c:170-187 writes jpg.ucPixelType) .... reference src/JPEGDEC.h:199-239
 *   - C entry points JPEG_* ........................ reference src/JPEGDEC.h:290-309, bodies src/jpeg.inl:564-738
 *   - C++ class JPEGDEC ............................ reference src/JPEGDEC.h:249-287, bodies src/JPEGDEC.cpp:64-273
 *
 * Differences a maintainer should know:
 *   - the C symbols are exported with C linkage from a shared library (the
 *     reference only defines them when a C TU #includes jpeg.inl);
 *   - all entropy decode / IDCT / colour conversion / dither work runs in
 *     hand-written sm_100a CUDA kernels; there is NO CPU fallback: decode fails
 *     with JPEG_ERROR_MEMORY if no CUDA device / kernel image is available;
 *   - open() reads the whole file (through the user's read callback) into memory;
 *     JPEG_close() releases it;
 *   - the arithmetic mode (which of the reference's two x86 builds to be
 *     bit-exact with) is selectable: JPEG_setArithMode(), default SSE2-build.
 */
#ifndef __JPEGDEC__
#define __JPEGDEC__

/* the reference header pulls these in for hosted builds and its example programs rely on that (src/JPEGDEC.h:16-30) */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#if !defined(PROGMEM)   /* hosted build: flash-resident arrays are ordinary arrays */
#define PROGMEM
#define memcpy_P(dst, src, n) memcpy((dst), (src), (n))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Decoder options (reference src/JPEGDEC.h:68-75) */
#define JPEG_AUTO_ROTATE 1
#define JPEG_SCALE_HALF 2
#define JPEG_SCALE_QUARTER 4
#define JPEG_SCALE_EIGHTH 8
#define JPEG_LE_PIXELS 16
#define JPEG_EXIF_THUMBNAIL 32
#define JPEG_LUMA_ONLY 64
#define JPEG_USES_DMA 128

#define MAX_BUFFERED_PIXELS 2048 /* reference src/JPEGDEC.h:65: draw-callback batching unit */

/* Supported decode modes (reference src/JPEGDEC.h:95-99) */
enum { JPEG_MODE_BASELINE = 0, JPEG_MODE_PROGRESSIVE, JPEG_MODE_INVALID };

/* Pixel types (reference src/JPEGDEC.h:102-111) */
enum { RGB565_LITTLE_ENDIAN = 0, RGB565_BIG_ENDIAN, RGB8888, EIGHT_BIT_GRAYSCALE, FOUR_BIT_DITHERED, TWO_BIT_DITHERED, ONE_BIT_DITHERED, INVALID_PIXEL_TYPE };

enum { JPEG_MEM_RAM = 0, JPEG_MEM_FLASH };

/* Error codes returned by getLastError() (reference src/JPEGDEC.h:119-126) */
enum { JPEG_SUCCESS = 0, JPEG_INVALID_PARAMETER, JPEG_DECODE_ERROR, JPEG_UNSUPPORTED_FEATURE, JPEG_INVALID_FILE, JPEG_ERROR_MEMORY };

/* Which reference build the pixels are bit-exact with (the reference has two
 * different x86-64 arithmetic paths, src/jpeg.inl:49-55). */
enum {
    JPEG_ARITH_SSE2 = 0,   /* default build: int16 column IDCT, SSE2 colour (B,G,R,A) */
    JPEG_ARITH_SCALAR = 1  /* -DNO_SIMD build: 32-bit IDCT, table colour (R,G,B,A)    */
};

typedef struct jpeg_file_tag {
    int32_t iPos;    /* current file position */
    int32_t iSize;   /* file size */
    uint8_t *pData;  /* memory file pointer */
    void *fHandle;   /* class pointer to File/SdFat or whatever you want */
} JPEGFILE;

typedef struct jpeg_draw_tag {
    int x, y;              /* upper left corner of current MCU group */
    int iWidth, iHeight;   /* size of this pixel block */
    int iWidthUsed;        /* clipped size for odd/edges */
    int iBpp;              /* bit depth of the pixels (1,2,4,8,16,32) */
    uint16_t *pPixels;     /* pixels (valid only during the callback) */
    void *pUser;
} JPEGDRAW;

typedef int32_t (JPEG_READ_CALLBACK)(JPEGFILE *pFile, uint8_t *pBuf, int32_t iLen);
typedef int32_t (JPEG_SEEK_CALLBACK)(JPEGFILE *pFile, int32_t iPosition);
typedef int (JPEG_DRAW_CALLBACK)(JPEGDRAW *pDraw);
typedef void *(JPEG_OPEN_CALLBACK)(const char *szFilename, int32_t *pFileSize);
typedef void (JPEG_CLOSE_CALLBACK)(void *pHandle);

/* Parsed-header state (internal to the library; sized so JPEGIMAGE stays a plain
 * caller-owned struct like the reference's). */
#define JD_HUFFVALS_BYTES 4128 /* mirrors the reference's DHT scratch = usPixels area (src/jpeg.inl:843) */
typedef struct jd_parsed_tag {
    int32_t scan_offset;           /* byte offset of entropy-coded data in the file */
    uint8_t comp_id[4], comp_quant[4], comp_dc[4], comp_ac[4];
    uint8_t huff_defined;          /* bit t: DC table t, bit 4+t: AC table t */
    uint8_t ncomp_in_scan, scan_start, scan_end;
    uint16_t quant_raw[4][64];     /* zigzag order, as read from DQT */
    uint8_t huffvals[JD_HUFFVALS_BYTES]; /* [t*273 + i]: 16 counts then symbols */
} JDPARSED;

/* Decoder state.  Caller-owned; public field names follow reference src/JPEGDEC.h:199-239. */
typedef struct jpeg_image_tag {
    int iWidth, iHeight;               /* image size */
    int iThumbWidth, iThumbHeight;     /* thumbnail size (if present) */
    int iThumbData;                    /* offset to thumbnail JPEG */
    int iXOffset, iYOffset;            /* placement on the display */
    int iCropX, iCropY, iCropCX, iCropCY;
    uint8_t ucBpp, ucSubSample, ucHuffTableUsed;
    uint8_t ucMode, ucOrientation, ucHasThumb, b11Bit;
    uint8_t ucComponentsInScan, ucNumComponents;
    uint8_t ucMemType, ucPixelType, ucArithMode;
    int iEXIF;
    int iError;
    int iOptions;
    int iResInterval;
    int iMaxMCUs;
    JPEG_READ_CALLBACK *pfnRead;
    JPEG_SEEK_CALLBACK *pfnSeek;
    JPEG_DRAW_CALLBACK *pfnDraw;
    JPEG_OPEN_CALLBACK *pfnOpen;
    JPEG_CLOSE_CALLBACK *pfnClose;
    JPEGFILE JPEGFile;
    void *pUser;
    uint8_t *pDitherBuffer;
    void *pFramebuffer;
    /* --- private --- */
    uint8_t *pFileData;     /* whole file in memory (user's buffer for openRAM) */
    int32_t iFileSize;
    int bOwnsFileData;      /* malloc'ed by open (file / callback I/O) */
    int iDevice;            /* CUDA device ordinal, -1 = current */
    JDPARSED parsed;
} JPEGIMAGE;

/* ---- C API (reference src/JPEGDEC.h:290-309) ---- */
int JPEG_openRAM(JPEGIMAGE *img, uint8_t *jpeg_bytes, int jpeg_size, JPEG_DRAW_CALLBACK *draw);
int JPEG_openFile(JPEGIMAGE *img, const char *path, JPEG_DRAW_CALLBACK *draw);
/* generic callback I/O open (what the reference's C++ open(name, cbs...) does, src/JPEGDEC.cpp:155-229) */
int JPEG_openCallbacks(JPEGIMAGE *img, const char *szFilename, void *fHandle, int iDataSize,
                       JPEG_OPEN_CALLBACK *pfnOpen, JPEG_CLOSE_CALLBACK *pfnClose,
                       JPEG_READ_CALLBACK *pfnRead, JPEG_SEEK_CALLBACK *pfnSeek,
                       JPEG_DRAW_CALLBACK *pfnDraw);
void JPEG_setFramebuffer(JPEGIMAGE *img, void *framebuffer);
void JPEG_setCropArea(JPEGIMAGE *img, int crop_x, int crop_y, int crop_w, int crop_h);
void JPEG_getCropArea(JPEGIMAGE *img, int *crop_x, int *crop_y, int *crop_w, int *crop_h);
int JPEG_getWidth(JPEGIMAGE *img);
int JPEG_getHeight(JPEGIMAGE *img);
int JPEG_decode(JPEGIMAGE *img, int x_offset, int y_offset, int options);
int JPEG_decodeDither(JPEGIMAGE *img, uint8_t *dither_rows, int options);
void JPEG_close(JPEGIMAGE *img);
int JPEG_getLastError(JPEGIMAGE *img);
int JPEG_getOrientation(JPEGIMAGE *img);
int JPEG_getBpp(JPEGIMAGE *img);
int JPEG_getSubSample(JPEGIMAGE *img);
int JPEG_getJPEGType(JPEGIMAGE *img);
int JPEG_hasThumb(JPEGIMAGE *img);
int JPEG_getThumbWidth(JPEGIMAGE *img);
int JPEG_getThumbHeight(JPEGIMAGE *img);
void JPEG_setPixelType(JPEGIMAGE *img, int pixel_type);
int JPEG_getPixelType(JPEGIMAGE *img);
void JPEG_setMaxOutputSize(JPEGIMAGE *img, int max_mcus);
void JPEG_setUserPointer(JPEGIMAGE *img, void *p);
/* extensions */
void JPEG_setArithMode(JPEGIMAGE *img, int iMode);   /* JPEG_ARITH_SSE2 | JPEG_ARITH_SCALAR */
void JPEG_setDevice(JPEGIMAGE *img, int iDevice);    /* CUDA ordinal for this handle */
int JPEG_sizeofImage(void);                            /* sizeof(JPEGIMAGE) for FFI callers */

#ifdef __cplusplus
} /* extern "C" */

/*
 * C++ class wrapper (reference src/JPEGDEC.h:249-287 / src/JPEGDEC.cpp): the
 * same thin forwarding layer, header-only.
 */
class JPEGDEC
{
  public:
    int openRAM(uint8_t *pData, int iDataSize, JPEG_DRAW_CALLBACK *pfnDraw)
    { return JPEG_openRAM(&_jpeg, pData, iDataSize, pfnDraw); }
    int openFLASH(const uint8_t *pData, int iDataSize, JPEG_DRAW_CALLBACK *pfnDraw)
    { int rc = JPEG_openRAM(&_jpeg, (uint8_t *)pData, iDataSize, pfnDraw); _jpeg.ucMemType = JPEG_MEM_FLASH; return rc; }
    int open(const char *szFilename, JPEG_OPEN_CALLBACK *pfnOpen, JPEG_CLOSE_CALLBACK *pfnClose,
             JPEG_READ_CALLBACK *pfnRead, JPEG_SEEK_CALLBACK *pfnSeek, JPEG_DRAW_CALLBACK *pfnDraw)
    { return JPEG_openCallbacks(&_jpeg, szFilename, NULL, 0, pfnOpen, pfnClose, pfnRead, pfnSeek, pfnDraw); }
    int open(const char *szFilename, JPEG_DRAW_CALLBACK *pfnDraw)
    { return JPEG_openFile(&_jpeg, szFilename, pfnDraw); }
    int open(void *fHandle, int iDataSize, JPEG_CLOSE_CALLBACK *pfnClose, JPEG_READ_CALLBACK *pfnRead,
             JPEG_SEEK_CALLBACK *pfnSeek, JPEG_DRAW_CALLBACK *pfnDraw)
    { return JPEG_openCallbacks(&_jpeg, NULL, fHandle, iDataSize, NULL, pfnClose, pfnRead, pfnSeek, pfnDraw); }
    void setFramebuffer(void *pFramebuffer) { JPEG_setFramebuffer(&_jpeg, pFramebuffer); }
    void setCropArea(int x, int y, int w, int h) { JPEG_setCropArea(&_jpeg, x, y, w, h); }
    void getCropArea(int *x, int *y, int *w, int *h) { JPEG_getCropArea(&_jpeg, x, y, w, h); }
    void close() { JPEG_close(&_jpeg); }
    int decode(int x, int y, int iOptions) { return JPEG_decode(&_jpeg, x, y, iOptions); }
    int decodeDither(uint8_t *pDither, int iOptions)
    { _jpeg.iXOffset = 0; _jpeg.iYOffset = 0; return JPEG_decodeDither(&_jpeg, pDither, iOptions); }
    int decodeDither(int x, int y, uint8_t *pDither, int iOptions)
    { _jpeg.iXOffset = x; _jpeg.iYOffset = y; return JPEG_decodeDither(&_jpeg, pDither, iOptions); }
    int getOrientation() { return JPEG_getOrientation(&_jpeg); }
    int getWidth() { return JPEG_getWidth(&_jpeg); }
    int getHeight() { return JPEG_getHeight(&_jpeg); }
    int getBpp() { return JPEG_getBpp(&_jpeg); }
    void setUserPointer(void *p) { JPEG_setUserPointer(&_jpeg, p); }
    int getSubSample() { return JPEG_getSubSample(&_jpeg); }
    int getJPEGType() { return JPEG_getJPEGType(&_jpeg); }
    int hasThumb() { return JPEG_hasThumb(&_jpeg); }
    int getThumbWidth() { return JPEG_getThumbWidth(&_jpeg); }
    int getThumbHeight() { return JPEG_getThumbHeight(&_jpeg); }
    int getLastError() { return JPEG_getLastError(&_jpeg); }
    void setPixelType(int iType) { JPEG_setPixelType(&_jpeg, iType); }
    int getPixelType() { return JPEG_getPixelType(&_jpeg); }
    void setMaxOutputSize(int iMaxMCUs) { JPEG_setMaxOutputSize(&_jpeg, iMaxMCUs); }
    void setArithMode(int iMode) { JPEG_setArithMode(&_jpeg, iMode); }
  private:
    JPEGIMAGE _jpeg;
};
#endif /* __cplusplus */

#endif /* __JPEGDEC__ */
