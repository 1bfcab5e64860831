/* oracle/jpegdec_oracle.h -- TEST INFRASTRUCTURE (see jpegdec_oracle.c). */
#ifndef JPEGDEC_ORACLE_H
#define JPEGDEC_ORACLE_H
#include <stdint.h>
enum { OR_RGB565_LE = 0, OR_RGB565_BE, OR_RGB8888, OR_GRAY8, OR_DITHER4, OR_DITHER2, OR_DITHER1 };
#define OR_SCALE_HALF 2
#define OR_SCALE_QUARTER 4
#define OR_SCALE_EIGHTH 8
#define OR_LUMA_ONLY 64
/* arith: 0 = the reference's default (SSE2) build, 1 = its -DNO_SIMD build.
 * out: tight ceil(w/s) x ceil(h/s) image, out_pitch bytes per row.  Returns 1 ok, 0 failure. */
int oracle_decode(const uint8_t *jpeg, int len, int pixel_type, int options, int arith,
                  uint8_t *out, int out_pitch, int *out_w, int *out_h);
void oracle_idct(const int16_t *coef, const int16_t *quant, unsigned flags, int arith, int mode, uint8_t *out);
#endif
