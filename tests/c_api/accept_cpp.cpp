/* tests/c_api/accept_cpp.cpp -- the C++ class used like linux/examples/jpeg_perf_test/main.cpp does. */
#include <stdio.h>
#include <vector>
#include "JPEGDEC.h"
static JPEGDEC jpeg;
static int ncb;
static int JPEGDraw(JPEGDRAW *) { ncb++; return 1; }
int main(int argc, char **argv)
{
    if (argc < 2) return 1;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 1;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)n); if (fread(buf.data(), 1, (size_t)n, f) != (size_t)n) return 1; fclose(f);
    const int opts[4] = {0, JPEG_SCALE_HALF, JPEG_SCALE_QUARTER, JPEG_SCALE_EIGHTH};
    for (int i = 0; i < 4; i++) {
        ncb = 0;
        if (!jpeg.openFLASH(buf.data(), (int)n, JPEGDraw)) { printf("open failed %d\n", jpeg.getLastError()); return 2; }
        int rc = jpeg.decode(0, 0, opts[i]);
        printf("scale 1/%d rc=%d err=%d %dx%d callbacks=%d\n", 1 << i, rc, jpeg.getLastError(), jpeg.getWidth(), jpeg.getHeight(), ncb);
        jpeg.close();
        if (!rc) return 3;
    }
    return 0;
}
