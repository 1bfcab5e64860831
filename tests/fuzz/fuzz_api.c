/* Host-side fuzz of the drop-in C API without a GPU: JPEG_openRAM, getters, crop snapping, setters, JPEG_close on mutated
 * files, built with ASan + UBSan (GPU entry points are left unresolved; this path never reaches them). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "JPEGDEC.h"
static uint8_t *rd(const char *p, int *n){FILE*f=fopen(p,"rb");fseek(f,0,SEEK_END);*n=ftell(f);fseek(f,0,SEEK_SET);uint8_t*b=malloc(*n);if(fread(b,1,*n,f)!=(size_t)*n)exit(2);fclose(f);return b;}
static int draw(JPEGDRAW *d){(void)d;return 1;}
int main(int argc,char**argv){
  unsigned seed=99; long ok=0,tot=0;
  const int iters = atoi(argv[1]);
  for(int a=2;a<argc;a++){
    int n; uint8_t*src=rd(argv[a],&n);
    for(int it=0;it<iters;it++){
      int m=n; if(it%5==0) m = 1 + rand_r(&seed)%n;
      uint8_t*buf=malloc(m); memcpy(buf,src,m);
      int k=rand_r(&seed)%4;
      for(int j=0;j<k;j++){ int lim = m<2200?m:2200; buf[rand_r(&seed)%lim]=(uint8_t)rand_r(&seed);}
      JPEGIMAGE img; memset(&img,0x5a,sizeof(img));
      tot++;
      if(JPEG_openRAM(&img,buf,m,draw)){ ok++;
        (void)JPEG_getWidth(&img); (void)JPEG_getHeight(&img); (void)JPEG_getBpp(&img); (void)JPEG_getSubSample(&img);
        (void)JPEG_hasThumb(&img); (void)JPEG_getThumbWidth(&img); (void)JPEG_getOrientation(&img); (void)JPEG_getJPEGType(&img);
        JPEG_setCropArea(&img,(int)(rand_r(&seed)%700)-50,(int)(rand_r(&seed)%700)-50,(int)(rand_r(&seed)%900)-50,(int)(rand_r(&seed)%900)-50);
        int x,y,w,h; JPEG_getCropArea(&img,&x,&y,&w,&h);
        JPEG_setPixelType(&img,rand_r(&seed)%9); JPEG_setMaxOutputSize(&img,(int)(rand_r(&seed)%40)-2);
        JPEG_close(&img);
      } else (void)JPEG_getLastError(&img);
      free(buf);
    }
    free(src);
  }
  printf("cases %ld opened %ld\n",tot,ok); return 0;}
