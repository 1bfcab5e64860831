/* Host-side fuzz (the reference's tests 11-12 corrupt the header area too, MacOS/JPEGDEC_Test/JPEGDEC_Test/main.cpp:262-300):
 * mutated / truncated files through jd_parse_header and the table builders, built with ASan + UBSan, exact-size heap blocks. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "JPEGDEC.h"
#include "jd_internal.h"
static uint8_t *rd(const char *p, int *n){FILE*f=fopen(p,"rb");fseek(f,0,SEEK_END);*n=ftell(f);fseek(f,0,SEEK_SET);uint8_t*b=malloc(*n);if(fread(b,1,*n,f)!=(size_t)*n)exit(2);fclose(f);return b;}
int main(int argc,char**argv){
  unsigned seed=12345; long ok=0,tot=0;
  const int iters = atoi(argv[1]);
  for(int a=2;a<argc;a++){
    int n; uint8_t*src=rd(argv[a],&n);
    for(int it=0;it<iters;it++){
      int m=n; if(it%7==0) m = 1 + rand_r(&seed)%n;             /* truncation */
      uint8_t*buf=malloc(m); memcpy(buf,src,m);                 /* exact-size heap block: ASan sees any over-read */
      int k=1+rand_r(&seed)%4;
      for(int j=0;j<k;j++){ int lim = m<2200?m:2200; int off=rand_r(&seed)%lim; buf[off]=(uint8_t)rand_r(&seed);}  /* header area */
      JDInfo info; memset(&info,0,sizeof(info));
      int r=jd_parse_header(buf,m,0,&info); tot++;
      if(r){ ok++; uint16_t *lut=malloc(JD_LUT_ENTRIES_H*2); jd_build_lut(&info,lut); int16_t q[192]; jd_build_quant(&info,q); (void)jd_tables_hash(&info); free(lut);
             if(info.has_thumb && info.thumb_data>0){ JDInfo t; memset(&t,0,sizeof(t)); jd_parse_header(buf,m,info.thumb_data,&t);} }
      free(buf);
    }
    free(src);
  }
  printf("cases %ld parsed-ok %ld\n",tot,ok); return 0;}
