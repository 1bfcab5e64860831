/* Corrupted / truncated entropy-coded data through the per-thread DEVICE code (jd_core.h, jd_chunk.h are __host__ __device__;
 * tests/hostsim steps them on the CPU), both the restart-segment and the chunk-parallel path, all scales, under ASan + UBSan:
 * what compute-sanitizer checks on the GPU, checked here where no GPU exists. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
int hostsim_decode(const uint8_t *data, int size, int pixel_type, int options, int arith, uint8_t *out, int out_pitch, int *out_w, int *out_h, int *n_events, int *err);
static uint8_t *rd(const char *p, int *n){FILE*f=fopen(p,"rb");fseek(f,0,SEEK_END);*n=ftell(f);fseek(f,0,SEEK_SET);uint8_t*b=malloc(*n);if(fread(b,1,*n,f)!=(size_t)*n)exit(2);fclose(f);return b;}
int main(int argc,char**argv){
  unsigned seed=7; long tot=0, okc=0; const int iters=atoi(argv[1]);
  for(int a=2;a<argc;a++){
    int n; uint8_t*src=rd(argv[a],&n);
    for(int it=0;it<iters;it++){
      int m=n; if(it%6==0) m = 700 + rand_r(&seed)%(n-700);
      uint8_t*buf=malloc(m); memcpy(buf,src,m);
      int k=1+rand_r(&seed)%5;
      for(int j=0;j<k;j++){ int off=600+rand_r(&seed)%(m-600); buf[off]=(uint8_t)rand_r(&seed);}   /* entropy-coded segment */
      int pt=(int)(rand_r(&seed)%4), opt=(int[]){0,2,4,8}[rand_r(&seed)%4]; if (it&1) opt|=0x20000;     /* odd: chunk-parallel path */
      uint8_t*out=malloc(4*1024*1024); int w,h,ne,err=0;
      int rc=hostsim_decode(buf,m,pt,opt,(int)(rand_r(&seed)&1),out,4096,&w,&h,&ne,&err); tot++; if(rc==1) okc++;
      free(out); free(buf);
    }
    free(src);
  }
  printf("cases %ld decoded-ok %ld\n",tot,okc); return 0;}
