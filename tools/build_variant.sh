#!/bin/bash
# development aid: build libjpegdec_b200 with extra -D flags into jpegdec_b200/_variants/<name>.so  (A/B runs: JPEGDEC_B200_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p jpegdec_b200/_variants
B=jpegdec_b200/_build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c jpegdec_b200/csrc/jd_device.cu -o /tmp/jd_device_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o jpegdec_b200/_variants/$name.so $B/jd_host.c.o $B/jd_api.c.o /tmp/jd_device_$name.o -lpthread
echo built jpegdec_b200/_variants/$name.so
