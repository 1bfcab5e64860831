#!/bin/bash
# development aid: same-box A/B of library variants:  tools/ab.sh <workload> name1 name2 ...   ("main" = the in-tree library)
wl=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset JPEGDEC_B200_LIB; else export JPEGDEC_B200_LIB=$PWD/jpegdec_b200/_variants/$v.so; fi
  python bench.py --workload $wl --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$v', 'total %.3f entropy %.3f idct %.3f' % (d['ms_per_step'], s['entropy'], s['idct']), d['parity_spot_check'][:5])"
done
