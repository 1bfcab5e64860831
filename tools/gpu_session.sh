#!/bin/bash
# development aid: one gpurun call = tests + A/B of library variants + ncu captures + bench lines.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh <tag> [steps...]'      steps: tests ab ncu bench uhd10k
tag=$1; shift
steps="$*"; [ -z "$steps" ] && steps="tests ab ncu bench"
: ${ABWORKLOADS:="hd1024 uhd"}
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${tag}_smi.txt 2>&1
lscpu | head -25 > $O/${tag}_lscpu.txt 2>&1
for s in $steps; do case $s in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log; tail -5 $O/${tag}_pytest.log ;;
ab)
  ab() { # name lib env workload
    ( [ "$2" != main ] && export JPEGDEC_B200_LIB=$PWD/jpegdec_b200/_variants/$2.so; [ -n "$3" ] && export $3
      timeout 600 python bench.py --workload $4 --no-cpu --no-e2e --steps 8 --warmup 3 --unique 32 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); s=d['stages_ms']
    print('$1 $4', 'total %.3f prescan %.3f entropy %.3f stitch %.3f idct %.3f' % (d['ms_per_step'], s['prescan'], s['entropy'], s['stitch'], s['idct']), str(d['parity_spot_check'])[:12])
except Exception as e: print('$1 $4 FAILED', e)" ) >> $O/${tag}_ab.txt 2>&1; }
  for wl in $ABWORKLOADS; do
    ab default main "" $wl
    for e in $ABENVS; do ab $e main $e $wl; done
    for v in $ABVARIANTS; do ab $v $v "" $wl; done
  done
  cat $O/${tag}_ab.txt ;;
ncu)
  B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --unique 16"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${tag}_launches_hd1024.csv $B --workload hd1024 > $O/${tag}_ncu_launches.log 2>&1
  prof() { # name regex workload
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o $O/${tag}_$1 $B --workload $3 > $O/${tag}_ncu_$1.log 2>&1
    python tools/ncu_summary.py $O/${tag}_$1.ncu-rep $O/${tag}_$1_summary.txt > /dev/null 2>&1
    sz=$(stat -c %s $O/${tag}_$1.ncu-rep 2>/dev/null || echo 0); [ "$sz" -gt 14000000 ] && rm -f $O/${tag}_$1.ncu-rep; }
  for spec in ${NCUSPECS:-entropy_hd1024:jdk_entropy:hd1024 idct_tb_hd1024:jdk_idct_tb:hd1024 idct_tb_uhd:jdk_idct_tb:uhd}; do
    IFS=: read n rx w <<< "$spec"; prof $n $rx $w
  done
  ls -la $O | tail -20 ;;
bench)
  timeout 900 python bench.py --pipelined > $O/${tag}_bench_hd1024_1gpu.json 2> $O/${tag}_bench_hd1024.err; tail -c 600 $O/${tag}_bench_hd1024_1gpu.json
  timeout 600 python bench.py --workload uhd --no-cpu --pipelined > $O/${tag}_bench_uhd_1gpu.json 2> $O/${tag}_bench_uhd.err ;;
norst)
  B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --unique 16"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${tag}_launches_hd_norst.csv $B --workload hd_norst > $O/${tag}_ncu_launches_norst.log 2>&1
  timeout 600 python bench.py --workload hd_norst --no-cpu --steps 5 --warmup 3 --unique 32 > $O/${tag}_bench_hd_norst_1gpu.json 2> $O/${tag}_bench_hd_norst.err
  tail -c 900 $O/${tag}_bench_hd_norst_1gpu.json; echo ;;
sanitize)
  CS=/usr/local/cuda/bin/compute-sanitizer
  K="fixture_batch or synthetic_formats or dither_batch or progressive or restart_free or corrupt or two_threads"
  timeout 1500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $O/${tag}_compute_sanitizer_memcheck.txt 2>&1; tail -3 $O/${tag}_compute_sanitizer_memcheck.txt
  timeout 900 $CS --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/${tag}_compute_sanitizer_racecheck.txt 2>&1; tail -3 $O/${tag}_compute_sanitizer_racecheck.txt
  timeout 900 $CS --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "restart_free or dither_batch or two_threads" > $O/${tag}_compute_sanitizer_synccheck.txt 2>&1; tail -3 $O/${tag}_compute_sanitizer_synccheck.txt ;;
onecall)
  for mb in ${ONECALL_MB:-32 64 128 192}; do JPEGDEC_B200_JOB_MB=$mb timeout 600 python tools/onecall_probe.py ${ONECALL_N:-625} >> $O/${tag}_onecall.txt 2>&1; done
  cat $O/${tag}_onecall.txt ;;
scale_n)
  for nimg in ${SCALE_NS:-64 256 1024}; do timeout 600 python bench.py --workload ${SCALE_WL:-dither} --images $nimg --no-cpu --no-e2e --steps 5 --warmup 3 --unique 32 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$nimg images', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})" >> $O/${tag}_scale_n.txt 2>&1; done; cat $O/${tag}_scale_n.txt ;;
others)
  for wl in ${OTHERS:-uhd_quarter uhd_eighth dither dither444 hd_norst}; do
    timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 5 --warmup 3 --unique 32 > $O/${tag}_bench_${wl}_1gpu.json 2> $O/${tag}_bench_${wl}.err
    tail -c 300 $O/${tag}_bench_${wl}_1gpu.json; echo
  done ;;
multi)
  # N-GPU lines (gpurun --gpus N): the north-star slice per GPU with every image verified, then the default workload with e2e
  N=${NGPUS:-8}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  nvidia-smi topo -m > $O/${tag}_topo.txt 2>&1
  timeout 1500 $TR bench.py --gpus $N --workload uhd10k --steps 3 --warmup 3 --no-cpu > $O/${tag}_bench_uhd10k_${N}gpu.json 2> $O/${tag}_bench_uhd10k_${N}gpu.err; tail -c 1200 $O/${tag}_bench_uhd10k_${N}gpu.json
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O/${tag}_bench_hd1024_${N}gpu.json 2> $O/${tag}_bench_hd1024_${N}gpu.err; tail -c 900 $O/${tag}_bench_hd1024_${N}gpu.json
  if [ -n "$NOBIND" ]; then JPEGDEC_B200_NO_BIND=1 timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > $O/${tag}_bench_hd1024_${N}gpu_nobind.json 2> $O/${tag}_bench_hd1024_${N}gpu_nobind.err; fi ;;
probe)
  N=${NGPUS:-4}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
  nvidia-smi topo -m > $O/${tag}_topo.txt 2>&1
  timeout 600 $TR tools/d2h_probe.py --out $O/${tag}_probe_bind > $O/${tag}_probe_bind.log 2>&1; cat $O/${tag}_probe_bind.rank* 
  JPEGDEC_B200_NO_BIND=1 timeout 600 $TR tools/d2h_probe.py --out $O/${tag}_probe_nobind > $O/${tag}_probe_nobind.log 2>&1; cat $O/${tag}_probe_nobind.rank*
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O/${tag}_bench_hd1024_${N}gpu.json 2> $O/${tag}_bench_hd1024_${N}gpu.err; tail -c 900 $O/${tag}_bench_hd1024_${N}gpu.json ;;
uhd10k)
  timeout 1200 python bench.py --workload uhd10k --no-cpu --steps 3 --warmup 3 > $O/${tag}_bench_uhd10k_1gpu.json 2> $O/${tag}_bench_uhd10k.err; tail -c 1500 $O/${tag}_bench_uhd10k_1gpu.json; tail -5 $O/${tag}_bench_uhd10k.err ;;
refarm)
  timeout 600 python bench.py --impl reference > $O/${tag}_bench_reference_arm.json 2> $O/${tag}_bench_reference.err ;;
esac; done
echo session $tag done
