#!/bin/bash
# development aid: short gpurun call: one ncu capture + the GPU tests.   bash tools/gpu_quick.sh <tag> <regex> <workload> [tests]
tag=$1; rx=$2; wl=$3; O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --unique 16"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o $O/${tag}_prof $B --workload $wl > $O/${tag}_ncu.log 2>&1
python tools/ncu_summary.py $O/${tag}_prof.ncu-rep $O/${tag}_summary.txt > /dev/null 2>&1
head -40 $O/${tag}_summary.txt
if [ "$4" = tests ]; then timeout 1500 python -m pytest tests -m gpu -x -q > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log; tail -5 $O/${tag}_pytest.log; fi
