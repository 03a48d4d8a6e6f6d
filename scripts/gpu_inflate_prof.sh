#!/bin/bash
# BAM leg on one B200: bench line, then one ncu capture of k_bgzf_inflate (and the launch list of the whole leg)
set -u
N=${1:-2000000}
python scripts/bench_bam.py --make $N || exit 1
timeout 300 python scripts/bench_bam.py 2>&1 | tail -1 | tee gpurun_out/r02_bench_bam_dev.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate -c 1 -o gpurun_out/r02_inflate -f python scripts/bench_bam.py > gpurun_out/ncu_inflate.log 2>&1
tail -2 gpurun_out/ncu_inflate.log
