#!/bin/bash
# round-2 closing visit on one B200: parity tests, smoke, full-size bench (fused and over rows), reference arm, workloads,
# BAM leg, ncu launch list + full captures.  Everything lands in gpurun_out/ as r02_final_*.
set -x
T=r02_final
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${T}_bench_n1_100m.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.err; cut -c1-700 gpurun_out/${T}_bench_n1_100m.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err
XM_ROWS=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_rows_100m.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_rows_100m.json
for w in pe pe_conservative_zs pe_cigar; do python bench.py --workload $w --records 40000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-200 gpurun_out/${T}_bench_$w.json; done
CMD="python bench.py --records 20000000 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 2 -o gpurun_out/${T}_prof -f $CMD > gpurun_out/ncu_full.log 2>&1
XM_ROWS=1 ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 5 -o gpurun_out/${T}_prof_rows -f $CMD > gpurun_out/ncu_full_rows.log 2>&1
python scripts/bench_bam.py --make 4000000 && python scripts/bench_bam.py > gpurun_out/${T}_bench_bam_4m.json 2> gpurun_out/bench_bam.err; cat gpurun_out/${T}_bench_bam_4m.json
XM_BAM_INFLATE=host python scripts/bench_bam.py > gpurun_out/${T}_bench_bam_4m_host_inflate.json 2>/dev/null
timeout 300 python scripts/bench_bgzf.py 2000000 > gpurun_out/${T}_bench_bgzf_device.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_bgzf_device.json
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${T}_gpu.txt
ls -la gpurun_out | tail -12
