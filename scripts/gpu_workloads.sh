# paired / conservative / cigar workloads (configs[2], [3] shapes) and the BAM leg (configs[4] shape), for profiles/; parity tests last, logged
for w in pe pe_conservative_zs pe_cigar; do python bench.py --workload $w --records 40000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']; print('$w', round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms classify', round(r['achieved']), 'GB/s', round(r['frac'],3), 'scan', round(r['scan_kernel']['achieved']), 'e2e', round(d['e2e']['value']/1e6,1))"; done
python scripts/bench_bam.py > gpurun_out/bench_bam.json 2> gpurun_out/bench_bam.err; cat gpurun_out/bench_bam.json; tail -3 gpurun_out/bench_bam.err
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
