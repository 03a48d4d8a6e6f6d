python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for sk in 1 0; do python bench.py --records 20000000 --steps 5 --warmup 3 --no-cpu --no-e2e --skip $sk 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('skip', $sk, round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms; classify', round(r['kernel_ms'],2), round(r['achieved']), 'scan', round(r['scan_kernel']['kernel_ms'],2), round(r['scan_kernel']['achieved']), 'launches', d['gpu_launches'])"; done
python bench.py --workload pe --records 20000000 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pe', round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms; classify', round(r['kernel_ms'],2), round(r['achieved']), 'scan', round(r['scan_kernel']['kernel_ms'],2), 'launches', d['gpu_launches'])"
