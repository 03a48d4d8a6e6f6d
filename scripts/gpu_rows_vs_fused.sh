# parity tests (both fast paths are parametrised there), then the row walk (XM_ROWS=1) against the fused pair on the bench workloads
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in ${WORKLOADS:-se pe}; do for rows in 1 0; do
XM_ROWS=$rows python bench.py --records ${RECORDS:-20000000} --steps 5 --warmup 3 --no-cpu --no-e2e --workload $wl 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl rows=$rows', round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms; scan(s)', round(r['scan_kernel']['kernel_ms'],2), 'classify/emit', round(r['kernel_ms'],2), 'whole', round(r['whole_path']['achieved']), round(r['whole_path']['frac'],4), r['kernels_run'], d.get('ms_kernel'))"
done; done 2>&1 | tee gpurun_out/rows_vs_fused.txt
