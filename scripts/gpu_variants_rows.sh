# every libxm_var_*.so once on the headline walk, fused and over rows (RECORDS, default 20 M)
for lib in xenomapper_b200/libxm_var_*.so; do for rows in 0 1; do
XM_ROWS=$rows XM_LIB_PATH=$PWD/$lib python bench.py --records ${RECORDS:-20000000} --steps 5 --warmup 3 --no-cpu --no-e2e ${WORKLOAD:+--workload $WORKLOAD} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib rows=$rows', round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms; scan', round(r['scan_kernel']['kernel_ms'],2), 'classify/emit', round(r['kernel_ms'],2), d['ms_kernel'])"
done; done 2>&1 | tee gpurun_out/variants_rows.txt
