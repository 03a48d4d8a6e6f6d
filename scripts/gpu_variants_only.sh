# every libxm_var_*.so once on the headline walk (RECORDS, default 10 M), no parity tests
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for lib in xenomapper_b200/libxm_var_*.so; do
for sk in ${SKIPS:-1}; do XM_LIB_PATH=$PWD/$lib python bench.py --records ${RECORDS:-10000000} --steps 5 --warmup 3 --no-cpu --no-e2e --skip $sk ${WORKLOAD:+--workload $WORKLOAD} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib skip=$sk', round(d['value']/1e6,1), 'Mreads/s', round(d['ms_per_step'],2), 'ms; classify', round(r['kernel_ms'],2), round(r['achieved']), 'scan', round(r['scan_kernel']['kernel_ms'],2), round(r['scan_kernel']['achieved']), 'whole', round(r['whole_path']['achieved']), round(r['whole_path']['frac'],4), r['kernels_run'])"; done; done 2>&1 | tee gpurun_out/variants_quick.txt
