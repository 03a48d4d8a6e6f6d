bash scripts/gpu_variants_bench.sh
CMD="python bench.py --records 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e"
XM_LIB_PATH=$PWD/xenomapper_b200/libxm_var_w8.so ncu --set full --clock-control none --import-source on -k regex:k_classify2 -s 3 -c 1 -o gpurun_out/r01_prof_w8 -f $CMD > gpurun_out/ncu_w8.log 2>&1
tail -1 gpurun_out/ncu_w8.log
