# ncu --set full capture of the two walk kernels on a 2 M-record walk (after the same command has exited 0 without ncu)
CMD="python bench.py --records 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2>/dev/null && \
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 2 -o gpurun_out/${NAME:-r02_prof} -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; cut -c1-300 gpurun_out/prof_plain.json
