# one GPU visit: parity tests, phase breakdown (profiling twin), 20M-record bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for sk in 1 0; do XM_LIB_PATH=$PWD/xenomapper_b200/libxenomapper_b200_prof.so python bench.py --records 20000000 --steps 1 --warmup 3 --no-cpu --no-e2e --skip $sk 2>&1 >/dev/null | grep "phases" | tail -2; done
python bench.py --records 20000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_20m.json 2> gpurun_out/bench_20m.err; cat gpurun_out/bench_20m.json | cut -c1-1800
