# end-of-milestone GPU visit: parity tests, smoke, full-size bench, reference arm, ncu launch list + full capture (2 M records)
# TAG names the outputs (default r01_final5)
set -x
TAG=${TAG:-r01_final5}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
CMD="python bench.py --records 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/prof_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 2 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
cut -c1-600 gpurun_out/bench_full.json
