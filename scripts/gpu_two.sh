# two-GPU visit (gpurun --gpus 2): sharded CLI over NCCL, bench at N=2 and N=1 on the same box
python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --records 50000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.err; cut -c1-700 gpurun_out/bench_n2.json
python bench.py --gpus 1 --records 50000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_50m.json 2> gpurun_out/bench_n1_50m.err; cut -c1-400 gpurun_out/bench_n1_50m.json
