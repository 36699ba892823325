python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/fin_bench_8gpus.json 2> gpurun_out/fin_bench_8gpus.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/fin_bench_2gpus.json 2> gpurun_out/fin_bench_2gpus.err
python -m pytest tests/test_ensemble_sharding.py -q -m gpu 2>&1 | tail -3 > gpurun_out/fin_pytest_gpu_2gpus.log
cut -c1-120 gpurun_out/fin_bench_8gpus.json; cut -c1-120 gpurun_out/fin_bench_2gpus.json; cat gpurun_out/fin_pytest_gpu_2gpus.log
