set -x
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c15_bench_n2.json 2> gpurun_out/c15_bench_n2.err
cat gpurun_out/c15_bench_n2.json | cut -c1-600; tail -3 gpurun_out/c15_bench_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c15_bench_n1.json 2> gpurun_out/c15_bench_n1.err
cat gpurun_out/c15_bench_n1.json | cut -c1-300
