cd /root/repo
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02g_bench_${N}gpu.json 2> gpurun_out/r02g_bench_${N}gpu.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/solve_batch.py --no-cpu --instances $((4096*N)) > gpurun_out/r02g_solve_${N}gpu.json 2> gpurun_out/r02g_solve_${N}gpu.err; echo "solve rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/solve_batch.py --no-cpu --instances $((4096*N)) --tail 0 > gpurun_out/r02g_solve_${N}gpu_notail.json 2> gpurun_out/r02g_solve_${N}gpu_notail.err; echo "solve notail rc=$?"
tail -n 1 gpurun_out/r02g_solve_${N}gpu.json gpurun_out/r02g_solve_${N}gpu_notail.json | cut -c1-900
