G=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/r2l_bench_$G.json 2> gpurun_out/r2l_bench_$G.err
tail -3 gpurun_out/r2l_bench_$G.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 tools/solve_batch.py --solver native --no-cpu --case ground --instances $((4096*G)) > gpurun_out/r2l_solve_$G.json 2> gpurun_out/r2l_solve_$G.err
tail -2 gpurun_out/r2l_solve_$G.json | cut -c1-400
nvidia-smi topo -m > gpurun_out/r2l_topo_$G.txt 2>&1
