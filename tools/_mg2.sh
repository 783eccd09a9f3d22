python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/solve_batch.py --solver native --no-cpu --case ground --instances 8192 2>/dev/null | tail -1 | cut -c1-700
python tools/solve_batch.py --solver native --no-cpu --case ground --instances 4096 | cut -c1-600
python tools/q5_allowance.py
