# Round-end verification on the GPU box: GPU tests, smoke, both bench arms, the ncu launch list of the bench command and one full
# capture per kernel of the round.  Outputs under gpurun_out/ (prefix r02f_).  Run from the repository root.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
for s in full packed computed; do python tools/run_eval.py --case ground4 --layout instance --slices $s --n 65536 --steps 200 --graph --ready; done > gpurun_out/r02f_slices.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_instance_major_cta -s 3 -c 1 -f -o gpurun_out/r02f_imc_computed_ground4_65536 python tools/run_eval.py --case ground4 --layout instance --slices computed --n 65536 --steps 3 --warmup 2 > gpurun_out/r02f_ncu_computed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tail -c 1 -f -o gpurun_out/r02f_k_tail python tools/solve_batch.py --no-cpu --instances 256 --repeats 1 > gpurun_out/r02f_ncu_tail.log 2>&1
tail -n 3 gpurun_out/r02f_tests_gpu.log; tail -n 2 gpurun_out/r02f_smoke.log; cut -c1-300 gpurun_out/r02f_bench.json; cat gpurun_out/r02f_slices.txt
