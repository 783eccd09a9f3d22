# Round-end verification on the GPU box: GPU tests, smoke, both bench arms, the variant table, the ncu launch list of the
# bench command and one full capture per kernel family.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/variant_table.py > gpurun_out/variants.md 2> gpurun_out/variants.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_instance_major -s 3 -c 1 -f -o gpurun_out/prof_instance_major_ground4_65536 python tools/run_eval.py --case ground4 --layout instance --n 65536 --steps 3 --warmup 2 > gpurun_out/ncu_instance.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_component_major -s 3 -c 1 -f -o gpurun_out/prof_component_major_superquadric8_65536 python tools/run_eval.py --case superquadric8 --layout component --n 65536 --steps 3 --warmup 2 > gpurun_out/ncu_sq8.log 2>&1
tail -3 gpurun_out/tests_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json | cut -c1-400
