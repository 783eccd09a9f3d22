timeout 900 python -m pytest tests/test_solve.py -x -q -m gpu -k "native" 2>&1 | tail -15
python tools/solve_batch.py --solver native --no-cpu --case ground --instances 4096
python tools/solve_batch.py --solver native --no-cpu --case com_planner --instances 4096
python tools/solve_batch.py --solver native --no-cpu --case superquadric --instances 1024
python tools/solve_batch.py --solver torch --no-cpu --case ground --instances 4096 --repeats 1
python tools/solve_batch.py --solver native --no-cpu --case ground --instances 65536 --repeats 2
