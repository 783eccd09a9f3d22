ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/r2i_solve_launches.csv python tools/solve_batch.py --solver native --no-cpu --case ground --instances 4096 --repeats 1 > gpurun_out/r2i_solve_ncu.log 2>&1
tail -2 gpurun_out/r2i_solve_ncu.log
