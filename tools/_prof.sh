python -m pytest tests -m gpu -q -x -k "batching or cube_field or multi_frame or config1" 2>&1 | tail -2
python tools/run_configs.py c1 c3 c5 10 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c3_r1s.csv python tools/profile_workloads.py c3 3 > gpurun_out/c3.log 2>&1
