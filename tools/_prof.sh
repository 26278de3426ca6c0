python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/run_configs.py c1 c3 c5 10 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c5_r1k.csv python tools/profile_workloads.py frame 3 > gpurun_out/frame.log 2>&1
