python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python tools/run_configs.py c1 c3 c5 10 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c3_r1p.csv python tools/profile_workloads.py c3 3 > gpurun_out/c3.log 2>&1
