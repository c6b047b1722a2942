mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1f.log 2>&1; tail -3 gpurun_out/pytest_r1f.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r1f.log 2>&1; tail -1 gpurun_out/smoke_r1f.log
timeout 900 python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 300 gpurun_out/bench_r1f.json
PROF_MEGA=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv python tools/prof_run.py > gpurun_out/prof_r1f.log 2>&1; tail -1 gpurun_out/prof_r1f.log
PROF_MEGA=1 PROF_WHAT=decode PROF_T=6 timeout 400 ncu --set full --clock-control none --import-source on -k decode_mega_kernel -s 7 -c 1 -f -o gpurun_out/r01_mega3 python tools/prof_run.py > gpurun_out/ncu_mega3.log 2>&1; tail -1 gpurun_out/ncu_mega3.log
