set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2f.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2f.log
python bench.py --impl reference --steps 12 --warmup 3 > gpurun_out/bench_r2f_ref.json 2>/dev/null
python bench.py > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err; tail -c 400 gpurun_out/bench_r2f_n1.json
EPDE_BINDING=ctypes ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_l.log 2>&1
EPDE_BINDING=ctypes ncu --set full --clock-control none --import-source on -k regex:k_pass2_tc --launch-skip 4 -c 1 -o gpurun_out/p2_r2f -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_p2.log 2>&1
EPDE_BINDING=ctypes ncu --set full --clock-control none --import-source on -k regex:k_pass1_tc --launch-skip 4 -c 1 -o gpurun_out/p1_r2f -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_p1.log 2>&1
python tools/perf_md.py > gpurun_out/perf_md_r2f.txt 2>&1; tail -3 gpurun_out/perf_md_r2f.txt
python tools/perf_smallB.py > gpurun_out/perf_smallB_r2f.txt 2>&1; tail -8 gpurun_out/perf_smallB_r2f.txt
python tools/perf_loop.py > gpurun_out/perf_loop_r2f.txt 2>&1; tail -4 gpurun_out/perf_loop_r2f.txt
