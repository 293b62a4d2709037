mkdir -p gpurun_out/x2
python -m pytest tests/test_gpu_layered_i8.py tests/test_gpu_pins.py -x -q -m gpu > gpurun_out/x2/pytest.log 2>&1; tail -3 gpurun_out/x2/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/x2/bench.json 2> gpurun_out/x2/bench.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:layered_i8s -c 2 --csv --log-file gpurun_out/x2/ncu_fixed.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-fixed10 --fixed-iters --frames 29600 > gpurun_out/x2/ncu_fixed.log 2>&1
