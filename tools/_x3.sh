mkdir -p gpurun_out/x3
python -m pytest tests -x -q -m gpu > gpurun_out/x3/pytest.log 2>&1; tail -3 gpurun_out/x3/pytest.log
python bench.py > gpurun_out/x3/bench.json 2> gpurun_out/x3/bench.err
python bench.py --discard-scratch --no-cpu > gpurun_out/x3/bench_discard.json 2> gpurun_out/x3/bench_discard.err
