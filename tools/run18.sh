mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=short 2>&1 | tail -30 > gpurun_out/t18.log
timeout 900 python bench.py --steps 6 --warmup 3 --cpu-frames 0 > gpurun_out/bench18.log 2>gpurun_out/bench18.err
timeout 900 python bench.py --steps 6 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/bench18_nograph.log 2>gpurun_out/bench18_nograph.err
