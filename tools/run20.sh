mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 --tb=short --durations=15 2>&1 | tail -40 > gpurun_out/t20.log
