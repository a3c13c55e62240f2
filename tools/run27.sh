mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 --tb=short --durations=5 2>&1 | tail -30 > gpurun_out/t27.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke27.log 2>&1
