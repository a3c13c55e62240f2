mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=short 2>&1 | tail -150 > gpurun_out/t3.log
timeout 280 python __graft_entry__.py smoke > gpurun_out/smoke3.log 2>&1
