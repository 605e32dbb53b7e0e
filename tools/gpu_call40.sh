mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/c43_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c43_smoke.log 2>&1
cat gpurun_out/c43_tests.log; tail -n 1 gpurun_out/c43_smoke.log
