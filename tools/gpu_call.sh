timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2j_pytest.log | cut -c1-250
