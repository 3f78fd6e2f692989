#!/bin/bash
# round-2 GPU call 6: full GPU parity suite incl. the wide-character texts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf -x -k "wide" > gpurun_out/r02_c6_pytest_wide.log 2>&1
echo "pytest wide rc=$?"; tail -30 gpurun_out/r02_c6_pytest_wide.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf -k "not wide" > gpurun_out/r02_c6_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_c6_pytest.log
