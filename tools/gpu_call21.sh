#!/bin/bash
# round-2 GPU call 21: the full parity suite and smoke() on the FINAL sources (hash b95595758709ee0b, k_offsets_emit on by default)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import bench; print('source hash', bench.source_hash())"
timeout 330 python -m pytest tests -m gpu -q --timeout=300 --tb=short -rf > gpurun_out/r02_c21_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_c21_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c21_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_c21_smoke.log
echo done
