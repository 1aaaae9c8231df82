#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r2i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
