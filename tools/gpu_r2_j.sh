#!/bin/bash
# Round-2 GPU call J (1 GPU): what the driver runs at round end -- the whole `pytest -m gpu` suite and smoke().
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2j_pytest.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/r2j_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1
echo "smoke exit $?"; tail -n 3 gpurun_out/r2j_smoke.log
