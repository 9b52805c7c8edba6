#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/r2c16_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c16_pytest.log
tail -3 gpurun_out/r2c16_pytest.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
