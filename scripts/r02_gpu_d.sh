#!/bin/bash
# round 2, call D: the whole GPU suite after splitting the case lists into items and tightening tolerances
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_d.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest_d.log | cut -c1-250
