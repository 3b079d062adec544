#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2k_tests.txt; cat gpurun_out/r2k_tests.txt
timeout 200 python -c "
import sys, json; sys.path.insert(0,'.')
import torch, bench
print(json.dumps(bench.offpolicy_extras(torch.device('cuda:0'))))" > gpurun_out/r2k_offpolicy.json 2> gpurun_out/r2k_offpolicy.err
tail -1 gpurun_out/r2k_offpolicy.json | cut -c1-700
