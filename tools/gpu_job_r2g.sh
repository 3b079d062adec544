#!/bin/bash
# GPU-box visit r2g: host-side probes of the default-order path (1 GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TS_B200_PERM_TRACE=1 timeout 120 python tools/host_pass_trace.py > gpurun_out/r2g_host_pass_trace.txt 2>&1
timeout 120 python tools/default_order_hosttrace.py > gpurun_out/r2g_default_order_hosttrace.txt 2>&1
timeout 200 python tools/default_order_probe.py > gpurun_out/r2g_default_order_probe.txt 2>&1
grep -c . gpurun_out/r2g_host_pass_trace.txt
grep "pass" gpurun_out/r2g_host_pass_trace.txt | tail -24
tail -13 gpurun_out/r2g_host_pass_trace.txt
cat gpurun_out/r2g_default_order_hosttrace.txt | tail -45
cat gpurun_out/r2g_default_order_probe.txt
