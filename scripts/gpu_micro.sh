#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/solve_micro.py > gpurun_out/solve_micro.log 2>&1; echo rc=$?
tail -3 gpurun_out/solve_micro.log
