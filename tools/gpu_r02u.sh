#!/bin/bash
# band pre-labelling + vectorised band morphology: the whole GPU suite, then the tail under load
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_pytest.log 2>&1
tail -15 gpurun_out/r02u_pytest.log
timeout -k 10 300 python tools/tail_probe.py > gpurun_out/r02u_tail_probe.txt 2>&1
cat gpurun_out/r02u_tail_probe.txt
