#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trace.py -m gpu -x -q 2>&1 | tail -3
python tools/trace_probe.py --frames 20 --lod --rounds 3 --ab 0:128,0:256,0:512,0:1024,0:128:256,0:128:512,0:128:1024 2>&1 | tee gpurun_out/r2k_trace_ab2.log
