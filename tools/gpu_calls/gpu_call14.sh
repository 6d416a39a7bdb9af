#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_gemm4_pair.py trace 4096x4096x4096 > gpurun_out/c14_trace.log 2>&1
grep -E "MMA thread|decode:|a-stage|epilogue|trace " gpurun_out/c14_trace.log
