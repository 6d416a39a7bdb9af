#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_llama_decode.py 8 > gpurun_out/c15_llama_prof.log 2>&1; echo "rc=$?"; grep -v Warning gpurun_out/c15_llama_prof.log | tail -32
