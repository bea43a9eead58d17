#!/bin/bash
mkdir -p gpurun_out
for t in 1 2 4 8 1; do ./tools/vmm_stall2 threads $t 8 128; done > gpurun_out/stall2_threads.log 2>&1
for r in 1 4 16 1; do ./tools/vmm_stall2 range $r 8 128; done > gpurun_out/stall2_range.log 2>&1
for r in 1 16; do ./tools/vmm_stall2 range $r 6 2; done >> gpurun_out/stall2_range.log 2>&1
