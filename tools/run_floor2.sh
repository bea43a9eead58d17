#!/bin/bash
# how much of cuMemUnmap's tail in the zero-page design is the ALIAS count of the one zero handle?
mkdir -p gpurun_out
for v in zero zero_multi8 nozero zero_multi64 zero zero_multi8; do
  echo "=== $v"; timeout 200 ./tools/vmm_floor tools/events_decode4096.txt $v 2 128 2>&1 | tee -a gpurun_out/floor2_$v.log | grep -E "variant|replay|Unmap|SetAccess"
done
