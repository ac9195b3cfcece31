#!/bin/bash
# Stage times of config #2 against the launch batch and the ECC cluster size (gpurun helper, not a test)
for B in 128 148 256 296 512 1024; do
  for C in auto 2 1; do
    if [ "$C" = auto ]; then unset SSK_ECC_CLUSTER; else export SSK_ECC_CLUSTER=$C; fi
    echo "B=$B cluster=$C"
    python tools/prof_step.py 3 $B 2>&1 | tail -2
  done
done
