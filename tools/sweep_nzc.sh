#!/bin/bash
# sweep of the z-chunk count of the ring kernels (tuning aid)
for shape in 512x512x64 512; do
  for nzc in "$@"; do
    echo -n "nzc=$nzc "; PHICUDA_RING_NZC=$nzc python tools/microbench.py $shape --ring-only 2>&1 | tail -1 | cut -c1-150
  done
done
