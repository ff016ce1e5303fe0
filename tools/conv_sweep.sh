#!/bin/bash
# A/B of the persistent kernel's tile width (SEG_TC_V2BN) and the wgrad tile width (SEG_TC_WGRAD_BN) on the C3 hot shapes.
# usage: bash tools/conv_sweep.sh > gpurun_out/conv_sweep.txt
cd "$(dirname "$0")/.."
SHAPES_FD="--shape 16,33,33,256,1024,1,1,1 --shape 16,33,33,1024,256,1,1,1 --shape 16,33,33,256,256,3,1,1 --shape 16,33,33,512,512,3,1,2 --shape 16,33,33,512,2048,1,1,1 --shape 16,33,33,2048,512,1,1,1 --shape 16,33,33,2048,256,3,1,12 --shape 16,65,65,128,512,1,1,1 --shape 16,65,65,512,128,1,1,1 --shape 16,129,129,64,256,1,1,1 --shape 16,129,129,304,256,3,1,1 --shape 16,129,129,256,256,3,1,1"
for bn in 128 256; do
  echo "=== fprop/dgrad SEG_TC_V2BN=$bn"
  SEG_TC_V2BN=$bn timeout -k 5 120 python tools/conv_micro.py $SHAPES_FD --kind fwd --iters 30
  SEG_TC_V2BN=$bn timeout -k 5 120 python tools/conv_micro.py $SHAPES_FD --kind dgrad --iters 30 --beta 1
done
for bn in 128 256; do
  echo "=== wgrad SEG_TC_WGRAD_BN=$bn"
  SEG_TC_WGRAD_BN=$bn timeout -k 5 120 python tools/conv_micro.py $SHAPES_FD --kind wgrad --iters 30
done
