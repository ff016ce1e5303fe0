S1=16,33,33,256,1024,1,1,1
S2=16,33,33,1024,256,1,1,1
for d in 0 16; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $S1 --kind fwd --iters 100 --stats 1; done
for d in 0 16; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $S2 --kind dgrad --iters 100 --beta 0; done
for d in 0 16; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $S2 --kind dgrad --iters 100 --beta 1; done
