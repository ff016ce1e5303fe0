# epilogue experiments on the persistent conv kernel (SEG_TC_DBG: 1 no global I/O, 2 no stats, 4 no staging, 16 manual stores instead of TMA)
S1=16,33,33,256,1024,1,1,1
S2=16,33,33,1024,256,1,1,1
for d in 0 16 1 2 7; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $S1 --kind fwd --iters 100 --stats 1; done
SEG_TC_DBG=0 python tools/conv_micro.py --shape $S1 --kind fwd --iters 100 --stats 0
for d in 0 16 1; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $S2 --kind dgrad --iters 100 --beta 1; done
SEG_TC_DBG=0 python tools/conv_micro.py --shape $S2 --kind dgrad --iters 100 --beta 0
for sh in 16,33,33,1024,256,1,1,1 16,33,33,256,256,3,1,1 16,129,129,64,256,1,1,1 16,65,65,128,512,1,1,1 16,129,129,256,256,3,1,1; do for d in 0 16; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $sh --kind fwd --iters 50 --stats 1; done; done
for sh in 16,129,129,256,64,1,1,1 16,65,65,512,128,1,1,1; do for d in 0 16; do SEG_TC_DBG=$d python tools/conv_micro.py --shape $sh --kind dgrad --iters 50 --beta 1; done; done
