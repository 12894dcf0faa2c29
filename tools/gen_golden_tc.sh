#!/bin/bash
# golden vectors of the reference's tensorcore/ program (built by oracle/build_ref_tensorcore.py)
REF=$PWD/oracle/_ref/ising_tensorcore_ref
OUT=$PWD/gpurun_out/golden_tc
rm -rf $OUT; mkdir -p $OUT
run() { name=$1; shift; mkdir -p $OUT/$name; (cd $OUT/$name && timeout 900 $REF "$@" > stdout.txt 2> stderr.txt; echo "$@" > args.txt); tail -4 $OUT/$name/stdout.txt; }
run t_small   -x 2 -y 2 -w 3 -n 5 -a 1.0 -s 1234 -o
run t_rect    -x 1 -y 3 -w 0 -n 4 -a 0.661030 -s 77 -o
run t_hot     -x 2 -y 1 -w 2 -n 2 -a 2.0 -o
run t_mid     -x 16 -y 16 -w 10 -n 20 -a 1.0
run t_mid_cold -x 16 -y 8 -w 5 -n 50 -a 0.661030 -s 4242
run t_big_T1.5 -x 128 -y 128 -w 100 -n 1000 -a 0.661030
run t_big_Tc   -x 128 -y 128 -w 100 -n 1000 -a 1.0
for d in $OUT/*; do if [ -f $d/final.txt ]; then python3 - $d/final.txt <<'PY'
import sys,hashlib,numpy as np
rows=[l.split() for l in open(sys.argv[1]) if l.strip()]
a=np.array(rows,dtype=np.int8)
np.save(sys.argv[1].replace('final.txt','final.npy'), np.packbits(a>0))
open(sys.argv[1].replace('final.txt','final.shape'),'w').write("%d %d %s\n"%(a.shape[0],a.shape[1],hashlib.sha256(open(sys.argv[1],'rb').read()).hexdigest()))
PY
rm $d/final.txt; fi; done
ls -la $OUT/*
