#!/bin/bash
# where does the variant start to differ from the reference tensorcore binary?
X=${1:-128}; Y=${2:-128}; A=${3:-1.0}
for K in 1 2 5 10 20 50 100 200 400 800 1100; do
  r=$(oracle/_ref/ising_tensorcore_ref -x $X -y $Y -w 0 -n $K -a $A | grep "average magnetism" | awk '{print $NF}')
  m=$(python - <<PY
import sys; sys.path.insert(0,'.')
from ising_gpu_b200.tensorcore import TensorcoreEngine
with TensorcoreEngine($X,$Y,alpha=$A) as T:
    T.init(); T.update($K); print("%g %d" % (abs(T.sum()/float(T.nx*T.ny)), T.sum()))
PY
)
  echo "K=$K ref=$r mine=$m"
done
