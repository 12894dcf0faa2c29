import sys; sys.path.insert(0,'.')
from ising_gpu_b200.tensorcore import TensorcoreEngine
for path in (0,0,1):
    with TensorcoreEngine(128,128,seed=1234,alpha=1.0,nn_path=path) as T:
        T.init(); s0=T.sum()
        out=[]
        for k in range(11):
            T.update(100); out.append(T.sum())
        print(path, s0, out, "%g"%abs(out[-1]/float(T.nx*T.ny)), flush=True)
