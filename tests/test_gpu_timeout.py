"""GPU: a neighbouring slab that never delivers its halo rows must not hang the GPU: the device-side waits are bounded
(ISING_B200_SPIN_TIMEOUT_MS), an error word is set and ising_sync reports ISING_ETIMEOUT."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def test_missing_neighbour_times_out_instead_of_hanging(tmp_path):
    """Slab 0 of a 2-slab lattice attaches slab 1 (a child process on the same GPU, CUDA IPC) as both neighbours;
    slab 1 never updates, so slab 0's white update of sweep 1 waits for a black halo push that never comes."""
    # CUDA IPC handles cannot be opened by the process that exported them, so this runs slab 1 in a child process
    # that attaches, signals readiness and then sleeps without ever updating.
    child = r'''
import sys, time, pickle
sys.path.insert(0, %r)
import ising_gpu_b200 as ig
from ising_gpu_b200.engine import F_NO_PEER_COPY
f = ig.IsingEngine(2048, 64, temp=2.0, slab_first=1, slab_count=2, flags=F_NO_PEER_COPY)
f.init_lattice()
pickle.dump(f.ipc_export(), open(sys.argv[1] + ".tmp", "wb")); import os; os.rename(sys.argv[1] + ".tmp", sys.argv[1])
while not os.path.exists(sys.argv[2]): time.sleep(0.05)
time.sleep(float(sys.argv[3]))
''' % ROOT
    parent = r'''
import sys, time, pickle, os
sys.path.insert(0, %r)
import ising_gpu_b200 as ig
from ising_gpu_b200.engine import F_NO_PEER_COPY
e = ig.IsingEngine(2048, 64, temp=2.0, slab_first=0, slab_count=2, flags=F_NO_PEER_COPY)
e.init_lattice()
while not os.path.exists(sys.argv[1]): time.sleep(0.05)
h = pickle.load(open(sys.argv[1], "rb"))
e.ipc_attach(h, h)
t0 = time.time()
try:
    e.sweeps(2)          # black of sweep 1 needs nothing; white of sweep 1 waits for the neighbour's black push
    print("NO ERROR")
except ig.IsingError as err:
    print("ERROR %%d %%.1f s: %%s" %% (err.code, time.time() - t0, err))
open(sys.argv[2], "w").write("done")
''' % ROOT
    hfile, dfile = str(tmp_path / "handles.pkl"), str(tmp_path / "done")
    env = dict(os.environ, ISING_B200_SPIN_TIMEOUT_MS="300")
    pc = subprocess.Popen([sys.executable, "-c", child, hfile, dfile, "0.2"], env=env)
    try:
        r = subprocess.run([sys.executable, "-c", parent, hfile, dfile], env=env, capture_output=True, text=True, timeout=120)
    finally:
        open(dfile, "w").write("done")
        pc.wait(timeout=60)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ERROR -6" in r.stdout and "did not arrive" in r.stdout, r.stdout + r.stderr[-500:]
