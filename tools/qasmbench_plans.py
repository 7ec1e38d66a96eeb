"""Plan every QASMBench circuit of the reference tree (CPU only): passes / rounds / swaps per circuit.
Reads /root/reference, so it runs in the build container, not on the GPU box."""
import sys, glob, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from quantum_simulations_b200.circuit.qasm import qasm_to_ops
from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
from quantum_simulations_b200.circuit import sharding
files=sorted(glob.glob('/root/reference/v3_hisvsim_spark/hisvsim_repo/QASMBench/**/*.qasm', recursive=True))
rows=[]
for f in files:
    name=os.path.basename(f)[:-5]
    try:
        n,ops=qasm_to_ops(open(f).read())
    except Exception as e:
        continue
    if len(ops)>60000: print(name,'skip',len(ops)); continue
    try:
        t0=time.time()
        fops=fuse_2q_blocks(ops, tol=1e-14)        # what kernel.cuda_dense.simulate_qasm uses
        g=max(0,n-30)
        if n<11:
            prog=sharding.plan_single(fops,n)
        elif g==0:
            prog=sharding.plan_single(fops,n)
        else:
            prog=sharding.plan(fops,n,n-g,swap_anywhere=True,rank_flips=True)
        dt=time.time()-t0
        st=prog.stats
        print(f"{name:22s} n={n:2d} gates={len(ops):6d} fused={len(fops):6d} passes={st['passes']:3d} rounds={st['rounds']:3d} swaps={st.get('swaps',0)} dense={st.get('dense2q_steps',0)+st.get('dense1q_steps',0)} plan_s={dt:.2f}", flush=True)
    except Exception as e:
        print(name,'PLAN FAIL',type(e).__name__,str(e)[:200], flush=True)
