"""Ad-hoc GPU probe (not part of the product): GEMM throughput vs cuBLAS and DiT step time at a few batch sizes."""
import os, sys, time, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llamole_b200 import GraphDiT, GraphCLIP, synth, _cabi
dev='cuda:0'
def ev_time(fn, iters):
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/iters
lib=_cabi.lib()
for (M,N,K,act) in [(204800,4096,1024,1),(204800,1024,4096,0),(204800,3072,1024,0),(204800,1024,1024,0),(25600,4096,1024,1),(4096,6144,1024,3)]:
    A=torch.randn(M,K,device=dev).bfloat16(); W=torch.randn(N,K,device=dev).bfloat16(); b=torch.randn(N,device=dev); C=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    f=lambda: _cabi.check(lib.llb_gemm_bf16(_cabi.ptr(A),K,_cabi.ptr(W),K,_cabi.ptr(b),_cabi.ptr(C),N,M,N,K,act,0,_cabi.stream_ptr()))
    f(); ms=ev_time(f,5); print(f"gemm M={M} N={N} K={K} act={act}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    t=lambda: torch.matmul(A,W.t()); t(); ms=ev_time(t,5); print(f"   cublas: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    del A,W,C
d=tempfile.mkdtemp()
cfg=synth.dit_config(); meta=synth.dit_meta(50)
t0=time.time(); synth.write_dit_checkpoint(d,cfg,meta); print('ckpt',time.time()-t0, flush=True)
m=GraphDiT(d+'/config.yaml', d+'/data.meta.json', torch.float32); m.init_model(d); m=m.to(dev)
for B in (16, 256, 2048):
    props,txt=synth.dit_conditions(B)
    n=torch.full((B,),50)
    eng=m.engine()
    m.generate_graphs(props,txt,-200,n_nodes=n,seed=1,steps=3); torch.cuda.synchronize(); l0=eng.launch_count()
    t0=time.time(); m.generate_graphs(props,txt,-200,n_nodes=n,seed=1,steps=5); torch.cuda.synchronize(); dt=time.time()-t0
    print(f"B={B} N=50: {dt/5*1000:.1f} ms/step -> {B/(500*dt/5):.2f} mol/s ; launches/step {(eng.launch_count()-l0)/5:.0f}; frac of 1405.6TF: {B*72.2e9/(dt/5)/1405.6e12:.3f}", flush=True)
