"""Ad-hoc GPU probe (not part of the product): GraphCLIP forward time and per-slot breakdown for the aggregation variant
chosen with LLB_GIN_AGG (0 = separate aggregate + pool kernels); writes the embeddings so two runs can be compared bit
for bit.  Usage: LLB_GIN_AGG=0 python tools/gin_probe.py out.pt"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llamole_b200 import _cabi, synth

dev = torch.device("cuda", 0)
g, enc, proj = bench.build_gin(dev)
G = 4096
x, ei, ea, batch = synth.molecular_graphs(G, seed=0)
xd, eid, ead, bd = (t.to(dev) for t in (x, ei, ea, batch))
eng = g.engine()
for _ in range(3):
    eng.bind(xd, eid, ead, bd, num_graphs=G)
    out = eng.encoder_forward()
torch.cuda.synchronize()
_cabi.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters):
    eng.bind(xd, eid, ead, bd, num_graphs=G)
    out = eng.encoder_forward()
e1.record()
torch.cuda.synchronize()
prof = _cabi.profile_read()
print("LLB_GIN_AGG=%s: %.3f ms/forward" % (os.environ.get("LLB_GIN_AGG", "default"), e0.elapsed_time(e1) / iters),
      {k: "%.3f ms x %d" % (v[0] / iters, v[1] / iters) for k, v in prof.items() if v[1]}, flush=True)
if len(sys.argv) > 1:
    prev = sys.argv[1]
    if os.path.exists(prev):
        ref = torch.load(prev)
        print("bit-identical to %s: %s (max |d| %.3g)" % (prev, torch.equal(ref, out.cpu()), float((ref - out.cpu()).abs().max())), flush=True)
    else:
        torch.save(out.cpu(), prev)
