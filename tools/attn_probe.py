"""Ad-hoc GPU probe (not part of the product): GraphDiT step time and the attention slot's share for the attention
variant chosen with LLB_ATTN (one variant per process).  Usage: LLB_ATTN=3 python tools/attn_probe.py [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llamole_b200 import _cabi, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
m, cfg, meta, sd = bench.build_dit(dev, small=False)
eng = m.engine()
B, N, T = 2048, m.max_n_nodes, 500
props, txt = synth.dit_conditions(B, seed=2024)
n_nodes = torch.full((B,), N, dtype=torch.int32)
if os.environ.get("PROBE_RAGGED"):
    n_nodes = torch.randint(5, N + 1, (B,), dtype=torch.int32, generator=torch.Generator().manual_seed(1))
props = props.to(dev)
props = torch.where(props == -200.0, torch.full_like(props, float("nan")), props).contiguous()
eng.begin(n_nodes, props, txt.to(dev).contiguous(), mol_index_base=0)
eng.init_state(7, None, None)
for i in range(3):
    eng.step(T - i, 7)
torch.cuda.synchronize()
_cabi.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    eng.step(T - 3 - i, 7)
e1.record()
torch.cuda.synchronize()
prof = _cabi.profile_read()
X, E = eng.get_state()
print("LLB_ATTN=%s step %.2f ms  attention %.2f ms/step  qkv %.2f proj %.2f fc1 %.2f fc2 %.2f  state-sum %d" % (
    os.environ.get("LLB_ATTN", "default"), e0.elapsed_time(e1) / steps, prof["attention"][0] / steps, prof["gemm_qkv"][0] / steps,
    prof["gemm_proj"][0] / steps, prof["gemm_fc1"][0] / steps, prof["gemm_fc2"][0] / steps, int(X.long().sum() + E.long().sum())), flush=True)
