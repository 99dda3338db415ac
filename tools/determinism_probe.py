"""Ad-hoc GPU probe (not part of the product): run-to-run determinism of the sampler at the benchmark shape.  The same
batch is sampled several times from the same seed in one process; the integer states must be identical."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llamole_b200 import synth

dev = torch.device("cuda", 0)
m, cfg, meta, sd = bench.build_dit(dev, small=False)
eng = m.engine()
B, N, T = int(os.environ.get("PROBE_BATCH", "2048")), m.max_n_nodes, 500
props, txt = synth.dit_conditions(B, seed=2024)
n_nodes = torch.full((B,), N, dtype=torch.int32)
if os.environ.get("PROBE_RAGGED"):
    n_nodes = torch.randint(5, N + 1, (B,), dtype=torch.int32, generator=torch.Generator().manual_seed(1))
props = props.to(dev)
props = torch.where(props == -200.0, torch.full_like(props, float("nan")), props).contiguous()
eng.begin(n_nodes, props, txt.to(dev).contiguous(), mol_index_base=0)
steps = int(os.environ.get("PROBE_STEPS", "4"))
ref = None
for rep in range(4):
    eng.init_state(7, None, None)
    for i in range(steps):
        eng.step(T - i, 7)
    lX, lE = eng.denoise(T - steps, False)
    X, E = eng.get_state()
    torch.cuda.synchronize()
    cur = (X.clone(), E.clone(), lX.clone(), lE.clone())
    if ref is None:
        ref = cur
    else:
        dx = int((cur[0] != ref[0]).sum()); de = int((cur[1] != ref[1]).sum())
        dl = float((cur[2] - ref[2]).abs().max()); dle = float((cur[3] - ref[3]).abs().max())
        print(f"rep {rep}: atoms differing {dx}, bonds differing {de}, max |d logits| X {dl:.3g} E {dle:.3g}", flush=True)
print("modes", {k: os.environ.get(k) for k in ("LLB_FUSED_LN", "LLB_ATTN", "LLB_ADALN_GROUPED", "LLB_GEMM_PAIR")}, flush=True)
