"""Ad-hoc GPU probe (not part of the product): device-side timeline of one small-batch reverse step.  Needs the library built
with LLB_EXTRA_NVCC_FLAGS=-DLLB_STEP_TRACE (python -m llamole_b200.build --force); CTA 0 of every kernel stamps %globaltimer at
entry (1x), when its dependency wait returns (2x), first / last operand stage ready (40 / 50), accumulators ready (60), end (3x)."""
import os, sys, ctypes, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llamole_b200 import synth, _cabi

dev = torch.device("cuda", 0)
m, cfg, meta, sd = bench.build_dit(dev, small=False)
eng = m.engine()
lib = _cabi.lib()
N, T = m.max_n_nodes, 500
B = int(os.environ.get("PROBE_B", "6"))
props, txt = synth.dit_conditions(B, seed=1)
n_nodes = torch.randint(10, N + 1, (B,), dtype=torch.int32, generator=torch.Generator().manual_seed(B))
props = torch.where(props.to(dev) == -200.0, torch.full_like(props.to(dev), float("nan")), props.to(dev)).contiguous()
eng.begin(n_nodes, props, txt.to(dev).contiguous(), mol_index_base=0)
eng.init_state(7, None, None)
for i in range(5):
    eng.step(T - i, 7)
torch.cuda.synchronize()
buf = torch.zeros(1 + 3 * 20000, dtype=torch.int64, device=dev)
for name in ("llb_trace_install_runtime", "llb_trace_install_rowops", "llb_trace_install_dit"):
    fn = getattr(lib, name)
    fn.argtypes = [ctypes.c_void_p]
    assert fn(buf.data_ptr()) == 0
torch.cuda.synchronize()
eng.step(T - 6, 7)
eng.step(T - 7, 7)
torch.cuda.synchronize()
h = buf.cpu().numpy()
n = int(h[0])
rec = sorted([(int(h[3 + 3 * i]), int(h[1 + 3 * i]), int(h[2 + 3 * i])) for i in range(min(n, 20000))])
print("records", n)
names = {0xA: "attn", 0xB: "rowln", 1: "gemm64", 2: "gemm128", 4: "gemm256"}
t0 = rec[0][0]
# print the timeline of the second traced step's blocks 10..11
deps = [(t, tag, info) for (t, tag, info) in rec]
start = len(deps) // 2 + len(deps) // 4
for (t, tag, info) in (deps if os.environ.get("TRACE_ALL") else deps[start:start + 90]):
    kind, what = tag >> 4, tag & 15
    label = {1: "entry", 2: "dep-ok", 3: "end", 4: "stage0", 5: "stageL", 6: "acc-ok"}.get(kind, "?")
    who = names.get(what, "gemm") if kind in (1, 2, 3) else "gemm"
    print("%9.2f us  %-7s %-8s N=%d K=%d" % ((t - t0) / 1e3, label, who, info >> 32, info & 0xffffffff))
# SM clock from the (clock64, globaltimer) pairs of the first / last operand stage of the same CTA
pairs = [(t, info) for (t, tag, info) in rec if (tag >> 4) in (4, 5)]
fr = []
for (t0_, c0), (t1_, c1) in zip(pairs[::2], pairs[1::2]):
    if t1_ > t0_ and 0 < c1 - c0 < 10**7:
        fr.append((c1 - c0) / (t1_ - t0_))
if fr:
    fr.sort()
    print("SM clock from k-loop stamps: median %.3f GHz (min %.3f max %.3f) over %d k-loops" % (fr[len(fr) // 2], fr[0], fr[-1], len(fr)))
