"""Ad-hoc GPU probe (not part of the product): GraphDiT reverse-step latency at small batch sizes (BASELINE.json
configs[0] B=16 and configs[4] B=6), wall clock vs device time, to see whether the step is launch-bound."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llamole_b200 import synth, _cabi

dev = torch.device("cuda", 0)
m, cfg, meta, sd = bench.build_dit(dev, small=False)
eng = m.engine()
N, T = m.max_n_nodes, 500
for B in [int(b) for b in os.environ.get('PROBE_B', '6,16,64,256,1024').split(',')]:
    props, txt = synth.dit_conditions(B, seed=1)
    n_nodes = torch.randint(10, N + 1, (B,), dtype=torch.int32, generator=torch.Generator().manual_seed(B))
    props = props.to(dev)
    props = torch.where(props == -200.0, torch.full_like(props, float("nan")), props).contiguous()
    eng.begin(n_nodes, props, txt.to(dev).contiguous(), mol_index_base=0)
    eng.init_state(7, None, None)
    for i in range(5):
        eng.step(T - i, 7)
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    steps = int(os.environ.get("PROBE_STEPS", "40"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        eng.step(T - 5 - i, 7)
    t_issue = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("B=%4d tokens=%6d: %.3f ms/step device, %.3f ms/step wall, host issue %.3f ms/step, %d launches/step -> %.1f molecules/s" % (
        B, int(n_nodes.sum()), e0.elapsed_time(e1) / steps, wall / steps * 1e3, t_issue / steps * 1e3, (eng.launch_count() - l0) // steps,
        B / (T * wall / steps)), flush=True)
    if os.environ.get("PROBE_BREAKDOWN"):
        _cabi.profile_enable(True)
        for i in range(10):
            eng.step(T - 50 - i, 7)
        torch.cuda.synchronize()
        prof = _cabi.profile_read()
        _cabi.profile_enable(False)
        print("   ", {k: "%.0f us x %d" % (v[0] / v[1] * 1e3, v[1] / 10) for k, v in prof.items() if v[1]}, flush=True)
