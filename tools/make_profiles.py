"""Turns the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked evidence under profiles/.

    python tools/make_profiles.py <tag> <round-name>

Writes profiles/<round>_launches.md (ncu launch list aggregated per kernel + the live CUDA-event shares of the same
bench command), profiles/<round>_launches.csv.gz (the raw list), profiles/<round>_ncu_<what>.txt (one line per
captured launch of each --set full report) and profiles/<round>_bench.json (the bench line the numbers belong to).
"""
import csv, gzip, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("LLB_PROFILES_OUT", os.path.join(ROOT, "profiles"))   # on the GPU box: gpurun_out/profiles (only gpurun_out/ travels back)
os.makedirs(P, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("llb::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<.*>)?\(", name)
    if m:
        t = m.group(2) or ""
        t = re.sub(r"\(bool\)", "", t)
        t = re.sub(r"\(int\)", "", t)
        t = re.sub(r"\(unsigned int\)", "", t)
        return m.group(1) + t
    return name[:90]


def bench_line(path):
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)
    return None


# ---------------------------------------------------------------- launch list
lpath = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lpath):
    text = open(lpath).read()
    body = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(body)))
    # the capture runs the GraphDiT section first, then the GIN encoder and the predictor: split at the first GIN kernel so that the
    # shares of the first table are shares of the GraphDiT part, comparable with the live table below
    agg, agg_gin, in_gin = {}, {}, False
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        k = short(r["Kernel Name"])
        in_gin = in_gin or k.startswith("gin_")
        a = (agg_gin if in_gin else agg).setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    lib = lambda d: {k: v for k, v in d.items() if not k.startswith("at::") and "at::native" not in k and "cub::" not in k and "nccl" not in k.lower()}
    ours, ours_gin = lib(agg), lib(agg_gin)
    tot = sum(v[1] for v in ours.values())
    b = bench_line(os.path.join(G, f"bench_{tag}.json"))
    with open(os.path.join(P, f"{rnd}_launches.md"), "w") as f:
        f.write(f"# {rnd}: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency` (1 B200)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised, so only the\n"
                "SHARES are comparable with the live CUDA-event numbers of the un-profiled bench (second table). Raw list: "
                f"`{rnd}_launches.csv.gz`.\n\n")
        f.write("## GraphDiT section of the capture (set-up, warm-up and timed steps, e2e call, ragged batch)\n\n")
        f.write("| kernel (library kernels only) | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tot:.1f}% |\n")
        if ours_gin:
            tg = sum(v[1] for v in ours_gin.values())
            f.write("\n## GIN encoder + predictor section of the capture (as far as the launch cap reached)\n\n")
            f.write("| kernel (library kernels only) | launches | total ms | share |\n|---|---:|---:|---:|\n")
            for k, v in sorted(ours_gin.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tg:.1f}% |\n")
        others = sum(v[1] for k, v in list(agg.items()) + list(agg_gin.items()) if k not in ours and k not in ours_gin)
        f.write(f"\nPyTorch plumbing kernels (copies / fills of the bench's own set-up) in the same capture: {others / 1e3:.2f} ms.\n")
        if b:
            kb = b["kernel_breakdown"]
            t = sum(v["ms_per_step"] for v in kb.values())
            f.write("\n## Live CUDA-event shares inside the timed region of the un-profiled bench (GraphDiT step)\n\n")
            f.write("| slot | ms/step | launches/step | share |\n|---|---:|---:|---:|\n")
            for k, v in sorted(kb.items(), key=lambda kv: -kv[1]["ms_per_step"]):
                f.write(f"| {k} | {v['ms_per_step']:.2f} | {v['launches_per_step']:.0f} | {100 * v['ms_per_step'] / t:.1f}% |\n")
            f.write(f"\nstep = {b['ms_per_step']:.2f} ms, {b['value']:.2f} molecules/s, clocks {b['clocks']}\n")
            g = b["gin"]["kernel_breakdown"]
            t = sum(v["ms_per_forward"] for v in g.values())
            f.write("\n## GIN encoder forward (4096 graphs)\n\n| slot | ms/forward | launches | share |\n|---|---:|---:|---:|\n")
            for k, v in sorted(g.items(), key=lambda kv: -kv[1]["ms_per_forward"]):
                f.write(f"| {k} | {v['ms_per_forward']:.3f} | {v['launches']:.0f} | {100 * v['ms_per_forward'] / t:.1f}% |\n")
    with gzip.open(os.path.join(P, f"{rnd}_launches.csv.gz"), "wt") as f:
        f.write(body)

# ---------------------------------------------------------------- --set full reports
for what in ("dit_block", "dit_step", "gin", "head", "gemm", "attn", "ln"):
    rep = os.path.join(G, f"prof_{what}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{rnd}_ncu_{what}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, capture '{what}' of bench.py (tag {tag}); one line per launch\n")
        f.write("# time_us is under the profiler (cold cache, serialised); dram_* are per launch; tensor% = sm__pipe_tensor_cycles_active\n")
        f.write(out)

# ---------------------------------------------------------------- DRAM traffic per launch of the dominant kernels (bench.py reads it)
def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for d in rows[2:]:
        def val(key):
            i = hdr.index(key)
            v = float(d[i].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
        out.append((d[hdr.index("Kernel Name")], val("dram__bytes_read.sum") + val("dram__bytes_write.sum")))
    return out


traffic = {}
rep = os.path.join(G, f"prof_dit_block_{tag}.ncu-rep")
if os.path.exists(rep):
    prev = None
    for name, byts in raw_rows(rep):
        slot = None
        if "EpiQKV" in name: slot = "gemm_qkv"
        elif "attention" in name: slot = "attention"
        elif "row_ln" in name: slot = "ln_mod_res"
        elif "gemm_ln" in name: slot = "gemm_proj" if prev == "attention" else "gemm_fc2"   # fused GEMM + LayerNorm tails
        elif "EpiBiasAct<1" in name: slot = "gemm_fc1"
        elif "EpiBiasAct<0" in name: slot = "gemm_fc2" if prev == "gemm_fc1" else "gemm_proj"
        prev = slot
        if slot:
            traffic.setdefault(slot, []).append(byts)
rep = os.path.join(G, f"prof_gin_{tag}.ncu-rep")
if os.path.exists(rep):
    for name, byts in raw_rows(rep):
        if "gin_aggregate" in name:
            traffic.setdefault("gin_aggregate", []).append(byts)
        elif "gemm_ln_pair_kernel" in name:
            traffic.setdefault("gin_tail", []).append(byts)
        elif "EpiLnGelu" in name:
            traffic.setdefault("gin_mlp0", []).append(byts)
        elif "EpiRowSq" in name:
            traffic.setdefault("gin_stats", []).append(byts)
rep = os.path.join(G, f"prof_head_{tag}.ncu-rep")
if os.path.exists(rep):
    for name, byts in raw_rows(rep):
        if "EpiHeadTopk" in name:
            traffic.setdefault("head_topk", []).append(byts)
if traffic:
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump({"source": f"{rnd}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over captured launches)",
                   "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, f, indent=1)

for name in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        b = bench_line(src)
        if b:
            with open(os.path.join(P, f"{rnd}_{name.replace('_' + tag, '')}"), "w") as f:
                json.dump(b, f, indent=1)
src = os.path.join(G, f"parity_{tag}.json")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"parity_{rnd}.json"))
print("\n".join(sorted(os.listdir(P))))
