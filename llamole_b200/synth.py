"""Synthetic checkpoints and inputs of the reference's shapes (there is no network for the real ones).

Everything here is seeded and independent of /root/reference: it writes the same *files* the
reference loaders read (`config.yaml`, `data.meta.json`, `model.pt`, `config.json`, `model_proj.pt`;
reference: src/model/loader.py:222-363) so that both the reference modules (in the authoring
container) and the drop-in classes of this package can be constructed from one directory.

State-dict keys / shapes follow SURVEY.md section 8b.  Two deliberate deviations from the reference's
`initialize_weights` (graph_decoder/transformer.py:66-84): the adaLN output layers and the virtual-node
embedding are drawn non-zero, otherwise every block is an identity and parity would be vacuous.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

X_CLASSES = 16
E_CLASSES = 5
Y_DIM = 10
TEXT_DIM = 768
ATOM_VOCAB = 118

PROPERTY_ORDER = ["BBBP", "HIV", "BACE", "CO2", "N2", "O2", "FFV", "TC", "SC", "SA"]
# data/property_ranges.json of the reference (min, max) -- used only to draw plausible magnitudes.
PROPERTY_RANGES = {
    "BBBP": (0.0, 1.0), "HIV": (0.0, 1.0), "BACE": (0.0, 1.0), "CO2": (0.94, 1019.265),
    "N2": (0.0, 73.417), "O2": (0.0, 122.94), "FFV": (0.324, 0.434), "TC": (0.117, 0.38),
    "SC": (1.0, 5.0), "SA": (1.0, 8.48),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _xavier(g, out_f, in_f):
    a = math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a


def _normal(g, *shape, std=0.02, mean=0.0):
    return torch.randn(*shape, generator=g) * std + mean


# ----------------------------------------------------------------------------------------------
# GraphDiT
# ----------------------------------------------------------------------------------------------
def dit_config(hidden=1024, depth=28, heads=16, mlp_ratio=4.0, T=500, guide_scale=2.0) -> dict:
    return {
        "diffusion_steps": T, "diffusion_noise_schedule": "cosine", "guide_scale": guide_scale,
        "hidden_size": hidden, "depth": depth, "num_heads": heads, "mlp_ratio": mlp_ratio,
        "drop_condition": 0.1, "lambda_train": [1, 10],
    }


def dit_meta(max_nodes=50, seed=7, min_nodes=5) -> dict:
    g = _gen(seed)
    active = sorted(torch.randperm(ATOM_VOCAB, generator=g)[:X_CLASSES].tolist())
    atom_dist = [0.0] * ATOM_VOCAB
    w = torch.rand(X_CLASSES, generator=g) + 0.05
    for k, a in enumerate(active):
        atom_dist[a] = float(w[k])
    n_dist = [0.0] * (max_nodes + 1)
    for n in range(min(min_nodes, max_nodes), max_nodes + 1):
        n_dist[n] = 1.0
    trans = (torch.rand(ATOM_VOCAB, ATOM_VOCAB, E_CLASSES, generator=g) + 0.01)
    symbols = [f"A{a}" for a in active]
    return {
        "active_atoms": symbols, "max_node": max_nodes, "n_atoms_per_mol_dist": n_dist,
        "bond_type_dist": [0.90, 0.05, 0.02, 0.005, 0.025], "transition_E": trans.tolist(),
        "atom_type_dist": atom_dist, "valencies": [0.0, 0.2, 0.3, 0.3, 0.2],
    }


def dit_state_dict(cfg: dict, max_nodes: int, seed=1234) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    H = cfg["hidden_size"]
    D = cfg["depth"]
    heads = cfg["num_heads"]
    dh = H // heads
    F = int(H * cfg["mlp_ratio"])
    d0 = X_CLASSES + E_CLASSES * max_nodes
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=True, small=False):
        sd[name + ".weight"] = _normal(g, out_f, in_f) if small else _xavier(g, out_f, in_f)
        if bias:
            sd[name + ".bias"] = _normal(g, out_f)

    def ln(name, n):
        sd[name + ".weight"] = _normal(g, n, std=0.1, mean=1.0)
        sd[name + ".bias"] = _normal(g, n, std=0.1)

    lin("x_embedder.0", H, d0, bias=False)
    ln("x_embedder.1", H)
    lin("t_embedder.mlp.0", H, 256)
    lin("t_embedder.mlp.2", H, H)
    sd["y_embedder.embedding_drop.weight"] = torch.randn(Y_DIM, H, generator=g)
    for d in range(Y_DIM):
        lin(f"y_embedder.mlps.{d}.0", H, 1)
        lin(f"y_embedder.mlps.{d}.2", H, H, bias=False)
    sd["txt_embedder.embedding_drop.weight"] = torch.randn(1, H, generator=g)
    lin("txt_embedder.linear", H, TEXT_DIM)
    for l in range(D):
        p = f"blocks.{l}."
        lin(p + "attn.qkv", 3 * H, H, bias=False)
        ln(p + "attn.q_norm", dh)
        ln(p + "attn.k_norm", dh)
        lin(p + "attn.proj", H, H)
        lin(p + "mlp.fc1", F, H)
        lin(p + "mlp.fc2", H, F)
        lin(p + "adaLN_modulation.0", H, H)
        lin(p + "adaLN_modulation.2", 6 * H, H, small=True)
    lin("output_layer.xedecoder.fc1", H, H)
    lin("output_layer.xedecoder.fc2", d0, H)
    lin("output_layer.adaLN_modulation.0", H, H)
    lin("output_layer.adaLN_modulation.2", 2 * d0, H, small=True)
    return sd


def write_dit_checkpoint(path: str, cfg: dict, meta: dict, sd: Optional[dict] = None, seed=1234) -> str:
    import yaml

    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    with open(os.path.join(path, "data.meta.json"), "w") as f:
        json.dump(meta, f)
    if sd is None:
        sd = dit_state_dict(cfg, meta["max_node"], seed)
    torch.save(sd, os.path.join(path, "model.pt"))
    return path


def dit_conditions(B: int, seed=2024, n_present=7) -> Tuple[torch.Tensor, torch.Tensor]:
    """properties (B,10) with -200 (= NO_LABEL_INDEX, extras/constants.py:24-25) for missing entries,
    and text embedding (B,768) = SiLU(N(0,1)) (the LLM connector ends in SiLU, modeling_llamole.py:211-214).

    "material" pattern: CO2..SA present, BBBP/HIV/BACE missing.
    """
    g = _gen(seed)
    props = torch.full((B, Y_DIM), -200.0)
    for d, name in enumerate(PROPERTY_ORDER):
        if d >= Y_DIM - n_present:
            lo, hi = PROPERTY_RANGES[name]
            props[:, d] = lo + (hi - lo) * torch.rand(B, generator=g)
    txt = torch.nn.functional.silu(torch.randn(B, TEXT_DIM, generator=g))
    return props, txt


# ----------------------------------------------------------------------------------------------
# GIN encoder / predictor
# ----------------------------------------------------------------------------------------------
def _gin_trunk(sd, g, L, H, affine_norms: bool):
    sd["atom_encoder.weight"] = torch.randn(ATOM_VOCAB, H, generator=g)
    sd["virtualnode_embedding.weight"] = _normal(g, 1, H, std=0.5)
    for l in range(L):
        p = f"convs.{l}."
        sd[p + "eps"] = _normal(g, 1, std=0.1)
        sd[p + "mlp.0.weight"] = _xavier(g, 4 * H, H)
        sd[p + "mlp.0.bias"] = _normal(g, 4 * H)
        sd[p + "mlp.1.weight"] = _normal(g, 4 * H, std=0.1, mean=1.0)
        sd[p + "mlp.1.bias"] = _normal(g, 4 * H, std=0.1)
        sd[p + "mlp.4.weight"] = _xavier(g, H, 4 * H)
        sd[p + "mlp.4.bias"] = _normal(g, H)
        sd[p + "bond_encoder.weight"] = torch.randn(5, H, generator=g)
        if affine_norms:
            sd[f"norms.{l}.weight"] = _normal(g, H, std=0.1, mean=1.0)
            sd[f"norms.{l}.bias"] = _normal(g, H, std=0.1)
        if l < L - 1:
            q = f"mlp_virtualnode_list.{l}."
            sd[q + "0.weight"] = _xavier(g, 4 * H, H)
            sd[q + "0.bias"] = _normal(g, 4 * H)
            sd[q + "1.weight"] = _normal(g, 4 * H, std=0.1, mean=1.0)
            sd[q + "1.bias"] = _normal(g, 4 * H, std=0.1)
            sd[q + "4.weight"] = _xavier(g, H, 4 * H)
            sd[q + "4.bias"] = _normal(g, H)


def gin_encoder_state_dicts(L=5, H=768, seed=11):
    g = _gen(seed)
    enc: Dict[str, torch.Tensor] = {}
    _gin_trunk(enc, g, L, H, affine_norms=True)
    proj = {
        "fc1.weight": _xavier(g, H, H), "fc1.bias": _normal(g, H),
        "norm1.weight": _normal(g, H, std=0.1, mean=1.0), "norm1.bias": _normal(g, H, std=0.1),
        "fc2.weight": _xavier(g, H, H), "fc2.bias": _normal(g, H),
    }
    return enc, proj


def gin_predictor_state_dict(L=5, H=768, out_dim=4096, text_dim=TEXT_DIM, seed=13):
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    _gin_trunk(sd, g, L, H, affine_norms=False)
    for l in range(L):
        sd[f"adapters.{l}.1.weight"] = _normal(g, 3 * H, text_dim, std=0.02)
        sd[f"adapters.{l}.1.bias"] = _normal(g, 3 * H, std=0.02)
    sd["text_dropping.weight"] = torch.randn(1, text_dim, generator=g)
    sd["decoder.0.weight"] = _xavier(g, 4 * H, H)
    sd["decoder.0.bias"] = _normal(g, 4 * H)
    sd["decoder.1.weight"] = _normal(g, 4 * H, std=0.1, mean=1.0)
    sd["decoder.1.bias"] = _normal(g, 4 * H, std=0.1)
    sd["decoder.4.weight"] = _xavier(g, out_dim, 4 * H)
    sd["decoder.4.bias"] = _normal(g, out_dim)
    return sd


def cost_mlp_state_dict(seed=17):
    g = _gen(seed)
    return {
        "layers.0.weight": _xavier(g, 128, 2048), "layers.0.bias": _normal(g, 128),
        "layers.3.weight": _xavier(g, 1, 128), "layers.3.bias": _normal(g, 1),
    }


def write_encoder_checkpoint(path: str, L=5, H=768, seed=11) -> str:
    os.makedirs(path, exist_ok=True)
    enc, proj = gin_encoder_state_dicts(L, H, seed)
    torch.save(enc, os.path.join(path, "model.pt"))
    torch.save(proj, os.path.join(path, "model_proj.pt"))
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({"num_layer": L, "hidden_size": H, "drop_ratio": 0.0}, f)
    return path


def write_predictor_checkpoint(path: str, L=5, H=768, out_dim=4096, seed=13, with_cost=True) -> str:
    os.makedirs(path, exist_ok=True)
    torch.save(gin_predictor_state_dict(L, H, out_dim, seed=seed), os.path.join(path, "model.pt"))
    if with_cost:
        torch.save(cost_mlp_state_dict(), os.path.join(path, "cost_model.pt"))
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({"num_layer": L, "hidden_size": H, "drop_ratio": 0.0, "num_task": out_dim}, f)
    return path


def molecular_graphs(num_graphs: int, seed=0, min_nodes=10, max_nodes=50):
    """Synthetic molecule-like graphs in the reference's PyG layout:
    random tree (parent among the previous 3 atoms) + n/8 ring closures, degree <= 4,
    x ~ U{0..117}, bond type ~ U{1..4}, both directions listed, `batch` sorted ascending.

    Returns x (sum_n,) int64, edge_index (2, sum_e) int64, edge_attr (sum_e,) int64, batch (sum_n,) int64.
    """
    rng = np.random.default_rng(seed)
    xs, srcs, dsts, eas, batches = [], [], [], [], []
    base = 0
    for gidx in range(num_graphs):
        n = int(rng.integers(min_nodes, max_nodes + 1))
        deg = np.zeros(n, dtype=np.int64)
        edges = set()
        for v in range(1, n):
            cands = [u for u in range(max(0, v - 3), v) if deg[u] < 4]
            if not cands:
                cands = [u for u in range(v) if deg[u] < 4]
            u = int(cands[int(rng.integers(len(cands)))])
            edges.add((u, v))
            deg[u] += 1
            deg[v] += 1
        for _ in range(n // 8):
            u, v = (int(t) for t in rng.integers(0, n, size=2))
            if u == v:
                continue
            a, b = min(u, v), max(u, v)
            if (a, b) in edges or deg[a] >= 4 or deg[b] >= 4:
                continue
            edges.add((a, b))
            deg[a] += 1
            deg[b] += 1
        xs.append(rng.integers(0, ATOM_VOCAB, size=n))
        for (a, b) in sorted(edges):
            t = int(rng.integers(1, 5))
            srcs += [a + base, b + base]
            dsts += [b + base, a + base]
            eas += [t, t]
        batches.append(np.full(n, gidx, dtype=np.int64))
        base += n
    x = torch.from_numpy(np.concatenate(xs).astype(np.int64))
    edge_index = torch.tensor([srcs, dsts], dtype=torch.int64)
    edge_attr = torch.tensor(eas, dtype=torch.int64)
    batch = torch.from_numpy(np.concatenate(batches))
    return x, edge_index, edge_attr, batch


def text_conditions(B: int, seed=5, text_dim=TEXT_DIM) -> torch.Tensor:
    return torch.nn.functional.silu(torch.randn(B, text_dim, generator=_gen(seed)))
