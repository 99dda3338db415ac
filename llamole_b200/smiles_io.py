"""Host-parallel edges of the hot path: integer graphs -> SMILES after sampling, SMILES -> PyG batch
before encoding / scoring.

The chemistry itself is the reference's and stays on RDKit (`graph_decoder/molecule_utils.graph_to_smiles`: valency correction
with a `SanitizeMol` per bond, molecule_utils.py:49-162; `GraphLLMForCausalMLM.smiles_to_graph`, modeling_llamole.py:720-760):
neither is re-implemented here and RDKit is not a dependency of this package.  What the reference lacks is parallelism -- both run
as serial Python loops on the caller's thread, and at the batch sizes the B200 sampler produces (2048 molecules per 71 s of GPU time
per GPU, 16k per 8-GPU node) a serial RDKit loop becomes the pipeline's tail.  This module provides

  * `wire_to_molecule_list`   host-side (numpy) inverse of `sharding.pack_graphs`: the gathered byte rows -> the
                              `[atom_types (n,), bond_types (n,n)]` pairs `graph_to_smiles` consumes, without a round trip
                              through (B, N, N) int64 tensors;
  * `graphs_to_smiles_parallel`  order-preserving map of ANY `graph_to_smiles`-shaped backend over chunks of the molecule list
                              in a process pool (RDKit holds the GIL); a chunk whose worker fails yields None entries, which is
                              what the reference returns for molecules it cannot fix (the caller's rollback handles them);
  * `smiles_to_graphs_parallel`  the same for a `smiles_to_graph`-shaped backend, plus `collate_graphs`, a vectorised
                              `Batch.from_data_list` (node offsets by cumulative sum instead of per-graph Python work).

The backends are injected (or resolved lazily from the reference's modules when they are importable), exactly like
`graph_decoder.set_smiles_backend`; tests use picklable stand-ins.
"""
from __future__ import annotations

import os
from concurrent.futures import ProcessPoolExecutor
from types import SimpleNamespace
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def wire_to_molecule_list(wire, max_nodes: int) -> List[List[torch.Tensor]]:
    """`sharding.pack_graphs` rows (B, 2 + N + N(N+1)/2) uint8 (tensor or ndarray, any device) -> [[atoms (n,), bonds (n,n)], ...]
    int64 host tensors, masked entries dropped (the slice `[:n]` of diffusion_model.py:297-300)."""
    w = wire.detach().cpu().numpy() if torch.is_tensor(wire) else np.asarray(wire)
    N = int(max_nodes)
    if w.ndim != 2 or w.shape[1] != 2 + N + N * (N + 1) // 2:
        raise ValueError(f"wire rows of {w.shape[-1]} bytes do not match N={N}")
    n = w[:, 0].astype(np.int64) | (w[:, 1].astype(np.int64) << 8)
    X = w[:, 2:2 + N].astype(np.int64) - 1
    iu = np.triu_indices(N)
    tri = w[:, 2 + N:].astype(np.int64) - 1
    out = []
    for b in range(w.shape[0]):
        k = int(n[b])
        E = np.zeros((N, N), dtype=np.int64)
        E[iu] = tri[b]
        E = E + np.triu(E, 1).T
        out.append([torch.from_numpy(X[b, :k].copy()), torch.from_numpy(E[:k, :k].copy())])
    return out


def _default_graph_to_smiles() -> Callable:
    for mod in ("src.model.graph_decoder.molecule_utils", "graph_decoder.molecule_utils"):
        try:
            return __import__(mod, fromlist=["graph_to_smiles"]).graph_to_smiles
        except Exception:
            continue
    raise ImportError("no graph_to_smiles backend: pass backend= (the reference's graph_decoder/molecule_utils.graph_to_smiles needs RDKit)")


def _run_chunk(args):
    backend, chunk, atom_decoder = args
    return backend(chunk, atom_decoder)


def graphs_to_smiles_parallel(molecule_list: Sequence, atom_decoder: list, backend: Optional[Callable] = None,
                              workers: Optional[int] = None, chunk: int = 64, executor=None) -> List[Optional[str]]:
    """`backend(molecule_list, atom_decoder) -> List[Optional[str]]` (the reference's `graph_to_smiles`) over chunks of `chunk`
    molecules in `workers` processes; same list, same order, as the serial call.  `workers` <= 1 (or a list no longer than one
    chunk) runs inline.  A chunk whose worker raises gives None for its molecules.  `executor` reuses a caller-owned pool."""
    backend = backend if backend is not None else _default_graph_to_smiles()
    mols = list(molecule_list)
    if workers is None:
        workers = min(32, os.cpu_count() or 1)
    if executor is None and (workers <= 1 or len(mols) <= chunk):
        return list(backend(mols, atom_decoder))
    chunks = [mols[i:i + chunk] for i in range(0, len(mols), chunk)]
    own = executor is None
    ex = executor if executor is not None else ProcessPoolExecutor(max_workers=min(workers, len(chunks)))
    try:
        futures = [ex.submit(_run_chunk, (backend, c, atom_decoder)) for c in chunks]
        out: List[Optional[str]] = []
        for c, f in zip(chunks, futures):
            try:
                res = list(f.result())
                if len(res) != len(c):
                    raise ValueError(f"backend returned {len(res)} results for {len(c)} molecules")
            except Exception:
                res = [None] * len(c)      # the reference's contract for molecules it cannot convert
            out.extend(res)
        return out
    finally:
        if own:
            ex.shutdown()


def collate_graphs(graphs: Sequence, device=None) -> Tuple[SimpleNamespace, List[int]]:
    """Vectorised `Batch.from_data_list` for PyG-like objects (x (n,), edge_index (2,e), edge_attr (e,)); None entries (invalid
    SMILES) are skipped.  Returns (batch with x / edge_index / edge_attr / batch / num_graphs, indices of the graphs kept)."""
    kept = [i for i, g in enumerate(graphs) if g is not None]
    gs = [graphs[i] for i in kept]
    if not gs:
        z = torch.zeros((0,), dtype=torch.int64)
        return SimpleNamespace(x=z, edge_index=torch.zeros((2, 0), dtype=torch.int64), edge_attr=z.clone(), batch=z.clone(), num_graphs=0), kept
    n = torch.tensor([int(g.x.shape[0]) for g in gs], dtype=torch.int64)
    e = torch.tensor([int(g.edge_attr.shape[0]) for g in gs], dtype=torch.int64)
    off = torch.cumsum(n, 0) - n
    x = torch.cat([g.x.reshape(-1).to(torch.int64) for g in gs])
    ei = torch.cat([g.edge_index.to(torch.int64).reshape(2, -1) for g in gs], dim=1) + torch.repeat_interleave(off, e).unsqueeze(0)
    ea = torch.cat([g.edge_attr.reshape(-1).to(torch.int64) for g in gs])
    batch = torch.repeat_interleave(torch.arange(len(gs), dtype=torch.int64), n)
    out = SimpleNamespace(x=x, edge_index=ei, edge_attr=ea, batch=batch, num_graphs=len(gs))
    if device is not None:
        for k in ("x", "edge_index", "edge_attr", "batch"):
            setattr(out, k, getattr(out, k).to(device))
    return out, kept


def _map_chunk(args):
    backend, chunk = args
    return [backend(s) for s in chunk]


def smiles_to_graphs_parallel(smiles_list: Sequence[str], backend: Callable, workers: Optional[int] = None, chunk: int = 256,
                              executor=None) -> List[Optional[object]]:
    """`backend(smiles) -> PyG-like Data or None` (the reference's `smiles_to_graph`) over a process pool, order preserved;
    a failing chunk gives None for its molecules."""
    items = list(smiles_list)
    if workers is None:
        workers = min(32, os.cpu_count() or 1)
    if executor is None and (workers <= 1 or len(items) <= chunk):
        return [backend(s) for s in items]
    chunks = [items[i:i + chunk] for i in range(0, len(items), chunk)]
    own = executor is None
    ex = executor if executor is not None else ProcessPoolExecutor(max_workers=min(workers, len(chunks)))
    try:
        futures = [ex.submit(_map_chunk, (backend, c)) for c in chunks]
        out: List[Optional[object]] = []
        for c, f in zip(chunks, futures):
            try:
                res = list(f.result())
            except Exception:
                res = [None] * len(c)
            out.extend(res)
        return out
    finally:
        if own:
            ex.shutdown()
