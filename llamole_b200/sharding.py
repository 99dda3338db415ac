"""Multi-GPU use of the hot path: independent molecules / graphs are sharded over the ranks of one node and the
results are exchanged with ONE all-gather at the end.  There is no data-path collective:
weights are replicated, every unit of work is independent, and the counter RNG is keyed by the GLOBAL molecule
index so the sampled graphs do not depend on the number of GPUs.

One process per GPU (torchrun); `torch.distributed` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, stop) of `total` units for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def balanced_graph_ranges(nodes_per_graph: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous graph ranges with (nearly) equal node counts: GIN cost is proportional to nodes, and contiguous
    ranges keep every rank's CSR segments contiguous."""
    counts = torch.as_tensor(nodes_per_graph, dtype=torch.int64)
    G = int(counts.numel())
    prefix = torch.cumsum(counts, 0)
    total = int(prefix[-1]) if G else 0
    bounds = [0]
    pf = prefix.double()
    for r in range(1, world):
        target = total * r / world
        cut = int(torch.searchsorted(pf, torch.tensor(target, dtype=torch.float64), right=True).item())   # graphs with prefix <= target
        # the boundary whose node prefix is nearest to the target: after `cut` graphs or after one more
        if cut < G and (cut == 0 or float(pf[cut]) - target < target - float(pf[cut - 1])):
            cut += 1
        bounds.append(min(max(cut, bounds[-1]), G))
    bounds.append(G)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def split_graph_batch(x, edge_index, edge_attr, batch, g0: int, g1: int):
    """Sub-batch of graphs [g0, g1) of a PyG-style batch (batch sorted ascending, edges never cross graphs)."""
    node_sel = (batch >= g0) & (batch < g1)
    idx = node_sel.nonzero().squeeze(1)
    if idx.numel() == 0:
        z = x.new_zeros((0,))
        return z, edge_index.new_zeros((2, 0)), edge_attr.new_zeros((0,)), batch.new_zeros((0,))
    n0 = int(idx[0])
    e_sel = node_sel[edge_index[1]]
    return x[idx], edge_index[:, e_sel] - n0, edge_attr[e_sel], batch[idx] - g0


def all_gather_rows(t: torch.Tensor, group=None, sizes: Optional[Sequence[int]] = None) -> torch.Tensor:
    """All-gather along dim 0 for per-rank tensors whose first dimension may differ (pads to the max).  `sizes` = every
    rank's row count when the caller knows it (contiguous shards): saves the size exchange and its host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    if sizes is None:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        got = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got, n, group=group)
        sizes = [int(s.item()) for s in got]
    sizes = list(sizes)
    if len(sizes) != world or sizes[dist.get_rank(group)] != t.shape[0]:
        raise ValueError(f"all_gather_rows: sizes {sizes} do not describe this rank's {t.shape[0]} rows")
    m = max(sizes)
    pad = t.new_zeros((m,) + tuple(t.shape[1:]))
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


_TRIU_CACHE = {}


def _triu(N: int, device) -> torch.Tensor:
    """Flat indices i * N + j of the upper triangle (with diagonal) of an N x N matrix, cached per (N, device)."""
    key = (N, str(device))
    if key not in _TRIU_CACHE:
        iu = torch.triu_indices(N, N, device=device)
        _TRIU_CACHE[key] = iu[0] * N + iu[1]
    return _TRIU_CACHE[key]


def pack_graphs(X: torch.Tensor, E: torch.Tensor, n: torch.Tensor, check: bool = True) -> torch.Tensor:
    """Compact wire format of sampled graphs: per molecule one row of
    `2 + N + N(N+1)/2` bytes = node count (little-endian u16) | atom classes + 1 | upper triangle (with diagonal) of
    the bond classes + 1, so the masked value -1 travels as 0.  1327 B per molecule at N=50 instead of 20.4 KB of int64.
    E must be symmetric (the sampler mirrors the upper triangle, diffusion_utils.py:316-349).  Any integer dtype (the engine's
    int8 state or the int64 tensors of generate_graphs); `check=False` skips the symmetry / range validation, which costs three
    host synchronisations (for tensors that come straight from the sampler)."""
    B, N = X.shape
    if E.shape != (B, N, N) or n.shape != (B,):
        raise ValueError(f"pack_graphs: X {tuple(X.shape)}, E {tuple(E.shape)}, n {tuple(n.shape)}")
    if check and B:
        if not torch.equal(E, E.transpose(1, 2)):
            raise ValueError("pack_graphs: E is not symmetric")
        lo, hi = int(min(X.min(), E.min())), int(max(X.max(), E.max()))
        if lo < -1 or hi > 254:
            raise ValueError("pack_graphs: classes must lie in [-1, 254]")
    nn_ = n.to(torch.int32)
    row = torch.empty((B, 2 + N + N * (N + 1) // 2), dtype=torch.uint8, device=X.device)
    row[:, 0] = (nn_ & 0xFF).to(torch.uint8)
    row[:, 1] = (nn_ >> 8).to(torch.uint8)
    row[:, 2:2 + N] = (X.to(torch.int16) + 1).to(torch.uint8)
    row[:, 2 + N:] = (E.reshape(B, N * N).index_select(1, _triu(N, E.device)).to(torch.int16) + 1).to(torch.uint8)
    return row


def unpack_graphs(wire: torch.Tensor, N: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Inverse of pack_graphs: (X (B,N), E (B,N,N), n (B,)) int64."""
    B = wire.shape[0]
    if wire.shape[1] != 2 + N + N * (N + 1) // 2:
        raise ValueError(f"unpack_graphs: row of {wire.shape[1]} bytes does not match N={N}")
    n = wire[:, 0].to(torch.int64) | (wire[:, 1].to(torch.int64) << 8)
    X = wire[:, 2:2 + N].to(torch.int64) - 1
    tri = wire[:, 2 + N:].to(torch.int64) - 1
    flat = _triu(N, wire.device)
    E = torch.empty((B, N * N), dtype=torch.int64, device=wire.device)
    E[:, flat] = tri
    E[:, (flat % N) * N + flat // N] = tri
    return X, E.view(B, N, N), n


def _comm_device(group=None) -> torch.device:
    """Device the collectives of `group` need their tensors on (NCCL: this rank's GPU; gloo: host)."""
    if dist.is_available() and dist.is_initialized() and "nccl" in str(dist.get_backend(group)).lower():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sample_graphs_sharded(generate_fn: Callable, properties: torch.Tensor, text_embedding: torch.Tensor, n_nodes: torch.Tensor,
                          seed: int = 0, group=None, wire: str = "compact", max_nodes: Optional[int] = None, **kw):
    """Shard a sampling batch over the ranks and gather the integer graphs.

    generate_fn(properties, text_embedding, n_nodes=..., seed=..., mol_index_base=...) -> (X, E, n) as
    GraphDiT.generate_graphs.  Every rank passes the FULL batch and receives the FULL result.  The one exchange is an
    all-gather of the compact byte rows of pack_graphs (`wire="full"` gathers the int64 tensors instead).

    A batch smaller than the world size leaves some ranks with an EMPTY shard (the tail chunk of a ConditionQueue flush,
    e.g. 2050 molecules in chunks of 2048 on 8 GPUs): such a rank does not call generate_fn at all (the sampler has no B=0
    launch) and joins the all-gather with zero rows; it needs `max_nodes` (or a bound `GraphDiT.generate_graphs`, whose
    module carries it) to shape them.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    s, e = shard_range(properties.shape[0], rank, world)
    if e > s:
        X, E, n = generate_fn(properties[s:e], text_embedding[s:e], n_nodes=n_nodes[s:e], seed=seed, mol_index_base=s, **kw)
    else:
        N = max_nodes if max_nodes is not None else getattr(getattr(generate_fn, "__self__", None), "max_n_nodes", None)
        if N is None:
            raise ValueError("sample_graphs_sharded: this rank's shard is empty (batch < world size); pass max_nodes=")
        dev = _comm_device(group)
        X = torch.zeros((0, int(N)), dtype=torch.int64, device=dev)
        E = torch.zeros((0, int(N), int(N)), dtype=torch.int64, device=dev)
        n = torch.zeros((0,), dtype=torch.int64, device=dev)
    if world == 1:
        return X, E, n
    sizes = [b - a for a, b in (shard_range(properties.shape[0], r, world) for r in range(world))]
    if wire == "full":
        return all_gather_rows(X, group, sizes), all_gather_rows(E, group, sizes), all_gather_rows(n, group, sizes)
    Xg, Eg, ng = unpack_graphs(all_gather_rows(pack_graphs(X, E, n, check=False), group, sizes), X.shape[1])
    return Xg.to(X.dtype), Eg.to(E.dtype), ng.to(n.dtype)


def encode_graphs_sharded(forward_fn: Callable, x, edge_index, edge_attr, batch, num_graphs: Optional[int] = None, group=None,
                          extra: Optional[torch.Tensor] = None):
    """Shard a graph batch by node count, run forward_fn(x, edge_index, edge_attr, batch[, extra_rows]) on the local
    range and gather the per-graph rows (embeddings, logits or top-k) from all ranks."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    G = int(batch[-1].item()) + 1 if num_graphs is None else num_graphs
    counts = torch.bincount(batch, minlength=G)
    g0, g1 = balanced_graph_ranges(counts.tolist(), world)[rank]
    sub = split_graph_batch(x, edge_index, edge_attr, batch, g0, g1)
    if g1 > g0:
        out = forward_fn(*sub, extra[g0:g1]) if extra is not None else forward_fn(*sub)
    else:
        probe = forward_fn(*split_graph_batch(x, edge_index, edge_attr, batch, 0, 1), *([extra[0:1]] if extra is not None else []))
        out = probe[:0]
    return all_gather_rows(out, group)
