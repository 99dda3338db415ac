"""Condition queue in front of the GraphDiT sampler.

The reference enters `GraphDiT.generate` once per dataloader batch of 6 prompts (eval/workflow.py:110-124,
config/generate/*.yaml: per_device_eval_batch_size 6; modeling_llamole.py `design_molecule`), i.e. 600 token rows per
GEMM, where the denoiser is bound by streaming its 1.15 GB of weights.  Molecules are independent, so conditions from
many prompts can be sampled in ONE batch at compute-bound size.  `ConditionQueue` accumulates `(properties,
text_embedding)` requests, runs them in chunks of at most `max_batch` molecules (sharded over the ranks of the process
group when one is initialised) and hands each request its own slice of the result.

Determinism contract: molecule i of the queue (in submission order since construction) is sampled with the counter RNG
keyed by the GLOBAL index `index_base + i`, and its node count comes from a generator keyed the same way, so a
request's result does not depend on what else was queued with it, on `max_batch`, or on the number of GPUs.
This is bit-exact as long as the runs being compared use the same block-tail kernels: below 2048 token rows per rank the
sampler switches to its latency-regime kernels, whose fp32 rounding differs in the last place from
the fused throughput kernels, so a request sampled alone in a tiny batch and the same request inside a large one agree
except on decisions whose margin is at that rounding level (~1e-6 of the decisions).  Set LLB_FUSED_LN to pin one mode.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import sharding


@dataclass(frozen=True)
class Ticket:
    """Handle of one submitted request: molecules [start, start+count) of the queue's global numbering."""
    start: int
    count: int


class ConditionQueue:
    def __init__(self, model, max_batch: int = 2048, seed: int = 0, index_base: int = 0, group=None,
                 generate_fn: Optional[Callable] = None):
        """`model` is a llamole_b200.GraphDiT (anything with `generate_graphs`, `sample_n_nodes`-compatible `node_prob`
        and `max_n_nodes`); `generate_fn` overrides `model.generate_graphs` (tests)."""
        if max_batch < 1:
            raise ValueError("max_batch must be positive")
        self.model = model
        self.max_batch = int(max_batch)
        self.seed = int(seed)
        self.group = group
        self._gen = generate_fn if generate_fn is not None else model.generate_graphs
        self._next = int(index_base)
        self._pending: List[Tuple[Ticket, torch.Tensor, torch.Tensor, torch.Tensor]] = []
        self._done: Dict[Ticket, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    # ------------------------------------------------------------------ submission
    def _draw_n_nodes(self, start: int, count: int) -> torch.Tensor:
        """Node counts for molecules [start, start+count): one categorical draw per molecule from the node-count
        histogram (diffusion_utils.py:157-162) by inverse CDF on a counter-based uniform keyed by (seed, global index)."""
        prob = self.model.node_prob.detach().to("cpu", torch.float64)
        cdf = torch.cumsum(prob / prob.sum(), 0)
        with np.errstate(over="ignore"):
            z = (np.arange(start, start + count, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
            z += np.uint64(self.seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0xD1B54A32D192ED03)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)     # splitmix64 finaliser
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
        u = torch.from_numpy((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))
        last = int((prob > 0).nonzero().max())
        return torch.searchsorted(cdf, u, right=True).clamp_(max=last)

    def submit(self, properties: torch.Tensor, text_embedding: torch.Tensor, no_label_index: float = -200,
               n_nodes: Optional[torch.Tensor] = None) -> Ticket:
        """Queue `b` conditions: properties (b,10) with `no_label_index` or NaN for missing, text_embedding (b,768)."""
        if properties.dim() != 2 or text_embedding.dim() != 2 or properties.shape[0] != text_embedding.shape[0]:
            raise ValueError(f"properties {tuple(properties.shape)} / text_embedding {tuple(text_embedding.shape)}: expected (b,P) and (b,D)")
        b = int(properties.shape[0])
        ticket = Ticket(self._next, b)
        props = properties.detach().to("cpu", torch.float32)
        props = torch.where(props == no_label_index, torch.full_like(props, float("nan")), props)
        txt = text_embedding.detach().to("cpu", torch.float32)
        if n_nodes is None:
            n_nodes = self._draw_n_nodes(ticket.start, b)
        n_nodes = n_nodes.detach().to("cpu", torch.int64)
        if n_nodes.shape != (b,):
            raise ValueError(f"n_nodes {tuple(n_nodes.shape)}: expected ({b},)")
        if b and (int(n_nodes.min()) < 0 or int(n_nodes.max()) > int(self.model.max_n_nodes)):
            raise ValueError("n_nodes out of range")
        self._next += b
        if b:
            self._pending.append((ticket, props, txt, n_nodes))
        return ticket

    def pending(self) -> int:
        """Molecules queued and not yet sampled."""
        return sum(t.count for t, *_ in self._pending)

    # ------------------------------------------------------------------ execution
    def flush(self, steps: Optional[int] = None) -> int:
        """Sample everything that is pending; returns the number of molecules sampled.  Collective when a process
        group is initialised: every rank must hold the same queue and call flush()."""
        if not self._pending:
            return 0
        tickets = [p[0] for p in self._pending]
        props = torch.cat([p[1] for p in self._pending])
        txt = torch.cat([p[2] for p in self._pending])
        n_nodes = torch.cat([p[3] for p in self._pending])
        # global index of every pending molecule (tickets need not be contiguous after partial flushes)
        gidx = torch.cat([torch.arange(t.start, t.start + t.count) for t in tickets])
        total = int(props.shape[0])
        Xs, Es, ns = [], [], []
        pos = 0
        kw = {} if steps is None else {"steps": steps}
        while pos < total:
            # a chunk must be a run of consecutive global indices: the RNG key is mol_index_base + row
            end = min(total, pos + self.max_batch)
            brk = (gidx[pos + 1:end] - gidx[pos:end - 1] != 1).nonzero()
            if brk.numel():
                end = pos + int(brk[0]) + 1
            base = int(gidx[pos])

            def gen(p, t, n_nodes, seed, mol_index_base, _base=base, **k):
                return self._gen(p, t, float("nan"), n_nodes=n_nodes, seed=seed, mol_index_base=_base + mol_index_base, **k)

            # a failure here (out of memory, a collective error) propagates with the queue untouched: `_pending` is only
            # cleared once every chunk has been sampled, so the caller can retry flush() and no ticket is lost
            X, E, n = sharding.sample_graphs_sharded(gen, props[pos:end], txt[pos:end], n_nodes[pos:end], seed=self.seed,
                                                     group=self.group, max_nodes=int(self.model.max_n_nodes), **kw)
            Xs.append(X.cpu())
            Es.append(E.cpu())
            ns.append(n.cpu())
            pos = end
        X, E, n = torch.cat(Xs), torch.cat(Es), torch.cat(ns)
        self._pending = []      # only now: every chunk has been sampled (a failure above leaves the queue as it was)
        pos = 0
        for t in tickets:
            self._done[t] = (X[pos:pos + t.count], E[pos:pos + t.count], n[pos:pos + t.count])
            pos += t.count
        return total

    def result(self, ticket: Ticket, keep: bool = False):
        """(X (b,N), E (b,N,N), n_nodes (b,)) int64 host tensors of the request; flushes if it is still pending."""
        if ticket.count == 0:
            N = int(self.model.max_n_nodes)
            z = torch.zeros
            return z((0, N), dtype=torch.int64), z((0, N, N), dtype=torch.int64), z((0,), dtype=torch.int64)
        if ticket not in self._done:
            if not any(t == ticket for t, *_ in self._pending):
                raise KeyError(f"unknown or already collected ticket {ticket}")
            self.flush()
        return self._done[ticket] if keep else self._done.pop(ticket)

    def smiles(self, ticket: Ticket, backend=None, workers: Optional[int] = None) -> List[Optional[str]]:
        """The request's molecules as SMILES (None where the conversion fails, like `GraphDiT.generate`), converted by `workers`
        processes (`smiles_io.graphs_to_smiles_parallel`; backend = the reference's RDKit `graph_to_smiles` unless given)."""
        from .graph_decoder import _smiles_backend
        from .smiles_io import graphs_to_smiles_parallel

        return graphs_to_smiles_parallel(self.molecules(ticket), self.model.atom_decoder, backend=backend if backend is not None else _smiles_backend(),
                                         workers=workers)

    def molecules(self, ticket: Ticket) -> List[List[torch.Tensor]]:
        """The request's graphs as `[atom_types (n,), bond_types (n,n)]` pairs, the input of `graph_to_smiles`
        (diffusion_model.py:297-304)."""
        X, E, n = self.result(ticket)
        return [[X[i, :int(n[i])], E[i, :int(n[i]), :int(n[i])]] for i in range(X.shape[0])]
