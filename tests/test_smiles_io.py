"""Host-parallel edges of the path (llamole_b200/smiles_io.py) with picklable stand-in backends (RDKit is
not installed here): the parallel drivers return exactly what the serial backend call returns, in order, with None where the
backend gives up or a worker fails; the host inverse of the wire format matches `sharding.unpack_graphs`; the vectorised collate
matches a naive per-graph concatenation."""
from types import SimpleNamespace

import pytest
import torch

from llamole_b200 import sharding, smiles_io, synth


def fake_graph_to_smiles(molecule_list, atom_decoder):
    out = []
    for atoms, bonds in molecule_list:
        n = int(atoms.numel())
        if n % 5 == 0:
            out.append(None)                       # "could not be fixed"
            continue
        if n == 13:
            raise RuntimeError("worker crash")     # a whole chunk fails
        out.append("".join(atom_decoder[int(a)] for a in atoms) + f"|{int(bonds.sum())}")
    return out


def fake_smiles_to_graph(smiles):
    if smiles.startswith("bad"):
        return None
    n = len(smiles)
    x = torch.tensor([ord(c) % 118 for c in smiles], dtype=torch.long)
    src = torch.arange(n - 1)
    ei = torch.stack([torch.cat([src, src + 1]), torch.cat([src + 1, src])]) if n > 1 else torch.empty((2, 0), dtype=torch.long)
    ea = torch.ones(ei.shape[1], dtype=torch.long)
    return SimpleNamespace(x=x, edge_index=ei, edge_attr=ea)


def _sampled_like(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(1, N + 1, (B,), generator=g)
    valid = torch.arange(N)[None] < n[:, None]
    X = torch.where(valid, torch.randint(0, 16, (B, N), generator=g), torch.full((B, N), -1))
    E = torch.randint(0, 5, (B, N, N), generator=g)
    E = torch.triu(E, 1)
    E = E + E.transpose(1, 2)
    E = torch.where(valid[:, :, None] & valid[:, None, :], E, torch.full_like(E, -1))
    return X, E, n


def test_wire_to_molecule_list_inverts_pack_graphs():
    N = 9
    X, E, n = _sampled_like(17, N, seed=3)
    wire = sharding.pack_graphs(X, E, n)
    mols = smiles_io.wire_to_molecule_list(wire, N)
    Xu, Eu, nu = sharding.unpack_graphs(wire, N)
    assert torch.equal(Xu, X) and torch.equal(Eu, E) and torch.equal(nu, n)
    for b, (atoms, bonds) in enumerate(mols):
        k = int(n[b])
        assert torch.equal(atoms, X[b, :k]) and torch.equal(bonds, E[b, :k, :k]) and atoms.dtype == torch.int64


@pytest.mark.parametrize("workers,chunk", [(1, 8), (3, 8), (4, 5)])
def test_graphs_to_smiles_parallel_equals_serial(workers, chunk):
    N = 12
    X, E, n = _sampled_like(40, N, seed=5)
    n[n == 13] = 12
    mols = smiles_io.wire_to_molecule_list(sharding.pack_graphs(X, E, n, check=False), N)
    dec = [chr(ord("A") + i) for i in range(16)]
    want = fake_graph_to_smiles(mols, dec)
    got = smiles_io.graphs_to_smiles_parallel(mols, dec, backend=fake_graph_to_smiles, workers=workers, chunk=chunk)
    assert got == want and any(s is None for s in got) and any(isinstance(s, str) for s in got)


def test_a_failing_worker_yields_none_for_its_chunk_only():
    dec = [chr(ord("A") + i) for i in range(16)]
    mols = [[torch.zeros(k, dtype=torch.long), torch.zeros((k, k), dtype=torch.long)] for k in (3, 4, 13, 6, 7, 8)]
    got = smiles_io.graphs_to_smiles_parallel(mols, dec, backend=fake_graph_to_smiles, workers=2, chunk=2)
    assert got[0] is not None and got[1] is not None and got[2] is None and got[3] is None and got[4] is not None and got[5] is not None


def test_smiles_to_graphs_parallel_and_collate():
    smiles = ["CCO", "bad1", "c1ccccc1", "N", "bad2", "CC(=O)O"] * 9
    serial = [fake_smiles_to_graph(s) for s in smiles]
    par = smiles_io.smiles_to_graphs_parallel(smiles, fake_smiles_to_graph, workers=3, chunk=7)
    assert [g is None for g in par] == [g is None for g in serial]
    for a, b in zip(par, serial):
        if a is not None:
            assert torch.equal(a.x, b.x) and torch.equal(a.edge_index, b.edge_index) and torch.equal(a.edge_attr, b.edge_attr)
    batch, kept = smiles_io.collate_graphs(par)
    assert kept == [i for i, g in enumerate(serial) if g is not None] and batch.num_graphs == len(kept)
    off = 0
    for gi, i in enumerate(kept):
        g = serial[i]
        k = g.x.numel()
        assert torch.equal(batch.x[off:off + k], g.x) and bool((batch.batch[off:off + k] == gi).all())
        sel = (batch.batch[batch.edge_index[0]] == gi)
        assert torch.equal(batch.edge_index[:, sel] - off, g.edge_index) and torch.equal(batch.edge_attr[sel], g.edge_attr)
        off += k
    empty, kept0 = smiles_io.collate_graphs([None, None])
    assert empty.num_graphs == 0 and kept0 == [] and empty.edge_index.shape == (2, 0)


def test_collate_matches_the_synthetic_generator_layout():
    x, ei, ea, b = synth.molecular_graphs(7, seed=2, min_nodes=1, max_nodes=9)
    graphs = []
    for g in range(7):
        sel = b == g
        first = int(sel.nonzero()[0])
        es = sel[ei[0]]
        graphs.append(SimpleNamespace(x=x[sel], edge_index=ei[:, es] - first, edge_attr=ea[es]))
    batch, _ = smiles_io.collate_graphs(graphs)
    assert torch.equal(batch.x, x) and torch.equal(batch.edge_index, ei) and torch.equal(batch.edge_attr, ea) and torch.equal(batch.batch, b)
