"""Property tests (hypothesis) of the multi-GPU host logic: every unit is owned by exactly one rank, graph ranges are
contiguous and cover the batch, the sub-batch split keeps edges inside their graphs, and the wire format is lossless."""
import pytest
import torch

pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from llamole_b200 import sharding, synth  # noqa: E402


@settings(max_examples=200, deadline=None)
@given(total=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_ranges_partition_the_batch(total, world):
    ranges = [sharding.shard_range(total, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=100, deadline=None)
@given(counts=st.lists(st.integers(1, 60), min_size=1, max_size=80), world=st.integers(1, 8))
def test_balanced_graph_ranges_are_contiguous_and_cover(counts, world):
    ranges = sharding.balanced_graph_ranges(counts, world)
    assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == len(counts)
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    # no rank holds much more than its share: at most the ideal share plus one graph
    total = sum(counts)
    for g0, g1 in ranges:
        assert sum(counts[g0:g1]) <= total / world + max(counts) + 1e-9


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10_000), graphs=st.integers(1, 12), world=st.integers(1, 4))
def test_split_graph_batch_keeps_graphs_whole(seed, graphs, world):
    x, ei, ea, b = synth.molecular_graphs(graphs, seed=seed, min_nodes=1, max_nodes=9)
    counts = torch.bincount(b, minlength=graphs).tolist()
    seen_nodes = seen_edges = 0
    for g0, g1 in sharding.balanced_graph_ranges(counts, world):
        sx, sei, sea, sb = sharding.split_graph_batch(x, ei, ea, b, g0, g1)
        seen_nodes += sx.numel()
        seen_edges += sea.numel()
        if sx.numel():
            assert int(sb.min()) == 0 and int(sb.max()) == g1 - g0 - 1
            if sea.numel():
                assert int(sei.min()) >= 0 and int(sei.max()) < sx.numel()
                assert torch.equal(sb[sei[0]], sb[sei[1]])      # edges stay inside their graph
    assert seen_nodes == x.numel() and seen_edges == ea.numel()


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10_000), B=st.integers(0, 6), N=st.integers(1, 20))
def test_wire_format_round_trip(seed, B, N):
    g = torch.Generator().manual_seed(seed)
    X = torch.randint(-1, 16, (B, N), generator=g)
    E = torch.randint(-1, 5, (B, N, N), generator=g)
    E = torch.triu(E) + torch.triu(E, 1).transpose(1, 2)
    n = torch.randint(0, N + 1, (B,), generator=g)
    w = sharding.pack_graphs(X, E, n)
    assert w.dtype == torch.uint8 and w.shape == (B, 2 + N + N * (N + 1) // 2)
    X2, E2, n2 = sharding.unpack_graphs(w, N)
    assert torch.equal(X, X2) and torch.equal(E, E2) and torch.equal(n, n2)
