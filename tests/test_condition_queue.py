"""Condition queue and the compact graph wire format (section 8e): host logic on CPU, with a
stand-in sampler whose output depends only on the GLOBAL molecule index and the molecule's own condition -- the
property the real sampler has (counter RNG keyed by global index; tested on the GPU in test_gpu_parity.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llamole_b200 import sharding, synth
from llamole_b200.condition_queue import ConditionQueue, Ticket

N = 6


class _FakeDit:
    max_n_nodes = N
    node_prob = torch.tensor([0.0, 0.1, 0.2, 0.3, 0.2, 0.1, 0.1])

    def __init__(self):
        self.calls = []

    def generate_graphs(self, props, txt, no_label_index=-200, n_nodes=None, noise=None, seed=0, steps=None, mol_index_base=0):
        B = props.shape[0]
        self.calls.append((B, mol_index_base))
        idx = torch.arange(B) + mol_index_base
        key = idx + torch.nan_to_num(props, nan=3.0).sum(1).round().long() + (txt.sum(1) * 10).round().long() + seed
        ar = torch.arange(N)
        X = (key[:, None] * 7 + ar[None]) % 16
        E = (key[:, None, None] + ar[None, :, None] * ar[None, None, :]) % 5
        valid = ar[None] < n_nodes[:, None]
        X = torch.where(valid, X, -1)
        E = torch.where(valid[:, :, None] & valid[:, None, :], E, -1)
        return X, E, n_nodes


def _requests(sizes, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in sizes:
        props = torch.randint(0, 5, (b, 10), generator=g).float()
        props[:, :3] = -200.0
        out.append((props, torch.randn(b, 8, generator=g)))
    return out


def test_queue_results_do_not_depend_on_batching():
    reqs = _requests([2, 5, 1, 0, 7, 3])
    ref = None
    for max_batch in (1, 4, 6, 2048):
        m = _FakeDit()
        q = ConditionQueue(m, max_batch=max_batch, seed=5)
        tickets = [q.submit(p, t) for p, t in reqs]
        assert q.pending() == 18 and [t.start for t in tickets] == [0, 2, 7, 8, 8, 15]
        assert q.flush() == 18 and q.pending() == 0
        assert all(b <= max_batch for b, _ in m.calls) and sum(b for b, _ in m.calls) == 18
        res = [q.result(t) for t in tickets]
        assert res[3][0].shape == (0, N) and res[3][1].shape == (0, N, N)
        if ref is None:
            ref = res
            # every request equals a direct call with its own global index base and node counts
            for (p, t), tk, r in zip(reqs, tickets, ref):
                if tk.count == 0:
                    continue
                pn = torch.where(p == -200.0, float("nan"), p)
                X, E, n = _FakeDit().generate_graphs(pn, t, n_nodes=r[2], seed=5, mol_index_base=tk.start)
                assert torch.equal(X, r[0]) and torch.equal(E, r[1])
        else:
            for a, b in zip(ref, res):
                assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_queue_partial_flush_and_lazy_result():
    reqs = _requests([3, 4, 2], seed=1)
    q_all = ConditionQueue(_FakeDit(), max_batch=64, seed=2)
    t_all = [q_all.submit(p, t) for p, t in reqs]
    want = [q_all.result(t) for t in t_all]
    m = _FakeDit()
    q = ConditionQueue(m, max_batch=64, seed=2)
    t0 = q.submit(*reqs[0])
    got0 = q.result(t0)              # lazy flush of one request
    t1 = q.submit(*reqs[1])
    t2 = q.submit(*reqs[2])
    got2 = q.result(t2, keep=True)   # flushes t1 and t2 together
    assert m.calls == [(3, 0), (6, 3)]
    got1 = q.result(t1)
    for a, b in zip(want, [got0, got1, got2]):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(q.result(t2)[0], got2[0])
    with pytest.raises(KeyError):
        q.result(t2)
    with pytest.raises(KeyError):
        q.result(Ticket(100, 1))
    mols = None
    t3 = q.submit(*reqs[0], n_nodes=torch.tensor([1, 6, 0]))
    mols = q.molecules(t3)
    assert [m_[0].shape[0] for m_ in mols] == [1, 6, 0] and mols[1][1].shape == (6, 6)


def test_queue_argument_checks():
    q = ConditionQueue(_FakeDit())
    p, t = _requests([2])[0]
    with pytest.raises(ValueError):
        q.submit(p, t[:1])
    with pytest.raises(ValueError):
        q.submit(p, t, n_nodes=torch.tensor([1, N + 1]))
    with pytest.raises(ValueError):
        q.submit(p, t, n_nodes=torch.tensor([1]))
    with pytest.raises(ValueError):
        ConditionQueue(_FakeDit(), max_batch=0)
    assert q.pending() == 0 and q.flush() == 0


def test_node_counts_follow_histogram_and_global_index():
    m = _FakeDit()
    a = ConditionQueue(m, seed=9)._draw_n_nodes(0, 400)
    b = torch.cat([ConditionQueue(m, seed=9)._draw_n_nodes(0, 150), ConditionQueue(m, seed=9)._draw_n_nodes(150, 250)])
    assert torch.equal(a, b)
    assert int(a.min()) >= 1 and int(a.max()) <= N
    freq = torch.bincount(a, minlength=N + 1).float() / 400
    assert float((freq - m.node_prob).abs().max()) < 0.08
    assert not torch.equal(a, ConditionQueue(m, seed=10)._draw_n_nodes(0, 400))


# ------------------------------------------------------------------------------------------------ wire format
def test_wire_format_round_trip_on_golden_graphs(dit_small):
    fx = dit_small
    X = fx["final_X"].long()
    E = fx["final_E"].long()
    n = fx["n_nodes"].long()
    w = sharding.pack_graphs(X, E, n)
    Nn = X.shape[1]
    assert w.dtype == torch.uint8 and w.shape == (X.shape[0], 2 + Nn + Nn * (Nn + 1) // 2)
    X2, E2, n2 = sharding.unpack_graphs(w, Nn)
    assert torch.equal(X, X2) and torch.equal(E, E2) and torch.equal(n, n2)


def test_wire_format_edge_cases():
    X = torch.full((3, 50), -1)
    E = torch.full((3, 50, 50), -1)
    n = torch.tensor([0, 50, 40000])
    X[1] = 15
    E[1] = 4
    w = sharding.pack_graphs(X, E, n)
    assert w.shape == (3, 1327)
    X2, E2, n2 = sharding.unpack_graphs(w, 50)
    assert torch.equal(X, X2) and torch.equal(E, E2) and torch.equal(n, n2)
    e = sharding.pack_graphs(X[:0], E[:0], n[:0])
    assert e.shape == (0, 1327) and sharding.unpack_graphs(e, 50)[1].shape == (0, 50, 50)
    E[1, 0, 1] = 0
    with pytest.raises(ValueError):
        sharding.pack_graphs(X, E, n)      # not symmetric
    with pytest.raises(ValueError):
        sharding.unpack_graphs(w[:, :-1], 50)
    with pytest.raises(ValueError):
        sharding.pack_graphs(X + 300, E.transpose(1, 2).clone().fill_(0), n)


# ------------------------------------------------------------------------------------------------ world-size 2
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        reqs = _requests([2, 5, 1, 0, 7, 3])
        m = _FakeDit()
        q = ConditionQueue(m, max_batch=5, seed=5)
        tickets = [q.submit(p, t) for p, t in reqs]
        q.flush()
        res = [q.result(t) for t in tickets]
        single = ConditionQueue(_FakeDit(), max_batch=2048, seed=5, generate_fn=None)
        # reference result without the process group: every molecule through one un-sharded call
        Xs, Es, ns = [], [], []
        for (p, t), tk, r in zip(reqs, tickets, res):
            if tk.count:
                X, E, n = _FakeDit().generate_graphs(torch.where(p == -200.0, float("nan"), p), t, n_nodes=r[2], seed=5,
                                                     mol_index_base=tk.start)
                Xs.append(torch.equal(X, r[0]) and torch.equal(E, r[1]) and torch.equal(n, r[2]))
        del single
        out.put((rank, all(Xs), max(b for b, _ in m.calls)))
    finally:
        dist.destroy_process_group()


def test_queue_sharded_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[0] for g in got] == [0, 1]
    assert all(g[1] for g in got)
    assert all(g[2] <= 3 for g in got)      # chunks of 5 molecules are split over the two ranks


def test_flush_failure_keeps_the_queue():
    """A sampler failure in the middle of a flush (out of memory, a collective error) must not lose any request: the
    queue is untouched and a later flush delivers everything."""
    fake = _FakeDit()
    boom = {"left": 1}

    def flaky(*a, **k):
        if boom["left"] > 0 and len(fake.calls) == 1:   # fail on the SECOND chunk of the first flush
            boom["left"] -= 1
            raise RuntimeError("CUDA out of memory (simulated)")
        return fake.generate_graphs(*a, **k)

    q = ConditionQueue(fake, max_batch=4, seed=1, generate_fn=flaky)
    reqs = _requests([3, 4, 2])
    tickets = [q.submit(p, t) for p, t in reqs]
    with pytest.raises(RuntimeError):
        q.flush()
    assert q.pending() == 9
    assert q.flush() == 9 and q.pending() == 0
    ref = ConditionQueue(_FakeDit(), max_batch=4, seed=1)
    rt = [ref.submit(p, t) for p, t in reqs]
    for a, b in zip(tickets, rt):
        for u, v in zip(q.result(a), ref.result(b)):
            assert torch.equal(u, v)


# ------------------------------------------------------------------------------------------------ property test
pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(sizes=st.lists(st.integers(0, 9), min_size=1, max_size=8), max_batch=st.integers(1, 12),
       flush_after=st.lists(st.booleans(), min_size=8, max_size=8), seed=st.integers(0, 50))
def test_queue_any_interleaving_of_submit_and_flush(sizes, max_batch, flush_after, seed):
    """Whatever the request sizes, chunk size and flush points, every request gets exactly what a direct call with its
    own global index base and node counts returns, and no sampler call exceeds max_batch."""
    reqs = _requests(sizes, seed=seed)
    m = _FakeDit()
    q = ConditionQueue(m, max_batch=max_batch, seed=seed)
    tickets = []
    for i, (p, t) in enumerate(reqs):
        tickets.append(q.submit(p, t))
        if flush_after[i]:
            q.flush()
    base = 0
    for (p, t), tk in zip(reqs, tickets):
        assert tk.start == base and tk.count == p.shape[0]
        base += p.shape[0]
        X, E, n = q.result(tk)
        assert X.shape[0] == tk.count
        if tk.count:
            Xd, Ed, _ = _FakeDit().generate_graphs(torch.where(p == -200.0, float("nan"), p), t, n_nodes=n, seed=seed, mol_index_base=tk.start)
            assert torch.equal(X, Xd) and torch.equal(E, Ed)
            assert torch.equal(n, ConditionQueue(_FakeDit(), seed=seed)._draw_n_nodes(tk.start, tk.count))
    assert q.pending() == 0
    assert all(b <= max_batch for b, _ in m.calls) and sum(b for b, _ in m.calls) == sum(sizes)
