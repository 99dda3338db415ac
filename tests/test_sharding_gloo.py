"""world_size-2 gloo test of the multi-GPU host logic (sharding + the single all-gather).  The compute function is the
CPU oracle here (the GPU path cannot run in this container); what is under test is llamole_b200/sharding.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llamole_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import llamole_oracle as O

        torch.set_num_threads(1)
        L, H = 2, 64
        enc, proj = synth.gin_encoder_state_dicts(L, H, seed=3)
        x, ei, ea, b = synth.molecular_graphs(9, seed=4, min_nodes=1, max_nodes=12)
        fwd = lambda *g: O.gin_encoder_forward(enc, proj, L, *g)  # noqa: E731
        full = fwd(x, ei, ea, b)
        got = sharding.encode_graphs_sharded(fwd, x, ei, ea, b)
        ok_enc = torch.allclose(got, full, atol=1e-5) and got.shape == full.shape

        # sampling: a fake generate_fn whose output depends on the GLOBAL molecule index only
        def gen(props, txt, n_nodes, seed, mol_index_base):
            idx = torch.arange(props.shape[0]) + mol_index_base
            X = (idx[:, None] * 7 + torch.arange(5)[None] + seed) % 16
            E = (idx[:, None, None] + torch.arange(5)[None, :, None] + torch.arange(5)[None, None, :]) % 5
            return X, E, n_nodes

        B = 7
        props, txt = synth.dit_conditions(B)
        n_nodes = torch.arange(B) + 1
        X, E, n = sharding.sample_graphs_sharded(gen, props, txt, n_nodes, seed=3)
        Xf, Ef, nf = gen(props, txt, n_nodes, 3, 0)
        ok_dit = torch.equal(X, Xf) and torch.equal(E, Ef) and torch.equal(n, nf)

        # a batch SMALLER than the world: rank 1's shard is empty; it must not call the sampler (which has no B = 0 launch)
        # and must still join the all-gather with zero rows (this used to hang the other ranks)
        calls = []

        def gen1(props, txt, n_nodes, seed, mol_index_base):
            assert props.shape[0] >= 1, "the sampler must never be called on an empty shard"
            calls.append(props.shape[0])
            return gen(props, txt, n_nodes, seed, mol_index_base)

        for wire in ("compact", "full"):
            X1, E1, n1 = sharding.sample_graphs_sharded(gen1, props[:1], txt[:1], n_nodes[:1], seed=3, max_nodes=5, wire=wire)
            Xs, Es, ns = gen(props[:1], txt[:1], n_nodes[:1], 3, 0)
            ok_dit = ok_dit and torch.equal(X1, Xs) and torch.equal(E1, Es) and torch.equal(n1, ns)
        ok_dit = ok_dit and len(calls) == (2 if rank == 0 else 0)
        q.put((rank, ok_enc, ok_dit))
    finally:
        dist.destroy_process_group()


def test_sharded_paths_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok_enc and ok_dit for _, ok_enc, ok_dit in res), res


def test_partition_helpers():
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_range(2, 3, 4) == (2, 2)
    r = sharding.balanced_graph_ranges([10, 10, 10, 10, 40, 1, 1], 2)
    assert r[0][0] == 0 and r[-1][1] == 7 and r[0][1] == r[1][0]
    x, ei, ea, b = synth.molecular_graphs(5, seed=1, min_nodes=2, max_nodes=6)
    xs, eis, eas, bs = sharding.split_graph_batch(x, ei, ea, b, 2, 4)
    assert int(bs.min()) == 0 and int(bs.max()) == 1 and int(eis.max()) < xs.numel() and int(eis.min()) >= 0
    assert xs.numel() == int(((b >= 2) & (b < 4)).sum())
