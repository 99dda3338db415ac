"""torchrun worker of tests/test_integration_config5.py::test_sharded_over_all_gpus_equals_single_gpu (one rank per GPU, NCCL).

Every rank builds the same three modules from the same seeds, then
  1. samples a molecule batch through the PRODUCT sharding API (ConditionQueue -> sharding.sample_graphs_sharded -> packed wire
     format -> one NCCL all-gather), including a tail chunk smaller than the world size (empty shards);
  2. encodes / scores a graph batch through sharding.encode_graphs_sharded (node-balanced contiguous graph ranges);
  3. recomputes the whole batch locally on its own GPU and requires bit-identical integer graphs and identical embeddings /
     top-k -- "results identical for any GPU count"  across REAL ranks.
The kernel selection is pinned (LLB_FUSED_LN=0, LLB_SPLITK=0) because the sampler otherwise switches block-tail kernels with
the number of token rows per rank, and those differ in the last place of their fp32 rounding.
"""
import os
import sys
import tempfile

os.environ.setdefault("LLB_FUSED_LN", "0")
os.environ.setdefault("LLB_SPLITK", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from llamole_b200 import GraphCLIP, GraphDiT, GraphPredictor, sharding, synth  # noqa: E402
from llamole_b200.condition_queue import ConditionQueue  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = tempfile.mkdtemp()
    cfg = synth.dit_config(hidden=256, depth=2, heads=4, T=12)
    meta = synth.dit_meta(50)
    synth.write_dit_checkpoint(d, cfg, meta, synth.dit_state_dict(cfg, 50, seed=3))
    dit = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    dit.init_model(d)
    dit = dit.to(dev)

    # ---- 1. sampling through the queue, sharded over the ranks
    B = 4 * world + 1 + (1 if world > 2 else 0)      # the second chunk below is smaller than the world: empty shards
    props, txt = synth.dit_conditions(B, seed=11)
    q = ConditionQueue(dit, max_batch=4 * world, seed=5, group=None)
    t1 = q.submit(props[:3], txt[:3].bfloat16(), -200)          # bf16 text, as the LLM connector hands it over
    t2 = q.submit(props[3:], txt[3:].bfloat16(), -200)
    assert q.flush() == B
    Xa, Ea, na = q.result(t1)
    Xb, Eb, nb = q.result(t2)
    X, E, n = torch.cat([Xa, Xb]), torch.cat([Ea, Eb]), torch.cat([na, nb])
    # the same molecules on THIS GPU alone (no process group involved)
    Xl, El, nl = dit.generate_graphs(props, txt.bfloat16(), -200, n_nodes=n, seed=5, mol_index_base=0)
    assert torch.equal(X, Xl.cpu()) and torch.equal(E, El.cpu()) and torch.equal(n, nl.cpu()), "sharded sampling differs from single-GPU sampling"
    valid = torch.arange(X.shape[1])[None] < n[:, None]
    assert bool((X[valid] >= 0).all()) and bool((X[~valid] == -1).all())

    # ---- 2. encoder and predictor over node-balanced graph shards
    L, H, D, k = 3, 256, 512, 10
    enc, proj = synth.gin_encoder_state_dicts(L, H, seed=2)
    clip = GraphCLIP(L, H, 0.0, {})
    clip.molecule_encoder.load_state_dict(enc)
    clip.molecule_projection.load_state_dict(proj)
    clip = clip.to(dev)
    pred = GraphPredictor(L, H, 0.0, D, {}, {})
    pred.predictor.load_state_dict(synth.gin_predictor_state_dict(L, H, D, seed=4))
    pred = pred.to(dev)
    G = 5 * world + 3
    x, ei, ea, b = (t.to(dev) for t in synth.molecular_graphs(G, seed=8, min_nodes=1, max_nodes=40))
    c = synth.text_conditions(G, seed=9).to(dev)
    emb = sharding.encode_graphs_sharded(clip, x, ei, ea, b, num_graphs=G)
    emb_l = clip(x, ei, ea, b)
    assert emb.shape == emb_l.shape and torch.equal(emb, emb_l), "sharded encoder embeddings differ"

    def topk(x_, ei_, ea_, b_, c_):
        p, i = pred.topk_templates(x_, ei_, ea_, b_, c_, k)
        return torch.cat([p, i.float()], dim=1)

    tk = sharding.encode_graphs_sharded(topk, x, ei, ea, b, num_graphs=G, extra=c)
    tk_l = topk(x, ei, ea, b, c)
    assert torch.equal(tk, tk_l), "sharded top-k differs"
    # fewer graphs than ranks: some ranks own nothing and still take part in the gather
    x1, ei1, ea1, b1 = (t.to(dev) for t in synth.molecular_graphs(1, seed=3, min_nodes=5, max_nodes=5))
    e1 = sharding.encode_graphs_sharded(clip, x1, ei1, ea1, b1, num_graphs=1)
    assert torch.equal(e1, clip(x1, ei1, ea1, b1))
    dist.barrier()
    out = os.environ.get("LLB_MGPU_OUT")
    if out:
        open(os.path.join(out, f"ok_rank{rank}"), "w").write("ok")
    print(f"rank {rank}/{world}: sharded results identical to single-GPU results ({B} molecules, {G} graphs)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
