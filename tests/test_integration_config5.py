"""BASELINE.json configs[4] substitute: the end-to-end qwen_material eval cannot run here (no
Qwen2 weights / peft / rdkit / PyG, no network), so a STUB LLM drives the three drop-in classes with exactly the calls,
shapes, dtypes and devices `GraphLLMForCausalMLM` makes:

  design_molecule     (modeling_llamole.py:585-657): graph_encoder(batch.x, .edge_index, .edge_attr, .batch) -> (n_mol, H);
                      graph_decoder.generate(properties (6,10) bf16 with -200, design_hidden (6,768) bf16 = SiLU(Linear(.)), -200)
  one_step_reaction   (:784-889): graph_encoder on a PyG Batch; graph_predictor.sample_templates(Data, retro_hidden (1,768) bf16,
                      product_smiles, topk)
  loader              (loader.py:245-247): every parameter cast in place to compute_dtype = bf16, module moved to the device.

Run with -m gpu.  The multi-GPU half launches torchrun over every visible GPU (skipped below 2) and checks that molecule
batches sharded through llamole_b200.sharding / ConditionQueue over NCCL reproduce the single-GPU results bit for bit.
"""
import os
import subprocess
import sys
import tempfile
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from llamole_b200 import GraphCLIP, GraphDiT, GraphPredictor, synth  # noqa: E402
from llamole_b200.graph_decoder import set_smiles_backend  # noqa: E402
from llamole_b200.graph_predictor import set_template_backend  # noqa: E402

DEV = "cuda:0"
NO_LABEL_INDEX = -200   # extras/constants.py:24


class PyGLike(SimpleNamespace):
    """Stand-in for torch_geometric.data.Data / Batch: attribute access + .to(device), which is all the callers use."""

    def to(self, device):
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class StubLLM(nn.Module):
    """Produces last-layer hidden states of the right shape / dtype (the real one is a HF causal LM in bf16)."""

    def __init__(self, hidden=896, vocab=1000):
        super().__init__()
        self.embed = nn.Embedding(vocab, hidden)
        self.mix = nn.Linear(hidden, hidden)

    def forward(self, input_ids):
        return torch.tanh(self.mix(self.embed(input_ids)))


def _cast_like_the_loader(module, dtype, device):
    for p in module.parameters():            # loader.py:245-247
        p.data = p.data.to(dtype)
    return module.to(device)


@pytest.fixture(scope="module")
def stack():
    torch.manual_seed(0)
    d = tempfile.mkdtemp()
    cfg = synth.dit_config(hidden=256, depth=2, heads=4, T=20)
    meta = synth.dit_meta(50)
    synth.write_dit_checkpoint(os.path.join(d, "dit"), cfg, meta, synth.dit_state_dict(cfg, 50, seed=3))
    dit = GraphDiT(os.path.join(d, "dit", "config.yaml"), os.path.join(d, "dit", "data.meta.json"), torch.bfloat16)
    dit.init_model(os.path.join(d, "dit"))
    dit.disable_grads()
    enc_dir = synth.write_encoder_checkpoint(os.path.join(d, "enc"), 3, 256, 5)
    clip = GraphCLIP(3, 256, 0.0, {})
    clip.init_model(enc_dir, verbose=False)
    clip.disable_grads()
    D = 333
    pred_dir = synth.write_predictor_checkpoint(os.path.join(d, "pred"), 3, 256, D, 6)
    pred = GraphPredictor(3, 256, 0.0, D, {"text_input_size": 768}, {i: f"T{i}" for i in range(D)}, ["C", "CC", "O"])
    pred.init_model(pred_dir)
    pred.init_neural_cost(pred_dir)
    pred.disable_grads()
    llm = StubLLM().to(torch.bfloat16).to(DEV)
    conn = SimpleNamespace(
        graph_to_lm=nn.Sequential(nn.Linear(256, 896), nn.SiLU()).to(torch.bfloat16).to(DEV),          # modeling_llamole.py:205-208
        lm_to_decoder=nn.Sequential(nn.Linear(896, dit.text_input_size), nn.SiLU()).to(torch.bfloat16).to(DEV),   # :211-214
        lm_to_predictor=nn.Sequential(nn.Linear(896, pred.text_input_size), nn.SiLU()).to(torch.bfloat16).to(DEV))
    mods = [_cast_like_the_loader(m, torch.bfloat16, DEV) for m in (dit, clip, pred)]
    return SimpleNamespace(dit=mods[0], clip=mods[1], pred=mods[2], llm=llm, conn=conn, D=D)


def _batch(n_graphs, seed):
    x, ei, ea, b = synth.molecular_graphs(n_graphs, seed=seed, min_nodes=3, max_nodes=28)
    return PyGLike(x=x, edge_index=ei, edge_attr=ea, batch=b).to(DEV)


def test_design_molecule_shaped_calls(stack):
    s = stack
    B = 6                                   # config/generate/qwen_material.yaml: per_device_eval_batch_size 6
    graphs = _batch(B, seed=1)
    with torch.no_grad():
        mol_embeds = s.clip(graphs.x, graphs.edge_index, graphs.edge_attr, graphs.batch)
        assert mol_embeds.shape == (B, s.clip.hidden_size) and mol_embeds.dtype == torch.bfloat16 and mol_embeds.device.type == "cuda"
        assert torch.allclose(mol_embeds.float().norm(dim=-1), torch.ones(B, device=DEV), atol=2e-2)
        lm_in = s.conn.graph_to_lm(mol_embeds)                                   # consumed by the LLM as input embeddings
        assert lm_in.dtype == torch.bfloat16
        ids = torch.randint(0, 1000, (B, 24), device=DEV)
        design_hidden = s.llm(ids)[:, -8:].mean(dim=1)                           # 8 query tokens, mean-pooled (:644)
        design_hidden = s.conn.lm_to_decoder(design_hidden)
        props, _ = synth.dit_conditions(B, seed=5)
        molecule_properties = props.to(DEV).type_as(design_hidden)               # bf16, -200 for missing (:652)
        assert design_hidden.shape == (B, 768) and design_hidden.dtype == torch.bfloat16 and molecule_properties.dtype == torch.bfloat16
    seen = {}

    def smiles_backend(molecule_list, atom_decoder):
        seen["n"] = len(molecule_list)
        out = []
        for atoms, bonds in molecule_list:
            assert atoms.dtype == torch.int64 and bonds.dtype == torch.int64 and atoms.device.type == "cpu"
            assert bonds.shape == (atoms.numel(), atoms.numel()) and int(atoms.min()) >= 0 and int(atoms.max()) < len(atom_decoder)
            assert torch.equal(bonds, bonds.t()) and int(bonds.min()) >= 0 and int(bonds.max()) < 5
            out.append(None if atoms.numel() % 7 == 0 else "".join(atom_decoder[int(a)] for a in atoms))   # None = invalid -> rollback
        return out

    set_smiles_backend(smiles_backend)
    try:
        torch.manual_seed(1)
        smiles = s.dit.generate(molecule_properties, design_hidden, NO_LABEL_INDEX)
        torch.manual_seed(1)
        again = s.dit.generate(molecule_properties, design_hidden, NO_LABEL_INDEX)
        other = s.dit.generate(molecule_properties, design_hidden, NO_LABEL_INDEX)
    finally:
        set_smiles_backend(None)
    assert seen["n"] == B and len(smiles) == B and all(x is None or isinstance(x, str) for x in smiles)
    assert smiles == again, "torch.manual_seed makes generate reproducible"
    assert smiles != other, "consecutive calls draw fresh noise (reference: multinomial on the global generator)"


def test_one_step_reaction_shaped_calls(stack):
    s = stack
    with torch.no_grad():
        x, ei, ea, _ = synth.molecular_graphs(1, seed=9, min_nodes=17, max_nodes=17)
        product_graph = PyGLike(x=x, edge_index=ei, edge_attr=ea).to(DEV)       # smiles_to_graph output (:720-760)
        all_graphs = _batch(3, seed=2)                                           # PyGBatch.from_data_list(previous + [product])
        mol_embeds = s.clip(all_graphs.x, all_graphs.edge_index, all_graphs.edge_attr, all_graphs.batch)
        assert mol_embeds.shape == (3, 256) and mol_embeds.dtype == torch.bfloat16
        retro_hidden = s.conn.lm_to_predictor(s.llm(torch.randint(0, 1000, (1, 40), device=DEV))[:, -8:].mean(dim=1))
        assert retro_hidden.shape == (1, 768) and retro_hidden.dtype == torch.bfloat16

    def run_template(template, smiles):
        k = int(template[1:])
        return [] if k % 4 == 0 else [f"C{k % 3}.O", f"O.C{k % 3}"]

    set_template_backend(run_template)
    try:
        reactants, scores, templates = s.pred.sample_templates(product_graph, retro_hidden, "CCO", 50)
        # the logits the reference returns for the same call, in bf16 like its parameters
        logits = s.pred(product_graph.x, product_graph.edge_index, product_graph.edge_attr,
                        torch.zeros(17, dtype=torch.long, device=DEV), retro_hidden)
    finally:
        set_template_backend(None)
    assert logits.shape == (1, s.D) and logits.dtype == torch.bfloat16
    assert len(reactants) == len(scores) == len(templates) > 0
    assert abs(sum(scores) - 1.0) < 1e-5 and scores == sorted(scores, reverse=True)
    assert all(isinstance(r, str) for r in reactants) and all(t.startswith("T") for t in templates)
    cost = s.pred.cost_from_fingerprints((torch.rand(2, 2048) > 0.9).float())
    assert cost.shape == (2, 1) and bool((cost > 0).all())


def test_sharded_over_all_gpus_equals_single_gpu():
    """torchrun over every visible GPU: ConditionQueue / sample_graphs_sharded / encode_graphs_sharded over NCCL give, on every
    rank, exactly what one GPU gives for the whole batch (RNG keyed by the global molecule index)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu_worker.py")
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, LLB_MGPU_OUT=d)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                            "--master-port", "29533", worker], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        oks = sorted(f for f in os.listdir(d) if f.startswith("ok_rank"))
        assert len(oks) == n, (oks, r.stdout[-2000:])
