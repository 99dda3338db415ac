"""Live pin of the oracle and of the drop-in classes against the UNMODIFIED reference modules.  Runs only where
/root/reference exists (the authoring container); the committed fixtures under tests/golden carry the same pin to
the GPU box."""
import os
import tempfile

import pytest
import torch

from llamole_b200 import GraphCLIP, GraphDiT, GraphPredictor, synth
from oracle import llamole_oracle as O
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present on this machine")


def _ref_dit(cfg, meta, sd):
    dm, du, _, _ = load_reference()
    d = tempfile.mkdtemp()
    synth.write_dit_checkpoint(d, cfg, meta, sd)
    m = dm.GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    m.init_model(d)
    return m.eval(), d


def test_state_dict_keys_and_shapes_match_the_reference_modules():
    dm, du, ge, gp = load_reference()
    cfg, meta = synth.dit_config(128, 3, 2, 4.0, 10, 2.0), synth.dit_meta(9, 7, 1)
    sd = synth.dit_state_dict(cfg, 9)
    ref, d = _ref_dit(cfg, meta, sd)
    mine = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    a, b = ref.denoiser.state_dict(), mine.denoiser.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert torch.equal(mine.x_marginals, ref.limit_dist.X) and torch.equal(mine.e_marginals, ref.limit_dist.E)
    assert torch.equal(mine.betas, ref.noise_schedule.betas) and torch.equal(mine.alphas_bar, ref.noise_schedule.alphas_bar)
    assert torch.equal(mine.xe_conditions, ref.transition_model.xe_conditions) and torch.equal(mine.ex_conditions, ref.transition_model.u_ex[0])
    assert mine.text_input_size == ref.text_input_size and mine.hidden_size == ref.hidden_size and mine.T == ref.T
    rc, mc = ge.GraphCLIP(3, 64, 0.0, {}), GraphCLIP(3, 64, 0.0, {})
    for x, y in ((rc.molecule_encoder, mc.molecule_encoder), (rc.molecule_projection, mc.molecule_projection)):
        assert {k: v.shape for k, v in x.state_dict().items()} == {k: v.shape for k, v in y.state_dict().items()}
    rp, mp_ = gp.GraphPredictor(3, 64, 0.0, 77, {}, {}), GraphPredictor(3, 64, 0.0, 77, {}, {})
    assert {k: v.shape for k, v in rp.predictor.state_dict().items()} == {k: v.shape for k, v in mp_.predictor.state_dict().items()}
    rcost = gp.CostMLP(1, 2048, 128, 0.1)
    assert set(rcost.state_dict()) == set(synth.cost_mlp_state_dict())


def test_oracle_denoiser_matches_reference_on_fresh_inputs():
    cfg, meta = synth.dit_config(128, 3, 2, 4.0, 10, 3.0), synth.dit_meta(9, 11, 1)
    sd = synth.dit_state_dict(cfg, 9, seed=5)
    ref, _ = _ref_dit(cfg, meta, sd)
    dm, du, _, _ = load_reference()
    B, N = 4, 9
    n_nodes = torch.tensor([9, 2, 5, 1])
    mask = torch.arange(N).unsqueeze(0) < n_nodes.unsqueeze(1)
    g = torch.Generator().manual_seed(0)
    tb = O.dit_tables(meta)
    ex = lambda *s: torch.empty(*s).exponential_(1.0, generator=g)  # noqa: E731
    X, E = O.initial_state(tb, mask, ex(B, N, 16), ex(B, N, N, 5), torch.float32)
    props, txt = synth.dit_conditions(B, seed=9)
    y = torch.where(props == -200.0, float("nan"), props)
    t = torch.full((B, 1), 7.0) / 10
    with torch.no_grad():
        for unc in (False, True):
            pred = ref._forward({"X_t": X, "E_t": E, "y_t": y, "t": t, "node_mask": mask}, txt, unconditioned=unc)
            lX, lE = O.denoiser_forward(sd, cfg, X, E, mask, y, txt, t, unc)
            assert torch.allclose(lX, pred.X, atol=2e-5) and torch.allclose(lE, pred.E, atol=2e-5)


def test_oracle_gin_matches_reference_on_fresh_inputs():
    _, _, ge, gp = load_reference()
    L, H, out_dim = 4, 128, 97
    x, ei, ea, b = synth.molecular_graphs(11, seed=21, min_nodes=1, max_nodes=30)
    enc, proj = synth.gin_encoder_state_dicts(L, H, seed=2)
    clip = ge.GraphCLIP(L, H, 0.0, {})
    clip.molecule_encoder.load_state_dict(enc)
    clip.molecule_projection.load_state_dict(proj)
    pred_sd = synth.gin_predictor_state_dict(L, H, out_dim, seed=3)
    pred = gp.GNNRetrosynthsizer(L, H, 768, 0.0, out_dim)
    pred.load_state_dict(pred_sd)
    c = synth.text_conditions(11, seed=8)
    with torch.no_grad():
        assert torch.allclose(O.gin_encoder_forward(enc, proj, L, x, ei, ea, b), clip.eval()(x, ei, ea, b), atol=1e-6)
        assert torch.allclose(O.gin_predictor_forward(pred_sd, L, x, ei, ea, b, c), pred.eval()(x, ei, ea, b, c), atol=1e-4)
        assert torch.allclose(O.gin_predictor_forward(pred_sd, L, x, ei, ea, b, None), pred(x, ei, ea, b, None), atol=1e-4)


@pytest.mark.parametrize("train", [False, True])
def test_graphdit_forward_loss_matches_reference(train):
    """GraphDiT.forward (the SFT loss, diffusion_model.py:148-250 + TrainLossDiscrete) of the drop-in class against the
    verbatim reference module under the same torch.manual_seed: same timestep / noise / condition-dropout draws, hence the
    same scalar (fp32 round-off) and the same gradients."""
    cfg, meta = synth.dit_config(128, 2, 2, 4.0, 20, 2.0), synth.dit_meta(12, 3, 1)
    sd = synth.dit_state_dict(cfg, 12, seed=8)
    ref, d = _ref_dit(cfg, meta, sd)
    mine = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    mine.init_model(d)
    ref.train(train), mine.train(train)
    x, ei, ea, b = synth.molecular_graphs(5, seed=4, min_nodes=1, max_nodes=12)
    active = (torch.tensor(meta["atom_type_dist"]) > 0).nonzero().squeeze()
    x = active[x % 16]
    x[3] = int((torch.tensor(meta["atom_type_dist"]) == 0).nonzero()[0])     # an atom outside the active set
    props, txt = synth.dit_conditions(5, seed=2)
    txt = txt.clone().requires_grad_(True)
    txt_r = txt.detach().clone().requires_grad_(True)
    for seed in (0, 1):
        torch.manual_seed(seed)
        want = ref(x, ei, ea, b, props, txt_r, -200)
        torch.manual_seed(seed)
        got = mine(x, ei, ea, b, props, txt, -200)
        assert got.shape == want.shape == () and torch.allclose(got, want, rtol=1e-5, atol=1e-6), (float(got), float(want))
    want.backward()
    got.backward()
    assert torch.allclose(txt.grad, txt_r.grad, rtol=1e-4, atol=1e-7)
    gr = dict(ref.denoiser.named_parameters())
    for k, p in mine.denoiser.named_parameters():
        if gr[k].grad is not None:
            assert p.grad is not None and torch.allclose(p.grad, gr[k].grad, rtol=2e-4, atol=1e-6), k
