"""Pin oracle/llamole_oracle.py against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

from llamole_b200 import synth
from oracle import llamole_oracle as O


def _weights(fx):
    P = fx["params"]
    sd = synth.dit_state_dict(fx["cfg"], P["max_nodes"], P["w_seed"])
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - fx["weights_checksum"]) < 1e-6 * fx["weights_checksum"], "synthetic weight generator drifted"
    return sd


def test_tables_match_reference(dit_small):
    fx = dit_small
    tb = O.dit_tables(fx["meta"])
    betas, abar = O.cosine_schedule(fx["cfg"]["diffusion_steps"])
    assert torch.equal(betas, fx["schedule_betas"])
    assert torch.equal(abar, fx["schedule_abar"])
    for a, b in ((tb.x_marg, fx["x_marg"]), (tb.e_marg, fx["e_marg"]), (tb.xe, fx["xe"]), (tb.ex, fx["ex"])):
        assert torch.equal(a, b)


def test_initial_state_matches_reference(dit_small):
    fx = dit_small
    tb = O.dit_tables(fx["meta"])
    N = tb.max_nodes
    mask = torch.arange(N).unsqueeze(0) < fx["n_nodes"].unsqueeze(1)
    X, E = O.initial_state(tb, mask, fx["qX0"], fx["qE0"], torch.float32)
    assert torch.equal(X, fx["X_T"]) and torch.equal(E, fx["E_T"])


def test_denoiser_logits_match_reference(dit_small):
    fx = dit_small
    sd = _weights(fx)
    cfg, T = fx["cfg"], fx["cfg"]["diffusion_steps"]
    tb = O.dit_tables(fx["meta"])
    N = tb.max_nodes
    B = fx["props"].shape[0]
    mask = torch.arange(N).unsqueeze(0) < fx["n_nodes"].unsqueeze(1)
    y = torch.where(fx["props"] == -200.0, float("nan"), fx["props"])
    # step 0 of the loop (t = T, zero-diagonal state) and a later one (one-hot diagonal)
    states = [(fx["X_T"], fx["E_T"], T)]
    Xs, Es = fx["cat_X"][3].long(), fx["cat_E"][3].long()
    X4, E4 = O.one_hot_state(Xs, Es, mask, torch.float32)
    states.append((X4, E4, T - 4))
    for (X, E, t), i in zip(states, (0, 4)):
        t_norm = (torch.full((B, 1), float(t)) / T)
        for unc, kx, ke in ((False, "logits_cond_X", "logits_cond_E"), (True, "logits_unc_X", "logits_unc_E")):
            lX, lE = O.denoiser_forward(sd, cfg, X, E, mask, y, fx["txt"], t_norm, unc)
            assert torch.allclose(lX, fx[kx][i], atol=2e-5, rtol=1e-5), (lX - fx[kx][i]).abs().max()
            assert torch.allclose(lE, fx[ke][i], atol=2e-5, rtol=1e-5), (lE - fx[ke][i]).abs().max()


def test_posterior_guidance_match_reference(dit_small):
    fx = dit_small
    cfg, T = fx["cfg"], fx["cfg"]["diffusion_steps"]
    tb = O.dit_tables(fx["meta"])
    U = O.union_transition(tb)
    betas, abar = O.cosine_schedule(T)
    N = tb.max_nodes
    mask = torch.arange(N).unsqueeze(0) < fx["n_nodes"].unsqueeze(1)
    X, E = fx["X_T"], fx["E_T"]
    for i in range(T):
        t = T - i
        args = (float(betas[t]), float(abar[t - 1]), float(abar[t]))
        pc = O.posterior_dense(tb, U, fx["logits_cond_X"][i], fx["logits_cond_E"][i], X, E, *args)
        pu = O.posterior_dense(tb, U, fx["logits_unc_X"][i], fx["logits_unc_E"][i], X, E, *args)
        pX = O.guidance(pc[0], pu[0], cfg["guide_scale"])
        pE = O.guidance(pc[1], pu[1], cfg["guide_scale"])
        assert torch.allclose(pX, fx["prob_X"][i], atol=1e-6, rtol=1e-5)
        assert torch.allclose(pE, fx["prob_E"][i], atol=1e-6, rtol=1e-5)
        Xs, Es = O.sample_categories(fx["prob_X"][i], fx["prob_E"][i], mask, fx["qX"][t - 1], fx["qE"][t - 1])
        assert torch.equal(Xs.to(torch.int8), fx["cat_X"][i]) and torch.equal(Es.to(torch.int8), fx["cat_E"][i])
        X, E = O.one_hot_state(Xs, Es, mask, torch.float32)


def test_full_trajectory_matches_reference(dit_small):
    fx = dit_small
    sd = _weights(fx)
    noise = {k: fx[k] for k in ("qX0", "qE0", "qX", "qE")}
    Xc, Ec = O.sample_graphs(sd, fx["cfg"], fx["meta"], fx["props"], fx["txt"], fx["n_nodes"], noise)
    assert torch.equal(Xc.to(torch.int8), fx["final_X"])
    assert torch.equal(Ec.to(torch.int8), fx["final_E"])


def test_gin_encoder_and_predictor_match_reference(gin_small):
    fx = gin_small
    P = fx["params"]
    enc, proj = synth.gin_encoder_state_dicts(P["L"], P["H"], P["enc_seed"])
    pred = synth.gin_predictor_state_dict(P["L"], P["H"], P["out_dim"], seed=P["pred_seed"])
    g = (fx["x"], fx["edge_index"], fx["edge_attr"], fx["batch"])
    pooled = O.gin_trunk(enc, P["L"], *g)
    assert torch.allclose(pooled, fx["encoder_pooled"], atol=1e-4, rtol=1e-5)
    emb = O.gin_encoder_forward(enc, proj, P["L"], *g)
    assert torch.allclose(emb, fx["encoder_embedding"], atol=1e-6, rtol=1e-5)
    lc = O.gin_predictor_forward(pred, P["L"], *g, fx["c"])
    ln = O.gin_predictor_forward(pred, P["L"], *g, None)
    assert torch.allclose(lc, fx["predictor_logits"], atol=1e-4, rtol=1e-5)
    assert torch.allclose(ln, fx["predictor_logits_dropped"], atol=1e-4, rtol=1e-5)
    tv, ti = O.predictor_topk(fx["predictor_logits"], 10)
    assert torch.equal(ti, fx["topk_indices"]) and torch.allclose(tv, fx["topk_probs"])
    cost = O.cost_mlp_forward(synth.cost_mlp_state_dict(), fx["fps"])
    assert torch.allclose(cost, fx["cost"], atol=1e-6)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    z = O.philox4x32(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    z = O.philox4x32(f, np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
