"""CPU-only checks: the C-ABI library loads and exports every symbol include/llamole_b200.h declares, the host
mirrors of the reference classes keep the reference's state-dict keys / file handling / error behaviour, and the
product path refuses to run without a B200 (no fallback)."""
import ctypes
import os
import re
import tempfile

import pytest
import torch

from llamole_b200 import GraphCLIP, GraphDiT, GraphPredictor, _cabi, synth
from llamole_b200.graph_decoder import cosine_schedule, state_from_onehot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "llamole_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(llb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.LIB_PATH):
        from llamole_b200 import build
        build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/llamole_b200.h but not exported"
    assert set(_cabi.SIGNATURES) == set(syms), set(_cabi.SIGNATURES) ^ set(syms)
    assert _cabi.lib().llb_version() >= 100


def test_no_cpu_fallback():
    d = tempfile.mkdtemp()
    cfg, meta = synth.dit_config(64, 1, 1, 4.0, 5, 2.0), synth.dit_meta(6, 7, 1)
    synth.write_dit_checkpoint(d, cfg, meta)
    m = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    m.init_model(d)
    with pytest.raises(_cabi.LlamoleB200Error):
        m.generate_graphs(torch.zeros(2, 10), torch.zeros(2, 768))
    g = GraphCLIP(2, 64, 0.0, {})
    x, ei, ea, b = synth.molecular_graphs(2, seed=1, min_nodes=2, max_nodes=5)
    with pytest.raises(_cabi.LlamoleB200Error):
        g(x, ei, ea, b)
    if not torch.cuda.is_available():
        assert _cabi.lib().llb_arch_check(0) != 0 and _cabi.lib().llb_last_error()


def test_missing_files_raise_like_the_reference():
    d = tempfile.mkdtemp()
    with pytest.raises(FileNotFoundError):
        GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    cfg, meta = synth.dit_config(64, 1, 1, 4.0, 5, 2.0), synth.dit_meta(6, 7, 1)
    synth.write_dit_checkpoint(d, cfg, meta)
    os.remove(os.path.join(d, "model.pt"))
    m = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    with pytest.raises(FileNotFoundError):
        m.init_model(d)
    with pytest.raises(FileNotFoundError):
        GraphCLIP(2, 64, 0.0, {}).init_model(d, verbose=False)
    p = GraphPredictor(2, 64, 0.0, 10, {}, {})
    with pytest.raises(FileNotFoundError):
        p.init_model(d)
    with pytest.raises(FileNotFoundError):
        p.init_neural_cost(d)
    with pytest.raises(ValueError):
        GraphCLIP(1, 64, 0.0, {})
    with pytest.raises(ValueError):
        p.estimate_cost("CCO")


def test_save_pretrained_round_trip():
    d, out = tempfile.mkdtemp(), tempfile.mkdtemp()
    cfg, meta = synth.dit_config(64, 2, 1, 4.0, 5, 2.0), synth.dit_meta(6, 7, 1)
    synth.write_dit_checkpoint(d, cfg, meta)
    m = GraphDiT(os.path.join(d, "config.yaml"), os.path.join(d, "data.meta.json"), torch.float32)
    m.init_model(d)
    m.save_pretrained(out)
    assert sorted(os.listdir(out)) == ["data.meta.json", "model.pt", "model_config.yaml"]
    m2 = GraphDiT(os.path.join(out, "model_config.yaml"), os.path.join(out, "data.meta.json"), torch.float32)
    m2.init_model(out)
    for (k1, v1), (k2, v2) in zip(m.denoiser.state_dict().items(), m2.denoiser.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert m.text_input_size == 768 and m.hidden_size == 64 and m.max_n_nodes == 6
    e = synth.write_encoder_checkpoint(os.path.join(d, "enc"), 3, 64)
    g = GraphCLIP(3, 64, 0.0, {"num_layer": 3})
    g.init_model(e, verbose=False)
    g.save_pretrained(out)
    assert {"model_proj.pt", "model_config.json"} <= set(os.listdir(out))
    p = synth.write_predictor_checkpoint(os.path.join(d, "pred"), 3, 64, 50)
    gp = GraphPredictor(3, 64, 0.0, 50, {}, {i: f"T{i}" for i in range(50)}, available=["CCO", "CC"])
    gp.init_model(p)
    gp.init_neural_cost(p)
    gp.save_pretrained(os.path.join(out, "pred"))
    assert {"model.pt", "cost_model.pt", "model_config.json", "label_to_template.csv.gz", "available.csv.gz"} <= set(os.listdir(os.path.join(out, "pred")))
    assert gp.text_input_size == 768 and gp.available == ["CCO", "CC"]


def test_schedule_and_state_helpers():
    betas, abar = cosine_schedule(500)
    assert betas.shape == (501,) and abar.shape == (501,) and betas.dtype == torch.float32
    assert 0 < float(betas[0]) < float(betas[-1]) <= 1 and float(abar[0]) > 0.99 and float(abar[-1]) < 1e-3
    X = torch.zeros(1, 3, 16)
    X[0, 0, 5] = 1
    E = torch.zeros(1, 3, 3, 5)
    E[0, 0, 1, 2] = E[0, 1, 0, 2] = 1
    Xs, Es = state_from_onehot(X, E)
    assert Xs.tolist() == [[5, -1, -1]] and Es[0, 0, 1] == 2 and Es[0, 0, 0] == -1


def test_synthetic_graphs_follow_the_reference_layout():
    x, ei, ea, b = synth.molecular_graphs(50, seed=0)
    assert x.dtype == ei.dtype == ea.dtype == b.dtype == torch.int64
    assert int(x.min()) >= 0 and int(x.max()) <= 117 and int(ea.min()) >= 1 and int(ea.max()) <= 4
    assert bool((b[1:] >= b[:-1]).all()) and int(b[-1]) == 49
    # both directions listed, no edge crosses a graph
    fwd = set(map(tuple, ei.t().tolist()))
    assert all((j, i) in fwd for i, j in fwd)
    assert bool((b[ei[0]] == b[ei[1]]).all())
    deg = torch.bincount(ei[1], minlength=x.numel())
    assert int(deg.max()) <= 4


# ------------------------------------------------------------------------------------------------ template merge (host)
def _reference_merge(topk_probs, templates, product_smiles, run):
    """Literal restatement of the host half of the reference's sample_templates (graph_predictor/model.py:187-228)."""
    from collections import defaultdict

    reactants_d = defaultdict(list)
    for prob, template in zip(topk_probs, templates):
        try:
            outcomes = run(template, product_smiles)
            if len(outcomes) == 0:
                continue
            outcomes = sorted(outcomes)
            for reactant in outcomes:
                if "." in reactant:
                    str_list = sorted(reactant.strip().split("."))
                    reactants_d[".".join(str_list)].append((prob / len(outcomes), template))
                else:
                    reactants_d[reactant].append((prob / len(outcomes), template))
        except Exception:
            pass
    if len(reactants_d) == 0:
        return [], [], []
    ret = []
    for reactant, l in reactants_d.items():
        ss, ts = zip(*l)
        ret.append((reactant, sum(ss), list(ts)[0]))
    reactants, scores, templates = zip(*sorted(ret, key=lambda item: item[1], reverse=True))
    total = sum(scores)
    return list(reactants), [s / total for s in scores], list(templates)


def _fake_rdchiral(template, smiles):
    """Deterministic stand-in for rdchiralRunText: outcomes depend on (template, product); some templates do not apply,
    some raise, several map to the same reactant set written in a different order."""
    k = int(template[1:])
    if k % 5 == 0:
        return []
    if k % 7 == 0:
        raise ValueError("template does not parse")
    if k % 3 == 0:
        return [f"C{len(smiles)}.N", f"N.C{len(smiles)}", "O"]      # two spellings of one reactant set
    return [f"C{k % 4}", f"O.C{k % 2}"]


def test_template_merge_matches_the_reference_logic():
    from llamole_b200.graph_predictor import set_template_backend

    gp = GraphPredictor(2, 64, 0.0, 40, {}, {i: f"T{i}" for i in range(40)})
    set_template_backend(_fake_rdchiral)
    try:
        g = torch.Generator().manual_seed(0)
        for trial in range(20):
            k = int(torch.randint(1, 12, (1,), generator=g))
            labels = torch.randperm(40, generator=g)[:k].tolist()
            probs = torch.rand(k, generator=g).sort(descending=True).values.tolist()
            smi = "C" * (trial % 5 + 1)
            got = gp._apply_templates(probs, labels, smi)
            want = _reference_merge(probs, [gp.label_to_template[i] for i in labels], smi, _fake_rdchiral)
            assert got[0] == want[0] and got[2] == want[2]
            assert all(abs(a - b) < 1e-12 for a, b in zip(got[1], want[1]))
            if got[1]:
                assert abs(sum(got[1]) - 1.0) < 1e-9
        assert gp._apply_templates([0.5, 0.5], [0, 5], "CC") == ([], [], [])     # nothing applies
    finally:
        set_template_backend(None)
    with pytest.raises(ValueError):
        gp.sample_templates_batch([object()], None, ["C", "CC"], 3)
    assert gp.sample_templates_batch([], None, [], 3) == []
