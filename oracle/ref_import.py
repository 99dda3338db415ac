"""Import the reference's three hot-path packages VERBATIM from /root/reference.

TEST INFRASTRUCTURE ONLY (never imported by the product path).

Works only where /root/reference exists (the authoring container).  It is used by
  * oracle/make_golden.py  -> writes tests/golden/*.pt (the fixtures that travel to the GPU box)
  * tests/test_oracle_vs_reference.py -> pins oracle/llamole_oracle.py against the real modules
Nothing under tests -m gpu, smoke() or bench.py may call this.

The reference packages have no __init__.py; they import as namespace packages once
/root/reference/src/model is on sys.path.  torch_geometric / rdkit / rdchiral
are not installed here, so the stand-ins under oracle/standins are put on sys.path first.
"""
from __future__ import annotations

import os
import sys
from contextlib import contextmanager

import torch

REFERENCE_ROOT = os.environ.get("LLAMOLE_REFERENCE_ROOT", "/root/reference")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model", "graph_decoder"))


_cache = {}


def load_reference():
    """Returns (diffusion_model, diffusion_utils, graph_encoder.model, graph_predictor.model)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    for name in ("torch_geometric", "rdkit", "rdchiral"):
        try:
            __import__(name)
        except Exception:
            if _STANDINS not in sys.path:
                sys.path.insert(0, _STANDINS)
    model_dir = os.path.join(REFERENCE_ROOT, "src", "model")
    if model_dir not in sys.path:
        sys.path.append(model_dir)
    import graph_decoder.diffusion_model as dm  # type: ignore
    import graph_decoder.diffusion_utils as du  # type: ignore
    import graph_encoder.model as ge  # type: ignore
    import graph_predictor.model as gp  # type: ignore

    _cache["mods"] = (dm, du, ge, gp)
    return _cache["mods"]


class NoiseTape:
    """Pre-drawn Exp(1) noise consumed by a patched Tensor.multinomial.

    torch's multinomial(1) is argmax(p / q), q ~ Exp(1) (probed); replacing the
    draw of q by a recorded tensor makes the reference sampler a deterministic function of the tape.
    `draws` is a list of tensors consumed in call order; each must match the probability tensor's shape.
    """

    def __init__(self, draws):
        self.draws = list(draws)
        self.pos = 0

    def next(self, like):
        q = self.draws[self.pos]
        self.pos += 1
        assert tuple(q.shape) == tuple(like.shape), (q.shape, like.shape)
        return q.to(like.dtype)


@contextmanager
def patched_multinomial(tape: NoiseTape):
    orig = torch.Tensor.multinomial

    def _mn(self, num_samples, replacement=False, *, generator=None):
        assert num_samples == 1
        q = tape.next(self)
        return (self / q).argmax(dim=-1, keepdim=True)

    torch.Tensor.multinomial = _mn
    try:
        yield tape
    finally:
        torch.Tensor.multinomial = orig
