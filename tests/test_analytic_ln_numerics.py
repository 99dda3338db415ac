"""Numerics of the analytic LayerNorm statistics used by the opt-in fused GIN node MLP (csrc/llb_gin.cu: EpiRowStats,
gin_ln_stats_kernel, EpiLnGelu): host restatement of exactly what those kernels compute -- mean = a . wbar + bbar,
E[z^2] = (a . (G a + 2 W^T b)) / 4H + |b|^2 / 4H with G = W^T W rounded to bf16 and a, W in bf16 -- against the two-pass
LayerNorm of z = W a + b.  The error of GELU(LN(z)) must stay far below the bf16 rounding of that output (the GEMM operand)."""
import torch

from llamole_b200 import synth


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def test_analytic_statistics_are_below_the_operand_rounding():
    L, H = 2, 256
    enc, _ = synth.gin_encoder_state_dicts(L, H, seed=11)
    x, ei, ea, _ = synth.molecular_graphs(64, seed=0)
    pre = "convs.0."
    h = enc["atom_encoder.weight"][x] + enc["virtualnode_embedding.weight"][0]
    msg = torch.nn.functional.gelu(_bf(h)[ei[0]] + enc[pre + "bond_encoder.weight"][ea])
    a = _bf(((1 + enc[pre + "eps"]) * h).index_add(0, ei[1], msg))          # the bf16 A operand (gin_aggregate_kernel)
    W, b = _bf(enc[pre + "mlp.0.weight"]), enc[pre + "mlp.0.bias"]           # the bf16 weight the tensor core multiplies by
    gamma, beta = enc[pre + "mlp.1.weight"], enc[pre + "mlp.1.bias"]
    rows = W.shape[0]
    z = a @ W.t() + b
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(z, (rows,), gamma, beta, 1e-5))
    # pack time (gin_gram_kernel, gin_stat_vectors_kernel)
    G = _bf(W.t() @ W)
    wbar, c2 = W.mean(0), 2.0 * (W.t() @ b)
    bbar, bb = b.mean(), (b * b).sum() / rows
    # statistics GEMM epilogue + finaliser
    mean = a @ wbar + bbar
    ez2 = ((a @ G + c2) * a).sum(1) / rows + bb
    rstd = torch.rsqrt((ez2 - mean * mean).clamp_min(0) + 1e-5)
    out = torch.nn.functional.gelu((z - mean[:, None]) * rstd[:, None] * gamma + beta)   # EpiLnGelu
    err = (out - ref).abs()
    rounding = (_bf(ref) - ref).abs()
    rms = lambda t: float(t.pow(2).mean().sqrt())  # noqa: E731
    assert float(((mean - z.mean(1)).abs() / z.std(1)).max()) < 1e-5
    assert rms(err) < rms(rounding) / 10, (rms(err), rms(rounding))
    assert float(err.max()) < float(rounding.max()) / 5, (float(err.max()), float(rounding.max()))
