"""Numerics of the analytic LayerNorm statistics of the fused GIN node MLP (csrc/llb_gin.cu: pack-time centring + Cholesky factor,
EpiRowSq, EpiLnGelu): host restatement of exactly what those kernels compute --
    Wc = bf16(W - column mean), bc = b - mean(b)                       (LayerNorm is shift-invariant: LN(Wc a + bc) = LN(W a + b))
    [Wc | bc]^T [Wc | bc] = Rt^T Rt (fp64 Cholesky), R = bf16(Rt[:H,:H]), r = Rt[:H,H], c0 = Rt[H,H]^2
    |zc|^2 = sum_k (R a + r)_k^2 + c0,  rstd = rsqrt(|zc|^2 / 4H + eps),  out = GELU(rstd (acc gamma + bc gamma) + beta)
-- against the two-pass LayerNorm.  Measured against the result with fp64 weights, the scheme must be as accurate as the
two-pass LayerNorm on bf16-rounded weights (the error of both is the bf16 rounding of the weights)."""
import torch

from llamole_b200 import synth


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def test_analytic_statistics_are_as_accurate_as_the_two_pass_layernorm():
    L, H = 2, 256
    enc, _ = synth.gin_encoder_state_dicts(L, H, seed=11)
    x, ei, ea, _ = synth.molecular_graphs(64, seed=0)
    pre = "convs.0."
    h = enc["atom_encoder.weight"][x] + enc["virtualnode_embedding.weight"][0]
    msg = torch.nn.functional.gelu(_bf(h)[ei[0]] + enc[pre + "bond_encoder.weight"][ea])
    a = _bf(((1 + enc[pre + "eps"]) * h).index_add(0, ei[1], msg))          # the bf16 A operand (gin_aggregate_kernel)
    W, b = enc[pre + "mlp.0.weight"], enc[pre + "mlp.0.bias"]
    gamma, beta = enc[pre + "mlp.1.weight"], enc[pre + "mlp.1.bias"]
    rows = W.shape[0]
    truth = torch.nn.functional.gelu(torch.nn.functional.layer_norm(a.double() @ W.double().t() + b.double(), (rows,), gamma.double(), beta.double(), 1e-5))
    two_pass = torch.nn.functional.gelu(torch.nn.functional.layer_norm(a @ _bf(W).t() + b, (rows,), gamma, beta, 1e-5))
    # pack time
    Wc = _bf(W - W.mean(0, keepdim=True))
    bc = b - b.mean()
    Wt = torch.cat([Wc.double(), bc.double()[:, None]], dim=1)
    Rt = torch.linalg.cholesky(Wt.t() @ Wt, upper=True)
    R, r, c0 = _bf(Rt[:H, :H].float()), Rt[:H, H].float(), float(Rt[H, H] ** 2)
    # statistics GEMM epilogue (EpiRowSq) and the first linear's epilogue (EpiLnGelu)
    q = ((a @ R.t() + r) ** 2).sum(1) + c0
    rstd = torch.rsqrt(q / rows + 1e-5)
    out = torch.nn.functional.gelu(rstd[:, None] * ((a @ Wc.t()) * gamma + bc * gamma) + beta)
    z = a @ Wc.t() + bc
    rstd_ref = torch.rsqrt(z.var(1, unbiased=False) + 1e-5)
    rms = lambda t: float(t.double().pow(2).mean().sqrt())  # noqa: E731
    assert float(((rstd - rstd_ref) / rstd_ref).abs().max()) < 2e-3
    assert float((z.mean(1).abs() * rstd_ref).max()) < 1e-3            # residual mean of the centred, bf16-rounded weight, in sigmas
    e_new, e_old = rms(out.double() - truth), rms(two_pass.double() - truth)
    assert e_new < 1.15 * e_old, (e_new, e_old)
    assert float((out.double() - truth).abs().max()) < 1.5 * float((two_pass.double() - truth).abs().max())
