"""CPU: the closed-form gradient of the weighted Kabsch block (3dm:696-758) that `head_train_backward_kernel`
implements (csrc/head.cu), restated in torch and checked against autograd of the oracle's softmax + Kabsch -- including
the det(R) < 0 branch, where the reference's own backward raises because `Vt[-1, :] *= -1` (3dm:750) modifies an output
of torch.linalg.svd in place.

    H = U S V^T,  R = V D U^T (D = diag(1, 1, +-1)),  t = ct - R cs
    A = V^T (dR - dt cs^T) U,   Mbar_ij = alpha_ij A_ij + beta_ji A_ji (i != j),   dH = U Mbar V^T
    equal-sign pair:  alpha_ij = -d / (s_i + s_j),  beta_ij = d / (s_i + s_j);   mixed pair:  alpha_ij = beta_ij = d_i / (s_i - s_j)
"""
import pytest
import torch


def forward(p, q, logits):
    a = torch.softmax(logits, -1)
    w = a / (a.sum() + 1e-6)                                                    # 3dm:718-724
    cs, ct = (w[:, None] * p).sum(0), (w[:, None] * q).sum(0)
    pc, qc = p - cs, q - ct
    H = (w[:, None, None] * pc[:, :, None] * qc[:, None, :]).sum(0) + 1e-6 * torch.eye(3, dtype=p.dtype)
    U, S, Vt = torch.linalg.svd(H)
    d3 = 1.0
    R = Vt.T @ U.T
    if torch.det(R) < 0:                                                        # out-of-place form of 3dm:749-751
        R = Vt.T @ torch.diag(torch.tensor([1.0, 1.0, -1.0], dtype=p.dtype)) @ U.T
        d3 = -1.0
    return R, ct - R @ cs, (w, cs, ct, pc, qc, U, S, Vt.T, d3)


def closed_form(p, q, logits, dR, dt):
    R, t, (w, cs, ct, pc, qc, U, S, V, d3) = forward(p, q, logits)
    GR = dR - torch.outer(dt, cs)
    dcs, dct = -R.T @ dt, dt.clone()
    A = V.T @ GR @ U
    d, s = [1.0, 1.0, d3], S.tolist()

    def coef(i, j, beta):
        if d[i] == d[j]:
            return (d[i] if beta else -d[i]) / (s[i] + s[j])
        return d[i] / (s[i] - s[j])

    Mb = torch.zeros(3, 3, dtype=p.dtype)
    for i in range(3):
        for j in range(3):
            if i != j:
                Mb[i, j] = coef(i, j, False) * A[i, j] + coef(j, i, True) * A[j, i]
    GH = U @ Mb @ V.T
    dcs = dcs - GH @ (w[:, None] * qc).sum(0)
    dct = dct - GH.T @ (w[:, None] * pc).sum(0)
    dw = torch.einsum("ni,ij,nj->n", pc, GH, qc) + p @ dcs + q @ dct
    dp = w[:, None] * ((GH @ qc.T).T + dcs)
    dq = w[:, None] * ((GH.T @ pc.T).T + dct)
    dlogits = w * (dw - (w * dw).sum() * (1 + 1e-6))
    return dp, dq, dlogits


@pytest.mark.parametrize("reflect", [False, True])
def test_closed_form_kabsch_gradient_equals_autograd(reflect):
    g = torch.Generator().manual_seed(3 + int(reflect))
    n = 60
    p = torch.randn(n, 3, generator=g, dtype=torch.float64)
    Q = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))[0]
    if (torch.det(Q) < 0) != reflect:
        Q[:, 0] *= -1
    q = p @ Q.T + 0.05 * torch.randn(n, 3, generator=g, dtype=torch.float64) + torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64)
    logits = torch.randn(n, generator=g, dtype=torch.float64)
    dR = torch.randn(3, 3, generator=g, dtype=torch.float64)
    dt = torch.randn(3, generator=g, dtype=torch.float64)
    leaves = [v.clone().requires_grad_(True) for v in (p, q, logits)]
    R, t, aux = forward(*leaves)
    assert (aux[-1] < 0) == reflect
    ((R * dR).sum() + (t * dt).sum()).backward()
    with torch.no_grad():
        got = closed_form(p, q, logits, dR, dt)
    for a, b in zip(got, leaves):
        assert float((a - b.grad).abs().max()) <= 1e-10 * float(b.grad.abs().max())
