"""GPU parity: attention core vs F.scaled_dot_product_attention in fp32 on the same bf16 inputs."""
import pytest
import torch
import torch.nn.functional as F

from open_clip_b200 import ops
from gpu_util import BF16, randn, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("batch,seq,heads", [(2, 17, 2), (3, 20, 2), (2, 50, 12), (2, 77, 8), (1, 197, 4), (1, 1, 1),
                                             (2, 130, 2),  # 80 < L < 192: the single-kernel recompute backward
                                             # long sequences: two-pass K,V-resident backward (L > 384), Q-streaming
                                             # forward (L > 528); 577 = ViT-L/14-336, 640 = the supported maximum
                                             (1, 400, 2), (2, 577, 3), (1, 640, 1)])
def test_attention_fwd_bwd(batch, seq, heads, causal):
    d = heads * 64
    qkv = randn(batch * seq, 3 * d, seed=seq, scale=1.0)
    o, lse = ops.attention_fwd(qkv, batch, seq, heads, causal)
    x = qkv.float().view(batch, seq, 3, heads, 64).requires_grad_(True)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    mask = torch.full((seq, seq), float("-inf"), device="cuda").triu_(1) if causal else None
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=64 ** -0.5)  # [B,H,L,64]
    ref_flat = ref.transpose(1, 2).reshape(batch * seq, d)
    assert rel_err(o, ref_flat) < 5e-3, rel_err(o, ref_flat)
    s = (q @ k.transpose(-1, -2)) * 64 ** -0.5
    if causal:
        s = s + mask
    assert rel_err(lse, torch.logsumexp(s, -1)) < 1e-4
    do = randn(batch * seq, d, seed=seq + 1)
    dbias = torch.zeros(3 * d, dtype=torch.float32, device="cuda")
    dqkv = ops.attention_bwd(qkv, o, do, lse, batch, seq, heads, causal, dbias=dbias)
    # fused in_proj_bias gradient == column sums of the (bf16) dqkv the kernel wrote
    want = dqkv.float().sum(0)
    assert float((dbias - want).norm() / (want.norm() + 1e-6 * dqkv.float().norm())) < 2e-3
    ref_flat.backward(do.float())
    got = dqkv.float().view(batch, seq, 3, heads, 64)
    for i, name in enumerate("qkv"):
        e = float((got[:, :, i] - x.grad[:, :, i]).norm() / (x.grad[:, :, i].norm() + 1e-3 * do.float().norm()))
        assert e < 1.5e-2, (name, e)
