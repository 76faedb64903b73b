"""GPU: the tcgen05/TMEM learned-projection PE (v3d_pe_mlp) against a plain torch fp32 evaluation of the
same module (llava/model/position_encoding.py:52-84) and against the reference's own output (golden)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP  # noqa: E402

DEV = torch.device("cuda:0")


def torch_reference(mod: PositionEmbeddingMLP, x: torch.Tensor, tokens=None) -> torch.Tensor:
    """fp32 math with the two roundings the kernel (and a bf16 model) makes: hidden activations and W2 in bf16,
    fp32 accumulation, Linear output rounded to bf16, then the bf16 add."""
    lin1, ln, lin2 = mod.mlp[0], mod.mlp[1], mod.mlp[3]
    h = torch.nn.functional.linear(x.float(), lin1.weight.float(), lin1.bias.float())
    h = torch.nn.functional.layer_norm(h, (512,), ln.weight.float(), ln.bias.float(), ln.eps).relu()
    h = h.to(torch.bfloat16).float()
    w2 = lin2.weight.to(torch.bfloat16).float()
    pe = (h @ w2.T + lin2.bias.float()).to(torch.bfloat16)
    return pe if tokens is None else tokens + pe


def make_module(E, seed=0, scale=1.0):
    torch.manual_seed(seed)
    mod = PositionEmbeddingMLP(E).to(DEV).eval()
    with torch.no_grad():
        mod.mlp[1].weight.uniform_(0.5, 1.5)
        mod.mlp[1].bias.uniform_(-0.5, 0.5)
        mod.mlp[3].weight.mul_(scale)
    return mod


def assert_close_bf16(got, ref, tokens=None):
    """Within one bf16 ulp of the largest quantity involved: the fp32 accumulation order differs from torch's, so
    the PE may land on the other side of a bf16 tie, and with the fused add that flip is measured at the
    magnitude of the PE / token, not of their (possibly cancelling) sum."""
    g, r = got.float(), ref.float()
    err = (g - r).abs()
    mag = r.abs().clamp(min=1.0)
    if tokens is not None:
        mag = torch.maximum(mag, torch.maximum(tokens.float().abs(), (r - tokens.float()).abs()))
    tol = (2.0 ** -6 if tokens is not None else 2.0 ** -7) * mag  # the second rounding (the add) can double the flip
    assert bool((err <= tol).all()), f"max err {err.max().item()} at ref {r.flatten()[err.argmax()].item()}"
    assert float((got.view(torch.int16) != ref.view(torch.int16)).float().mean()) < 0.02


@pytest.mark.parametrize("E,M", [(1152, 729 * 2), (3584, 196 * 3), (192, 300), (128, 128), (64, 1), (1152, 127)])
@pytest.mark.parametrize("add", [False, True])
def test_pe_mlp_matches_torch(E, M, add):
    mod = make_module(E, seed=E + M)
    g = torch.Generator(device="cpu").manual_seed(M)
    x = (torch.rand((1, M, 3), generator=g) * 20 - 10).to(DEV)
    tokens = torch.randn((1, M, E), generator=g).to(DEV).to(torch.bfloat16) if add else None
    with torch.no_grad():
        got = mod(x.to(torch.bfloat16).float(), add_to=tokens)
        ref = torch_reference(mod, x.to(torch.bfloat16).float(), tokens)
    assert got.shape == (1, M, E)
    assert_close_bf16(got.to(torch.bfloat16), ref, tokens)


def test_pe_mlp_large_and_persistent_tiles():
    """More M tiles than SMs (persistent loop, accumulator / ring phases wrap many times)."""
    E, M = 1152, 148 * 128 * 2 + 77
    mod = make_module(E, seed=3)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.rand((1, M, 3), generator=g) * 30 - 15).to(DEV)
    tokens = torch.randn((1, M, E), generator=g).to(DEV).to(torch.bfloat16)
    with torch.no_grad():
        got = mod(x, add_to=tokens)
        ref = torch_reference(mod, x, tokens)
        again = mod(x, add_to=tokens)
    assert_close_bf16(got.to(torch.bfloat16), ref, tokens)
    assert torch.equal(got, again)  # deterministic


def test_pe_mlp_golden(golden_dir):
    """Against the reference module's own fp32 output (tests/golden/frontend.npz, E=192)."""
    G = np.load(os.path.join(golden_dir, "frontend.npz"))
    mod = PositionEmbeddingMLP(192).to(DEV).eval()
    sd = {k: torch.from_numpy(G["mlp_" + k.replace(".", "_")]) for k in mod.state_dict().keys()}
    mod.load_state_dict(sd)
    x = torch.from_numpy(G["avg14"].reshape(2, 729, 3)).to(DEV)
    with torch.no_grad():
        got = mod(x)
    assert got.dtype == torch.float32 and got.shape == (2, 729, 192)
    np.testing.assert_allclose(got[:, ::37].cpu().numpy(), G["mlp_out_sel"], rtol=2e-2, atol=2e-2)


def test_pe_mlp_training_path_uses_autograd():
    mod = make_module(128, seed=9).train()
    x = torch.randn(1, 16, 3, device=DEV)
    out = mod(x)
    out.sum().backward()
    assert mod.mlp[3].weight.grad is not None and out.requires_grad
