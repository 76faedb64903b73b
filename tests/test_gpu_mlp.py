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
    """A bf16 module (what the recipe trains and evaluates): the tcgen05 kernel serves exactly this dtype."""
    torch.manual_seed(seed)
    mod = PositionEmbeddingMLP(E).to(DEV).eval()
    with torch.no_grad():
        mod.mlp[1].weight.uniform_(0.5, 1.5)
        mod.mlp[1].bias.uniform_(-0.5, 0.5)
        mod.mlp[3].weight.mul_(scale)
    return mod.to(torch.bfloat16)


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
    x = (torch.rand((1, M, 3), generator=g) * 20 - 10).to(DEV).to(torch.bfloat16)
    tokens = torch.randn((1, M, E), generator=g).to(DEV).to(torch.bfloat16) if add else None
    with torch.no_grad():
        got = mod(x, add_to=tokens)
        ref = torch_reference(mod, x, tokens)
    assert got.shape == (1, M, E) and got.dtype == torch.bfloat16
    assert_close_bf16(got, ref, tokens)


def test_pe_mlp_large_and_persistent_tiles():
    """More M tiles than SMs (persistent loop, accumulator / ring phases wrap many times)."""
    E, M = 1152, 148 * 128 * 2 + 77
    mod = make_module(E, seed=3)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.rand((1, M, 3), generator=g) * 30 - 15).to(DEV).to(torch.bfloat16)
    tokens = torch.randn((1, M, E), generator=g).to(DEV).to(torch.bfloat16)
    with torch.no_grad():
        got = mod(x, add_to=tokens)
        ref = torch_reference(mod, x, tokens)
        again = mod(x, add_to=tokens)
    assert_close_bf16(got, ref, tokens)
    assert torch.equal(got, again)  # deterministic


def fp32_module_and_bound(mod32: PositionEmbeddingMLP, x32: torch.Tensor):
    """The reference module's fp32 output and the element-wise error bound of evaluating it with W2 and the hidden
    activations rounded to bf16 (each product within 2^-8 relative: two half-ulp roundings of 2^-9) and the output rounded
    to bf16 once more (2^-9 relative): |err| <= 2^-8 * (sum_k |w2_k| |h_k|) + 2^-9 * |out|, plus fp32 accumulation noise."""
    lin1, ln, lin2 = mod32.mlp[0], mod32.mlp[1], mod32.mlp[3]
    h = torch.nn.functional.layer_norm(torch.nn.functional.linear(x32, lin1.weight, lin1.bias), (512,), ln.weight, ln.bias, ln.eps).relu()
    out = torch.nn.functional.linear(h, lin2.weight, lin2.bias)
    mass = h.abs() @ lin2.weight.abs().T
    return out, 2.0 ** -8 * mass + 2.0 ** -9 * out.abs() + 1e-6 * mass


@pytest.mark.parametrize("E,M", [(1152, 729 * 4), (3584, 196 * 8)])
def test_pe_mlp_error_against_the_fp32_module(E, M, capsys):
    """north_star: bf16 PE within bf16 rounding of the reference's fp32 module.  The kernel keeps W2 and the hidden
    activations in bf16 (tensor-core operands); the measured error stays inside the rounding bound of exactly those
    roundings and is reported relative to the output scale."""
    torch.manual_seed(E)
    mod32 = PositionEmbeddingMLP(E).to(DEV).eval()  # nn.Linear's default init: the scale a trained projection starts from
    x = (torch.rand((1, M, 3), device=DEV) * 20 - 10)
    # a bf16 copy sees bf16-rounded parameters and coordinates; compare on those inputs so that only the kernel's own
    # roundings (W2, hidden, output) are measured
    mod16 = PositionEmbeddingMLP(E).to(DEV).eval()
    mod16.load_state_dict(mod32.state_dict())
    mod16 = mod16.to(torch.bfloat16)
    mod32.load_state_dict({k: v.float() for k, v in mod16.state_dict().items()})
    x16 = x.to(torch.bfloat16)
    with torch.no_grad():
        got = mod16(x16).float()
        ref, bound = fp32_module_and_bound(mod32, x16.float())
    err = (got - ref).abs()
    assert bool((err <= bound).all()), f"max err/bound {(err / bound).max().item():.3f}"
    rms = ref.pow(2).mean().sqrt().item()
    with capsys.disabled():
        print(f"\n[mlp-pe E={E}] max |err| = {err.max().item():.3e}, output rms = {rms:.3e}, max err / rms = {err.max().item() / rms:.3e} "
              f"(2^-8 = {2 ** -8:.3e}), max err / bound = {(err / bound).max().item():.3f}")
    assert err.max().item() <= 2.0 ** -6 * max(rms, ref.abs().max().item())


def test_pe_mlp_golden(golden_dir):
    """Against the reference module's own fp32 output (tests/golden/frontend.npz, E=192), within the rounding bound of
    the bf16 operands (fp32_module_and_bound) -- no blanket tolerance."""
    G = np.load(os.path.join(golden_dir, "frontend.npz"))
    mod = PositionEmbeddingMLP(192).to(DEV).eval()
    sd = {k: torch.from_numpy(G["mlp_" + k.replace(".", "_")]) for k in mod.state_dict().keys()}
    mod.load_state_dict(sd)
    x = torch.from_numpy(G["avg14"].reshape(2, 729, 3)).to(DEV)
    with torch.no_grad():
        ref32, bound = fp32_module_and_bound(mod, x)
        # an fp32 module keeps the reference's fp32 operator sequence (no silent bf16): same values as the reference
        got32 = mod(x)
    assert got32.dtype == torch.float32 and got32.shape == (2, 729, 192)
    np.testing.assert_allclose(got32[:, ::37].cpu().numpy(), G["mlp_out_sel"], rtol=1e-5, atol=1e-5)
    # the bf16 copy runs the tcgen05 kernel: parameters and coordinates rounded to bf16 add their own (input) error,
    # bounded through the fp32 module evaluated on the rounded inputs
    mod16 = PositionEmbeddingMLP(192).to(DEV).eval()
    mod16.load_state_dict(sd)
    mod16 = mod16.to(torch.bfloat16)
    mod_r = PositionEmbeddingMLP(192).to(DEV).eval()
    mod_r.load_state_dict({k: v.float() for k, v in mod16.state_dict().items()})
    x16 = x.to(torch.bfloat16)
    with torch.no_grad():
        got = mod16(x16)
        ref_r, bound_r = fp32_module_and_bound(mod_r, x16.float())
    assert got.dtype == torch.bfloat16
    assert bool(((got.float() - ref_r).abs() <= bound_r).all())
    # and the whole thing stays within bf16-level distance of the reference's golden output
    gold = torch.from_numpy(G["mlp_out_sel"]).to(DEV)
    scale = gold.abs().max().item()
    assert float((got[:, ::37].float() - gold).abs().max()) <= 2.0 ** -5 * scale


def test_pe_mlp_autograd_kernel_forward_torch_backward():
    """Training (train_3d.py:1818-1821 makes the module trainable): forward = the tcgen05 kernel, backward = the reference's
    operator sequence differentiated by torch.  Gradients equal those of the plain torch module up to bf16 noise."""
    import copy

    E, M = 256, 512
    mod = make_module(E, seed=11).train()
    ref = copy.deepcopy(mod)
    x = (torch.rand(1, M, 3, device=DEV) * 10 - 5).to(torch.bfloat16)
    tok = torch.randn(1, M, E, device=DEV).to(torch.bfloat16).requires_grad_(True)
    tok_r = tok.detach().clone().requires_grad_(True)
    w = torch.randn(1, M, E, device=DEV).to(torch.bfloat16)
    out = mod(x, add_to=tok)
    assert out.requires_grad and out.dtype == torch.bfloat16
    (out.float() * w.float()).sum().backward()
    out_r = tok_r + ref.mlp(x)
    (out_r.float() * w.float()).sum().backward()
    assert torch.equal(tok.grad, tok_r.grad)  # identity on the tokens
    for (n, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        d = (p.grad.float() - q.grad.float()).abs().max().item()
        s_ = q.grad.float().abs().max().item()
        assert d <= 0.05 * s_ + 1e-3, (n, d, s_)
    # the forward went through the library, not through nn.Sequential
    with torch.no_grad():
        assert_close_bf16(out.detach(), torch_reference(mod, x, tok.detach()), tok.detach())


def test_pe_mlp_under_torch_compile():
    """`--torch_compile True` (scripts/3d/train/train_multi.sh:75): the custom op is an opaque node with a fake
    implementation, so dynamo traces through the module without touching ctypes."""
    E, M = 128, 256
    mod = make_module(E, seed=13)
    x = (torch.rand(1, M, 3, device=DEV) * 10 - 5).to(torch.bfloat16)
    tok = torch.randn(1, M, E, device=DEV).to(torch.bfloat16)
    with torch.no_grad():
        eager = mod(x, add_to=tok)
        compiled = torch.compile(mod, backend="aot_eager", fullgraph=True)(x, add_to=tok)
    assert torch.equal(eager, compiled)
