"""GPU: the projector-tail kernel (get_2dPool -> PE add -> add_token_per_grid) through the
host mirror -> ctypes -> C ABI, against the CPU restatement of the reference's operator sequence
(oracle/torch_port.py, pinned on the reference's own methods in tests/test_oracle_golden.py) and the golden vectors.

Bar: newline rows, max pooling and the grid layout bit-exact; pooled / PE-added bf16 values equal except for a small
fraction of elements that differ by one bf16 ulp (fp32 evaluation order of the bilinear taps, CUDA vs SLEEF sin/cos)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import torch_port as TP  # noqa: E402
from test_gpu_parity import assert_bf16_close, bf16_bits  # noqa: E402
from video_3d_llm_b200 import cabi, projector_tail as PT, synthetic  # noqa: E402
from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D  # noqa: E402

DEV = torch.device("cuda:0")


def assert_pooled_close(got: torch.Tensor, ref: torch.Tensor, max_frac=5e-4):
    """Pooled tokens: equal except for a small fraction of elements that differ by at most one bf16/fp16 ulp of the
    reference value -- or, where the four taps cancel, by the fp32 rounding error of the weighted sum (taps are
    N(0,1): 4 operations x 2^-24 x |tap| < 2^-19)."""
    g, r = got.cpu().float(), ref.float()
    d = (g - r).abs()
    rel = 2.0 ** -7 if ref.dtype == torch.bfloat16 else 2.0 ** -10
    tol = torch.maximum(r.abs() * rel, torch.tensor(2.0 ** -19))
    assert bool((d <= tol).all()), f"max error {float((d / tol).max()):.2f} x tolerance"
    assert float((d > 0).float().mean()) <= max_frac, f"{float((d > 0).float().mean()):.3e} of the elements differ"


TAIL_CASES = ["bilinear_1152", "bilinear_3584", "average_1152", "max_1152", "bilinear_16_264"]


def _case(golden_dir, tag):
    G = np.load(os.path.join(golden_dir, "tail.npz"))
    V, side, E, Go = (int(v) for v in G[f"{tag}_meta"])
    tokens = synthetic.make_tokens(70, V, side * side, E)
    coords = torch.from_numpy(G[f"{tag}_coords"]).to(torch.bfloat16)
    newline = torch.from_numpy(G[f"{tag}_newline"]).view(torch.bfloat16)
    return G, V, side, E, Go, tokens, coords, newline, tag.split("_")[0]


@pytest.mark.parametrize("tag", TAIL_CASES)
def test_pool_pe_newline_matches_reference(golden_dir, tag):
    G, V, side, E, Go, tokens, coords, newline, mode = _case(golden_dir, tag)
    ref_pooled = TP.get_2dpool(tokens, 2, mode)
    ref = TP.projector_tail(tokens, coords, E, newline, 2, mode)
    # pooling alone, under the reference's name
    got_pooled = PT.get_2dPool(tokens.to(DEV), 2, mode)
    assert got_pooled.shape == ref_pooled.shape and got_pooled.dtype == torch.bfloat16
    if mode == "max":
        assert torch.equal(got_pooled.cpu(), ref_pooled)
    else:
        assert_pooled_close(got_pooled, ref_pooled)
    # the fused tail
    pe = PositionEmbeddingSine3D(E)
    got = PT.pool_pe_newline(tokens.to(DEV), stride=2, mode=mode, coords=coords.to(DEV), position_embedding=pe, image_newline=newline.to(DEV))
    assert got.shape == (V * Go * (Go + 1), E) and got.dtype == torch.bfloat16
    # a pooled value on the other side of a bf16 tie moves the sum by one ulp OF THE POOLED VALUE (the PE is <= 1 in
    # magnitude); where the two nearly cancel that is several ulps of the small result
    g4 = got.view(V, Go, Go + 1, E)
    r4 = ref.view(V, Go, Go + 1, E)
    d = (g4[:, :, :Go].cpu().float() - r4[:, :, :Go].float()).abs().reshape(V, Go * Go, E)
    ulp = torch.maximum(ref_pooled.float().abs(), torch.tensor(1.0)) * 2.0 ** -7
    assert bool((d <= ulp).all()), f"max error {float((d / ulp).max()):.2f} ulp of the pooled value"
    assert float((d > 0).float().mean()) <= 1e-3
    assert torch.equal(g4[:, :, Go, :].cpu(), newline.expand(V, Go, E))  # add_token_per_grid: newline closes every row
    # against the vectors written by the reference itself
    gold = torch.from_numpy(G[f"{tag}_out_sample"].copy()).view(torch.bfloat16).float()
    dg = (got.cpu()[::7, ::5].float() - gold).abs()
    assert float((dg > 0).float().mean()) <= 2e-3 and float(dg.max()) <= 2.0 ** -7 * 8
    # the two-step composition (pool kernel, then the PE module's fused add, then the newline kernel) agrees exactly
    two_step = PT.add_token_per_grid(pe(coords.to(DEV), add_to=got_pooled), newline.to(DEV))
    assert torch.equal(two_step, got)


def test_tail_other_dtypes_layouts_and_continuous_coords():
    V, side, E = 2, 12, 192
    tokens = synthetic.make_tokens(5, V, side * side, E)
    gen = torch.Generator().manual_seed(3)
    coords = torch.rand(V, 36, 3, generator=gen) * 20  # non-integer: the PE is evaluated directly, no table rows
    coords[0, :5] = torch.tensor([3.0, 7.0, 250.0])  # integer-valued: served from the table
    newline = torch.randn(E, generator=gen)
    for dt, frac in ((torch.float32, None), (torch.float16, 2e-3), (torch.bfloat16, 2e-3)):
        t, c, nl = tokens.to(dt), coords.to(dt), newline.to(dt)
        ref = TP.projector_tail(t, c, E, nl, 2, "bilinear")
        got = PT.pool_pe_newline(t.to(DEV), coords=c.to(DEV), position_embedding=PositionEmbeddingSine3D(E), image_newline=nl.to(DEV))
        assert got.dtype == dt and got.shape == ref.shape
        if dt == torch.float32:
            torch.testing.assert_close(got.cpu(), ref, rtol=0, atol=2e-6)
        else:
            d = (got.cpu().float() - ref.float()).abs()
            ulp = torch.maximum(ref.float().abs(), torch.tensor(1e-3)) * (2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10)
            assert bool((d <= ulp).all()) and float((d > 0).float().mean()) < frac
    # no newline, no PE, not flattened: (V, Go*Go, E)
    out = PT.pool_pe_newline(tokens.to(DEV), stride=3, mode="bilinear")
    assert out.shape == (V, 16, E)
    ref = TP.get_2dpool(tokens, 3, "bilinear")
    assert_pooled_close(out, ref)
    # newline only (add_token_per_grid on its own) is a pure copy
    nl = newline.to(torch.bfloat16)
    assert torch.equal(PT.add_token_per_grid(tokens.to(DEV), nl.to(DEV)).cpu(), TP.add_token_per_grid(tokens, nl))
    # NaN propagation of max pooling (a later NaN replaces the maximum, like ATen)
    tn = tokens.clone()
    tn[0, 13, 7] = float("nan")
    a, b = PT.get_2dPool(tn.to(DEV), 2, "max").cpu(), TP.get_2dpool(tn, 2, "max")
    assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


def test_tail_errors_and_empty():
    tokens = synthetic.make_tokens(1, 1, 36, 64).to(DEV)
    with pytest.raises(ValueError):
        PT.pool_pe_newline(tokens, mode="median")
    with pytest.raises(ValueError):
        PT.pool_pe_newline(tokens[:, :35])
    with pytest.raises(ValueError):
        PT.pool_pe_newline(tokens, coords=torch.zeros(1, 9, 3, device=DEV))
    with pytest.raises(NotImplementedError):
        PT.pool_pe_newline(tokens, stride=3, mode="average")
    with pytest.raises(RuntimeError):
        PT.pool_pe_newline(tokens.cpu())
    assert PT.pool_pe_newline(tokens[:0]).shape == (0, 9, 64)


def test_tail_full_size_properties():
    """Default-recipe size (16 scenes x 32 frames, 27x27 -> 14x14, E = 3584): properties that hold without the oracle."""
    V, side, E, Go = 512, 27, 3584, 14
    g = torch.Generator(device=DEV).manual_seed(11)
    tokens = torch.randn((V, side * side, E), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    coords = torch.randint(0, 101, (V, Go * Go, 3), generator=g, device=DEV).to(torch.bfloat16)
    newline = torch.randn(E, generator=g, device=DEV).to(torch.bfloat16)
    pe = PositionEmbeddingSine3D(E)
    out = PT.pool_pe_newline(tokens, coords=coords, position_embedding=pe, image_newline=newline)
    assert out.shape == (V * Go * (Go + 1), E)
    assert torch.equal(out, PT.pool_pe_newline(tokens, coords=coords, position_embedding=pe, image_newline=newline))  # deterministic
    o4 = out.view(V, Go, Go + 1, E)
    assert torch.equal(o4[:, :, Go], newline.expand(V, Go, E))
    # the fused result is the composition of its parts: pooled tokens + the module's PE, rounded once more
    pooled = PT.get_2dPool(tokens, 2, "bilinear")
    assert torch.equal(o4[:, :, :Go].reshape(V, Go * Go, E), pooled + pe(coords))
    # identity pooling + newline is a copy; a constant image pools to the same constant (weights sum to one)
    ident = PT.pool_pe_newline(tokens[:8], mode="none", image_newline=newline).view(8, side, side + 1, E)
    assert torch.equal(ident[:, :, :side].reshape(8, side * side, E), tokens[:8])
    const = torch.full((4, side * side, E), 0.625, device=DEV, dtype=torch.bfloat16)
    for mode in ("bilinear", "average", "max"):
        assert torch.equal(PT.get_2dPool(const, 2, mode), torch.full((4, (13 if mode != "bilinear" else 14) ** 2, E), 0.625, device=DEV, dtype=torch.bfloat16))
    # first frames against the CPU restatement of the reference
    ref = TP.projector_tail(tokens[:2].cpu(), coords[:2].cpu(), E, newline.cpu(), 2, "bilinear").view(2, Go, Go + 1, E)
    d = (o4[:2].cpu().float() - ref.float()).abs()
    assert float((d > 0).float().mean()) <= 1e-3 and float(d.max()) <= 2.0 ** -7 * 8


@pytest.mark.parametrize("mode", ["bilinear", "average"])
def test_pool_commutes_with_the_last_linear_of_the_projector(mode):
    """projector_tail.project_pool_pe_newline runs the projector's last Linear on the pooled tokens (3.7x fewer GEMM rows).
    In fp32 the two orders agree to rounding noise (the pool is a convex combination and the bias passes through); in bf16
    they differ by the roundings that moved: the result must stay within 4 bf16 ulps of the output scale of the reference
    order (mm_projector -> get_2dPool -> + PE -> add_token_per_grid), element-wise, and agree to 2^-8 of that scale in the
    mean."""
    from video_3d_llm_b200 import projector_tail as PT
    from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D

    torch.manual_seed(11)
    V, side, C, E = 3, 27, 256, 1152
    Go = 14 if mode == "bilinear" else 13  # math.ceil(27 / 2) for the interpolation, 27 // 2 for the 2x2 windows (llava_arch.py:196-206)
    proj32 = torch.nn.Sequential(torch.nn.Linear(C, E), torch.nn.GELU(), torch.nn.Linear(E, E)).to(DEV).eval()
    x32 = torch.randn((V, side * side, C), device=DEV)
    coords = torch.randint(0, 101, (V, Go * Go, 3), device=DEV).float()
    newline32 = torch.randn(E, device=DEV)
    pe = PositionEmbeddingSine3D(E)
    with torch.no_grad():
        # fp32: algebraic identity
        a = PT.pool_pe_newline(proj32(x32), stride=2, mode=mode, coords=coords, position_embedding=pe, image_newline=newline32)
        b = PT.project_pool_pe_newline(x32, proj32, stride=2, mode=mode, coords=coords, position_embedding=pe, image_newline=newline32)
        assert a.shape == b.shape == (V * Go * (Go + 1), E)
        scale = a.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-5 * scale
        # bf16: the recipe's dtype
        proj16 = torch.nn.Sequential(torch.nn.Linear(C, E), torch.nn.GELU(), torch.nn.Linear(E, E)).to(DEV).eval()
        proj16.load_state_dict(proj32.state_dict())
        proj16 = proj16.to(torch.bfloat16)
        x16, nl16, c16 = x32.to(torch.bfloat16), newline32.to(torch.bfloat16), coords.to(torch.bfloat16)
        ref = PT.pool_pe_newline(proj16(x16), stride=2, mode=mode, coords=c16, position_embedding=pe, image_newline=nl16).float()
        got = PT.project_pool_pe_newline(x16, proj16, stride=2, mode=mode, coords=c16, position_embedding=pe, image_newline=nl16).float()
        scale = ref.abs().max().item()
        err = (got - ref).abs()
        assert err.max().item() <= 4 * 2.0 ** -8 * scale, err.max().item() / scale
        assert err.mean().item() <= 2.0 ** -8 * scale
        # newline rows are untouched by the reordering
        rows = ref.view(V, Go, Go + 1, E)[:, :, Go]
        assert torch.equal(got.view(V, Go, Go + 1, E)[:, :, Go], rows)
        # 'max' does not commute: the reference order is kept
        cm = c16[:, : 13 * 13]
        m1 = PT.pool_pe_newline(proj16(x16), stride=2, mode="max", coords=cm, position_embedding=pe, image_newline=nl16)
        m2 = PT.project_pool_pe_newline(x16, proj16, stride=2, mode="max", coords=cm, position_embedding=pe, image_newline=nl16)
        assert torch.equal(m1, m2)
