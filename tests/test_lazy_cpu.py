"""CPU: mechanics of the deferred values (lazy.py) and of batched lazy handles (video_utils.LazyWorldCoords) that do not
need a device -- operator routing, shapes/dtypes, per-scene clamps, the fake (meta) implementations of the custom ops."""
import pytest
import torch

from video_3d_llm_b200 import lazy, ops, position_encoding  # noqa: F401
from video_3d_llm_b200 import video_utils as VU


class Const(lazy.DeferredTensor):
    """A deferred value whose materialisation is a given CPU tensor (counts how often it is asked for)."""

    def __init__(self, t):
        self._t, self.n = t, 0

    shape = property(lambda self: self._t.shape)
    dtype = property(lambda self: self._t.dtype)
    device = property(lambda self: self._t.device)

    def materialize(self):
        self.n += 1
        return self._t


def test_deferred_tensor_behaves_like_its_value():
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    c = Const(t)
    assert c.shape == (2, 3, 4) and c.size(1) == 3 and c.dim() == 3 and c.ndim == 3 and len(c) == 2 and c.n == 0
    assert torch.equal(c + 1, t + 1) and torch.equal(1 + c, t + 1) and torch.equal(c * 2, t * 2) and torch.equal(-c, -t)
    x = torch.ones(2, 3, 4)
    assert torch.equal(x + c, x + t) and torch.equal(torch.add(x, c), x + t) and torch.equal(torch.cat([x, c]), torch.cat([x, t]))
    assert torch.equal(c.flatten(1, 2), t.flatten(1, 2)) and torch.equal(c[1, :2], t[1, :2]) and torch.equal(torch.stack(list(c)), t)
    assert torch.equal(c.reshape(-1, 4), t.reshape(-1, 4)) and c.detach() is c and not c.requires_grad
    # index assignment with a deferred value (llava_arch.py:726: new_coords[:, :, :W, :] = coords)
    z = torch.zeros(2, 3, 5)
    z[:, :, :4] = c
    assert torch.equal(z[:, :, :4], t) and bool((z[:, :, 4] == 0).all())
    y = torch.ones(2, 3, 4)
    y += c
    assert torch.equal(y, t + 1)
    assert torch.equal(c, t) and float(torch.max(c)) == 23.0


def _handle(V=2, clamp=None, dtype=torch.float32):
    return VU.LazyWorldCoords(torch.zeros((V, 480, 640), dtype=torch.int16), torch.eye(4).repeat(V, 1, 1), torch.eye(4).repeat(V, 1, 1),
                              clamp=clamp, dtype=dtype)


def test_batched_handle_keeps_one_clamp_per_scene():
    """under a '-norm' strategy every scene is clamped to ITS OWN mesh AABB before stacking (video_utils.py:232-234)."""
    a, b, c = _handle(clamp=([-1, -2, -3], [1, 2, 3])), _handle(clamp=([-4, -5, -6], [4, 5, 6])), _handle(clamp=None)
    m = VU.merge_video_dict([{"world_coords": h, "box_input": None} for h in (a, b, c)])["world_coords"]
    assert m.batched and m.shape == (3, 2, 384, 384, 3) and isinstance(m.clamp, list)
    kids = list(m)
    assert [k.clamp for k in kids] == [a.clamp, b.clamp, None] and m[1].clamp == b.clamp and not kids[0].batched
    assert [k.clamp for k in m.half().to("cpu")] == [a.clamp, b.clamp, None] and m.half().dtype == torch.float16
    # all-None clamps collapse to None; handles that cannot be stacked are refused
    assert VU.merge_video_dict([{"world_coords": _handle(), "box_input": None}] * 2)["world_coords"].clamp is None
    with pytest.raises(ValueError):
        VU.merge_video_dict([{"world_coords": _handle(), "box_input": None}, {"world_coords": _handle(dtype=torch.float16), "box_input": None}])
    with pytest.raises(TypeError):
        VU.LazyWorldCoords(torch.zeros((2, 2, 480, 640), dtype=torch.int16), torch.eye(4).repeat(2, 2, 1, 1), torch.eye(4).repeat(2, 2, 1, 1),
                           clamp=([0, 0, 0], [1, 1, 1]), batched=True)


def test_deferred_chain_launches_nothing_until_consumed():
    from video_3d_llm_b200 import coords as C

    h = _handle(V=3, dtype=torch.bfloat16)
    p = C.average_coordinate_in_patch(h)
    assert isinstance(p, lazy.LazyPooled) and p.shape == (3, 14, 14, 3) and p.dtype == torch.bfloat16
    d = C.discrete_coords(p, None, voxel_size=0.1, min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5])
    assert isinstance(d, lazy.LazyPooled) and d.disc["voxel_size"] == 0.1 and p.disc is None
    f = d.flatten(1, 2).detach()
    assert f.shape == (3, 196, 3) and f.flat
    pe = position_encoding.PositionEmbeddingSine3D(3584)(f)
    assert isinstance(pe, lazy.LazyPE) and pe.shape == (3, 196, 3584) and pe.dtype == torch.bfloat16
    assert C.sample_n_points(h, 9).shape == (3, 14, 14, 9, 3) and C.sample_n_points(h, 9).flatten(1, 2).shape == (3, 196, 9, 3)
    assert C.minmax_coordinate_in_patch(h).shape == (3, 14, 14, 2, 3) and C.average_coordinate_in_patch(h, 14).shape == (3, 27, 27, 3)
    pe9 = position_encoding.PositionEmbeddingSine3D(3584, n_points=9)(C.sample_n_points(h, 9).flatten(1, 2))
    assert pe9.shape == (3, 196, 3584)
    # deferred pooled tokens: `+ PE` keeps deferring when shapes / dtypes allow the fused kernel
    tok = torch.zeros(3, 729, 3584, dtype=torch.bfloat16)
    pool = lazy.LazyTokenPool(tok, 2, "bilinear")
    assert pool.shape == (3, 196, 3584) and pool.dtype == torch.bfloat16 and pool.out_grid == 14
    assert lazy.LazyTokenPool(tok, 2, "average").out_grid == 13
    assert pool.with_pe(pe) is None  # CPU tokens: no fused path (and materialising them raises: no CPU path)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pool.materialize()


def test_custom_ops_have_fake_implementations():
    """torch.compile / make_fx trace the ops through their fake implementations: shapes and dtypes without the library."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        x = torch.empty(2, 196, 3, dtype=torch.bfloat16, device="cuda")
        tok = torch.empty(2, 196, 3584, dtype=torch.bfloat16, device="cuda")
        out = torch.ops.v3d.pe_sine3d(x, tok, 3584, 1, 10000.0, 512, torch.bfloat16)
        assert out.shape == (2, 196, 3584) and out.dtype == torch.bfloat16 and out.device.type == "cuda"
        w1, b1, w2, b2 = (torch.empty(s, dtype=torch.bfloat16, device="cuda") for s in ((512, 3), (512,), (1152, 512), (1152,)))
        o2 = torch.ops.v3d.pe_mlp(torch.empty(64, 3, device="cuda"), w1, b1, b1, b1, 1e-5, w2, b2, None)
        assert o2.shape == (64, 1152) and o2.dtype == torch.bfloat16
        t729 = torch.empty(2, 729, 64, dtype=torch.float16, device="cuda")
        nl = torch.empty(64, dtype=torch.float16, device="cuda")
        assert torch.ops.v3d.pool_pe_grid(t729, 2, 0, None, 1, 10000.0, 0, nl).shape == (2, 14, 15, 64)
        assert torch.ops.v3d.pool_pe_grid(t729, 2, 1, None, 1, 10000.0, 0, None).shape == (2, 13, 13, 64)
        assert torch.ops.v3d.pool_pe_grid(t729, 2, 3, None, 1, 10000.0, 0, nl).shape == (2, 27, 28, 64)
    # autograd formulas are registered: identity on the tokens for the PE ops
    for name in ("frontend_pe", "pe_sine3d", "pe_mlp", "pool_pe_grid", "frontend_coords"):
        assert hasattr(torch.ops.v3d, name)
