"""GPU: the path the UNCHANGED reference scripts execute under the overlay launcher.

The reference's own `prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:336-836; inherited unmodified from
/root/reference, or from the copy oracle/make_ref.py put under oracle/_ref/ on the GPU box) is driven twice on the same
on-disk ScanNet-shaped sample:

  reference arm   the reference's VideoProcessor + merge_video_dict + its own helpers / PE module (torch ops)
  overlay arm     launch.install(): the package's VideoProcessor(lazy=True), merge_video_dict, patched helpers; the
                  method's call sequence must end in the fused kernels (checked through the library's launch log)

and the resulting input embeddings must agree.  SURVEY.md section 4 item 3 / VERDICT r1 "Missing #1, #3".
"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import ref_import  # noqa: E402

if not ref_import.available():
    pytest.skip("reference sources neither mounted nor copied (python oracle/make_ref.py)", allow_module_level=True)

import fixtures  # noqa: E402
import ref_harness  # noqa: E402
from video_3d_llm_b200 import launch  # noqa: E402
from video_3d_llm_b200 import video_utils as VU  # noqa: E402

DEV = torch.device("cuda:0")
VIDEO = "scannet/scene0003_00"  # 480x640, 4 frames


def _vp_kwargs(root, strategy):
    return dict(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"), voxel_size=0.1,
                min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5], frame_sampling_strategy=strategy)


def _sample(vp, merge, frames=4, box_input=None):
    vd = vp.process_3d_video(VIDEO, fixtures.FakeImageProcessor(384), force_sample=True, frames_upbound=frames)
    images = vd.pop("images")
    vd.pop("video_size")
    vd["box_input"] = box_input
    return images, merge([vd])


class LaunchLog:
    """Names of the library entry points called while active (wraps the ctypes functions of the loaded library)."""

    NAMES = ("v3d_frontend_pe", "v3d_pe_sine3d", "v3d_pool_pe_grid", "v3d_unproject", "v3d_pool_coords", "v3d_discretize", "v3d_pe_mlp",
             "v3d_object_patch_masks_depth", "v3d_object_patch_masks_depth_bounds", "v3d_object_patch_masks", "v3d_masked_token_mean")

    def __enter__(self):
        from video_3d_llm_b200 import cabi

        self.lib, self.calls, self.saved = cabi.load(), [], {}
        for n in self.NAMES:
            fn = getattr(self.lib, n)
            self.saved[n] = fn

            def wrapped(*a, _fn=fn, _n=n):
                self.calls.append(_n)
                return _fn(*a)

            setattr(self.lib, n, wrapped)
        return self

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.lib, n, fn)


def _arms(dataset, monkeypatch, wpe, dtype, strategy="uniform", pool_mode="bilinear", newline_position="grid", E=256, use_objects=False,
          object_feature_type="patch14"):
    """Returns (reference result on the GPU, overlay result, launch log of the overlay arm)."""
    monkeypatch.chdir(dataset)  # the reference hard-codes data/metadata relative to the working directory
    arch = ref_import.llava_arch()
    launch.uninstall()
    ref_vu = ref_import.video_utils()
    ref_pe = ref_import.position_encoding()
    # ---- reference arm: the reference's loader on the host, its model-side code as torch ops on the GPU
    images, vd = _sample(ref_vu.VideoProcessor(**_vp_kwargs(dataset, strategy)), ref_vu.merge_video_dict)
    model = ref_harness.build_model(arch, ref_pe, E=E, wpe=wpe, pool_mode=pool_mode, newline_position=newline_position,
                                    object_feature_type=object_feature_type).to(DEV).to(dtype)
    vd_ref = {k: (v.to(DEV).to(dtype) if torch.is_floating_point(v) else v.to(DEV)) for k, v in vd.items()}
    with torch.no_grad():
        ref = ref_harness.run(model, images.to(DEV).to(dtype), vd_ref, use_object_proposals=use_objects, device=DEV)
    # ---- overlay arm: same weights, the package underneath the same method
    launch.install(lazy=True)
    try:
        pe_mod = importlib.import_module("llava.model.position_encoding")
        model2 = ref_harness.build_model(arch, pe_mod, E=E, wpe=wpe, pool_mode=pool_mode, newline_position=newline_position,
                                         object_feature_type=object_feature_type).to(DEV).to(dtype)
        model2.load_state_dict(model.state_dict())
        vu = importlib.import_module("llava.video_utils")
        images2, vd2 = _sample(vu.VideoProcessor(**_vp_kwargs(dataset, strategy)), vu.merge_video_dict)
        assert isinstance(vd2["world_coords"], VU.LazyWorldCoords)
        # what the eval drivers do to every entry (model_scanqa.py:164-165); the HF Trainer moves tensors only
        vd2 = {k: (v.to(dtype).to(DEV) if k == "world_coords" or torch.is_floating_point(v) else v.to(DEV)) for k, v in vd2.items()}
        assert torch.equal(images2, images)
        with LaunchLog() as log, torch.no_grad():
            got = ref_harness.run(model2, images2.to(DEV).to(dtype), vd2, use_object_proposals=use_objects, device=DEV)
            torch.cuda.synchronize()
    finally:
        launch.uninstall()
    return ref, got, log.calls


def _assert_embeds(ref, got, dtype, frac=2e-3):
    e_ref, l_ref = ref[0], ref[1]
    e_got, l_got = got[0], got[1]
    assert e_got.shape == e_ref.shape and e_got.dtype == e_ref.dtype == dtype
    assert torch.equal(l_got, l_ref)
    d = (e_got.float() - e_ref.float()).abs()
    scale = e_ref.float().abs().clamp(min=1.0)
    ulp = {torch.float32: 2.0 ** -20, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}[dtype]
    # pooled tokens: our bilinear taps are ATen's, the fp32 evaluation order is not (within one ulp of the token dtype,
    # twice that after the PE add); PE: sin/cos of CUDA libm on both arms here -> the same values
    bad = d > 4 * ulp * scale
    assert not bool(bad.any()), (f"{int(bad.sum())} of {d.numel()} elements differ by more than 4 ulp: max diff {d.max().item():.3e} at "
                                 f"ref {e_ref.flatten()[d.argmax()].item():.4f}, first bad index {torch.nonzero(bad)[0].tolist()}")
    nz = float((d > 0).float().mean())
    assert nz <= (frac if dtype != torch.float32 else 1.0), f"{nz:.3e} of the elements differ"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_default_recipe_through_the_reference_method(dataset, monkeypatch, dtype):
    """avg-discrete-sin3d, bilinear 2x2 pool, grid newlines (scripts/3d/train/train_multi.sh:82-87): two launches --
    depth -> discretised patch coordinates, then pool + PE + add + newline over the projector's tokens."""
    ref, got, calls = _arms(dataset, monkeypatch, "avg-discrete-sin3d", dtype)
    assert got[0].shape[1] == 3 + 2 + 4 * 14 * 15  # text + 4 frames x 14 rows x (14 tokens + newline)
    _assert_embeds(ref, got, dtype)
    assert calls == ["v3d_frontend_pe", "v3d_pool_pe_grid"], calls


@pytest.mark.parametrize("wpe", ["avg-discrete-sin3d", "avg-sin3d", "sample9-discrete-sin3d", "sample5-sin3d", "sample1-discrete-sin3d", "minmax-discrete-sin3d"])
def test_pooling_variants_without_newlines(dataset, monkeypatch, wpe):
    """mm_newline_position='no_token': the `+` of llava_arch.py:515 is the last thing that happens to the tokens, so the
    deferred pooled tokens are materialised by the pooling kernel with the PE fused in (two launches again)."""
    ref, got, calls = _arms(dataset, monkeypatch, wpe, torch.bfloat16, newline_position="no_token", E=288)
    _assert_embeds(ref, got, torch.bfloat16, frac=5e-3)
    assert calls == ["v3d_frontend_pe", "v3d_pool_pe_grid"], calls


def test_norm_strategy_and_mc_frames(dataset, monkeypatch):
    """mc-ratio95-norm: MC frame list from the offline JSON + per-scene clamp to the mesh AABB (video_utils.py:232-234)."""
    ref, got, calls = _arms(dataset, monkeypatch, "avg-discrete-sin3d", torch.bfloat16, strategy="mc-ratio95-norm")
    _assert_embeds(ref, got, torch.bfloat16)
    assert calls == ["v3d_frontend_pe", "v3d_pool_pe_grid"], calls


def test_mlp_pe_and_object_proposals(dataset, monkeypatch):
    """'mlp' PE (tcgen05 kernel behind the module) and the object-proposal branch (llava_arch.py:351-378, 479-503), which
    slices the coordinate tensor: the lazy handle materialises on demand and the result still matches."""
    ref, got, calls = _arms(dataset, monkeypatch, "avg-mlp", torch.bfloat16, use_objects=True)
    e_ref, e_got = ref[0].float(), got[0].float()
    # the reference's bf16 module rounds the hidden activations after every layer, the kernel keeps them in fp32 until
    # the tensor-core operand: compare at the PE's own scale
    assert float((e_got - e_ref).abs().max()) <= 2.0 ** -5 * float(e_ref.abs().max())
    assert torch.equal(got[1], ref[1])
    o_ref, o_got = ref[2].float(), got[2].float()
    assert o_got.shape == o_ref.shape and float((o_got - o_ref).abs().max()) <= 2.0 ** -5 * max(1.0, float(o_ref.abs().max()))
    assert "v3d_pe_mlp" in calls


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("object_feature_type", ["patch14", "patch27"])
def test_object_proposal_loop_ends_in_one_mask_launch(dataset, monkeypatch, object_feature_type, dtype):
    """The reference's inlined object loop (llava_arch.py:351-378) on a lazy handle: slice, reshape, transpose, flatten, the two
    comparisons, all / sum / >= are recorded (lazy_objects.py) and the first `.view()` of a mask (:485 / :487) evaluates every
    object's mask in ONE v3d_object_patch_masks_depth_bounds launch straight from depth -- no (V,384,384,3) tensor
    (no v3d_unproject), and object features equal to the reference arm's."""
    ref, got, calls = _arms(dataset, monkeypatch, "avg-discrete-sin3d", dtype, use_objects=True, object_feature_type=object_feature_type)
    assert calls.count("v3d_object_patch_masks_depth_bounds") == 1, calls
    assert "v3d_unproject" not in calls, calls
    _assert_embeds(ref, got, dtype)
    o_ref, o_got = ref[2].float(), got[2].float()
    assert o_got.shape == o_ref.shape
    tol = {torch.float32: 2.0 ** -18, torch.bfloat16: 2.0 ** -6}[dtype]
    assert float((o_got - o_ref).abs().max()) <= tol * max(1.0, float(o_ref.abs().max()))
    assert float(o_ref.abs().max()) > 0  # some object selected some patches


def test_mrope_positions_from_deferred_coordinates(dataset, monkeypatch):
    """'mrope' consumes the discretised coordinates as 3-D position ids (llava_arch.py:711-729, 795-805): the deferred value
    is written into a tensor by index assignment."""
    ref, got, calls = _arms(dataset, monkeypatch, "avg-discrete-sin3d-mrope", torch.float32)
    assert got[3] is not None and torch.equal(got[3], ref[3])
    _assert_embeds(ref, got, torch.float32)


def test_unpooled_tokens_take_the_single_fused_launch(dataset, monkeypatch):
    """Without spatial pooling of the tokens the whole front end is ONE launch of v3d_frontend_pe with tokens_in: the
    vision tokens (27x27 grid, 14-pixel windows -- BASELINE configs[1]) are read once and written once."""
    from video_3d_llm_b200 import coords as C
    from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D

    monkeypatch.chdir(dataset)
    ref_vu = ref_import.video_utils()
    arch_shim = ref_import.ArchShim()
    ref_pe = ref_import.position_encoding().PositionEmbeddingSine3D(1152)
    images, vd = _sample(ref_vu.VideoProcessor(**_vp_kwargs(dataset, "uniform")), ref_vu.merge_video_dict)
    tokens = torch.randn(4, 729, 1152, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(DEV)
    wc = vd["world_coords"].to(DEV).to(torch.bfloat16)
    pooled = [arch_shim.average_coordinate_in_patch(c, patch_size=14) for c in wc]
    disc = [arch_shim.discrete_coords(c, None) for c in pooled]
    ref = tokens + ref_pe(disc[0].flatten(1, 2).detach())
    vp = VU.VideoProcessor(**_vp_kwargs(dataset, "uniform"), lazy=True, metadata_dir=os.path.join(dataset, "data", "metadata"))
    _, vd2 = _sample(vp, VU.merge_video_dict)
    h = vd2["world_coords"].to(torch.bfloat16).to(DEV)
    pe = PositionEmbeddingSine3D(1152)
    with LaunchLog() as log:
        p2 = [C.average_coordinate_in_patch(c, patch_size=14) for c in h]
        d2 = [C.discrete_coords(c, None, div_mode=1) for c in p2]
        got = tokens + pe(d2[0].flatten(1, 2).detach())
        torch.cuda.synchronize()
    assert log.calls == ["v3d_frontend_pe"], log.calls
    assert got.dtype == torch.bfloat16 and torch.equal(got.view(torch.int16), ref.view(torch.int16))


def test_training_step_gradients_flow_through_the_overlay(dataset, monkeypatch):
    """The projector and image_newline are trained (train_multi.sh mm_tunable_parts): gradients through the fused pool +
    PE + newline kernel equal those of the reference's operator sequence."""
    monkeypatch.chdir(dataset)
    arch = ref_import.llava_arch()
    launch.uninstall()
    ref_vu, ref_pe = ref_import.video_utils(), ref_import.position_encoding()
    images, vd = _sample(ref_vu.VideoProcessor(**_vp_kwargs(dataset, "uniform")), ref_vu.merge_video_dict)
    model = ref_harness.build_model(arch, ref_pe, E=256).to(DEV).to(torch.bfloat16)
    vd_ref = {k: (v.to(DEV).to(torch.bfloat16) if torch.is_floating_point(v) else v.to(DEV)) for k, v in vd.items()}
    w = torch.randn(1, 3 + 2 + 4 * 210, 256, device=DEV)
    emb = ref_harness.run(model, images.to(DEV).to(torch.bfloat16), vd_ref, device=DEV)[0]
    (emb.float() * w).sum().backward()
    launch.install(lazy=True)
    try:
        pe_mod = importlib.import_module("llava.model.position_encoding")
        vu = importlib.import_module("llava.video_utils")
        model2 = ref_harness.build_model(arch, pe_mod, E=256).to(DEV).to(torch.bfloat16)
        model2.load_state_dict(model.state_dict())
        _, vd2 = _sample(vu.VideoProcessor(**_vp_kwargs(dataset, "uniform")), vu.merge_video_dict)
        vd2 = {k: (v.to(torch.bfloat16).to(DEV) if k == "world_coords" or torch.is_floating_point(v) else v.to(DEV)) for k, v in vd2.items()}
        emb2 = ref_harness.run(model2, images.to(DEV).to(torch.bfloat16), vd2, device=DEV)[0]
        assert emb2.requires_grad
        (emb2.float() * w).sum().backward()
    finally:
        launch.uninstall()
    for (n, p), (_, q) in zip(model.named_parameters(), model2.named_parameters()):
        if p.grad is None:
            assert q.grad is None, n
            continue
        assert q.grad is not None, n
        d, s = float((p.grad.float() - q.grad.float()).abs().max()), float(p.grad.float().abs().max())
        assert d <= 0.02 * s + 1e-4, (n, d, s)


def test_deferred_handles_behave_like_the_reference_tensors(dataset, monkeypatch):
    """Anything else done to the deferred values yields the tensor the reference would hold (same shape, dtype, values)."""
    from video_3d_llm_b200 import coords as C

    monkeypatch.chdir(dataset)
    vp_l = VU.VideoProcessor(**_vp_kwargs(dataset, "uniform"), lazy=True, metadata_dir=os.path.join(dataset, "data", "metadata"))
    vp_e = VU.VideoProcessor(**_vp_kwargs(dataset, "uniform"), lazy=False, metadata_dir=os.path.join(dataset, "data", "metadata"))
    _, ml = _sample(vp_l, VU.merge_video_dict)
    _, me = _sample(vp_e, VU.merge_video_dict)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        hl, we = ml["world_coords"].to(DEV).to(dt), me["world_coords"].to(DEV).to(dt)
        for fn, kw in ((C.average_coordinate_in_patch, {}), (C.average_coordinate_in_patch, {"patch_size": 14}), (C.minmax_coordinate_in_patch, {}),
                       (C.sample_n_points, {"n_points": 9}), (C.sample_n_points, {"n_points": 5}), (C.sample_n_points, {"n_points": 1})):
            lazy_v, eager_v = fn(hl[0], **kw), fn(we[0], **kw)
            assert lazy_v.shape == eager_v.shape and lazy_v.dtype == eager_v.dtype == dt
            assert torch.equal(lazy_v.materialize(), eager_v), (fn.__name__, kw, dt)
            for mode in (0, 1):
                dl, de = C.discrete_coords(lazy_v, None, div_mode=mode), C.discrete_coords(eager_v, None, div_mode=mode)
                assert torch.equal(dl.materialize(), de), (fn.__name__, kw, dt, mode)
                # and equal to torch's own element-wise kernels on this device for the reciprocal mode (what the reference runs)
                if mode == 1:
                    lo, hi = torch.tensor([-15, -15, -5], device=DEV), torch.tensor([15, 15, 5], device=DEV)
                    t = ((torch.minimum(torch.maximum(eager_v, lo), hi) - lo) / 0.1).round()
                    assert torch.equal(de, t), (fn.__name__, kw, dt)
        assert torch.equal(torch.stack(list(C.average_coordinate_in_patch(hl[0]))), C.average_coordinate_in_patch(we[0]))
