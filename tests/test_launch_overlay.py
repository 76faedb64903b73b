"""CPU: the overlay launcher rebinds exactly the reference entry points it claims to (needs the reference
mounted at /root/reference; skipped on machines without it, e.g. the GPU box)."""
import pytest

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference not mounted")


def test_install_patches_reference_modules():
    import importlib

    arch = ref_import.llava_arch()  # registers the stub `llava` package (SURVEY 8c recipe)
    from video_3d_llm_b200 import launch, position_encoding, video_utils

    done = launch.install(lazy=True)
    vu = importlib.import_module("llava.video_utils")
    pe = importlib.import_module("llava.model.position_encoding")
    assert vu.unproject is video_utils.unproject and vu.merge_video_dict is video_utils.merge_video_dict
    assert issubclass(vu.VideoProcessor, video_utils.VideoProcessor)
    assert pe.PositionEmbeddingSine3D is position_encoding.PositionEmbeddingSine3D
    assert pe.PositionEmbeddingMLP is position_encoding.PositionEmbeddingMLP
    cls = arch.LlavaMetaForCausalLM
    for name in ("average_coordinate_in_patch", "minmax_coordinate_in_patch", "sample_n_points", "discrete_coords", "get_2dPool",
                 "add_token_per_grid"):
        assert getattr(cls, name) is done[f"llava.model.llava_arch.LlavaMetaForCausalLM.{name}"]
    # LlavaMetaModel.__init__ imports the PE classes lazily from the patched module (llava_arch.py:49)
    import types

    class Base:
        def __init__(self, config):
            self.config = config

    Model = type("Model", (arch.LlavaMetaModel, Base), {})
    m = Model(types.SimpleNamespace(world_position_embedding_type="avg-discrete-sin3d", hidden_size=3584))
    assert isinstance(m.world_position_embedding, position_encoding.PositionEmbeddingSine3D)
    m9 = Model(types.SimpleNamespace(world_position_embedding_type="sample9-mlp", hidden_size=96))
    assert isinstance(m9.world_position_embedding, position_encoding.PositionEmbeddingMLP) and m9.world_position_embedding.n_points == 9


def test_patched_projector_tail_methods():
    """get_2dPool / add_token_per_grid are rebound; like every other entry point of the package they refuse CPU tensors
    (no CPU path), and the slow-fast variant that the package does not implement keeps the reference's own method."""
    import types

    import torch

    arch = ref_import.llava_arch()
    from oracle import torch_port as TP
    from video_3d_llm_b200 import launch

    launch.install(lazy=True)
    cls = arch.LlavaMetaForCausalLM
    x = torch.randn(2, 36, 16).to(torch.bfloat16)
    newline = torch.randn(16).to(torch.bfloat16)
    shim = types.SimpleNamespace(config=types.SimpleNamespace(mm_spatial_pool_mode="bilinear", add_faster_video=False),
                                 model=types.SimpleNamespace(image_newline=newline),
                                 get_vision_tower=lambda: types.SimpleNamespace(num_patches_per_side=6))
    with pytest.raises(RuntimeError, match="no CPU path"):
        cls.get_2dPool(shim, x, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cls.add_token_per_grid(shim, x)
    shim.config.add_faster_video = True
    out = cls.add_token_per_grid(shim, x)  # the reference's slow-fast layout (V, G*(G+1), E)
    assert out.shape == (2, 6 * 7, 16) and torch.equal(out.flatten(0, 1), TP.add_token_per_grid(x, newline))
