"""TEST INFRASTRUCTURE -- a tiny stand-in model that drives the REFERENCE's own
`LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:336-836), as SURVEY.md section 8c
found feasible: a fake vision tower (V,3,384,384) -> (V,729,Dv), a 2-layer projector, the world position embedding,
`image_newline` and an embedding table.  The reference method is inherited unmodified (from /root/reference in the
build container, from the copy under oracle/_ref/ on the GPU box); whether its callees are the reference's own or the
CUDA package's depends on `launch.install()` / `launch.uninstall()`.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

IMAGE_TOKEN_INDEX = -200


class FakeTower(nn.Module):
    """27x27 patches of 14x14 pixels (SigLIP-so400m geometry): per-patch channel means -> Linear(3, Dv)."""

    num_patches_per_side = 27

    def __init__(self, Dv: int):
        super().__init__()
        self.proj = nn.Linear(3, Dv)

    def forward(self, images):
        V = images.shape[0]
        p = images[:, :, :378, :378].reshape(V, 3, 27, 14, 27, 14).mean(dim=(3, 5))  # (V,3,27,27)
        return self.proj(p.permute(0, 2, 3, 1).reshape(V, 729, 3))


def build_model(arch, pe_module, E=256, Dv=32, wpe="avg-discrete-sin3d", object_feature_type="patch14", pool_mode="bilinear",
                newline_position="grid", seed=0):
    """arch: the reference's llava.model.llava_arch module; pe_module: the module whose PositionEmbedding* classes to
    instantiate (the reference's or the package's)."""
    torch.manual_seed(seed)
    n_points = 9 if "sample9" in wpe else 5 if "sample5" in wpe else 2 if "minmax" in wpe else 1

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_tower = FakeTower(Dv)
            self.mm_projector = nn.Sequential(nn.Linear(Dv, E), nn.GELU(), nn.Linear(E, E))
            self.embed_tokens = nn.Embedding(64, E)
            self.image_newline = nn.Parameter(torch.randn(E) * 0.1)
            if "mlp" in wpe:
                self.world_position_embedding = pe_module.PositionEmbeddingMLP(E, n_points=n_points)
            elif "sin3d" in wpe:
                self.world_position_embedding = pe_module.PositionEmbeddingSine3D(E, n_points=n_points)

        def get_vision_tower(self):
            return self.vision_tower

    class Model(nn.Module, arch.LlavaMetaForCausalLM):
        def __init__(self):
            super().__init__()
            self.model = Inner()
            self.config = types.SimpleNamespace(
                world_position_embedding_type=wpe, voxel_size=0.1, min_xyz_range=[-15, -15, -5], max_xyz_range=[15, 15, 5],
                object_feature_type=object_feature_type, mm_spatial_pool_mode=pool_mode, mm_spatial_pool_stride=2,
                mm_patch_merge_type="spatial_unpad", image_aspect_ratio="anyres_max_9", mm_newline_position=newline_position,
                tokenizer_padding_side="right", tokenizer_model_max_length=None, hidden_size=E)

        def get_model(self):
            return self.model

        @property
        def device(self):
            return self.model.image_newline.device

    return Model().eval()


def run(model, images, video_dict, use_object_proposals=False, device="cpu"):
    """One call of the reference method on a [text, <video>, text] sample; returns (embeds, labels, object_features, position_ids)."""
    input_ids = torch.tensor([[1, 2, IMAGE_TOKEN_INDEX, 3, 4, 5]], device=device)
    labels = input_ids.clone()
    out = model.prepare_inputs_labels_for_multimodal(input_ids, None, None, None, labels, [images], modalities=["video"], image_sizes=None,
                                                     video_dict=video_dict, use_object_proposals=use_object_proposals)
    _, position_ids, _, _, embeds, new_labels, object_features, _ = out
    return embeds, new_labels, object_features, position_ids
