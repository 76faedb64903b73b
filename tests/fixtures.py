"""On-disk ScanNet/EmbodiedScan-shaped fixture writer (formats: SURVEY.md Appendix B).

Writes, under `root`:
  data/embodiedscan/embodiedscan_infos_{train,val,test}.pkl
  data/scannet/posed_images/<scene>/<idx>.{jpg,png,txt}   (RGB, uint16 mm depth, %f pose)
  data/metadata/scannet_{train_gt,val_pred}_box.json
  data/metadata/pcd_discrete_0.1.pkl                      (mesh universe, list of int triples)
Used by oracle/gen_golden.py (to run the unmodified reference) and by the tests of the
host mirror, so both sides read byte-identical files.
"""
from __future__ import annotations

import json
import os
import pickle
import sys

import numpy as np
from PIL import Image

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from video_3d_llm_b200 import synthetic  # noqa: E402


def scene_name(seed: int) -> str:
    return f"scene{seed:04d}_00"


def write_scene(root: str, seed: int, n_total: int, H: int, W: int, split: str = "train"):
    """Render scene `seed` with n_total frames and write its files.  Returns the info entry."""
    sc = synthetic.make_scene(seed, n_total, n_total=n_total, H=H, W=W, device="cpu")
    name = scene_name(seed)
    d = os.path.join(root, "data", "scannet", "posed_images", name)
    os.makedirs(d, exist_ok=True)
    depth = sc.depth_u16()
    images = []
    rng = np.random.default_rng(seed)
    for i in range(n_total):
        stem = f"{i * 10:05d}"
        Image.fromarray(depth[i]).save(os.path.join(d, stem + ".png"))
        rgb = rng.integers(0, 255, (H, W, 3), dtype=np.uint8)
        Image.fromarray(rgb).save(os.path.join(d, stem + ".jpg"), quality=60)
        np.savetxt(os.path.join(d, stem + ".txt"), sc.sensor_poses[i], fmt="%f")
        images.append({"img_path": f"scannet/posed_images/{name}/{stem}.jpg"})
    info = {
        "sample_idx": f"scannet/{name}",
        "images": images,
        "axis_align_matrix": sc.axis_align.tolist(),
        "depth_cam2img": sc.intrinsics[0].double().numpy().tolist(),
    }
    boxes = [[float((b[0, a] + b[1, a]) / 2) for a in range(3)] + [float(b[1, a] - b[0, a]) for a in range(3)] for b in sc.boxes]
    if not boxes:
        boxes = [[0.0, 0.0, 0.5, 1.0, 1.0, 1.0]]
    pc = synthetic.mesh_voxels(sc.room, sc.boxes, seed)
    return info, boxes, pc, sc


def write_dataset(root: str, scenes, H: int = 480, W: int = 640):
    """scenes: list of (seed, n_total, split).  Returns {scene_name: Scene}."""
    infos = {"train": [], "val": [], "test": []}
    boxes = {"train": {}, "val": {}}
    pcs, out = {}, {}
    for seed, n_total, split in scenes:
        info, bx, pc, sc = write_scene(root, seed, n_total, H, W, split)
        infos[split].append(info)
        boxes["val" if split == "val" else "train"][info["sample_idx"]] = bx
        pcs[scene_name(seed)] = [tuple(int(t) for t in v) for v in pc]
        out[scene_name(seed)] = sc
    os.makedirs(os.path.join(root, "data", "embodiedscan"), exist_ok=True)
    os.makedirs(os.path.join(root, "data", "metadata"), exist_ok=True)
    for split in infos:
        with open(os.path.join(root, "data", "embodiedscan", f"embodiedscan_infos_{split}.pkl"), "wb") as f:
            pickle.dump({"data_list": infos[split]}, f)
    with open(os.path.join(root, "data", "metadata", "scannet_train_gt_box.json"), "w") as f:
        json.dump(boxes["train"], f)
    with open(os.path.join(root, "data", "metadata", "scannet_val_pred_box.json"), "w") as f:
        json.dump(boxes["val"], f)
    with open(os.path.join(root, "data", "metadata", "pcd_discrete_0.1.pkl"), "wb") as f:
        pickle.dump(pcs, f)
    return out


class FakeImageProcessor:
    """Stands in for SigLipImageProcessor (siglip_encoder.py:34-67): crop_size + preprocess()."""

    def __init__(self, size: int = 384):
        self.crop_size = {"height": size, "width": size}

    def preprocess(self, images, return_tensors="pt"):
        import torch

        arr = np.stack([np.asarray(im.convert("RGB").resize((self.crop_size["width"], self.crop_size["height"])), dtype=np.float32) for im in images])
        arr = (arr / 255.0 - 0.5) / 0.5
        return {"pixel_values": torch.from_numpy(arr.transpose(0, 3, 1, 2).copy())}


def disk_poses(sc) -> np.ndarray:
    """Poses as the reference sees them after the %f round trip through the pose .txt files:
    float32(axis_align(f64) @ loadtxt(pose)) (llava/video_utils.py:221-230)."""
    sensor = np.array([[float("%f" % x) for x in row] for row in sc.sensor_poses.reshape(-1, 4)]).reshape(-1, 4, 4)
    return (sc.axis_align[None] @ sensor).astype(np.float32)


def mc_candidates(n_total: int) -> list:
    """max_coverage_sampling.py:30-32: every second frame, all frames if that leaves < 32."""
    idx = list(range(n_total))[::2]
    return idx if len(idx) >= 32 else list(range(n_total))
