"""Maximum-coverage frame sampler on the CUDA library: mirror of
scripts/3d/preprocessing/max_coverage_sampling.py (and of convert_pcd_to_voxel.py for the mesh
universe).  Same CLI flags and the same JSON schema (:98-104) as the reference script.

Tie-break protocol: the reference breaks ties with an *unseeded* random.choice (:84).  Here every
scene gets `random.seed(seed)` semantics right before its greedy loop (seed = scene index unless
given), reproduced on the device with CPython's MT19937, so results are reproducible and
bit-identical to the reference run under that protocol.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pickle

import numpy as np
import torch

from . import cabi


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class MCSampler:
    """Batched sampler: voxelise + greedy for many scenes in three kernel launches."""

    def __init__(self, voxel_size: float = 0.1, max_pick: int = 32, div_mode: int = cabi.DIV_IEEE, device=None,
                 depth_scale: float = 1000.0):
        self.lib = cabi.load()
        self.voxel_size, self.max_pick, self.div_mode, self.depth_scale = float(voxel_size), int(max_pick), div_mode, depth_scale
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("MCSampler needs a CUDA device: this path has no CPU implementation")

    def plan(self, scenes):
        """Allocate workspace + records for `scenes` (list of dicts with device tensors depth (F,H,W),
        poses (F,4,4), intrinsics (F,4,4), pc_voxels (n_pc,3) int32, and an int seed).  Returns a
        plan that can be executed repeatedly (bench) without touching the allocator."""
        n = len(scenes)
        if n == 0:
            raise ValueError("no scenes")
        H, W = scenes[0]["depth"].shape[-2:]
        dcode = cabi.depth_code(scenes[0]["depth"])
        sizes, total = [], 0
        for sc in scenes:
            for k in ("depth", "poses", "intrinsics", "pc_voxels"):
                cabi.require_cuda(sc[k], k)
            if tuple(sc["depth"].shape[-2:]) != (H, W) or cabi.depth_code(sc["depth"]) != dcode:
                raise ValueError("all scenes of one call must share H, W and the depth dtype")
            F, n_pc = sc["depth"].shape[0], sc["pc_voxels"].shape[0]
            slots = self.lib.v3d_mc_hash_slots(n_pc)
            words = (n_pc + 31) // 32
            o_hash = total
            total += _align(16 * slots)
            o_bits = total
            total += _align(4 * F * words)
            sizes.append((F, n_pc, slots, words, o_hash, o_bits))
        work = torch.empty(total + 256, dtype=torch.uint8, device=self.device)
        base = _align(work.data_ptr()) - work.data_ptr()
        records = torch.empty((n, cabi.MC_RECORD_INTS), dtype=torch.int32, device=self.device)
        scratch = torch.empty(n * C.sizeof(cabi.McScene) + 64, dtype=torch.uint8, device=self.device)
        arr = (cabi.McScene * n)()
        keep = []
        for i, (sc, (F, n_pc, slots, words, o_hash, o_bits)) in enumerate(zip(scenes, sizes)):
            depth = sc["depth"].contiguous()
            poses = sc["poses"].to(torch.float32).contiguous()
            intr = sc["intrinsics"].to(torch.float32).contiguous()
            pc = sc["pc_voxels"].to(torch.int32).contiguous()
            keep.append((depth, poses, intr, pc))
            a = arr[i]
            a.depth, a.poses, a.intr, a.pc_voxels = depth.data_ptr(), poses.data_ptr(), intr.data_ptr(), pc.data_ptr()
            a.n_frames, a.n_pc, a.seed = F, n_pc, int(sc.get("seed", i)) & 0xFFFFFFFFFFFFFFFF
            a.hash_table, a.hash_slots = work.data_ptr() + base + o_hash, slots
            a.bitsets = work.data_ptr() + base + o_bits
            a.record = records.data_ptr() + i * cabi.MC_RECORD_INTS * 4
        return {"arr": arr, "n": n, "H": H, "W": W, "dcode": dcode, "work": work, "records": records, "scratch": scratch,
                "keep": keep, "sizes": sizes, "base": base}

    def run(self, plan, stage: str = "all"):
        """Enqueue the kernels of a plan on the current stream (no sync)."""
        lib, p = self.lib, plan
        with torch.cuda.device(self.device):
            st = cabi.stream_ptr(self.device)
            if stage == "all":
                rc = lib.v3d_mc_sample(p["arr"], p["n"], p["H"], p["W"], p["dcode"], self.depth_scale, self.voxel_size,
                                       self.div_mode, self.max_pick, p["scratch"].data_ptr(), st)
            elif stage == "voxelize":
                rc = lib.v3d_mc_voxelize(p["arr"], p["n"], p["H"], p["W"], p["dcode"], self.depth_scale, self.voxel_size,
                                         self.div_mode, p["scratch"].data_ptr(), st)
            elif stage == "greedy":
                rc = lib.v3d_mc_greedy(p["arr"], p["n"], self.max_pick, p["scratch"].data_ptr(), st)
            else:
                raise ValueError(stage)
            cabi.check(rc, f"v3d_mc_{stage}")
        return p["records"]

    @staticmethod
    def decode(records: torch.Tensor):
        """records (n, 68) int32 -> list of dicts {order, gains, n_all, n_sel} (host)."""
        rec = records.cpu().numpy()
        out = []
        for r in rec:
            k = int(r[2 * cabi.MC_MAX_PICK])
            out.append({"order": r[:k].tolist(), "gains": r[cabi.MC_MAX_PICK:cabi.MC_MAX_PICK + k].tolist(),
                        "n_all": int(r[2 * cabi.MC_MAX_PICK + 1]), "n_sel": int(r[2 * cabi.MC_MAX_PICK + 2])})
        return out

    def bitsets(self, plan, i: int) -> np.ndarray:
        """Coverage bitsets (F, words) uint32 of scene i after a run (tests)."""
        F, n_pc, slots, words, o_hash, o_bits = plan["sizes"][i]
        off = plan["base"] + o_bits
        raw = plan["work"][off:off + 4 * F * words].cpu().numpy()
        return raw.view(np.uint32).reshape(F, words)

    def sample(self, scenes):
        plan = self.plan(scenes)
        return self.decode(self.run(plan))

    def sample_and_encode(self, plan, frontend, strategy: str = "mc-ratio95", frames_upbound: int = 32, out=None):
        """BASELINE configs[3] for the scenes of a plan: max-coverage order (voxelise + greedy) -> coverage-ratio prefix cut
        (llava/video_utils.py:131-159) -> sin-3D PE of the SELECTED frames with the fused front end (the frames a training
        sample built with this strategy would feed to llava_arch.py:396-515).
        Returns (records (n,68) int32 on the device, selected candidate indices per scene, PE (n_selected, G*G, E))."""
        rec = self.run(plan)
        picks = [ratio_cut(r, strategy, frames_upbound) for r in self.decode(rec)]  # one D2H of 272 bytes per scene
        depth = torch.cat([plan["keep"][i][0][torch.as_tensor(p, device=self.device)] for i, p in enumerate(picks)])
        poses = torch.cat([plan["keep"][i][1][torch.as_tensor(p, device=self.device)] for i, p in enumerate(picks)])
        intr = torch.cat([plan["keep"][i][2][torch.as_tensor(p, device=self.device)] for i, p in enumerate(picks)])
        pe = frontend(depth, poses, intr, out=out)
        return rec, picks, pe


def ratio_cut(record: dict, strategy: str, frames_upbound: int = 32) -> list:
    """llava/video_utils.py:131-159 on candidate indices: the first frames_upbound picks, cut at the first prefix whose
    cumulative gain reaches ratio * num_all_voxels (0.95 / 0.90 by substring, else no cut), sorted by frame order."""
    order, gains = record["order"][:frames_upbound], record["gains"][:frames_upbound]
    ratio = 0.95 if "ratio95" in strategy else (0.9 if "ratio90" in strategy else 1.0)
    if ratio != 1.0:
        need, kept, covered = record["n_all"] * ratio, [], 0
        for f, n in zip(order, gains):
            kept.append(f)
            covered += n
            if covered >= need:
                break
        order = kept
    return sorted(order)


def voxel_ids(world: torch.Tensor, voxel_size: float = 0.1, div_mode: int = cabi.DIV_IEEE) -> torch.Tensor:
    """int(round(world / voxel_size)) (max_coverage_sampling.py:44-45) on the device."""
    cabi.require_cuda(world, "world")
    lib = cabi.load()
    w = world.to(torch.float32).contiguous()
    out = torch.empty(w.shape, dtype=torch.int32, device=w.device)
    with torch.cuda.device(w.device):
        cabi.check(lib.v3d_mc_voxel_ids(w.data_ptr(), w.numel() // 3, float(voxel_size), div_mode, out.data_ptr(),
                                        cabi.stream_ptr(w.device)), "v3d_mc_voxel_ids")
    return out


def mesh_to_voxels(xyz: torch.Tensor, voxel_size: float = 0.1) -> list:
    """convert_pcd_to_voxel.py:69-82: unique round(xyz / voxel) triples of the mesh vertices."""
    ids = voxel_ids(xyz.to(torch.device("cuda", torch.cuda.current_device())), voxel_size)
    uniq = torch.unique(ids.reshape(-1, 3), dim=0)
    return [tuple(v) for v in uniq.cpu().tolist()]


def main(args):
    """Drop-in for max_coverage_sampling.py:main -- same inputs (data/metadata/pcd_discrete_<voxel>.pkl,
    the EmbodiedScan infos, posed_images) and the same output JSON.  Under torchrun (WORLD_SIZE > 1) the scenes are
    sharded round-robin over the ranks (sharding.shard_indices, like the reference's eval drivers split their questions,
    llava/eval/model_scanqa.py:244-245), every rank samples its own scenes on its own GPU, ONE all-gather of the 272-byte
    records brings them to every rank in scene order and rank 0 writes the JSON.  Scene i is sampled under
    random.seed(seed + i) whatever the sharding, so the file does not depend on the number of GPUs."""
    import torch.distributed as dist

    from . import sharding
    from .video_utils import VideoProcessor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    started = False
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        started = True
    rank, world = sharding.world()
    with open(f"data/metadata/pcd_discrete_{args.voxel_size}.pkl", "rb") as f:
        pc_data = pickle.load(f)
    vp = VideoProcessor(voxel_size=args.voxel_size, min_xyz_range=None, max_xyz_range=None,
                        pack_dir=getattr(args, "pack_dir", None) or os.environ.get("V3D_PACK_DIR"))
    dev = torch.device("cuda", torch.cuda.current_device())
    div = getattr(args, "div_mode", None)
    div_mode = cabi.div_mode_default() if div is None else {"ieee": cabi.DIV_IEEE, "reciprocal": cabi.DIV_RECIPROCAL}[div]
    sampler = MCSampler(args.voxel_size, device=dev, div_mode=div_mode)
    scene_ids = sorted(os.listdir(args.video_dir)) if (getattr(args, "sorted", False) or world > 1) else os.listdir(args.video_dir)

    def candidates(scene_id):
        info = vp.scene[f"scannet/{scene_id}"]
        files = [os.path.join(vp.video_folder, img["img_path"]) for img in info["images"]][::2]
        if len(files) < 32:
            files = [os.path.join(vp.video_folder, img["img_path"]) for img in info["images"]]
        return files

    mine = sharding.shard_indices(len(scene_ids), rank, world)
    batch = max(1, int(getattr(args, "batch_scenes", 8)))
    local = torch.full((len(mine), cabi.MC_RECORD_INTS), -1, dtype=torch.int32, device=dev)
    for b0 in range(0, len(mine), batch):
        group = []
        for i in mine[b0:b0 + batch]:
            scene_id = scene_ids[i]
            depth, poses, intr = vp.load_frames(f"scannet/{scene_id}", candidates(scene_id))
            pc = torch.tensor(np.asarray(pc_data[scene_id], dtype=np.int32).reshape(-1, 3), device=dev)
            group.append({"depth": depth.to(dev), "poses": poses.to(dev), "intrinsics": intr.to(dev), "pc_voxels": pc,
                          "seed": getattr(args, "seed", 0) + i})
        # scenes of one launch share H, W; fall back to one scene per launch otherwise
        shapes = {tuple(g["depth"].shape[-2:]) for g in group}
        o = b0
        for r in ([group] if len(shapes) == 1 else [[g] for g in group]):
            local[o:o + len(r)] = sampler.run(sampler.plan(r))
            o += len(r)
    records = sharding.all_gather_records(local, len(scene_ids))  # (n_scenes, 68) in scene order on every rank
    all_data = []
    if rank == 0:
        for scene_id, res in zip(scene_ids, MCSampler.decode(records)):
            video_id, files = f"scannet/{scene_id}", candidates(scene_id)
            print(f"[{scene_id}] all voxels: {res['n_sel']}/{res['n_all']} | ratio: "
                  f"{res['n_sel'] / max(res['n_all'], 1)} | image num: {len(res['order'])}/{len(files)}")
            all_data.append({"video_id": video_id, "frame_files": [files[i] for i in res["order"]], "voxel_nums": res["gains"],
                             "num_all_voxels": res["n_all"], "num_select_voxels": res["n_sel"]})
        with open(args.output_file, "w") as f:
            json.dump(all_data, f)
    if started:
        dist.barrier()
        dist.destroy_process_group()
    return all_data


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--voxel_size", type=float, default=0.1)
    parser.add_argument("--video_dir", type=str, default="data/scannet/posed_images")
    parser.add_argument("--output_file", type=str, default="data/metadata/scannet_select_frames.json")
    parser.add_argument("--seed", type=int, default=0, help="scene i is sampled under random.seed(seed + i)")
    parser.add_argument("--batch_scenes", type=int, default=8)
    parser.add_argument("--div_mode", choices=["ieee", "reciprocal"], default=None,
                        help="x / voxel_size semantics: 'reciprocal' = ATen's CUDA kernel (what the reference script runs, default; "
                             "or V3D_DIV_MODE), 'ieee' = torch CPU")
    parser.add_argument("--pack_dir", default=None, help="packed scenes (scene_pack.py) instead of PNG/txt files")
    parser.add_argument("--sorted", action="store_true", help="process scenes in sorted order (always on under torchrun)")
    main(parser.parse_args())
