#!/usr/bin/env python
"""bench.py -- headline measurement of the B200-native Video-3D LLM front end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic ScanNet-shaped input:
BASELINE.json configs[1] -- 16 scenes x 32 uniformly sampled frames, depth 480x640 uint16 ->
unproject -> nearest resize/crop -> 14x14-pixel patch means on the 27x27 SigLIP grid -> clamp +
discretise -> sin-3D PE (E=1152) added to the bf16 vision tokens, ONE fused kernel launch.

Printed JSON line (rank 0):
  value      frames/s of the whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the public API with HOST (pinned) depth/pose/intrinsics buffers: H2D copies
             and a D2H read of the per-token voxel coordinates inside the timed region.  The pinned depth is
             what the packed-scene loader emits (ScenePack.load(resample=...)): the reference's nearest resize +
             crop applied on the host to the raw uint16 depth, i.e. the 378x378 samples per frame the kernel reads
  roofline   algorithmic bytes per launch / average launch time vs the measured HBM copy peak
  cpu_baseline  the reference's torch-CPU operator sequence (oracle/torch_port.py) on the host cores
  mc         secondary metric: max-coverage sampled scenes/s (voxelise + greedy, BASELINE configs[2]); mc.e2e: the
             same with all candidate frames uploaded from pinned host memory every step
  default_recipe_geometry / mlp_pe / projector_tail / host_ingest   secondary objects: the reference-default 14x14 grid
             with E=3584, the tcgen05 learned-projection PE, the pool + PE + newline kernel, PNG/txt vs packed ingest
`--impl reference` times the CPU port only (rank 0), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "frames/sec for fused unproject+pool+3D-PE (+token add)"
SCENES, FRAMES, PATCH, GRID, EMB = 16, 32, 14, 27, 1152
H, W = 480, 640
# algorithmic bytes per frame, BASELINE.md section 4: referenced depth samples + pose + (fx,fy,cx,cy)
# + token read + token write
BYTES_IN_PER_FRAME = 378 * 378 * 2 + 64 + 16
BYTES_PE_PER_FRAME = GRID * GRID * EMB * 2
BYTES_PER_FRAME_FUSED_ADD = BYTES_IN_PER_FRAME + 2 * BYTES_PE_PER_FRAME  # 3 645 080
WORKLOAD = (f"{SCENES} scenes x {FRAMES} frames uniform sampling, depth {H}x{W} u16 -> fused unproject+pool+PE+add "
            f"into {EMB}-d tokens on the {GRID}x{GRID} grid (BASELINE configs[1])")


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f).get("k_fused_dram_bytes_per_launch")
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------
def cpu_reference_leg(steps: int, warmup: int, with_boundary: bool = False):
    """Time the torch-CPU port of the reference on a bounded sample: 1 scene x 32 frames per step
    (the reference processes scenes one at a time, so frames/s does not depend on the batch)."""
    from oracle import torch_port as TP
    from video_3d_llm_b200 import synthetic

    torch.set_num_threads(os.cpu_count() or 1)
    sc = synthetic.make_scene(100, FRAMES, n_total=300, device="cpu")
    depth = torch.from_numpy(sc.depth_u16().astype(np.int32))  # the reference holds int32 depth (video_utils.py:217)
    tokens = synthetic.make_tokens(100, FRAMES, GRID * GRID, EMB)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        TP.frontend(depth, sc.poses, sc.intrinsics, PATCH, EMB, tokens=tokens, with_boundary=with_boundary)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    best, mean = min(times), sum(times) / len(times)
    return {"frames_per_s_best": FRAMES / best, "frames_per_s_mean": FRAMES / mean, "ms_per_step": mean * 1e3,
            "cores": torch.get_num_threads(), "sample": f"1 scene x {FRAMES} frames per step, {steps} steps after {warmup} warm-up"}


def cpu_c_oracle_leg():
    """The C/OpenMP oracle on the same sample (informational: a hand-optimised CPU port)."""
    from oracle import oracle as O
    from video_3d_llm_b200 import synthetic

    sc = synthetic.make_scene(100, FRAMES, n_total=300, device="cpu")
    tokens = synthetic.make_tokens(100, FRAMES, GRID * GRID, EMB).view(torch.int16).numpy().view(np.uint16)
    d, p, k = sc.depth_u16(), sc.poses.numpy(), sc.intrinsics.numpy()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        O.frontend(d, p, k, patch_size=PATCH, embedding_size=EMB, tokens_bf16_bits=tokens)
        best = min(best, time.perf_counter() - t0)
    return {"value": FRAMES / best, "unit": "frames/s", "cores": O.num_threads(), "kind": "port (C/OpenMP oracle)"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    r = cpu_reference_leg(max(args.steps, 1), max(args.warmup, 1))
    v = r["frames_per_s_mean"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference path = torch-CPU operator sequence (oracle/torch_port.py, pinned bit-exact "
                   "against the reference); the reference itself is Python and /root/reference does not travel to the GPU box"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------
def make_batch(rank: int, dev):
    from video_3d_llm_b200 import synthetic

    scenes = [synthetic.make_scene(1000 * rank + 100 + s, FRAMES, n_total=300, device=dev) for s in range(SCENES)]
    depth = torch.cat([s.depth for s in scenes]).contiguous()
    poses = torch.cat([s.poses for s in scenes]).contiguous()
    intr = torch.cat([s.intrinsics for s in scenes]).contiguous()
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    tokens = torch.randn((SCENES * FRAMES, GRID * GRID, EMB), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    return depth, poses, intr, tokens


def mc_leg(rank: int, world: int, dev, steps: int, warmup: int):
    """Secondary metric: MC-sampled scenes/s (BASELINE configs[2]): 300-frame scenes -> 150 candidate
    frames, 0.1 m voxels, 32 picks; voxelise + greedy, plus the all-gather of the 272-byte records."""
    import torch.distributed as dist

    from video_3d_llm_b200 import synthetic
    from video_3d_llm_b200.mc_sampling import MCSampler
    from video_3d_llm_b200 import sharding

    n_scenes, cand = 16, 150
    scenes = []
    for s in range(n_scenes):
        seed = 1000 * rank + 500 + s
        sc = synthetic.make_scene(seed, cand, n_total=300, frame_indices=np.arange(0, 300, 2), device=dev)
        pc = torch.from_numpy(synthetic.mesh_voxels(sc.room, sc.boxes, seed)).to(dev)
        scenes.append({"depth": sc.depth, "poses": sc.poses, "intrinsics": sc.intrinsics, "pc_voxels": pc, "seed": seed})
    sampler = MCSampler(device=dev)
    plan = sampler.plan(scenes)
    for _ in range(max(warmup, 1)):
        rec = sampler.run(plan)
        sharding.all_gather_records(rec)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rec = sampler.run(plan)
        allrec = sharding.all_gather_records(rec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # stage split (one extra run each, untimed region)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    sampler.run(plan, "voxelize")
    ev[1].record()
    sampler.run(plan, "greedy")
    ev[2].record()
    torch.cuda.synchronize()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms = float(t.item())
    depth_bytes = sum(s["depth"].numel() * 2 + s["depth"].shape[0] * 80 + s["pc_voxels"].numel() * 4 + 272 for s in scenes)
    res = sampler.decode(rec)
    # the same with the candidate frames coming from pinned host memory each step (reported separately: SURVEY 8d)
    e2e = None
    if world == 1:
        h_depth = [s["depth"].cpu().pin_memory() for s in scenes]
        for _ in range(2):
            for s, h in zip(scenes, h_depth):
                s["depth"].copy_(h, non_blocking=True)
            sampler.decode(sampler.run(plan))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = 3
        for _ in range(n_e2e):
            for s, h in zip(scenes, h_depth):
                s["depth"].copy_(h, non_blocking=True)
            sampler.decode(sampler.run(plan))  # records read back to the host
        e2e_s = time.perf_counter() - t0
        e2e = {"value": n_scenes * n_e2e / e2e_s, "unit": "scenes/s", "h2d_bytes_per_step": int(sum(h.numel() * 2 for h in h_depth)),
               "d2h_bytes_per_step": int(rec.numel() * 4),
               "note": "all 150 candidate frames of every scene uploaded in full each step (92 MB per scene): PCIe-bound"}
        del h_depth
    return {
        "metric": "MC-sampled scenes/sec (voxelise + greedy max-coverage, 32 picks)", "value": world * n_scenes * steps / (ms / 1e3),
        "unit": "scenes/s", "ms_per_step": ms / steps, "scenes_per_rank": n_scenes, "candidate_frames": cand, "voxel_size": 0.1,
        "voxelize_ms": ev[0].elapsed_time(ev[1]), "greedy_ms": ev[1].elapsed_time(ev[2]),
        "algorithmic_bytes_per_step": depth_bytes,
        "hbm_gbs_achieved_voxelize": depth_bytes / (ev[0].elapsed_time(ev[1]) / 1e3) / 1e9,
        "e2e": e2e, "allgather_bytes": int(allrec.numel() * 4), "mean_picks": float(np.mean([len(r["order"]) for r in res])),
        "mean_coverage": float(np.mean([r["n_sel"] / max(r["n_all"], 1) for r in res])),
    }


def tail_leg(dev, n_frames: int = 512, steps: int = 10):
    """Secondary: the projector tail on the default recipe -- (V,27*27,3584) bf16 projector output ->
    bilinear 2x2 pool -> + sin-3D PE of the discretised patch coordinates -> newline per grid row, one kernel
    (k_pool_pe_grid), against the same operator sequence of the reference run by torch on the same GPU."""
    from video_3d_llm_b200 import projector_tail as PT
    from video_3d_llm_b200.position_encoding import PositionEmbeddingSine3D

    E, side, Go = 3584, 27, 14
    g = torch.Generator(device=dev).manual_seed(7)
    tokens = torch.randn((n_frames, side * side, E), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    coords = torch.randint(0, 101, (n_frames, Go * Go, 3), generator=g, device=dev).to(torch.bfloat16)
    newline = torch.randn(E, generator=g, device=dev).to(torch.bfloat16)
    pe = PositionEmbeddingSine3D(E)

    def ours():
        return PT.pool_pe_newline(tokens, coords=coords, position_embedding=pe, image_newline=newline)

    def torch_ops():  # llava_arch.py:191-210, 515, 307-328 as written, on the GPU
        x = tokens.view(n_frames, side, side, -1).permute(0, 3, 1, 2).contiguous()
        x = torch.nn.functional.interpolate(x, size=[Go, Go], mode="bilinear")
        x = x.permute(0, 2, 3, 1).view(n_frames, -1, E)
        x = x + pe(coords)
        x = x.view(n_frames, 1, Go, Go, -1).permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)
        x = torch.cat((x, newline[:, None, None].expand(*x.shape[:-1], 1)), dim=-1)
        return x.flatten(1, 2).transpose(0, 1)

    res = {}
    for name, fn, warm, reps in (("ours", ours, 3, steps), ("torch_gpu_ops", torch_ops, 1, 2)):  # the torch sequence takes ~1 s per pass
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / reps
    bytes_per_frame = side * side * E * 2 + Go * (Go + 1) * E * 2 + Go * Go * 12
    res.update({"kernel": "k_pool_pe_grid", "frames": n_frames, "E": E, "algorithmic_bytes_per_frame": bytes_per_frame,
                "hbm_gbs": bytes_per_frame * n_frames / res["ours_ms"] / 1e6,
                "frac_of_measured_hbm_peak": bytes_per_frame * n_frames / res["ours_ms"] / 1e6 / measured_peaks()[0],
                "frames_per_s": n_frames / res["ours_ms"] * 1e3, "speedup_vs_torch_gpu_ops": res["torch_gpu_ops_ms"] / res["ours_ms"]})
    return res


def ingest_leg(dev, n_frames: int = 32):
    """Secondary: host ingest of one 32-frame sample -- the reference's per-frame PNG decode + pose text
    parse + float64 axis-align product (llava/video_utils.py:212-230) against the packed scene binary
    (video_3d_llm_b200/scene_pack.py), both ending with the depth stack resident on the device."""
    import shutil
    import tempfile

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import fixtures
    from video_3d_llm_b200 import scene_pack
    from video_3d_llm_b200 import video_utils as VU

    root = tempfile.mkdtemp(prefix="v3d_ingest_")
    try:
        fixtures.write_dataset(root, [(900, 2 * n_frames, "train")])
        vp = VU.VideoProcessor(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"),
                               voxel_size=0.1, frame_sampling_strategy="uniform", metadata_dir=os.path.join(root, "data", "metadata"))
        video_id = "scannet/" + fixtures.scene_name(900)
        packs = os.path.join(root, "packs")
        os.makedirs(packs)
        scene_pack.pack_scene(vp, video_id, scene_pack.pack_path(packs, video_id))
        vp_pack = VU.VideoProcessor(video_folder=os.path.join(root, "data"), annotation_dir=os.path.join(root, "data", "embodiedscan"),
                                    voxel_size=0.1, frame_sampling_strategy="uniform", metadata_dir=os.path.join(root, "data", "metadata"),
                                    pack_dir=packs)
        files = vp.sample_frame_files(video_id, force_sample=True, frames_upbound=n_frames)
        out = {"frames": n_frames, "depth": "480x640 uint16"}
        for name, proc in (("png_txt", vp), ("packed", vp_pack)):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                d, p_, k = proc.load_frames(video_id, files)
                d = d.to(dev, non_blocking=True)
                p_ = p_.to(dev, non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            out[name + "_ms"] = best * 1e3
            out[name + "_frames_per_s"] = n_frames / best
        out["speedup"] = out["png_txt_ms"] / out["packed_ms"]
        return out
    finally:
        shutil.rmtree(root, ignore_errors=True)


def mlp_leg(dev, steps: int):
    """Secondary: the learned-projection PE variant (PositionEmbeddingMLP) on tcgen05 at the config-2 token count:
    M = 16*32*729 tokens, K = 512, N = 1152, fused bias + token add.  Kernel-only time through the C ABI."""
    from video_3d_llm_b200 import cabi
    from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP

    lib = cabi.load()
    M, E = SCENES * FRAMES * GRID * GRID, EMB
    torch.manual_seed(0)
    mod = PositionEmbeddingMLP(E).to(dev).eval()
    x = torch.rand((M, 3), device=dev) * 20 - 10
    tok = torch.randn((M, E), device=dev).to(torch.bfloat16)
    out = torch.empty_like(tok)
    l1, ln, l2 = mod.mlp[0], mod.mlp[1], mod.mlp[3]
    w2 = l2.weight.detach().to(torch.bfloat16).contiguous()
    a = (x.data_ptr(), M, E, l1.weight.data_ptr(), l1.bias.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps),
         w2.data_ptr(), l2.bias.data_ptr(), tok.data_ptr(), out.data_ptr(), cabi.stream_ptr(dev))
    for _ in range(3):
        cabi.check(lib.v3d_pe_mlp(*a), "v3d_pe_mlp")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lib.v3d_pe_mlp(*a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = 2.0 * M * 512 * E
    byts = 2.0 * M * E * 2 + M * 12 + E * 512 * 2
    peak_tf = 1599.5
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf = float(json.load(f)["bf16_tflops"])
    except OSError:
        pass
    return {"kernel": "k_pe_mlp (tcgen05.mma cta_group::1 M128 N128 K16, TMEM accumulators)", "tokens": M, "K": 512, "N": E,
            "ms": ms, "tflops": flops / ms / 1e9, "frac_of_measured_bf16_peak": flops / ms / 1e9 / peak_tf,
            "hbm_gbs": byts / ms / 1e6, "frames_per_s": SCENES * FRAMES / (ms / 1e3)}


def run_ours(args):
    import torch.distributed as dist

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from video_3d_llm_b200 import cabi
    from video_3d_llm_b200.frontend import FrontEnd

    cabi.load()
    K, Wm = args.steps, max(args.warmup, 3)
    depth, poses, intr, tokens = make_batch(rank, dev)
    n_frames = depth.shape[0]
    fe = FrontEnd(device=dev, patch_size=PATCH, embedding_size=EMB)
    out = torch.empty_like(tokens)

    # ---- device-resident timing: exactly K launches between two events -----------------------
    for _ in range(Wm):
        fe(depth, poses, intr, tokens=tokens, out=out)
    clocks = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fe(depth, poses, intr, tokens=tokens, out=out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the timed region lasts milliseconds; keep the identical launches going for ~1 s (untimed) so that the
    # 200 ms nvidia-smi sampler sees clocks and throttle reasons under this load
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 1.0:
        for _ in range(50):
            fe(depth, poses, intr, tokens=tokens, out=out)
        torch.cuda.synchronize()
    clk = clocks.stop()
    clk["window"] = "timed region + 1 s continuation of the same launches"
    ms_local = e0.elapsed_time(e1)
    t = torch.tensor([ms_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n_frames * K / (ms / 1e3)

    # ---- end to end: host (pinned) depth/pose/intrinsics in, per-token voxel coordinates out ---
    # pinned host inputs as the packed-scene loader emits them (ScenePack.load(resample=...)): the nearest-neighbour resize +
    # crop of the reference's preprocess (video_utils.py:291-308) applied to the raw uint16 depth on the host -- per frame
    # exactly the 378 x 378 samples the kernel reads (47 % of the frame), contiguous
    h_depth, h_poses, h_intr = fe.pack_resampled(depth.cpu()), poses.cpu().pin_memory(), intr.cpu().pin_memory()
    h_coords = [torch.empty((n_frames, GRID, GRID, 3), dtype=torch.float32).pin_memory() for _ in range(2)]

    # The step is split into chunks so that the upload of chunk i+1 (copy stream) overlaps the scatter + kernel of chunk i,
    # and the device buffers exist twice so that the uploads of step i+1 start while step i's last kernels and its result
    # read-back are still running.  Every step copies its inputs from pinned host memory and reads its result back.
    n_chunks = 8
    bounds = [(i * n_frames // n_chunks, (i + 1) * n_frames // n_chunks) for i in range(n_chunks)]
    up = torch.cuda.Stream(device=dev)
    sets = [{"depth": torch.empty_like(depth), "poses": torch.empty_like(poses), "intr": torch.empty_like(intr),
             "stage": torch.empty(h_depth.shape, dtype=h_depth.dtype, device=dev),
             "coords": torch.empty((n_frames, GRID, GRID, 3), dtype=torch.float32, device=dev),
             "free": None, "read": None} for _ in range(2)]
    step_no = [0]

    def e2e_step():
        cur = torch.cuda.current_stream(dev)
        i = step_no[0]
        step_no[0] += 1
        S = sets[i % 2]
        if S["read"] is not None:
            S["read"].synchronize()  # the host has the result of step i-2 before its buffers are reused
        events = []
        with torch.cuda.stream(up):
            if S["free"] is not None:
                up.wait_event(S["free"])  # step i-2's kernels are done with this buffer set
            S["poses"].copy_(h_poses, non_blocking=True)
            S["intr"].copy_(h_intr, non_blocking=True)
            for a, b in bounds:
                S["stage"][a:b].copy_(h_depth[a:b], non_blocking=True)  # one contiguous DMA per chunk
                ev = torch.cuda.Event()
                ev.record(up)
                events.append(ev)
        for (a, b), ev in zip(bounds, events):
            cur.wait_event(ev)
            fe.scatter_resampled(S["stage"][a:b], S["depth"][a:b])  # back to the pixels the samples were taken from
            _, _, pe_coords = fe(S["depth"][a:b], S["poses"][a:b], S["intr"][a:b], tokens=tokens[a:b], out=out[a:b], want_coords=True)
            S["coords"][a:b].copy_(pe_coords)
        S["free"] = torch.cuda.Event()
        S["free"].record(cur)
        h_coords[i % 2].copy_(S["coords"], non_blocking=True)
        S["read"] = torch.cuda.Event()
        S["read"].record(cur)

    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.barrier()
    e2e_value = world * n_frames * Ke / float(te.item())
    h2d = h_depth.numel() * h_depth.element_size() + h_poses.numel() * 4 + h_intr.numel() * 4
    d2h = h_coords[0].numel() * 4

    # secondary: the reference's default recipe geometry (27-pixel windows -> 14x14 grid, E = 3584)
    recipe = None
    if rank == 0:
        fe2 = FrontEnd(device=dev, patch_size=27, embedding_size=3584)
        tok2 = torch.randn((n_frames, 196, 3584), device=dev, dtype=torch.float32).to(torch.bfloat16)
        out2 = torch.empty_like(tok2)
        for _ in range(3):
            fe2(depth, poses, intr, tokens=tok2, out=out2)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(10):
            fe2(depth, poses, intr, tokens=tok2, out=out2)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / 10
        rbytes = (BYTES_IN_PER_FRAME + 2 * 196 * 3584 * 2) * n_frames
        recipe = {"workload": "same depth batch, patch 27 -> 14x14 grid, E=3584, fused add (train_multi.sh:82-87 defaults)",
                  "ms": rms, "frames_per_s": n_frames / (rms / 1e3), "hbm_gbs": rbytes / rms / 1e6,
                  "frac_of_measured_hbm_peak": rbytes / rms / 1e6 / measured_peaks()[0]}
        del tok2, out2
    mc = mc_leg(rank, world, dev, steps=max(2, min(K, 5)), warmup=1)
    mlp = mlp_leg(dev, steps=5) if rank == 0 else None
    tail = None
    if rank == 0:
        try:
            tail = tail_leg(dev)
        except Exception as e:
            tail = {"error": repr(e)}
    ingest = None
    if rank == 0 and world == 1:
        try:
            ingest = ingest_leg(dev)
        except Exception as e:  # host-side secondary line; never fails the bench
            ingest = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel_s = (ms_local / 1e3) / K
        achieved = BYTES_PER_FRAME_FUSED_ADD * n_frames / kernel_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_rank": n_frames, "patch": PATCH, "grid": GRID, "embedding": EMB,
                       "pe_dtype": "bf16", "world_position_embedding_type": "avg-discrete-sin3d",
                       "l2": "inputs larger than L2: one step touches 1.87 GB (depth 315 MB + tokens 860 MB in + 860 MB out)",
                       "parallelism": f"scene-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": BYTES_PER_FRAME_FUSED_ADD * n_frames, "kernel": "k_fused",
                         "avg_launch_us": kernel_s * 1e6},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": Ke, "note": "pinned host depth as the packed-scene loader emits it (nearest resize + crop applied to the raw uint16 depth on the host: the 378x378 samples per frame the kernel reads; contiguous DMAs + device-side scatter into the frame layout)/pose/intrinsics -> device each step, device buffers double-buffered so that the uploads of the next step overlap the tail of the current one; vision tokens are produced on the "
                    "device by the vision tower and stay resident; per-token voxel coordinates read back"},
            "gpu_launches": K,  # timed region: K launches of k_fused (the e2e leg launches 4 per step, the mc leg 4 per step)
            "clocks": clk,
            "default_recipe_geometry": recipe,
            "mc": mc,
            "mlp_pe": mlp,
            "projector_tail": tail,
            "host_ingest": ingest,
        }
        if world == 1 and not args.no_cpu:
            r = cpu_reference_leg(steps=3, warmup=1)
            line["cpu_baseline"] = {"value": r["frames_per_s_mean"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"] + " (torch-CPU operator sequence of the reference, oracle/torch_port.py)",
                                    "best": r["frames_per_s_best"]}
            try:
                line["cpu_baseline_c_oracle"] = cpu_c_oracle_leg()
            except Exception as e:  # the C oracle is optional for the bench
                line["cpu_baseline_c_oracle"] = {"error": str(e)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line: dict) -> None:
    """Write the one JSON line to the real stdout (see main: fd 1 is pointed at stderr while the bench runs, so that
    banners printed by libraries -- e.g. NCCL's version line -- cannot end up next to it)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT, data)


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
