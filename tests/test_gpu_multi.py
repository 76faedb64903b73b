"""GPU, world size 2 (needs two devices: run with `gpurun --gpus 2`; skipped on a single-GPU box): the sharded product
path of BASELINE configs[3] -- `python -m video_3d_llm_b200.mc_sampling` under two NCCL ranks writes the same JSON as the
single-GPU run, whatever the sharding."""
import json
import os
import socket
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, root, out_file):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    os.chdir(root)
    from video_3d_llm_b200 import mc_sampling

    args = types.SimpleNamespace(voxel_size=0.1, video_dir=os.path.join(root, "data", "scannet", "posed_images"), output_file=out_file,
                                 seed=7, batch_scenes=2, div_mode="ieee")
    mc_sampling.main(args)


def test_sharded_mc_cli_equals_the_single_gpu_run(dataset, tmp_path):
    from video_3d_llm_b200 import mc_sampling

    os.chdir(dataset)
    single = str(tmp_path / "single.json")
    args = types.SimpleNamespace(voxel_size=0.1, video_dir=os.path.join(dataset, "data", "scannet", "posed_images"), output_file=single,
                                 seed=7, batch_scenes=2, div_mode="ieee", sorted=True)
    ref = mc_sampling.main(args)
    assert len(ref) == 3
    sharded = str(tmp_path / "sharded.json")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dataset, sharded)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    with open(sharded) as f:
        got = json.load(f)
    assert got == ref
    assert not dist.is_initialized()
