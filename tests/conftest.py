import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _dim_t_like_the_cpu_reference(request):
    """The golden vectors and the CPU oracle hold the reference's results on the CPU, where dim_t comes from torch's CPU pow;
    the parity tests therefore build the PE tables from that.  tests/test_gpu_overlay.py compares with the reference running
    on the GPU and switches to the device-evaluated dim_t (the package default)."""
    from video_3d_llm_b200 import position_encoding

    prev = position_encoding.set_dim_t_source("device" if request.module.__name__.endswith("test_gpu_overlay") else "cpu")
    yield
    position_encoding.set_dim_t_source(prev)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def D(golden_dir):
    import json

    with open(os.path.join(golden_dir, "dataset.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def dataset(tmp_path_factory, D):
    """The on-disk fixture of the golden run, plus the reference's own select-frames JSON."""
    import json

    import fixtures

    root = str(tmp_path_factory.mktemp("scannet_like"))
    metas = D["scenes"]
    infos = {"train": [], "val": [], "test": []}
    boxes = {"train": {}, "val": {}}
    pcs = {}
    import pickle
    for name, m in metas.items():
        info, bx, pc, sc = fixtures.write_scene(root, m["seed"], m["n_total"], m["H"], m["W"], m["split"])
        infos[m["split"]].append(info)
        boxes["val" if m["split"] == "val" else "train"][info["sample_idx"]] = bx
        pcs[name] = [tuple(int(t) for t in v) for v in pc]
    os.makedirs(os.path.join(root, "data", "embodiedscan"), exist_ok=True)
    os.makedirs(os.path.join(root, "data", "metadata"), exist_ok=True)
    for split in infos:
        with open(os.path.join(root, "data", "embodiedscan", f"embodiedscan_infos_{split}.pkl"), "wb") as f:
            pickle.dump({"data_list": infos[split]}, f)
    for k, fn in (("train", "scannet_train_gt_box.json"), ("val", "scannet_val_pred_box.json")):
        with open(os.path.join(root, "data", "metadata", fn), "w") as f:
            json.dump(boxes[k], f)
    with open(os.path.join(root, "data", "metadata", "pcd_discrete_0.1.pkl"), "wb") as f:
        pickle.dump(pcs, f)
    with open(os.path.join(root, "data", "metadata", "scannet_select_frames.json"), "w") as f:
        json.dump(D["select_frames"], f)
    return root


