"""CPU check of the error bound behind k_mc_voxelize's approximate-then-verify voxel ids.

The kernel evaluates t_i = d*(A_i*um + B_i*vm + C_i) + D_i (coefficients = pose row / (f*scale*voxel),
rounded from fp64) with fp32 FMAs and accepts rint(t) whenever |t - rint(t)| < 0.5 - (k1*d + k0); otherwise
the pixel goes through the exact operator sequence.  This script emulates that arithmetic in numpy on
synthetic scenes and reports (a) how far t is from the exact quotient in units of the bound (must be < 1),
(b) how many pixels fall into the guard band, (c) that accepted ids equal the oracle's.  Test infrastructure
only (imports oracle/).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib

from oracle import oracle as O

synthetic = importlib.import_module("video_3d_llm_b200.synthetic")

U = 2.0 ** -24
K1, K0, SLACK = 40.0, 12.0, 2.0 ** -20


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def check(seed, n_frames, H=480, W=640, voxel=0.1, shift=None):
    sc = synthetic.make_scene(seed, n_frames, n_total=n_frames, H=H, W=W, device="cpu")
    poses = sc.poses.numpy().copy()
    if shift is not None:
        poses[:, :3, 3] += np.asarray(shift, np.float32)
    K = sc.intrinsics.numpy()
    d16 = sc.depth_u16()
    world = O.unproject(K, poses, d16)
    ids = O.mc_voxel_ids(world, voxel, O.DIV_IEEE)
    q_exact = (world.astype(np.float32) / np.float32(voxel)).astype(np.float64)
    v32 = float(np.float32(voxel))
    worst, band, wrong, total = 0.0, 0, 0, 0
    for f in range(n_frames):
        fx, fy, cx, cy = (np.float32(K[f, 0, 0]), np.float32(K[f, 1, 1]), np.float32(K[f, 0, 2]), np.float32(K[f, 1, 2]))
        den = [float(fx) * 1000.0 * v32, float(fy) * 1000.0 * v32, 1000.0 * v32, v32]
        coef64 = poses[f, :3, :].astype(np.float64) / np.asarray(den)[None, :]
        coef = coef64.astype(np.float32)
        ummax = max(abs(0 - float(cx)), abs(W - 1 - float(cx)))
        vmmax = max(abs(0 - float(cy)), abs(H - 1 - float(cy)))
        rmax = (np.abs(coef64[:, 0]) * ummax + np.abs(coef64[:, 1]) * vmmax + np.abs(coef64[:, 2])).max()
        k1 = np.float32(K1 * U * rmax)
        k0 = np.float32(K0 * U * np.abs(coef64[:, 3]).max() + SLACK)
        df = d16[f].astype(np.float32)
        c0 = (np.arange(W) // 8 * 8).astype(np.float32)
        e = (np.arange(W) % 8).astype(np.float32)
        um0 = (c0 - cx).astype(np.float32)[None, :]
        vm = (np.arange(H, dtype=np.float32) - cy).astype(np.float32)[:, None]
        ok = np.ones((H, W), bool)
        n = np.empty((H, W, 3), np.float32)
        bound = (k1 * df + k0).astype(np.float64)
        for i in range(3):
            A, B, Cc, D = (np.float32(x) for x in coef[i])
            rowc = fma32(np.broadcast_to(B, vm.shape), vm, np.broadcast_to(Cc, vm.shape))
            ray0 = fma32(np.broadcast_to(A, (H, W)), np.broadcast_to(um0, (H, W)), np.broadcast_to(rowc, (H, W)))
            ray = fma32(np.broadcast_to(A, (H, W)), np.broadcast_to(e[None, :], (H, W)), ray0)
            t = fma32(df, ray, np.broadcast_to(D, (H, W)))
            n[..., i] = np.rint(t)
            ok &= np.abs(t - n[..., i]) < (np.float32(0.5) - k1 * df - k0)
            worst = max(worst, float((np.abs(t.astype(np.float64) - q_exact[f, ..., i]) / bound).max()))
        band += int((~ok).sum())
        wrong += int((n[ok].astype(np.int32) != ids[f][ok]).any(axis=-1).sum())
        total += H * W
    return worst, band / total, wrong


if __name__ == "__main__":
    for seed, nf, shift in ((21, 6, None), (22, 6, None), (23, 6, (0.05, 0.05, 0.05)), (24, 4, (1.25, -0.35, 0.45))):
        worst, frac, wrong = check(seed, nf, shift=shift)
        print(f"seed {seed} shift {shift}: max |t - q| / bound = {worst:.3f}, guard-band pixels = {frac:.2e}, "
              f"accepted ids differing from the oracle = {wrong}")
        assert worst < 1.0 and wrong == 0
