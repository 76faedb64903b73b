"""Experiment: can the page-cache mapping of a packed scene be registered with CUDA (cudaHostRegister) so that the H2D DMA
reads it directly (no CPU copy into a pinned buffer)?  Prints the outcome and the copy rates."""
import ctypes, mmap, os, sys, tempfile, time
import numpy as np, torch
dev = torch.device("cuda:0"); torch.cuda.init()
rt = torch.cuda.cudart()
n = 512 * 378 * 378 * 2  # one config-2 step of resampled depth
path = os.path.join(tempfile.gettempdir(), "v3d_reg_test.bin")
with open(path, "wb") as f:
    f.write(np.random.randint(0, 60000, n // 2, dtype=np.uint16).tobytes())
for mode, prot, flag in (("r+", None, 0), ("r", None, 8), ("r", None, 0)):
    a = np.memmap(path, dtype=np.uint8, mode=mode)
    _ = a.sum()  # fault the pages in
    ptr = a.ctypes.data
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(ptr, n, flag)
    t1 = time.perf_counter()
    print(f"memmap mode={mode} flag={flag}: cudaHostRegister -> {rc} in {(t1 - t0) * 1e3:.1f} ms", flush=True)
    if int(rc) != 0:
        continue
    import warnings
    warnings.simplefilter("ignore")
    t = torch.from_numpy(a) if mode != "r" else torch.frombuffer(a, dtype=torch.uint8)
    print("  torch sees pinned:", t.is_pinned())
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for reps in (1, 5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): d.copy_(t, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        print(f"  whole-buffer H2D: {n / dt / 1e9:.1f} GB/s")
    # 512 frame-sized pieces (286 KB each), as a batch gathers them
    piece = 378 * 378 * 2
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(512): d[i * piece:(i + 1) * piece].copy_(t[i * piece:(i + 1) * piece], non_blocking=True)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"  512 pieces: {n / dt / 1e9:.1f} GB/s (issue {t_issue * 1e3:.2f} ms, total {dt * 1e3:.2f} ms)")
    assert torch.equal(d.cpu(), torch.from_numpy(np.array(a)))
    rt.cudaHostUnregister(ptr)
    del t, a
os.unlink(path)
