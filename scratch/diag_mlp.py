import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_mlp import make_module, torch_reference, DEV
E, M = 1152, 148 * 128 * 2 + 77
mod = make_module(E, seed=3)
g = torch.Generator(device="cpu").manual_seed(5)
x = (torch.rand((1, M, 3), generator=g) * 30 - 15).to(DEV)
tokens = torch.randn((1, M, E), generator=g).to(DEV).to(torch.bfloat16)
with torch.no_grad():
    for tk in (None, tokens):
        got = mod(x, add_to=tk).float(); ref = torch_reference(mod, x, tk).float()
        err = (got - ref).abs()
        mag = ref.abs().clamp(min=1.0)
        if tk is not None: mag = torch.maximum(mag, torch.maximum(tk.float().abs(), (ref - tk.float()).abs()))
        bad = err > 2.0**-7 * mag
        print('add' if tk is not None else 'pe ', 'n_bad', int(bad.sum()), 'of', bad.numel(), 'max err', float(err.max()), 'max rel', float((err / mag).max()))
        idx = bad[0].nonzero()[:20]
        for m, e in idx.tolist():
            print('  m', m, 'tile', m // 128, 'row', m % 128, 'e', e, 'got', float(got[0, m, e]), 'ref', float(ref[0, m, e]), 'x', x[0, m].tolist())
        rows = bad[0].any(1).nonzero().flatten()
        print('  bad rows', rows[:30].tolist(), 'count', len(rows))
        # is the hidden layer the source?  compare PE rows for those tokens in fp32
