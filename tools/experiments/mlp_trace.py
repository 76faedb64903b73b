import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from video_3d_llm_b200 import cabi
from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP
DEV = torch.device('cuda:0'); lib = cabi.load()
for E, M in ((1152, 512 * 729), (3584, 512 * 196)):
    mod = PositionEmbeddingMLP(E).to(DEV).eval()
    x = (torch.rand((M, 3), device=DEV) * 20 - 10)
    tok = torch.randn((M, E), device=DEV).to(torch.bfloat16); o = torch.empty_like(tok)
    l1, ln, l2 = mod.mlp[0], mod.mlp[1], mod.mlp[3]
    w2 = l2.weight.detach().to(torch.bfloat16).contiguous()
    args = (x.data_ptr(), M, E, l1.weight.data_ptr(), l1.bias.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps), w2.data_ptr(), l2.bias.data_ptr(), tok.data_ptr(), o.data_ptr(), cabi.stream_ptr(DEV))
    for _ in range(3): cabi.check(lib.v3d_pe_mlp(*args))
    buf = (C.c_ulonglong * 8)()
    lib.v3d_debug_mlp_trace.argtypes = [C.c_void_p]
    lib.v3d_debug_mlp_trace(buf)
    t = list(buf)
    print(E, 'total', t[3], 'wait A', t[0], f'{100*t[0]/t[3]:.1f}%', 'wait acc', t[1], f'{100*t[1]/t[3]:.1f}%', 'wait W2', t[2], f'{100*t[2]/t[3]:.1f}%', 'm tiles', (M+127)//128/148)
