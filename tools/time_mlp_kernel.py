import sys, json, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from video_3d_llm_b200 import cabi
from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP
DEV = torch.device('cuda:0'); lib = cabi.load()
out = {}
for E, M in ((1152, 512 * 729), (3584, 512 * 196)):
    mod = PositionEmbeddingMLP(E).to(DEV).eval()
    x = (torch.rand((M, 3), device=DEV) * 20 - 10)
    tok = torch.randn((M, E), device=DEV).to(torch.bfloat16); o = torch.empty_like(tok)
    l1, ln, l2 = mod.mlp[0], mod.mlp[1], mod.mlp[3]
    w2 = l2.weight.detach().to(torch.bfloat16).contiguous()
    args = (x.data_ptr(), M, E, l1.weight.data_ptr(), l1.bias.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps), w2.data_ptr(), l2.bias.data_ptr(), tok.data_ptr(), o.data_ptr(), cabi.stream_ptr(DEV))
    for _ in range(3): cabi.check(lib.v3d_pe_mlp(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lib.v3d_pe_mlp(*args)
    e1.record(); torch.cuda.synchronize()
    out[E] = e0.elapsed_time(e1) / 10
print(json.dumps(out))
