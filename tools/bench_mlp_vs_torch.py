"""Timing of the tcgen05 MLP-PE kernel at config-2 size (M = 512*729 tokens, E = 1152) and E = 3584."""
import sys, json, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP
DEV = torch.device('cuda:0')
res = {}
for E, M in ((1152, 512 * 729), (3584, 512 * 196)):
    mod = PositionEmbeddingMLP(E).to(DEV).eval()
    x = (torch.rand((1, M, 3), device=DEV) * 20 - 10).to(torch.bfloat16)  # coordinates arrive in the model dtype
    tok = torch.randn((1, M, E), device=DEV).to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(3): mod(x, add_to=tok)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): mod(x, add_to=tok)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        # torch baseline (cuBLAS + elementwise) for context
        mod_bf = PositionEmbeddingMLP(E).to(DEV).to(torch.bfloat16).eval(); mod_bf.load_state_dict(mod.state_dict())
        xb = x
        for _ in range(3): tok + mod_bf.mlp(xb)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): tok + mod_bf.mlp(xb)
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
    flops = 2.0 * M * 512 * E
    byts = 2.0 * M * E * 2 + M * 12
    res[E] = {"M": M, "ms": ms, "tflops": flops / ms / 1e9, "gbs": byts / ms / 1e6, "torch_ms": ms_t}
print(json.dumps(res))
