import sys, torch
sys.path.insert(0, '.')
from video_3d_llm_b200.position_encoding import PositionEmbeddingMLP
DEV = torch.device('cuda:0')
E, M = 1152, 512 * 729
mod = PositionEmbeddingMLP(E).to(DEV).eval()
x = (torch.rand((1, M, 3), device=DEV) * 20 - 10)
tok = torch.randn((1, M, E), device=DEV).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(3): mod(x, add_to=tok)
torch.cuda.synchronize()
print('ok')
