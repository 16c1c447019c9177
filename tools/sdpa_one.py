import torch, torch.nn.functional as F
n, h = 29640, 40
q = torch.randn(1, h, n, 128, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(3):
    o = F.scaled_dot_product_attention(q, k, v)
torch.cuda.synchronize(); print("ok")
