import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uniadapter_b200.gemm import attention_tf32x3
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, H = int(os.environ.get("PB", 30)), 513, 6
qkv = torch.randn(B * N, 3 * H * 64, device=dev)
for _ in range(3):
    attention_tf32x3(qkv, B, N, H)
torch.cuda.synchronize()
print("ok")
