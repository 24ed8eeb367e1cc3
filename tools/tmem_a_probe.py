"""Check the assumed TMEM layout of a 16-bit A operand (tcgen05.mma with A in tensor memory). Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmg_b200 import _lib
torch.manual_seed(0)
A = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
D = torch.zeros(128, 64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
_lib.check(_lib.lib.pmg_debug_tmem_a(A.data_ptr(), B.data_ptr(), D.data_ptr(), st), "probe")
torch.cuda.synchronize()
ref = A.float() @ B.float().T
print("rel err random:", ((D - ref).norm() / ref.norm()).item())
# one-hot rows: which K element does packed position k land on?
perm = []
for k0 in range(64):
    A1 = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16); A1[:, k0] = 1
    _lib.check(_lib.lib.pmg_debug_tmem_a(A1.data_ptr(), B.data_ptr(), D.data_ptr(), st), "probe")
    torch.cuda.synchronize()
    # D[i][n] should equal B[n][perm(k0)]
    d = D[5]
    match = [(k, (d - B[:, k].float()).abs().max().item()) for k in range(64)]
    k, e = min(match, key=lambda t: t[1])
    perm.append(k if e < 1e-3 else -1)
print("perm:", perm)
print("identity:", perm == list(range(64)))
