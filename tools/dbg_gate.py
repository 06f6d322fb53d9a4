import torch, sys
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import ops
torch.manual_seed(0)
for T, E, k in [(33, 16, 1), (33, 16, 1), (32, 16, 1), (32, 16, 1), (31, 16, 1), (256, 4, 1), (64, 4, 1), (777, 5, 2)]:
    lg = torch.randn(T, E).cuda()
    out = ops.router_gate_from_logits(lg, k)
    torch.cuda.synchronize()
    st = out[5].cpu()
    ws = list(ops._ws_cache.values())[0]
    print(T, E, k, "ticket", ws[:4].view(torch.int32).item() if False else ws[:16].view(torch.int32).tolist(), "colsum_sum", float(st[:E].sum()), "cnt", float(st[E:2*E].sum()), "z", float(st[2*E]))
