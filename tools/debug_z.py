import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matcouply_b200 import _lib, _ops
for variant in (_lib.VARIANT_FMA, _lib.VARIANT_DMMA):
  for (N, K, R) in [(513, 1030, 8), (513, 1030, 16), (513, 1024, 8), (513, 1040, 16), (64, 1030, 8), (513, 600, 8), (513, 2048, 32), (3000, 1024, 20), (513,512,8)]:
    rs = np.random.RandomState(1)
    ld = _ops.padded_ld(K, torch.float64)
    Xh = rs.standard_normal(size=(N, K))
    X = torch.zeros((N, ld), dtype=torch.float64, device="cuda"); X[:, :K] = torch.as_tensor(Xh).cuda()
    W = rs.standard_normal(size=(N, R))
    ldw = _ops.z_ldw(R, torch.float64, variant) if hasattr(_ops, "z_ldw") else R
    Wpad = torch.zeros(((N + 31) // 32 * 32, ldw), dtype=torch.float64, device="cuda"); Wpad[:N, :R] = torch.as_tensor(W).cuda()
    Z = torch.full((K, R), float("nan"), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    try:
        _ops.xstream_z(X, N, K, Wpad, Z, ws, variant)
        torch.cuda.synchronize()
    except Exception as e:
        print("variant", variant, (N, K, R), "EXC", e); continue
    ref = Xh.T @ W
    err = np.abs(Z.cpu().numpy() - ref).max(axis=1)
    bad = np.nonzero(~(err < 1e-9))[0]
    print("variant", variant, (N, K, R), "bad rows:", len(bad), (bad[:6], bad[-6:]) if len(bad) else "")
