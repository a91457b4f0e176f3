"""Diagnose a GEMM tile plan: full-output comparison against torch for one shape, mismatch pattern by tile."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
libpath, M, K, N, act = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cfgs = sys.argv[6:]
lib = C.CDLL(libpath)
lib.dfv_last_error.restype = C.c_char_p
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
bias = torch.randn(N, device="cuda") * 0.1
ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for i in range(0, M, 65536):
    r = a[i:i + 65536].float() @ w.float().t() + bias
    if act: r = r * torch.sigmoid(r)
    ref[i:i + 65536] = r.bfloat16()
vp = lambda t: C.c_void_p(t.data_ptr())
for cfg in cfgs:
    os.environ["DFV_GEMM_FORCE"] = cfg
    for rep in range(3):
        y = torch.full((M, N), 777.0, device="cuda", dtype=torch.bfloat16)
        rc = lib.dfv_pw_gemm_fwd(vp(a), vp(w), vp(bias), None, 0, None, vp(y), 1, C.c_longlong(M), K, N, act, None)
        torch.cuda.synchronize()
        if rc: print(cfg, "rc", rc, lib.dfv_last_error()); break
        bad = ((y.float() - ref.float()).abs() > 0.05 * (ref.float().abs() + 1.0))
        nb = int(bad.sum())
        msg = f"{cfg} rep{rep}: bad {nb}"
        if nb:
            rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
            un = int((y == 777.0).sum())
            mt = torch.unique(rows // 128)
            msg += f" unwritten {un}; rows {int(rows.min())}..{int(rows.max())} ({len(rows)}), m-tiles {len(mt)} first {mt[:8].tolist()} last {mt[-4:].tolist()}; cols {int(cols.min())}..{int(cols.max())} ({len(cols)}); row%128 uniq {torch.unique(rows % 128)[:10].tolist()}"
        print(msg)
