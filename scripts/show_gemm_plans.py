"""Print the tensor-core GEMM tile plans chosen for the B4 1x1 convolutions (host only; no GPU needed).
Rows are (M, K, N, gated) AFTER row folding, batch from argv[1] (default 256)."""
import ctypes as C, os, sys
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "deepfake_vit_b200", "libdfvit.so"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
def fold(M, K, N, gated, rpi):
    if K > 48: return 1
    for f in (4, 2):
        if M % f or (gated and rpi % f): continue
        exact4 = f == 4 and not gated and f * N <= 768 and (f * N) % 192 == 0 and N % 48 == 0 and N < 192
        if f * N > (256 if f == 4 else 512) and not exact4: continue
        return f
    return 1
# (name, H, K, N, gated)
layers = [("b0 project", 190, 48, 24, 1), ("b1 project", 190, 24, 24, 1), ("b2 expand", 190, 24, 144, 0), ("b2 project", 95, 144, 32, 1),
          ("b3 expand", 95, 32, 192, 0), ("b3 project", 95, 192, 32, 1), ("b6 project", 48, 192, 56, 1), ("b7 expand", 48, 56, 336, 0),
          ("b7 project", 48, 336, 56, 1), ("b10 project", 24, 336, 112, 1), ("b11 expand", 24, 112, 672, 0), ("b11 project", 24, 672, 112, 1),
          ("b16 project", 24, 672, 160, 1), ("b17 expand", 24, 160, 960, 0), ("b17 project", 24, 960, 160, 1), ("b22 project", 12, 960, 272, 1),
          ("b23 expand", 12, 272, 1632, 0), ("b23 project", 12, 1632, 272, 1), ("b30 project", 12, 1632, 448, 1), ("b31 expand", 12, 448, 2688, 0),
          ("b31 project", 12, 2688, 448, 1), ("head", 12, 448, 1792, 0)]
out = (C.c_int * 10)()
print(f"{'layer':12s} {'M':>9s} {'K':>5s} {'N':>5s} g f |  BN res stg nbuf grid t/cta   smem ntn  cl")
for name, H, K, N, g in layers:
    M = B * H * H
    f = fold(M, K, N, bool(g), H * H)
    Mf, Kf, Nf = M // f, K * f, N * f
    rc = lib.dfv_gemm_plan_info(C.c_longlong(Mf), Kf, Nf, g, out)
    print(f"{name:12s} {Mf:9d} {Kf:5d} {Nf:5d} {g} {f} |", rc, " ".join(f"{v:4d}" for v in out))
