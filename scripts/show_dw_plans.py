"""Print the depthwise tile plans chosen for the B4 layers (host only; no GPU needed)."""
import ctypes as C, os, sys
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "deepfake_vit_b200", "libdfvit.so"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layers = [(48,190,3,1,1,1),(24,190,3,1,1,1),(144,190,3,2,0,1),(192,95,3,1,1,1),(192,95,5,2,2,2),(336,48,5,1,2,2),(336,48,3,2,0,1),
          (672,24,3,1,1,1),(672,24,5,1,2,2),(960,24,5,1,2,2),(960,24,5,2,1,2),(1632,12,5,1,2,2),(1632,12,3,1,1,1),(2688,12,3,1,1,1)]
out = (C.c_int * 10)()
print("    C   H k s | CB  L TW TH thr   smem tw th parts grid")
for (c, h, k, s, pl, ph) in layers:
    rc = lib.dfv_dwconv_plan_info(1, B, h, h, c, k, s, pl, ph, out)
    print(f"{c:5d} {h:3d} {k} {s} |", rc, " ".join(f"{v:3d}" for v in out))
