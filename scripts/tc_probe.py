import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stable_nerf_b200 import _lib
_lib.use_debug_library()  # hooks live in libsnerf_b200_dbg.so
lib = _lib.load()
for variant in (0, 1, 2, 3):
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    _lib.check(lib.snerf_tc_probe(_lib.ptr(out), variant, _lib.stream()), "probe")
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    print("variant", variant, "(N=%d, A from %s)" % (64 if variant & 1 else 128, "TMEM" if variant >= 2 else "smem"))
    for e, n in enumerate([0, 1, 2, 4, 8, 16, 32, 64]):
        print(f"  {n:3d} MMAs: issue {o[3*e]:6d}  done(t0) {o[3*e+1]:6d}  done(t255) {o[3*e+2]:6d}")
    print("  fence+bar", o[24], " ldtm64", o[25], " epilogue+bar", o[26])
