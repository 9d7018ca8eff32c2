import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from pathed_b200 import load_scene
W, H = 1920, 1080
ctx = load_scene("scenes/teapot.json", W, H)
found = []
s = 0
while s < 8192 and len(found) < 3:
    img = ctx.render(0x5EED, s, 16, 0, 10)
    bad = ~np.isfinite(img).all(-1)
    if bad.any():
        for k in range(16):
            one = ctx.render(0x5EED, s + k, 1, 0, 10)
            b1 = ~np.isfinite(one).all(-1)
            for (r, c) in np.argwhere(b1):
                found.append((s + k, int(r), int(c), one[r, c].tolist()))
    s += 16
print("non-finite samples:", found)
if found:
    from oracle_binding import oracle_scene
    o = oracle_scene("scenes/teapot.json", W, H)
    o.set_option("brute_force", 0)
    for (smp, r, c, val) in found[:3]:
        ref = o.render(0x5EED, smp, 1, 0, 10)
        print("sample", smp, "pixel", r, c, "cuda", val, "oracle", ref[r, c].tolist())
        for lb in range(0, 11):
            g = ctx.render(0x5EED, smp, 1, 0, lb)[r, c]
            print("  lastBounce", lb, g.tolist(), o.render(0x5EED, smp, 1, 0, lb)[r, c].tolist() if lb in (1, 2, 3, 10) else "")
