import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ["PTC_BUILD_TIMING"] = "1"
from pathed_b200 import load_scene
for i in range(2):
    t = time.time(); ctx = load_scene("scenes/dragon.json", 1024, 1024); print("load_scene", time.time() - t, ctx.stats().bvh_build_ms); ctx.close()
