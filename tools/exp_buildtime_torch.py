"""Scene build times inside a process that has torch's CUDA context up (as bench.py does): cold and warm (development)."""
import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
torch.cuda.init(); x = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
import toymeshpathtracer_b200 as tm
from bench import scene_obj_path
path = scene_obj_path("sponza"); tris, mn, mx = tm.load_scene(path)
t0 = time.perf_counter(); tm.Scene(tris[:2]).close(); print("warm-up (2 triangles) wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
for k in range(4):
    t0 = time.perf_counter(); s = tm.Scene(tris); wall = (time.perf_counter() - t0) * 1e3
    i = s.info(); print("build %d: build_ms %.2f wall %.2f" % (k, i["build_ms"], wall)); s.close()
