import sys, time; sys.path.insert(0,'/root/repo')
import toymeshpathtracer_b200 as tm
from bench import scene_obj_path
path = scene_obj_path("sponza"); tris, mn, mx = tm.load_scene(path)
tm.Scene(tris[:2]).close()
for flags in (0, 1, 0, 1):
    t0 = time.perf_counter(); s = tm.Scene(tris, flags=flags); wall = (time.perf_counter()-t0)*1e3
    i = s.info(); print("builder", i["builder"], "build_ms %.2f wall %.2f nodes %d depth %d" % (i["build_ms"], wall, i["node_count"], i["max_depth"])); s.close()
