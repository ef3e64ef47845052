"""Wall-clock breakdown of the host-buffer call sequence (create / render / destroy) for one config."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from raytracer_go_b200 import api, scenes

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
data, opts = scenes.build_config(cfg)
cam = api.camera_from_options(opts)
for it in range(4):
    t0 = time.perf_counter()
    sc = api.Scene(data, 0)
    t1 = time.perf_counter()
    rgb, _, st = sc.render(cam)
    t2 = time.perf_counter()
    sc.close()
    t3 = time.perf_counter()
    print(f"{cfg} iter {it}: create {1e3 * (t1 - t0):.1f} ms, render {1e3 * (t2 - t1):.1f} ms (device {st.ms_render:.1f}, "
          f"call {st.ms_total:.1f}), destroy {1e3 * (t3 - t2):.1f} ms", flush=True)
