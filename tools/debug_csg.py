"""Dump where the GPU and the oracle disagree on the csg stress scene (per object kind / node)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location('t', os.path.join(ROOT, 'tests/test_gpu_rtc.py')); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
from oracle import oracle as orc
from rendering_learning_b200 import Context, rtc
ctx = Context(0)
sc = m.csg_stress_scene()
desc = sc.world.lower(); ctx.scene_upload(desc); cam = sc.camera.abi()
img, _ = ctx.render_rtc(cam, 1)
ref = orc.rtc_render(desc, cam, 1)
bad = (np.abs(m.u8(img.astype(np.float64)) - m.u8(ref)) > 1).any(axis=2)
rays = orc.rtc_camera_rays(cam, 1)
node, t, _ = orc.rtc_trace(desc, rays)
hits = ctx.trace_batch(rays[:, 0:3].astype(np.float32), rays[:, 3:6].astype(np.float32))
print("bad frac", bad.mean(), "hit id mismatches", (hits["node"] != node).sum())
node2 = node.reshape(bad.shape)
for n in np.unique(node2):
    sel = node2 == n
    print("node", n, "kind", desc.nodes[n][0] if n >= 0 else None, "pixels", sel.sum(), "bad", (bad & sel).sum())
mm = np.nonzero(hits["node"] != node)[0][:10]
for i in mm: print("ray", i, "gpu", hits["node"][i], hits["t"][i], "oracle", node[i], t[i])
np.savez_compressed(os.path.join(ROOT, "gpurun_out/csg_dbg.npz"), img=img, ref=ref, bad=bad)
