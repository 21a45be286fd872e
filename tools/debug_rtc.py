import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from rendering_learning_b200 import Context, rtc
import test_gpu_rtc as T
ctx = Context(0)
sc = T.all_shapes_scene()
desc = sc.world.lower()
ctx.scene_upload(desc)
cam = sc.camera.abi()
img, st = ctx.render_rtc(cam, 1)
ref = orc.rtc_render(desc, cam, 1)
d = np.abs(T.u8(img.astype(np.float64)) - T.u8(ref)).max(axis=2)
rays = orc.rtc_camera_rays(cam, 1)
node, t, _ = orc.rtc_trace(desc, rays)
node = node.reshape(d.shape)
bad = d > 1
print('bad frac', bad.mean(), 'max', d.max())
kinds = {i: n[0] for i, n in enumerate(sd for sd in desc.nodes)}
for nid in np.unique(node[bad]):
    m = bad & (node == nid)
    print('node', nid, 'kind', desc.nodes[nid][0] if nid >= 0 else None, 'bad px', m.sum(), 'of', (node == nid).sum(), 'maxdiff', d[m].max())
ys, xs = np.nonzero(bad)
for y, x in list(zip(ys, xs))[:12]:
    print(x, y, 'node', node[y, x], 'gpu', img[y, x], 'ref', ref[y, x])
# per depth
for depth in (0, 1):
    sc.world.max_reflection_depth = depth
    dd = sc.world.lower(); ctx.scene_upload(dd)
    i2, _ = ctx.render_rtc(cam, 1); r2 = orc.rtc_render(dd, cam, 1)
    d2 = np.abs(T.u8(i2.astype(np.float64)) - T.u8(r2)).max(axis=2)
    print('depth', depth, 'bad frac', (d2 > 1).mean())
    for nid in np.unique(node[d2 > 1]):
        m = (d2 > 1) & (node == nid)
        print('   node', nid, 'kind', desc.nodes[nid][0] if nid >= 0 else None, 'bad px', m.sum())
