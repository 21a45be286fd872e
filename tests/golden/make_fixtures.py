"""Generate the committed fixtures from the read-only reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  python tests/golden/make_fixtures.py

Writes
  tests/golden/rtc_{mirror,obj,csg}.npz, tests/golden/ow_test.npz
      the reference's golden PPMs (RTC/tests/expectations/*.ppm, OW/tests/expectations/test.ppm)
      decoded to uint8 [H,W,3] + md5 of the original PPM text (so the byte-exact P3 encoders are
      checked too) — these pin the oracle;
  rendering_learning_b200/assets/{teapot_low,spot}.npz, spot_texture.npz
      the reference's objs/ parsed by the mirrored OBJ parsers into per-triangle arrays
      (inputs of BASELINE configs C3 and C5).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "rendering_learning_b200", "assets")


def ppm_fixture(src, dst):
    text = open(src, "rb").read()
    toks = text.split()
    assert toks[0] == b"P3"
    w, h, mx = int(toks[1]), int(toks[2]), int(toks[3])
    px = np.array(toks[4:], dtype=np.int64).reshape(h, w, 3).astype(np.uint8)
    np.savez_compressed(dst, pixels=px, md5=np.array(hashlib.md5(text).hexdigest()),
                        nbytes=np.array(len(text)))
    print(dst, px.shape, hashlib.md5(text).hexdigest())


def main():
    os.makedirs(ASSETS, exist_ok=True)
    rtc_exp = os.path.join(REF, "ray-tracer-challenge/tests/expectations")
    for name in ("mirror", "obj", "csg"):
        ppm_fixture(os.path.join(rtc_exp, f"test_{name}_scene.ppm"), os.path.join(GOLD, f"rtc_{name}.npz"))
    ppm_fixture(os.path.join(REF, "ray-tracing-one-weekend/tests/expectations/test.ppm"),
                os.path.join(GOLD, "ow_test.npz"))

    from rendering_learning_b200 import rtc
    obj = rtc.WavefrontObj.parse(open(os.path.join(REF, "objs/teapot-low.obj")).read())
    tris = obj.triangles()
    P = np.array([[list(p) for p in t.points] for t in tris], np.float64)
    N = np.array([[list(n) for n in (t.normals or [(0, 0, 0)] * 3)] for t in tris], np.float64)
    S = np.array([t.normals is not None for t in tris])
    np.savez_compressed(os.path.join(ASSETS, "teapot_low.npz"), tri_p=P, tri_n=N, tri_smooth=S)
    print("teapot", P.shape, int(S.sum()), "smooth; ignored lines", obj.ignored)

    from rendering_learning_b200 import ow
    o = ow.WavefrontObj.parse(open(os.path.join(REF, "objs/spot_triangulated.obj")).read())
    T = o.tris()
    P = np.array([t[0] for t in T], np.float64)
    UV = np.array([t[1] if t[1] is not None else [(0, 0)] * 3 for t in T], np.float64)
    N = np.array([t[2] if t[2] is not None else [(0, 0, 0)] * 3 for t in T], np.float64)
    HUV = np.array([t[1] is not None for t in T])
    HN = np.array([t[2] is not None for t in T])
    np.savez_compressed(os.path.join(ASSETS, "spot.npz"), tri_p=P, tri_uv=UV, tri_n=N, has_uv=HUV, has_n=HN)
    print("spot", P.shape, int(HUV.sum()), "with uv", int(HN.sum()), "with normals")

    # the OBJ TEXT itself (input data named by BASELINE.json's north_star), gzip-compressed: what the device-side OBJ
    # ingest parses on the GPU box, where /root/reference does not exist
    import gzip
    for name in ("teapot-low.obj", "spot_triangulated.obj"):
        raw = open(os.path.join(REF, "objs", name), "rb").read()
        with gzip.GzipFile(os.path.join(GOLD, name + ".gz"), "wb", mtime=0) as f:
            f.write(raw)
        print(name, len(raw), "bytes")

    from PIL import Image
    im = np.array(Image.open(os.path.join(REF, "objs/spot_texture.png")).convert("RGB"), np.uint8)
    np.savez_compressed(os.path.join(ASSETS, "spot_texture.npz"), rgb8=im)
    print("spot texture", im.shape)


if __name__ == "__main__":
    main()
