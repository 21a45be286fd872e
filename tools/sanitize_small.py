"""Tiny renders of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import Context, ow, scenes
ctx = Context(0)
w, p = scenes.ow_cover_world(), scenes.ow_cover_params(image_width=64, samples_per_pixel=70, max_depth=12)
ctx.scene_upload(ow.lower_world(w)); img, st = ctx.render_ow(p.abi()); print("cover", img.mean(), st.overflow, flush=True)
w, p = scenes.ow_cow_world(), scenes.ow_cow_params(image_width=48, samples_per_pixel=8, max_depth=8)
ctx.scene_upload(ow.lower_world(w)); img, st = ctx.render_ow(p.abi()); print("cow", img.mean(), st.overflow, flush=True)
w, p = scenes.ow_cornell_smoke(); p.image_width = 32; p.samples_per_pixel = 8; p.max_depth = 8
ctx.scene_upload(ow.lower_world(w)); img, st = ctx.render_ow(p.abi()); print("smoke", img.mean(), st.overflow, flush=True)
sc = scenes.rtc_csg_scene(48, 32); ctx.scene_upload(sc.world.lower()); img, st = ctx.render_rtc(sc.camera.abi(), 1); print("csg", img.mean(), flush=True)
sc = scenes.rtc_obj_scene(48, 32); ctx.scene_upload(sc.world.lower()); img, st = ctx.render_rtc(sc.camera.abi(), 1); print("teapot", img.mean(), flush=True)
ctx.close(); print("done", flush=True)
