"""SURVEY.md §8(f) rank 4 — the callers either side of the path: the JSON scene loader (mray_b200/host/scene_loader.cpp,
a SceneLoaderI) and the head-less run command (mray_b200/host/run_main.cpp). Both speak only TracerI, so on the CPU they are
checked with the UNMODIFIED REFERENCE TRACER behind them (oracle/_ref/libTracerDLL_CPU.so) against the reference-rendered golden
images and against the array-driven TracerI driver; the -m gpu tests put the B200 plugin behind the same commands."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "mray_b200", "lib")
RUN, LOADER = os.path.join(LIB, "mray_b200_run"), os.path.join(LIB, "libSceneLoaderB200.so")
PLUGIN = os.path.join(LIB, "libTracerDLL_B200.so")
REF_DLL = os.path.join(ROOT, "oracle", "_ref", "libTracerDLL_CPU.so")
DOC_SCENE = os.path.join(GOLDEN, "scene_cornell_doc.json")
have_tools = os.path.exists(RUN) and os.path.exists(LOADER)
needs_ref = pytest.mark.skipif(not (have_tools and os.path.exists(REF_DLL) and O.driver_available()), reason="host tools / reference tracer were not prebuilt")
needs_plugin = pytest.mark.skipif(not (have_tools and os.path.exists(PLUGIN)), reason="host tools / plugin were not prebuilt")
bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, 3).mean(axis=(1, 3))
rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def run(tracer, scene_path, out, res=64, spp=256, extra=()):
    r = subprocess.run([RUN, "--tracer", tracer, "--loader", LOADER, "--scene", scene_path, "-r", f"{res}x{res}", "--spp", str(spp),
                        "--rr", "2,20", "--out", out, *extra], capture_output=True, text=True)
    stats = None
    for line in r.stdout.splitlines():
        if line.startswith("{\"scene_load_ms\""):
            stats = json.loads(line)
    return r, stats


def flatten_doc_scene():
    """The documentation-style fixture turned into arrays + per-surface matrices by an independent reading of the format
    (comments, arrayed nodes, matrix / trs layouts): -> cornell-style dict with `transforms` [surface, 3, 4]."""
    text = open(DOC_SCENE).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r",(\s*[\]}])", r"\1", text)
    s = json.loads(text)

    def items(nodes):
        for n in nodes:
            if isinstance(n["id"], list):
                for k, i in enumerate(n["id"]):
                    yield i, {key: (v[k] if key not in ("id", "type", "layout", "tag") else v) for key, v in n.items()}
            else:
                yield n["id"], n

    def trs(t, r, sc):
        rx, ry, rz = np.deg2rad(r)
        S = np.diag(list(sc) + [1.0])
        def rot(a, ax):
            c, s_ = np.cos(a), np.sin(a)
            m = np.eye(4)
            i, j = [(1, 2), (2, 0), (0, 1)][ax]
            m[i, i], m[i, j], m[j, i], m[j, j] = c, -s_, s_, c
            return m
        T = np.eye(4); T[:3, 3] = t
        return T @ rot(rz, 2) @ rot(ry, 1) @ rot(rx, 0) @ S

    xf = {}
    for i, n in items(s["Transforms"]):
        if n["type"] == "Identity":
            xf[i] = np.eye(4)
        elif n["layout"] == "matrix":
            xf[i] = np.array(n["matrix"], float).reshape(4, 4)
        else:
            xf[i] = trs(n.get("translate", [0, 0, 0]), n.get("rotate", [0, 0, 0]), n.get("scale", [1, 1, 1]))
    prims = {}
    for i, n in items(s["Primitives"]):
        p = np.array(n["position"], float)
        idx = np.array(n["index"]) if n["tag"] == "nodeTriangleIndexed" else np.arange(len(p)).reshape(-1, 3)
        prims[i] = (p, idx)
    mats = {i: n["albedo"] for i, n in items(s["Materials"])}
    # one batch per surface in the primitive's LOCAL space + its (T)Single matrix, the light last — the same TracerI calls the
    # loader makes, minus the sharing of primitive batches (the reference keeps shading frames in local space under a
    # (T)Single transform, so a flattened world-space copy would NOT be the same scene for it)
    pos, idx, mat, albedo, xforms = [], [], [], [], []
    nv = 0
    for k, sf in enumerate(s["Surfaces"]):
        p, t = prims[sf["primitive"]]
        pos.append(p); idx.append(t + nv); mat += [k] * len(t); nv += len(p)
        albedo.append(mats[sf["material"]]); xforms.append(xf[sf["transform"]][:3])
    light = next(n for i, n in items(s["Lights"]) if n["type"] == "Primitive")
    ls = s["LightSurfaces"][0]
    p, t = prims[light["primitive"]]
    light_id = len(s["Surfaces"])
    pos.append(p); idx.append(t + nv); mat += [light_id] * len(t)
    albedo.append([0, 0, 0]); xforms.append(xf[ls["transform"]][:3])
    cam = s["Cameras"][0]
    return dict(positions=np.ascontiguousarray(np.concatenate(pos), np.float32), indices=np.ascontiguousarray(np.concatenate(idx), np.uint32),
                material=np.array(mat, np.uint32), albedo=np.array(albedo, np.float32), transforms=np.array(xforms, np.float32),
                light_material=light_id, radiance=np.array(light["radiance"], np.float32),
                camera=dict(eye=cam["position"], gaze=cam["gaze"], up=cam["up"], fov_y_deg=cam["fov"]))


def check_against_golden(tracer, tmp_path, spp, tol, res=64):
    c = scenes.cornell_box()
    sp, op = str(tmp_path / "cornell.json"), str(tmp_path / "o.pfm")
    open(sp, "w").write(scenes.mray_scene_json(c, res, res))
    r, st = run(tracer, sp, op, res=res, spp=spp, extra=("--seed", "5"))
    assert r.returncode == 0, r.stderr[-800:]
    assert st["paths"] == res * res * spp and st["surfaces"] == 3 and st["aabb"] == [-1, 0, -1, 1, 2, 1]
    img = scenes.read_pfm(op)
    ref = np.load(os.path.join(GOLDEN, "render_cornell64_spp16384.npz"))["img"].astype(np.float32)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.015), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel(bm(img, 8), bm(ref, 8))
    assert e <= tol, e


def check_doc_scene(tracer, tmp_path, spp, tol):
    op = str(tmp_path / "doc.pfm")
    r, st = run(tracer, DOC_SCENE, op, spp=spp, extra=("--seed", "7"))
    assert r.returncode == 0, r.stderr[-800:]
    # one plane + one cube, instanced (the reference reports 2 concrete accelerators; the plugin keeps the light's un-culled
    # copy of the plane apart from the walls' culled one: 3)
    assert st["surfaces"] == 7 and st["instances"] == 8 and 2 <= st["accelerators"] <= 3
    img = scenes.read_pfm(op)
    c = flatten_doc_scene()
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    ref, w, _ = O.driver_render(tracer, b, c["albedo"], c["light_material"], c["radiance"], c["camera"], 64, 64, spp, seed=11, near_far=(0.005, 90.0),
                                batch_transforms=c["transforms"], burst_size=64 if tracer == PLUGIN else 1)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.02), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel(bm(img, 8), bm(ref, 8))
    assert e <= tol, e


@needs_ref
def test_loader_and_run_command_with_the_reference_tracer(tmp_path):
    check_against_golden(REF_DLL, tmp_path, 256, 1e-3)


@needs_ref
def test_documentation_style_scene_with_the_reference_tracer(tmp_path):
    check_doc_scene(REF_DLL, tmp_path, 128, 3e-3)


@pytest.mark.skipif(not have_tools or not os.path.exists(REF_DLL), reason="host tools / reference tracer were not prebuilt")
def test_scene_errors_are_reported(tmp_path):
    cases = {"{ \"Cameras\": [ ": "json:", "{}": "does not contain", "[1, 2]": "one JSON object"}
    good = json.loads(scenes.mray_scene_json(scenes.cornell_box(), 8, 8))
    bad = dict(good); bad["Surfaces"] = [{"transform": 0, "material": 77, "primitive": 0}]
    cases[json.dumps(bad)] = "Material(77) is not defined"
    bad = dict(good); bad["Primitives"] = [dict(good["Primitives"][0], type="Sphere")]
    cases[json.dumps(bad)] = "not supported by this loader"
    for k, (text, expect) in enumerate(cases.items()):
        sp = str(tmp_path / f"bad{k}.json")
        open(sp, "w").write(text)
        r, _ = run(REF_DLL, sp, str(tmp_path / "x.pfm"), res=8, spp=1)
        assert r.returncode != 0 and expect in r.stderr, (text[:40], r.stderr[-300:])
    r, _ = run(REF_DLL, str(tmp_path / "missing.json"), str(tmp_path / "x.pfm"), res=8, spp=1)
    assert r.returncode != 0 and "not found" in r.stderr


@pytest.mark.gpu
@needs_plugin
def test_loader_and_run_command_with_the_b200_plugin(tmp_path):
    check_against_golden(PLUGIN, tmp_path, 16384, 6e-4)


@pytest.mark.gpu
@needs_plugin
def test_documentation_style_scene_with_the_b200_plugin(tmp_path):
    if not O.driver_available():
        pytest.skip("TracerI driver was not prebuilt")
    check_doc_scene(PLUGIN, tmp_path, 8192, 6e-4)


@pytest.mark.gpu
@needs_plugin
def test_textures_alpha_maps_and_skysphere_through_the_loader(tmp_path):
    """PFM textures ("Pf" alpha map on a surface, "PF" radiance map on a Skysphere boundary light) through the scene file, against
    the reference's own renders of the same scenes."""
    c = scenes.cornell_alpha()
    scenes.write_pfm(str(tmp_path / "alpha.pfm"), c["alpha_texture"]["data"])
    tex = [dict(id=5, file="alpha.pfm", interpolation="Nearest", edgeResolve="Clamp", isColor=False)]
    sp, op = str(tmp_path / "alpha.json"), str(tmp_path / "alpha_out.pfm")
    open(sp, "w").write(scenes.mray_scene_json(c, 64, 64, textures=tex, alpha_map=[None, None, None, None, 5], uvs=c["uvs"]))
    r, st = run(PLUGIN, sp, op, spp=16384, extra=("--seed", "3", "--burst", "64"))
    assert r.returncode == 0, r.stderr[-800:]
    ref = np.load(os.path.join(GOLDEN, "render_cornell64_alpha_spp16384.npz"))["img"].astype(np.float32)
    img = scenes.read_pfm(op)
    assert rel(bm(img, 2), bm(ref, 2)) <= 1e-3
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
    # skysphere with an HDR map (concentric-octahedral) next to the area light, spectral renderer
    path = os.path.join(GOLDEN, "render_cornell64_sky_coocta_spectral_spp16384.npz")
    from mray_b200 import spectral
    if not os.path.exists(path) or not spectral.available():
        pytest.skip("golden image / spectral LUT was not generated")
    c = scenes.cornell_open(keep_light=True)
    scenes.write_pfm(str(tmp_path / "sky.pfm"), scenes.sky_texture()["data"][..., :3])
    tex = [dict(id=9, file="sky.pfm", interpolation="Linear", edgeResolve="Wrap", isIlluminant=True)]
    sp, op = str(tmp_path / "sky.json"), str(tmp_path / "sky_out.pfm")
    open(sp, "w").write(scenes.mray_scene_json(c, 64, 64, textures=tex, boundary=dict(type="Skysphere_CoOcta", texture=9)))
    r, st = run(PLUGIN, sp, op, spp=16384, extra=("--seed", "4", "--burst", "64", "--renderer", "PathTracerSpectral"))
    assert r.returncode == 0, r.stderr[-800:]
    ref = np.load(path)["img"].astype(np.float32)
    img = scenes.read_pfm(op)
    assert rel(bm(img, 4), bm(ref, 4)) <= 1e-3
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)


@needs_plugin
@pytest.mark.gpu
def test_generated_mips_through_the_loader_and_run_command(tmp_path):
    """`--genMips` (TracerParameters.genMips + mipGenFilter) on a scene file whose white material reads a strongly tiled PFM texture:
    the plugin completes the chain at upload and the path's ray cones pick the levels — the same image as the C-ABI renderer given
    the same texture with gen_mips, and not the image of the single-level texture."""
    import mray_b200
    from mray_b200 import capi
    c = scenes.cornell_box()
    rng = np.random.default_rng(21)
    tex = rng.random((64, 64, 3)).astype(np.float32)
    g = np.arange(64) // 16
    tex[..., 0] = 0.5 * tex[..., 0] + 0.15 * ((g[:, None] + g[None, :]) % 3)        # low-frequency structure: coarse levels differ
    tex[..., 1] = 0.5 * tex[..., 1] + 0.2 * ((2 * g[:, None] + g[None, :]) % 3)
    uvs = np.tile(np.array([[0, 0], [96, 0], [96, 96], [0, 96]], np.float32), (c["positions"].shape[0] // 4, 1))
    scenes.write_pfm(str(tmp_path / "albedo.pfm"), tex)
    nodes = [dict(id=7, file="albedo.pfm", interpolation="Linear", edgeResolve="Wrap")]
    sp, op = str(tmp_path / "mips.json"), str(tmp_path / "mips_out.pfm")
    open(sp, "w").write(scenes.mray_scene_json(c, 64, 64, textures=nodes, albedo_texture=[7, None, None, None], uvs=uvs))
    r, st = run(PLUGIN, sp, op, spp=16384, extra=("--seed", "5", "--burst", "64", "--genMips", "Gaussian,2"))
    assert r.returncode == 0, r.stderr[-800:]
    img = scenes.read_pfm(op)
    # the same through the C-ABI
    ctx = mray_b200.Context(0)
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    rgba = np.concatenate([tex, np.ones((64, 64, 1), np.float32)], axis=-1)
    imgs = {}
    for name, t in (("mips", dict(data=rgba, gen_mips=("Gaussian", 2.0))), ("flat", dict(data=rgba))):
        rr = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 16384, seed=6,
                           rr_range=(2, 20), textures=[t], albedo_texture=[0, -1, -1], vertex_uvs=uvs)
        imgs[name], s2 = rr.render(batch=64); rr.close()
    acc.close(); ctx.close()
    e = rel(bm(img, 2), bm(imgs["mips"], 2))
    assert e <= 1e-3, e
    assert np.allclose(img.mean(axis=(0, 1)), imgs["mips"].mean(axis=(0, 1)), rtol=0.01)
    # without --genMips the run command renders the single-level image
    r, st = run(PLUGIN, sp, op, spp=4096, extra=("--seed", "7", "--burst", "64"))
    assert r.returncode == 0, r.stderr[-800:]
    flat = scenes.read_pfm(op)
    assert rel(bm(flat, 2), bm(imgs["flat"], 2)) <= 1.5e-3
