"""CPU-only tests: layouts against the C header, the C-ABI library's exported symbols, host-logic known answers
restated from the reference's own unit tests, oracle properties, and the world_size-2 sharding logic over gloo."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from rend3_b200 import glam, layouts
from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL, ENTRY_POINTS, CUDA_LIB_PATH
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings, per_camera_header
from rend3_b200.scenes import cloud_camera, cube_field_scene, object_cloud_records
from rend3_b200.world import allocate_shadow_atlas, frustum_from_matrix

from oracle import load_oracle_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ layouts
def test_numpy_layouts_match_c_header():
    """sizeof/offsetof of every struct in include/r3_layouts.h, probed with the system C compiler."""
    probes = {
        "r3_object": (layouts.OBJECT_DTYPE, ["transform", "sphere_center", "sphere_radius", "first_index", "index_count", "material_index", "attr_offset", "enabled"]),
        "r3_camera_header": (layouts.CAMERA_HEADER_DTYPE, ["view", "view_proj", "shadow_index", "frustum", "resolution", "flags", "object_count"]),
        "r3_object_matrices": (layouts.OBJECT_MATRICES_DTYPE, ["model_view", "model_view_proj"]),
        "r3_object_culling_info": (layouts.CULLING_INFO_DTYPE, ["invocation_start", "invocation_end", "object_id", "region_id", "base_region_invocation", "local_region_id", "previous_global_invocation", "atomic_capable"]),
        "r3_batch_data": (layouts.BATCH_DTYPE, ["total_objects", "total_invocations", "batch_base_invocation", "object_culling_information"]),
        "r3_region": (layouts.REGION_DTYPE, ["job_index", "bind_group_index", "material_key"]),
        "r3_indirect_call": (layouts.INDIRECT_CALL_DTYPE, ["vertex_count", "instance_count", "base_index", "vertex_offset", "base_instance"]),
        "r3_frame_uniforms": (layouts.FRAME_UNIFORMS_DTYPE, ["view", "view_proj", "origin_view_proj", "inv_view", "inv_view_proj", "inv_origin_view_proj", "frustum", "ambient", "resolution"]),
        "r3_directional_light": (layouts.DIRECTIONAL_LIGHT_DTYPE, ["view_proj", "color", "direction", "inv_resolution", "atlas_offset", "atlas_size"]),
        "r3_point_light": (layouts.POINT_LIGHT_DTYPE, ["position", "color", "radius"]),
        "r3_material": (layouts.MATERIAL_DTYPE, ["textures", "uv_transform0", "uv_transform1", "albedo", "emissive", "roughness", "metallic", "reflectance", "clear_coat", "clear_coat_roughness", "anisotropy", "ambient_occlusion", "alpha_cutout", "flags"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/r3_layouts.h"', "int main(void){"]
    for name, (_dt, fields) in probes.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write("\n".join(lines))
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, src, "-o", exe], check=True)
        out = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, (dt, fields) in probes.items():
        assert int(out[name]) == dt.itemsize, name
        for f in fields:
            assert int(out[f"{name}.{f}"]) == dt.fields[f][1], f"{name}.{f}"


# ------------------------------------------------------------------ the C ABI library
def test_cuda_library_exports_every_entry_point_and_fails_loudly_without_a_gpu():
    assert os.path.exists(CUDA_LIB_PATH), "librend3_b200.so is not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(CUDA_LIB_PATH)
    header = open(os.path.join(ROOT, "include", "rend3_b200.h")).read()
    for name in ENTRY_POINTS:
        assert hasattr(lib, "r3_" + name), f"r3_{name} not exported"
        assert f"r3_{name}(" in header, f"r3_{name} not declared in include/rend3_b200.h"
    lib.r3_abi_version.restype = ctypes.c_uint32
    assert lib.r3_abi_version() == 1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        ctx = ctypes.c_void_p()
        rc = lib.r3_ctx_create(0, ctypes.byref(ctx))
        assert rc == -4 and not ctx.value, "context creation must fail with R3_E_NO_DEVICE: there is no CPU fallback"


def test_cpp_host_driver_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """include/rend3_b200.hpp + rend3_b200/host/r3_frame: the native host mirror links against the C ABI; without a CUDA device it
    must stop with the library's error (no CPU path behind it)."""
    import torch

    from rend3_b200.scene_io import dump_scene

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rend3_b200", "host", "r3_frame")
    assert os.path.exists(exe), "__graft_entry__.build() compiles the host driver"
    scene = tmp_path / "s.r3s"
    ev = cube_field_scene(n_objects=20, seed=3, resolution=(64, 48), shadow_resolution=64, shadow_distance=40.0, pull_back=4.0, extent=4.0)
    dump_scene(str(scene), ev, (64, 48))
    assert os.path.getsize(scene) > 128 * 20
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the driver is exercised by tests/test_gpu_parity.py")
    r = subprocess.run([exe, str(scene), str(tmp_path / "o.r3o")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CUDA device" in r.stderr, r.stderr


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may reference oracle/."""
    pkg = os.path.join(ROOT, "rend3_b200")
    for d, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "r3o_" not in text and "libr3_oracle" not in text, f"{f} references the oracle"


# ------------------------------------------------------------------ known answers restated from the reference's unit tests
def test_round_up_and_pot_known_answers():
    """rend3/src/util/math.rs:66-96 (round_up / div_round_up) as used by batch_objects' 256-padding."""
    r = lambda v, m: (v + m - 1) // m * m
    assert [r(0, 256), r(1, 256), r(256, 256), r(257, 256), r(12, 256)] == [0, 256, 256, 512, 256]


def test_shadow_atlas_known_answers():
    """rend3/src/managers/directional/shadow_alloc.rs:146-320 — the packing cases of the reference's own tests."""
    assert allocate_shadow_atlas([]) is None
    dims, maps = allocate_shadow_atlas([(0, 16)], 16)
    assert dims == (16, 16) and maps == [(0, 0, 16, 0)]
    dims, maps = allocate_shadow_atlas([(0, 16), (1, 16)], 32)
    assert dims == (32, 16) and maps == [(0, 0, 16, 0), (16, 0, 16, 1)]
    dims, maps = allocate_shadow_atlas([(0, 16), (1, 16)], 16)   # single column available -> stacked
    assert dims == (16, 32) and maps == [(0, 0, 16, 0), (0, 16, 16, 1)]
    dims, maps = allocate_shadow_atlas([(0, 16), (1, 8), (2, 8)], 32)   # two 8s share a quadtree root
    assert dims == (32, 16) and (0, 0, 16, 0) in maps and (16, 0, 8, 1) in maps and (24, 0, 8, 2) in maps
    dims, maps = allocate_shadow_atlas([(i, 2048) for i in range(4)], 8192)
    assert dims == (8192, 2048) and [m[0] for m in maps] == [0, 2048, 4096, 6144]


def test_frustum_planes_are_normalised_and_cull_like_the_reference():
    cam = cloud_camera()
    fr = frustum_from_matrix(cam.view_proj)
    assert np.allclose(np.linalg.norm(fr[:, :3], axis=1), 1.0, atol=1e-6)
    # a sphere at the camera target is inside, one far behind the camera is outside (util/frustum.rs:148-161)
    eye = cam.location()
    fwd = -eye / np.linalg.norm(eye)
    inside = lambda c, r: all(float(np.dot(p[:3], c) + p[3]) >= -r for p in fr)
    assert inside(eye + 10 * fwd, 0.1) and not inside(eye - 50 * fwd, 1.0)


# ------------------------------------------------------------------ oracle properties at sizes the oracle finishes in seconds
def test_oracle_cull_bake_properties():
    n = 100_000
    rec = object_cloud_records(n, seed=3)
    cam = cloud_camera()
    b = load_oracle_backend()
    b.set_objects(rec)
    b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cam, CAMERA_VIEWPORT, (1920, 1080), 1, n), CB_BAKE | CB_CULL)
    vis = b.readback_visible(CAMERA_VIEWPORT)
    # the visible set is exactly the enabled spheres inside the 5 planes, evaluated independently in float64 away from the boundary
    fr = cam.world_frustum.astype(np.float64)
    dist = rec["sphere_center"].astype(np.float64) @ fr[:, :3].T + fr[:, 3]
    margin = dist + rec["sphere_radius"].astype(np.float64)[:, None]
    clearly_in = (margin > 1e-3).all(axis=1) & (rec["enabled"] != 0)
    clearly_out = (margin < -1e-3).any(axis=1) | (rec["enabled"] == 0)
    mask = np.zeros(n, dtype=bool)
    mask[vis] = True
    assert mask[clearly_in].all() and not mask[clearly_out].any()
    assert np.all(np.diff(vis.astype(np.int64)) > 0)
    # linearity of the bake: MVP == view_proj * T to f32 accuracy against float64
    mats = b.readback_object_matrices(CAMERA_VIEWPORT, 0, 64)
    for i in range(64):
        if rec["enabled"][i]:
            t = rec["transform"][i].reshape(4, 4).astype(np.float64)
            ref = (cam.view_proj.astype(np.float64).T @ t.T).T
            assert np.allclose(mats["model_view_proj"][i].reshape(4, 4), ref, rtol=1e-5, atol=1e-3)


def test_oracle_batches_are_sorted_and_padded_like_the_reference():
    res = (320, 180)
    ev = cube_field_scene(n_objects=700, seed=2, resolution=res, n_dir_lights=0, pull_back=10.0, extent=20.0, subdivisions=(1, 3), material_count=2)
    b = load_oracle_backend()
    BaseRenderGraph(b).add_to_graph(ev, res, 1, BaseRenderGraphSettings())
    batches, regions = b.readback_batches(CAMERA_VIEWPORT)
    vis = set(int(v) for v in b.readback_visible(CAMERA_VIEWPORT))
    seen, base = [], 0
    eye = ev.camera.location()
    for bd in batches:
        assert bd["batch_base_invocation"] == base and bd["total_objects"] <= 256
        inv = 0
        for info in bd["object_culling_information"][: bd["total_objects"]]:
            assert info["invocation_start"] == inv and info["invocation_start"] % 256 == 0        # batching.rs:235
            tris = int(ev.object_buffer["index_count"][info["object_id"]]) // 3
            assert info["invocation_end"] - info["invocation_start"] == tris
            inv += (tris + 255) // 256 * 256
            seen.append(int(info["object_id"]))
            assert info["previous_global_invocation"] == 0xFFFFFFFF                                # first frame (batching.rs:226)
        assert inv == bd["total_invocations"]
        base += inv
    assert set(seen) == vis and len(seen) == len(vis)
    d2 = np.sum((ev.object_location[seen] - eye) ** 2, axis=1)
    assert np.all(np.diff(d2) >= -1e-3), "opaque objects are sorted front to back (batching.rs:53-79)"
    assert len(regions) >= 1 and all(r["material_key"] == 0 for r in regions)


def test_culling_buffers_ping_pong_like_suballoc():
    """InputOutputBuffer (suballoc.rs:66-222): partitions flip every cull; the previous frame's visibility bits feed
    the residual decision, so a static second frame has an empty residual list."""
    res = (256, 144)
    ev = cube_field_scene(n_objects=300, seed=4, resolution=res, n_dir_lights=0, pull_back=8.0, extent=15.0)
    b = load_oracle_backend()
    g = BaseRenderGraph(b)
    g.add_to_graph(ev, res, 1, BaseRenderGraphSettings())
    first_pred = b.readback_draw_calls(CAMERA_VIEWPORT, 0)["vertex_count"].sum()
    first_resid = b.readback_draw_calls(CAMERA_VIEWPORT, 1)["vertex_count"].sum()
    img1 = b.readback_ldr().copy()
    assert first_pred == first_resid > 0                      # frame 1: everything is residual
    g.add_to_graph(ev, res, 1, BaseRenderGraphSettings(), upload=False)
    assert b.readback_draw_calls(CAMERA_VIEWPORT, 1)["vertex_count"].sum() == 0
    assert b.readback_draw_calls(CAMERA_VIEWPORT, 0)["vertex_count"].sum() <= first_pred   # hi-Z occlusion may only remove
    assert np.array_equal(img1, b.readback_ldr()), "a static scene renders identically through the predicted pass"


@pytest.mark.parametrize("far_first", [True, False])
@pytest.mark.parametrize("samples", [1, 4])
def test_oracle_blend_routine_known_answer(far_first, samples):
    """pbr_forward_rendering_transparent (base.rs:181): ALPHA_BLENDING in back-to-front object order with depth test AND
    depth write; the reference holds no golden for it, so the oracle is pinned to the blend equation evaluated by hand."""
    import blend_case

    orc = load_oracle_backend()
    r = blend_case.build(orc, far_first)
    r.render_frame(64, samples)
    want = blend_case.expected(far_first)
    hdr = orc.readback_hdr_f32()
    assert np.array_equal(hdr.reshape(-1, 4), np.broadcast_to(want, (64 * 64, 4))), (hdr[32, 32], want)
    assert np.allclose(orc.readback_depth(), 0.6, rtol=0, atol=1e-6)   # the nearest transparent layer wrote depth
    st = orc.forward_stats()
    assert st[3] == 64 * 64 * samples * (2 if far_first else 1)
    _, regions = orc.readback_batches(CAMERA_VIEWPORT)
    assert sorted(int(k) for k in regions["material_key"]) == [0, 2]


@pytest.mark.parametrize("sample_type", ["linear", "nearest"])
@pytest.mark.parametrize("srgb", [False, True])
def test_oracle_texture_sampling_known_answers(sample_type, srgb):
    """textureSampleGrad of the bindless d2 table (rule R9): texel centres on pixel centres reproduce the texture; half the
    resolution selects mip 1 exactly; uv_transform0 x2 tiles it (Repeat); *Srgb formats decode before filtering."""
    import texture_case as tcase
    from rend3_b200.world import Texture

    data = tcase.checker_texture(64, seed=3)
    tex = Texture(data, srgb=srgb)
    decode = tcase.srgb_decode if srgb else (lambda a: a.astype(np.float64) / 255.0)
    orc = load_oracle_backend()
    tcase.build(orc, tex, sample_type).render_frame(64)
    hdr = orc.readback_hdr_f32()
    assert np.abs(hdr - decode(data)).max() < 2e-5, "level 0 at one texel per pixel"
    # one texel of mip 1 per pixel
    orc = load_oracle_backend()
    tcase.build(orc, tex, sample_type).render_frame(32)
    assert np.abs(orc.readback_hdr_f32() - decode(tex.levels()[1])).max() < 2e-5, "lambda = 1 selects mip 1"
    # two tiles per axis at 128 pixels: still one texel per pixel
    orc = load_oracle_backend()
    tcase.build(orc, tex, sample_type, uv_scale=2.0).render_frame(128)
    assert np.abs(orc.readback_hdr_f32() - np.tile(decode(data), (2, 2, 1))).max() < 2e-5, "Repeat addressing"
    if sample_type == "linear":
        # magnification x2: pixel centres sit a quarter texel off the texel centres -> bilinear mix of four texels
        orc = load_oracle_backend()
        tcase.build(orc, Texture(data, srgb=srgb, mips="none"), "linear").render_frame(128)
        d = decode(data)
        pad = np.pad(d, ((1, 1), (1, 1), (0, 0)), mode="wrap")
        y, x = np.mgrid[0:128, 0:128]
        fx, fy = (x + 0.5) / 2 - 0.5, (y + 0.5) / 2 - 0.5
        x0, y0 = np.floor(fx).astype(int), np.floor(fy).astype(int)
        wx, wy = (fx - x0)[..., None], (fy - y0)[..., None]
        want = (pad[y0 + 1, x0 + 1] * (1 - wx) + pad[y0 + 1, x0 + 2] * wx) * (1 - wy) + (pad[y0 + 2, x0 + 1] * (1 - wx) + pad[y0 + 2, x0 + 2] * wx) * wy
        assert np.abs(orc.readback_hdr_f32() - want).max() < 2e-5, "bilinear magnification"


def test_oracle_skybox_known_answers():
    """SkyboxRoutine + skybox.wgsl (rule R10): a 90-degree camera at the origin sees exactly one cube face per axis direction, with
    Vulkan's (s, t) orientation — looking down +Z with +Y up the image IS the +Z face, texel for pixel."""
    import skybox_case as sk

    for f in range(6):
        orc = load_oracle_backend()
        sk.build(orc, sk.solid_faces(), f).render_frame(32)
        want = np.append(sk.FACE_COLOURS[f][:3].astype(np.float32) / 255.0, 1.0)
        assert np.array_equal(orc.readback_hdr_f32().reshape(-1, 4), np.broadcast_to(want, (32 * 32, 4))), f
    rng = np.random.default_rng(5)
    faces = [rng.integers(0, 256, (32, 32, 4), dtype=np.uint8) for _ in range(6)]
    orc = load_oracle_backend()
    sk.build(orc, faces, 4, mips="none").render_frame(32)          # +Z, one texel per pixel
    got = orc.readback_hdr_f32()
    assert np.abs(got[..., :3] - faces[4][..., :3].astype(np.float64) / 255.0).max() < 2e-5 and np.all(got[..., 3] == 1.0)
    # an opaque object in front keeps its pixels: the skybox only fills samples still at the clear depth
    orc = load_oracle_backend()
    r = sk.build(orc, sk.solid_faces(), 4)
    mat = r.add_unlit_material((0.25, 0.5, 0.75, 1.0))
    r.cube(mat, glam.from_translation((0.0, 0.0, 5.0)))
    r.render_frame(32)
    h = orc.readback_hdr_f32()
    assert np.allclose(h[16, 16], (0.25, 0.5, 0.75, 1.0)) and np.allclose(h[0, 0], (0.0, 1.0, 1.0, 1.0))


def test_shard_and_tile_partitions_cover_everything():
    """parallel.shard_range / tile_rows: contiguous, disjoint, complete for every world size, including ragged ones."""
    from rend3_b200.parallel import shard_range, tile_rows

    for n in (0, 1, 7, 1000, 1_000_003, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:])) and all(lo <= hi for lo, hi in edges)
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
    for h in (1, 270, 1080, 2160, 2161):
        for world in (1, 2, 4, 8):
            rows = [tile_rows(h, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    # shadow maps split by light: every light has exactly one owner, whatever the ratio of lights to ranks
    from rend3_b200.parallel import ForwardSplit
    for world in (1, 2, 4, 8):
        splits = [ForwardSplit(None, None, None, r, world, (64, 64), 5) for r in range(world)]
        for light in range(11):
            assert sum(1 for sp in splits if sp.owns_shadow(light)) == 1
        assert [sp.assembles_on(sp.rank) for sp in splits].count(True) == 1


def test_oracle_frustum_boundary_and_degenerate_records():
    """Frustum::contains_sphere is inclusive (`distance >= -radius`, util/frustum.rs:148-161); NaN centres fail every comparison and are
    culled; zero-radius spheres on a plane stay; disabled slots neither bake nor list."""
    n = 6
    rec = np.zeros(n, dtype=layouts.OBJECT_DTYPE)
    rec["transform"] = np.tile(glam.identity().reshape(16), (n, 1))
    rec["enabled"] = 1
    hdr = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (64, 64), 1, n)
    plane = hdr["frustum"][0]                                          # (a, b, c, d) of the left plane, unit normal
    on_plane = (-plane[3] * plane[:3]).astype(np.float32)              # a point with distance exactly 0 (up to rounding)
    inward = plane[:3]
    rec["sphere_center"][0] = on_plane + inward * 5.0; rec["sphere_radius"][0] = 1.0      # clearly inside the left plane
    rec["sphere_center"][1] = on_plane - inward * 2.0; rec["sphere_radius"][1] = 1.0      # clearly outside
    rec["sphere_center"][2] = on_plane - inward * 1.0; rec["sphere_radius"][2] = 1.5      # centre outside, sphere reaches in
    rec["sphere_center"][3] = np.nan; rec["sphere_radius"][3] = 1.0
    rec["sphere_center"][4] = on_plane + inward * 5.0; rec["sphere_radius"][4] = 1.0; rec["enabled"][4] = 0
    rec["sphere_center"][5] = on_plane + inward * 5.0; rec["sphere_radius"][5] = 0.0
    orc = load_oracle_backend()
    orc.set_objects(rec)
    orc.object_uniform_upload(CAMERA_VIEWPORT, hdr, CB_BAKE | CB_CULL)
    vis = set(int(v) for v in orc.readback_visible(CAMERA_VIEWPORT))
    # whether 0, 2 and 5 survive the OTHER four planes depends on the camera; membership relative to each other is what is pinned
    assert 1 not in vis and 3 not in vis and 4 not in vis
    assert (0 in vis) == (5 in vis) and ((0 not in vis) or (2 in vis))
    mats = orc.readback_object_matrices(CAMERA_VIEWPORT, 0, n)
    raw = mats.view(np.uint8).reshape(n, 128)
    assert not raw[4].any(), "disabled slots are not baked (uniform_prep.wgsl:18-20)"
    assert raw[0].any()


def test_oracle_skinning_matches_float64_blend():
    """skinning.wgsl:37-94 restated: skinned positions equal the float64 4-joint blend, normals are unit length, and
    identity joints leave the mesh untouched."""
    import skinning_case

    words, inputs, joints, expect = skinning_case.build(seed=1)
    b = load_oracle_backend()
    b.set_mesh_buffer(words)
    b.skin(inputs, joints)
    out = b.readback_mesh_buffer(len(words))
    for rec, exp in zip(inputs, expect):
        nv = int(rec["vertex_count"])
        p = out[rec["updated_position_offset"] // 4:][: nv * 3].view(np.float32).reshape(nv, 3)
        n = out[rec["updated_normal_offset"] // 4:][: nv * 3].view(np.float32).reshape(nv, 3)
        assert np.allclose(p, exp, rtol=1e-5, atol=1e-5)
        assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (len(joints), 1))
    b.set_mesh_buffer(words)
    b.skin(inputs, ident)
    out = b.readback_mesh_buffer(len(words))
    for rec in inputs:
        nv = int(rec["vertex_count"])
        src = words[rec["base_position_offset"] // 4:][: nv * 3].view(np.float32)
        dst = out[rec["updated_position_offset"] // 4:][: nv * 3].view(np.float32)
        assert np.allclose(dst, src, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ multi-GPU sharding logic over gloo (world_size 2, CPU)
SHARD_SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["R3_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import load_oracle_backend
from rend3_b200.backend import CAMERA_VIEWPORT
from rend3_b200.routines import per_camera_header
from rend3_b200.scenes import cloud_camera, object_cloud_records
from rend3_b200.parallel import shard_range, allgather_visible
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 20001
rec = object_cloud_records(n, seed=6)
lo, hi = shard_range(n, rank, world)
b = load_oracle_backend()
b.set_objects(rec[lo:hi])
b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, hi - lo))
local = torch.from_numpy(b.readback_visible(CAMERA_VIEWPORT).astype(np.int64))
merged = allgather_visible(local, lo, n, world).numpy()
full = load_oracle_backend()
full.set_objects(rec)
full.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n))
assert np.array_equal(merged, full.readback_visible(CAMERA_VIEWPORT).astype(np.int64)), "sharded visible list differs from the single-rank list"
dist.destroy_process_group()
print("rank", rank, "ok", len(merged))
'''


def test_two_rank_object_sharding_over_gloo():
    """Object-range shards + all-gather of the visible lists reproduce the single-GPU list bit for bit."""
    with tempfile.TemporaryDirectory() as d:
        script = os.path.join(d, "shard.py")
        open(script, "w").write(SHARD_SCRIPT)
        env = dict(os.environ, R3_ROOT=ROOT, OMP_NUM_THREADS="2")
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", "29541", script], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("ok") == 2


def test_rule_r9_log2_is_exact_at_powers_of_two_and_close_elsewhere():
    """Rule R9's log2 (the mip-level selection of textureSampleGrad): a fixed sequence of IEEE f32 operations.  Exact at powers of two
    (so integer footprints select exact levels), within 4e-7 of the real log2 elsewhere (inside Vulkan's 2^-21 bound on [0.5, 2])."""
    import ctypes

    import oracle
    lib = ctypes.CDLL(oracle.build())
    lib.r3o_log2_r9.restype = ctypes.c_float
    lib.r3o_log2_r9.argtypes = [ctypes.c_float]
    for e in range(-20, 40):
        assert lib.r3o_log2_r9(float(2.0 ** e)) == float(e)
    rng = np.random.default_rng(5)
    xs = np.exp(rng.uniform(np.log(1e-3), np.log(1e6), 20000)).astype(np.float32)
    got = np.array([lib.r3o_log2_r9(float(x)) for x in xs], dtype=np.float64)
    want = np.log2(xs.astype(np.float64))
    assert np.abs(got - want).max() < 4e-7 + 1.2e-7 * np.abs(want).max()
    assert lib.r3o_log2_r9(0.0) == -np.inf and np.isnan(lib.r3o_log2_r9(-1.0)) and lib.r3o_log2_r9(float("inf")) == np.inf


def test_oracle_sort_orders_nan_distances_last_like_ordered_float():
    """ShaderJobSortingKey compares OrderedFloat distances (batching.rs:37): NaN is greater than every number and equal to itself.
    Objects with NaN locations must land at the end of their (material, reason) group, in handle order."""
    from rend3_b200.routines import per_camera_header
    from rend3_b200.scenes import cloud_camera, object_cloud_records

    n = 600
    rec = object_cloud_records(n, seed=21, extent=50.0, disabled_fraction=0.0)
    rec["sphere_radius"][:] = 1.0e6          # everything visible
    loc = rec["sphere_center"].copy()
    bad = np.array([5, 17, 300, 599])
    loc[bad] = np.nan
    b = load_oracle_backend()
    b.set_objects(rec)
    b.set_object_sort_info(np.zeros(n, dtype=np.uint64), np.full(n, 3, dtype=np.uint8), loc)
    b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (640, 480), 1, n))
    b.batch_objects(CAMERA_VIEWPORT, np.zeros(3, dtype=np.float32))
    batches, _ = b.readback_batches(CAMERA_VIEWPORT)
    order = np.concatenate([bt["object_culling_information"]["object_id"][: int(bt["total_objects"])] for bt in batches])
    assert len(order) == n and list(order[-4:]) == list(bad)
    d2 = ((loc[order[:-4]].astype(np.float32)) ** 2).sum(axis=1)
    assert np.all(np.diff(d2) >= 0)


def test_oracle_narrow_texture_formats_read_missing_channels_as_zero_zero_one():
    """R8Unorm / Rg8Unorm in the bindless table (split AO / metallic / roughness maps, bicomponent normal maps): the stored channels
    come back as unorm, the missing ones as (0, 0, 1) — one texel per pixel through the unlit albedo slot."""
    import texture_case as tcase
    from rend3_b200.world import Texture

    data = tcase.checker_texture(32, seed=5)
    for channels in (1, 2):
        b = load_oracle_backend()
        tcase.build(b, Texture(data, channels=channels, mips="none"), "nearest").render_frame(32)
        want = np.zeros((32, 32, 4))
        want[..., :channels] = data[..., :channels].astype(np.float64) / 255.0
        want[..., 3] = 1.0
        got = b.readback_hdr_f32().astype(np.float64)
        assert np.abs(got - want).max() < 2e-6, channels


def _known_blocks():
    """Hand-assembled blocks and the texels the published palettes give (D3D11 functional spec 19.5 / Khronos Data Format 1.3 ch. 18-20)."""
    import struct

    codes = [0, 1, 2, 3] * 4                                              # texel t = 4 py + px uses code t % 4
    two = sum(c << (2 * t) for t, c in enumerate(codes))
    cases = []
    # BC1, c0 = pure red 0xF800 > c1 = pure blue 0x001F: four-colour mode
    cases.append(("bc1", struct.pack("<HHI", 0xF800, 0x001F, two), [(1, 0, 0, 1), (0, 0, 1, 1), (2 / 3, 0, 1 / 3, 1), (1 / 3, 0, 2 / 3, 1)]))
    # BC1, c0 <= c1: three-colour mode, code 3 = transparent black
    cases.append(("bc1", struct.pack("<HHI", 0x001F, 0xF800, two), [(0, 0, 1, 1), (1, 0, 0, 1), (0.5, 0, 0.5, 1), (0, 0, 0, 0)]))
    # BC2 ignores the endpoint order (always four colours); alpha nibble of texel t = t
    alpha4 = sum(t << (4 * t) for t in range(16))
    cases.append(("bc2", struct.pack("<QHHI", alpha4, 0x001F, 0xF800, two), [((0, 0, 1), (1, 0, 0), (1 / 3, 0, 2 / 3), (2 / 3, 0, 1 / 3))[t % 4] + (t / 15,) for t in range(16)]))
    # BC4 unorm, e0 = 255 > e1 = 0: eight values; texel t uses code t % 8
    three = sum((t % 8) << (3 * t) for t in range(16))
    cases.append(("bc4", bytes([255, 0]) + three.to_bytes(6, "little"), [(v, 0, 0, 1) for v in (1, 0, 6 / 7, 5 / 7, 4 / 7, 3 / 7, 2 / 7, 1 / 7)]))
    # BC4 unorm, e0 = 50 <= e1 = 150: six values, then 0 and 1
    cases.append(("bc4", bytes([50, 150]) + three.to_bytes(6, "little"), [(v / 255, 0, 0, 1) for v in (50, 150, 70, 90, 110, 130)] + [(0, 0, 0, 1), (1, 0, 0, 1)]))
    # BC4 snorm: -128 reads as -1 like -127; e0 = 127 > e1 = -128
    cases.append(("bc4s", bytes([127, 0x80]) + three.to_bytes(6, "little"), [(v, 0, 0, 1) for v in (1, -1, 5 / 7, 3 / 7, 1 / 7, -1 / 7, -3 / 7, -5 / 7)]))
    # BC4 snorm six-value mode: e0 = -64 <= e1 = 64, code 6 = -1, code 7 = +1
    cases.append(("bc4s", bytes([0xC0, 0x40]) + three.to_bytes(6, "little"), [(v / 127, 0, 0, 1) for v in (-64, 64, -38.4, -12.8, 12.8, 38.4)] + [(-1, 0, 0, 1), (1, 0, 0, 1)]))
    # BC3 = BC4 alpha block + four-colour block; BC5 = two BC4 blocks
    cases.append(("bc3", bytes([255, 0]) + three.to_bytes(6, "little") + struct.pack("<HHI", 0x001F, 0xF800, two), None))
    cases.append(("bc5", bytes([255, 0]) + three.to_bytes(6, "little") + bytes([50, 150]) + three.to_bytes(6, "little"), None))
    return cases


def test_block_decode_known_answers():
    """Rule R11 on hand-assembled blocks: the numpy decoder of rend3_b200/bc.py and the ORACLE's texel fetch (through a 4 x 4 texture drawn
    one texel per pixel) both give the palettes of the published format descriptions."""
    import texture_case as tcase
    from rend3_b200 import bc
    from rend3_b200.world import Texture

    for name, block, palette in _known_blocks():
        img = bc.decode(name, np.frombuffer(block, dtype=np.uint8), 4, 4)
        if palette is not None:
            want = np.array([palette[t % len(palette)] for t in range(16)], dtype=np.float64).reshape(4, 4, 4)
            assert np.abs(img - want).max() < 1e-12, name
        elif name == "bc3":
            assert np.allclose(img[..., 3].reshape(-1)[:8], [1, 0, 6 / 7, 5 / 7, 4 / 7, 3 / 7, 2 / 7, 1 / 7]) and np.allclose(img[0, 2, :3], [1 / 3, 0, 2 / 3])
        else:
            assert np.allclose(img[..., 0].reshape(-1)[:8], [1, 0, 6 / 7, 5 / 7, 4 / 7, 3 / 7, 2 / 7, 1 / 7]) and np.allclose(img[..., 1].reshape(-1)[6:8], [0, 1])
        b = load_oracle_backend()
        tex = Texture(np.zeros((4, 4, 4), dtype=np.uint8), mips="none", block_format=name, block_levels=[np.frombuffer(block, dtype=np.uint8)])
        tcase.build(b, tex, "nearest").render_frame(4)
        assert np.abs(b.readback_hdr_f32().astype(np.float64) - img).max() < 1e-7, name


def test_oracle_block_compressed_formats_match_the_published_palettes():
    """BC1 - BC5 entries of the bindless table (what rend3-gltf's ktx2 / dds loaders pass to add_texture_2d): random images — punch-through
    BC1 blocks, both BC4 palettes, signed variants, sRGB variants, a size that is not a multiple of the block — encoded by bc.py, drawn one
    texel per pixel by the oracle and compared with the float64 decode of the same blocks."""
    import texture_case as tcase
    from rend3_b200 import bc
    from rend3_b200.world import Texture

    for size in (32, 30):
        data = tcase.checker_texture(size, seed=5)
        for name, (_, srgb_format, _) in bc.BLOCK_FORMATS.items():
            for srgb in ((False, True) if srgb_format is not None else (False,)):
                t = Texture(data, srgb=srgb, mips="none", block_format=name)
                b = load_oracle_backend()
                tcase.build(b, t, "nearest").render_frame(size)
                want = bc.decode(name, t.stored_levels()[0], size, size, srgb)
                assert np.abs(b.readback_hdr_f32().astype(np.float64) - want).max() < 5e-7, (size, name, srgb)
    # the encoder is good enough that the scenes keep their look: the decode stays close to the source image
    img = np.clip(np.rint(np.add.outer(np.linspace(0, 200, 32), np.linspace(0, 55, 32))), 0, 255).astype(np.uint8)
    rgba = np.stack([img, img[::-1], img.T, 255 - img // 2], axis=-1)
    for name in ("bc1", "bc3", "bc5"):
        dec = bc.decode(name, bc.encode(name, rgba), 32, 32)
        ch = {"bc1": 3, "bc3": 4, "bc5": 2}[name]
        assert np.abs(dec[..., :ch] - rgba[..., :ch] / 255.0).max() < 0.1, name


def test_block_compressed_mip_chains_are_validated():
    """r3o_set_textures sizes the chain of a block format in blocks (ceil(w / 4) x ceil(h / 4) x 8 or 16 bytes per level)."""
    from rend3_b200 import layouts

    b = load_oracle_backend()
    d = np.zeros(1, dtype=layouts.TEXTURE_DESC_DTYPE)
    d["width"], d["height"], d["mip_count"], d["format"] = 10, 6, 3, layouts.TEXFMT_BC3_RGBA_UNORM          # 3x2 + 2x1 + 1x1 blocks of 16 bytes
    need = (6 + 2 + 1) * 16
    b.set_textures(d, np.zeros(need, dtype=np.uint8))
    with pytest.raises(Exception):
        b.set_textures(d, np.zeros(need - 16, dtype=np.uint8))
    d["format"] = 31                                                                                         # BC6H and anything unknown: rejected
    with pytest.raises(Exception):
        b.set_textures(d, np.zeros(4096, dtype=np.uint8))


def _pillow_decode(name, blocks, w, h):
    """Pillow's DDS reader as an independent BCn decoder (8-bit texels): the blocks wrapped in a DX10 DDS header."""
    import io
    import struct

    Image = pytest.importorskip("PIL.Image")
    code = {"bc1": 71, "bc2": 74, "bc3": 77, "bc4": 80, "bc5": 83, "bc7": 98}[name]
    pf = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    hdr = struct.pack("<IIIIIII44x", 124, 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000, h, w, len(bytes(blocks)), 0, 1) + pf + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0)
    im = Image.open(io.BytesIO(b"DDS " + hdr + struct.pack("<IIIII", code, 3, 0, 1, 0) + bytes(blocks)))
    im.load()
    a = np.asarray(im)
    return a[..., None] if a.ndim == 2 else a


def test_block_decoders_agree_with_pillow():
    """The decode rules pinned on an implementation that is not ours: Pillow's BCn decoder.  BC7 is specified to the integer, so the 8-bit
    texels of random blocks of every mode are IDENTICAL; BC1 - BC5 are decoded by Pillow in 8-bit integer arithmetic, so rule R11's exact
    ratios agree with it to within one 8-bit step."""
    import texture_case as tcase
    from rend3_b200 import bc

    blocks = tcase.random_bc7_blocks(8 * 360, seed=7)                       # modes 0..7 (the reserved mode: the specification says zeros)
    blocks = blocks[np.arange(len(blocks)) % 9 != 8]
    n = len(blocks)
    got = _pillow_decode("bc7", blocks.reshape(-1), 4 * n, 4)
    mine = np.stack([bc.decode_bc7_block(bytes(b)) for b in blocks]).reshape(n, 4, 4, 4).transpose(1, 0, 2, 3).reshape(4, 4 * n, 4)
    assert np.array_equal(got, mine)
    assert not bc.decode_bc7_block(bytes(16)).any()
    img = tcase.checker_texture(32, seed=11)
    for name in ("bc1", "bc2", "bc3", "bc4", "bc5", "bc7"):
        data = bc.encode(name, img)
        ref = _pillow_decode(name, data, 32, 32).astype(np.float64) / 255.0
        assert np.abs(bc.decode(name, data, 32, 32)[..., : ref.shape[-1]] - ref).max() <= 1.0 / 255.0 + 1e-12, name


def test_bc7_tables_of_the_header_match_the_python_copy_and_the_probe():
    """include/r3_bc7_tables.h (what the CUDA code and the oracle decode with) holds the same constants as rend3_b200/bc.py, and both equal
    what tools/derive_bc7_tables.py reads off Pillow's decoder block by block."""
    import re
    from rend3_b200 import bc

    text = open(os.path.join(ROOT, "include", "r3_bc7_tables.h")).read()

    def table(name):
        body = re.search(name + r"\[[^=]*=\s*\{(.*?)\};", text, re.S).group(1)
        return [int(x.rstrip("u"), 0) for x in re.findall(r"0x[0-9A-Fa-f]+u?|\d+", body)]

    assert table("r3_bc7_partition2") == list(bc.BC7_P2) and table("r3_bc7_partition3") == list(bc.BC7_P3)
    assert table("r3_bc7_anchor2") == list(bc.BC7_A2) and table("r3_bc7_anchor3a") == list(bc.BC7_A3A) and table("r3_bc7_anchor3b") == list(bc.BC7_A3B)
    assert table("r3_bc7_modes") == [x for m in bc.BC7_MODES for x in m]
    pytest.importorskip("PIL.Image")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import derive_bc7_tables

    p2, p3, a2, a3a, a3b = derive_bc7_tables.derive()
    assert (p2, p3, a2, a3a, a3b) == (list(bc.BC7_P2), list(bc.BC7_P3), list(bc.BC7_A2), list(bc.BC7_A3A), list(bc.BC7_A3B))


def test_oracle_bc7_blocks_of_every_mode():
    """Random BC7 blocks (every mode, the reserved one included) in the bindless table, drawn one texel per pixel by the oracle, against
    the Python statement of the format; linear and sRGB variant."""
    import texture_case as tcase
    from rend3_b200 import bc
    from rend3_b200.world import Texture

    blocks = tcase.random_bc7_blocks(64, seed=3).reshape(-1)
    for srgb in (False, True):
        t = Texture(np.zeros((32, 32, 4), dtype=np.uint8), srgb=srgb, mips="none", block_format="bc7", block_levels=[blocks])
        b = load_oracle_backend()
        tcase.build(b, t, "nearest").render_frame(32)
        assert np.abs(b.readback_hdr_f32().astype(np.float64) - bc.decode("bc7", blocks, 32, 32, srgb)).max() < 5e-7, srgb


def test_oracle_ktx2_formats_read_back_as_the_sampler_defines():
    """Formats 17 - 30 of include/r3_layouts.h (snorm8, Bgra8 / sRGB, Rgb10a2, 16 / 32-bit float, unorm16; what the ktx2 loader can pass,
    rend3-gltf/src/lib.rs:1195-1285): drawn one texel per pixel by the oracle against the float64 unpack of the stored bytes — -128 snorm
    reads -1, subnormal halves survive, missing channels read (0, 0, 1)."""
    import texture_case as tcase
    from rend3_b200 import texformats as tf
    from rend3_b200.world import Texture

    data = tcase.checker_texture(32, seed=5)
    data[0, 0], data[0, 1] = 0, 7                                            # -128 for the snorm formats, the subnormal marker of the half formats
    for name, (fmt, bpp) in tf.STORAGE.items():
        t = Texture(data, mips="none", storage=name)
        stored = t.stored_levels()[0]
        assert len(stored) == 32 * 32 * bpp and t.format() == fmt
        b = load_oracle_backend()
        tcase.build(b, t, "nearest").render_frame(32)
        want = tf.unpack(name, stored, 32, 32)
        assert np.abs(b.readback_hdr_f32().astype(np.float64) - want).max() < 5e-7, name
        if name.endswith("8s"):
            assert want[0, 0, 0] == -1.0
    from rend3_b200 import layouts
    d = np.zeros(1, dtype=layouts.TEXTURE_DESC_DTYPE)
    d["width"], d["height"], d["mip_count"], d["format"] = 4, 4, 3, layouts.TEXFMT_RGBA16_FLOAT            # 16 + 4 + 1 texels of 8 bytes
    b = load_oracle_backend()
    b.set_textures(d, np.zeros(21 * 8, dtype=np.uint8))
    with pytest.raises(Exception):
        b.set_textures(d, np.zeros(21 * 8 - 8, dtype=np.uint8))
    d["format"] = 31                                                                                         # integer formats and anything unknown
    with pytest.raises(Exception):
        b.set_textures(d, np.zeros(4096, dtype=np.uint8))
