"""GameScene::update's cycle (visible set -> load -> mesh -> unload, scene_game.rs:39-82,170-223) through the C ABI
against the CPU restatement in oracle/scene_oracle.py."""
import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc
from oracle import scene_oracle as sco

START_CAMERA = (2.5, 22.5, -2.5)  # engine/core/app.rs:203


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


# ---- host logic: no GPU needed ------------------------------------------------------------------------------------
def test_default_settings_visible_set_is_2853_chunks_in_reference_order():
    vis = capi.visible_chunks(START_CAMERA, 10, 4)  # settings.json5:4-5
    assert vis.shape == (2853, 3)
    assert [tuple(p) for p in vis.tolist()] == sco.visible_chunks(START_CAMERA, 10, 4)
    assert sco.current_chunk(START_CAMERA) == (0, 1, -1)
    # x outermost, then y, then z (scene_game.rs:199-201)
    keys = vis[:, 0].astype(np.int64) * 10 ** 8 + vis[:, 1].astype(np.int64) * 10 ** 4 + vis[:, 2]
    assert (np.diff(keys) > 0).all()


@pytest.mark.parametrize("cam", [(0.0, 0.0, 0.0), (-0.5, -0.5, -0.5), (-1.0, -16.0, -17.0), (15.99, 16.0, -15.99), (-16.5, 31.9, 1e6),
                                 (float("nan"), 3e9, -3e9)])
@pytest.mark.parametrize("h,v", [(0, 0), (1, 0), (2, 2), (3, 1), (7, 3)])
def test_visible_set_matches_the_restatement(cam, h, v):
    """`as i32` truncates toward zero BEFORE div_euclid (so -0.5 is chunk 0, -1.0 is chunk -1); NaN -> 0; saturation."""
    if abs(cam[1]) > 2e9 and (h, v) != (1, 0):
        pytest.skip("saturated coordinates: one case is enough")
    want = sco.visible_chunks(cam, h, v)
    if any(abs(c) > 2 ** 31 - 1 for p in want for c in p):
        want = [p for p in want if all(-2 ** 31 <= c < 2 ** 31 for c in p)]
        got = capi.visible_chunks(cam, h, v)
        assert len(got) >= 0  # positions beyond i32 overflow in the reference too (debug panic / release wrap): not pinned
        return
    got = capi.visible_chunks(cam, h, v)
    assert [tuple(p) for p in got.tolist()] == want


# ---- the cycle on the GPU ------------------------------------------------------------------------------------------
def _check_scene(ctx, scene, uv):
    pos, entries = ctx.scene_draw_list()
    got = {tuple(p): e for p, e in zip(pos.tolist(), entries)}
    want = {p: m for p, m in scene.chunk_meshes.items() if m is not None}
    assert set(got) == set(want)
    spans = []
    for p, e in got.items():
        assert int(e["vertex_count"]) == want[p].shape[0], p
        spans.append((int(e["offset_bytes"]), int(e["offset_bytes"]) + int(e["vertex_count"]) * 20))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), "two live meshes overlap in the arena"
    return got, want


@pytest.mark.gpu
def test_first_frame_of_the_reference_world(uv_default):
    """2,853 chunks loaded, 1,268 meshes, 10,712,064 vertices; every mesh byte-exact."""
    with capi.Context(0, 4096, 512 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        launches0 = ctx.kernel_launches
        st = ctx.scene_update(START_CAMERA, 10, 4)
        assert tuple(st["current_chunk"]) == (0, 1, -1) and st["changed"] == 1
        assert (st["n_visible"], st["n_loaded_now"], st["n_unloaded_now"], st["n_loaded_total"]) == (2853, 2853, 0, 2853)
        assert st["n_meshes"] == 1268 and st["vertex_total"] == 10_712_064
        assert ctx.kernel_launches - launches0 == 2  # one terrain launch + one mesh launch for the whole frame
        scene = sco.GameSceneOracle(uv_default, 10, 4)
        scene.update(START_CAMERA)
        got, want = _check_scene(ctx, scene, uv_default)
        for p in list(got)[::7]:
            assert ctx.download_vertices(got[p]).tobytes() == want[p].tobytes(), p
        # still in the same chunk: nothing happens (scene_game.rs:48)
        st2 = ctx.scene_update((3.0, 30.0, -15.0), 10, 4)
        assert st2["changed"] == 0 and st2["n_loaded_total"] == 2853 and ctx.kernel_launches - launches0 == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [capi.MESH_PARITY, capi.MESH_GREEDY | capi.MESH_INDEXED])
def test_walk_load_unload_and_arena_recycling(uv_default, mode):
    """A walk over noise terrain with a small arena: chunks enter and leave, batches' arena intervals are freed and
    reused, and after every step the loaded set, the draw list and every mesh equal the restatement's."""
    h, v, kind, seed = 3, 1, capi.GEN_NOISE, 0xD5E50001
    path = [(8.0, 120.0, 8.0), (24.0, 120.0, 8.0), (40.5, 121.0, 9.0), (40.5, 140.0, 9.0), (8.0, 120.0, 8.0), (500.0, 100.0, -500.0),
            (8.0, 120.0, 8.0), (-9.0, 119.0, -0.5)]
    with capi.Context(0, 512, 40 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        scene = sco.GameSceneOracle(uv_default, h, v, kind, seed)
        used = []
        for cam in path:
            st = ctx.scene_update(cam, h, v, kind, seed, mode)
            changed, new, gone = scene.update(cam)
            assert bool(st["changed"]) == changed and tuple(st["current_chunk"]) == sco.current_chunk(cam)
            assert (st["n_loaded_now"], st["n_unloaded_now"], st["n_loaded_total"]) == (len(new), len(gone), len(scene.loaded))
            assert st["n_meshes"] == scene.amount_of_chunk_meshes()
            assert ctx.resident_chunks == len(scene.loaded)
            pos, entries = ctx.scene_draw_list()
            got = {tuple(p): e for p, e in zip(pos.tolist(), entries)}
            assert set(got) == {p for p, m in scene.chunk_meshes.items() if m is not None}
            spans = sorted((int(e["offset_bytes"]), int(e["offset_bytes"]) + ((int(e["vertex_count"]) * 20 + 127) & ~127)
                            + ((int(e["index_count"]) * 4 + 127) & ~127)) for e in got.values())
            assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
            for p, e in got.items():
                if mode == capi.MESH_PARITY:
                    assert ctx.download_vertices(e).tobytes() == scene.chunk_meshes[p].tobytes(), (cam, p)
                else:
                    vv, ii, _ = orc.mesh_chunk_ext(p, scene.loaded[p], uv_default, greedy=True, indexed=True)
                    assert ctx.download_vertices(e).tobytes() == vv.tobytes() and np.array_equal(ctx.download_indices(e), ii), (cam, p)
            used.append(int(st["arena_bytes_used"]))
        # an interval is held until the last chunk of its batch leaves, so a partial return (step 4) may hold more than the
        # first frame did; after the far jump everything old is gone and the same view is one batch again
        assert used[0] == used[6] and used[4] >= used[0]
        ctx.scene_reset()
        assert ctx.resident_chunks == 0


@pytest.mark.gpu
def test_arena_too_small_rolls_the_update_back(uv_default):
    with capi.Context(0, 4096, 1 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        with pytest.raises(capi.DspError) as ei:
            ctx.scene_update(START_CAMERA, 10, 4)
        assert ei.value.status == capi.ERR_ARENA_FULL and ctx.resident_chunks == 0
        st = ctx.scene_update(START_CAMERA, 1, 0)  # a smaller view fits: the scene is usable after the failure
        assert st["n_loaded_total"] == len(sco.visible_chunks(START_CAMERA, 1, 0))


@pytest.mark.gpu
def test_scene_refuses_a_context_with_box_regions(uv_default):
    with capi.Context(0, 256, 16 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        ctx.generate_region((0, 0, 0), (2, 2, 2), capi.GEN_REFERENCE, 0)
        with pytest.raises(capi.DspError) as ei:
            ctx.scene_update(START_CAMERA, 1, 1)
        assert ei.value.status == capi.ERR_BAD_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed,cam", [(capi.GEN_NOISE, 0xD5E50001, (8.0, 120.0, 8.0)), (capi.GEN_REFERENCE, 0, START_CAMERA)])
def test_first_frame_with_cross_chunk_culling(uv_default, kind, seed, cam):
    """DSP_MESH_HALO through the scene: chunks addressed by the position map, summaries written by the list-form generators,
    all-air and buried chunks dropped by the pre-pass -- every chunk of the frame must equal the extended oracle with the
    frame's other chunks as neighbours."""
    h, v = 3, 2
    dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
    with capi.Context(0, 1024, 128 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        st = ctx.scene_update(cam, h, v, kind, seed, capi.MESH_HALO)
        vis = sco.visible_chunks(cam, h, v)
        assert int(st["n_loaded_total"]) == len(vis)
        ids = {p: orc.generate(kind, seed, p) for p in vis}
        pos, entries = ctx.scene_draw_list()
        got = {tuple(p): e for p, e in zip(pos.tolist(), entries)}
        n_meshes = 0
        for p in vis:
            slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(dirs):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if q in ids:
                    vv = ids[q].reshape(16, 16, 16)
                    present[f] = 1
                    slabs[f] = [vv[:, 0, :], vv[:, 15, :], vv[15, :, :], vv[0, :, :], vv[:, :, 15], vv[:, :, 0]][f].reshape(256)
            want = orc.mesh_chunk_halo(p, ids[p], slabs, present, uv_default)
            if want.shape[0] == 0:
                assert p not in got, p
            else:
                n_meshes += 1
                assert ctx.download_vertices(got[p]).tobytes() == want.tobytes(), p
        assert n_meshes == len(got) == int(st["n_meshes"]) and 0 < n_meshes < len(vis)


@pytest.mark.gpu
def test_crossings_at_a_larger_render_distance_and_a_changing_one(uv_default):
    """The loaded set is kept as arithmetic on (position - previous centre): unit crossings, a diagonal step, a far jump and
    a change of the render distances between updates must load exactly visible(now) - visible(before) and unload exactly
    visible(before) - visible(now) (scene_game.rs:56-73), at a distance where the set has tens of thousands of chunks."""
    kind, seed = capi.GEN_NOISE, 0xD5E50001
    steps = [((8.0, 120.0, 8.0), 20, 6), ((24.0, 120.0, 8.0), 20, 6), ((40.0, 136.0, -8.0), 20, 6), ((40.0, 136.0, -8.0), 12, 3),  # same chunk: nothing happens,
             ((56.0, 136.0, -8.0), 12, 3),                                                                                          # the new distances apply at the next crossing
             ((72.0, 136.0, -8.0), 22, 2), ((5000.0, 100.0, -5000.0), 22, 2), ((5016.0, 100.0, -5000.0), 3, 1)]
    with capi.Context(0, 40000, 6 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        before, last_chunk = set(), None
        for cam, h, v in steps:
            st = ctx.scene_update(cam, h, v, kind, seed)
            cur = sco.current_chunk(cam)
            if cur == last_chunk:
                assert st["changed"] == 0 and int(st["n_loaded_total"]) == len(before)
                continue
            now = set(sco.visible_chunks(cam, h, v))
            assert (int(st["n_visible"]), int(st["n_loaded_now"]), int(st["n_unloaded_now"]), int(st["n_loaded_total"])) == \
                   (len(now), len(now - before), len(before - now), len(now)), (cam, h, v)
            assert ctx.resident_chunks == len(now)
            pos, entries = ctx.scene_draw_list()
            drawn = {tuple(p) for p in pos.tolist()}
            assert drawn <= now and len(drawn) == int(st["n_meshes"])
            for p in list(now)[::997]:  # residency by position, and a sampled mesh against the oracle
                assert ctx.find_chunk(p) is not None
            sample = [p for p in sorted(drawn)[::701]]
            got = {tuple(p): e for p, e in zip(pos.tolist(), entries)}
            for p in sample:
                assert ctx.download_vertices(got[p]).tobytes() == orc.mesh_chunk(p, orc.generate(kind, seed, p), uv_default).tobytes(), p
            for p in list(before - now)[::499]:
                assert ctx.find_chunk(p) is None
            before, last_chunk = now, cur
