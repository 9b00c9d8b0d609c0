"""the draw-record buffer and the count the indirect draw reads (BlitzenVulkan/vulkanDraw.cpp:469-471) are
exported as file descriptors and consumed IN PLACE by a second process -- the CUDA stand-in for the renderer importing them with
VkImportMemoryFdInfoKHR (no Vulkan loader in this image).  The consumer's view of the list must equal the oracle's list."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from blitzen_b200 import capi
from conftest import view_at

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_child(e, fence, rec_bytes):
    cmd = [sys.executable, os.path.join(HERE, "interop_child.py"), "0", str(e.draws_fd), str(e.draws_alloc_bytes), str(e.counts_fd),
           str(e.counts_alloc_bytes), str(e.count_offset_bytes), str(rec_bytes), fence.hex()]
    r = subprocess.run(cmd, pass_fds=(e.draws_fd, e.counts_fd), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("INTEROP")][-1].split()
    return int(line[1]), int(line[2]), line[3]


@pytest.mark.parametrize("fmt", [capi.REC_VK24, capi.REC_DX32])
def test_second_process_reads_the_exported_lists_in_place(built, small_scene, fmt):
    sc = small_scene
    tables = sc
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    rec_bytes = 24 if fmt == capi.REC_VK24 else 32
    ctx = capi.CullContext(0)
    try:
        ctx.upload_scene(sc["objs"], sc["transforms"], tables["surfaces"], tables["lods"])
        ctx.set_view(view)
        ctx.frustum_lod(fmt=fmt)                                   # a pass BEFORE the export: the buffers are swapped under a live context
        before, _ = ctx.read_draws()
        e = ctx.export_outputs()
        assert e.draws_fd >= 0 and e.counts_fd >= 0 and e.draws_alloc_bytes >= len(sc["objs"]) * 32 and e.generation >= 1
        fence = ctx.export_fence()
        ctx.frustum_lod(fmt=fmt)
        ctx.signal_fence()
        exp, total, _ = O.cull(sc["objs"], sc["transforms"], tables["surfaces"], tables["lods"], view, O.PASS_FRUSTUM, rec_words=rec_bytes // 4)
        written, tot, sha = run_child(e, fence, rec_bytes)
        assert (written, tot) == (len(exp), total) and total > 0
        assert sha == hashlib.sha256(np.ascontiguousarray(exp).tobytes()).hexdigest()
        got, _ = ctx.read_draws()                                  # the producer's own read-back still works on the exported buffers
        assert np.array_equal(got.view(np.uint32).reshape(-1), np.ascontiguousarray(exp).view(np.uint32).reshape(-1))
        assert np.array_equal(got.view(np.uint32), before.view(np.uint32))
        # a different pass into the same exported buffers; same descriptors, new fence value
        ctx.reset_visibility()
        ctx.early(fmt)
        ctx.signal_fence()
        written, tot, sha = run_child(e, fence, rec_bytes)
        assert (written, tot) == (0, 0)                            # nothing was visible last frame
        # exporting twice hands out fresh descriptors of the same allocations
        e2 = ctx.export_outputs()
        assert e2.generation == e.generation and e2.draws_alloc_bytes == e.draws_alloc_bytes
        for fd in (e.draws_fd, e.counts_fd, e2.draws_fd, e2.counts_fd):
            os.close(fd)
    finally:
        ctx.close()


def test_export_errors(built, small_scene):
    tables = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    ctx = capi.CullContext(0)
    try:
        with pytest.raises(capi.BlzError):
            ctx.export_outputs()                                    # no scene yet
        with pytest.raises(capi.BlzError):
            ctx.signal_fence()                                      # no fence exported
        with pytest.raises(capi.BlzError):
            ctx._check(ctx._lib.blz_cull_signal_semaphore(ctx._h, 1))   # no semaphore imported
        with pytest.raises(capi.BlzError):
            ctx._check(ctx._lib.blz_cull_import_semaphore(ctx._h, -1, 1))
        sc = small_scene
        ctx.upload_scene(sc["objs"], sc["transforms"], tables["surfaces"], tables["lods"])
        ctx.set_view(view)
        ctx.frustum_lod()
        ctx.read_draws()                                            # context still usable
    finally:
        ctx.close()


def test_read_draws_refuses_the_wrong_record_format(built, small_scene):
    tables = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    """ADVICE r01: a DX32 list must not be copied into a buffer sized for VK24 records."""
    sc = small_scene
    ctx = capi.CullContext(0)
    try:
        ctx.upload_scene(sc["objs"], sc["transforms"], tables["surfaces"], tables["lods"])
        ctx.set_view(view)
        ctx.frustum_lod(fmt=capi.REC_DX32)
        with pytest.raises(capi.BlzError):
            ctx.read_draws(capi.REC_VK24)
        got, _ = ctx.read_draws(capi.REC_DX32)
        assert got.dtype.itemsize == 32
        ctx.frustum_lod(fmt=capi.REC_VK24)
        with pytest.raises(capi.BlzError):
            ctx.read_draws(capi.REC_DX32)
    finally:
        ctx.close()
