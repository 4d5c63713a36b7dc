"""Host-side logic that needs no GPU: the reference-interface mirrors' error paths, the batch sharder,
a world_size-2 gloo run of the multi-GPU plumbing, and the synthetic-input generator."""
import asyncio
import io
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_degradation_types_and_factory():
    from irp_b200 import ClassifierService, createClassifierService, SCORE_KEYS

    t = ClassifierService.getDegradationTypes()
    assert list(t.keys()) == list(SCORE_KEYS)  # classifier.js:17-25 order == analyze() key order :62-70
    t["blur"] = "mutated"
    assert ClassifierService.getDegradationTypes()["blur"] != "mutated"  # returns a copy (:342-344)
    assert isinstance(createClassifierService({"logger": None}), ClassifierService)


def test_analyze_rejects_garbage_buffers_like_the_health_probe():
    """restorator.js:294 probes the classifier with Buffer.alloc(100) and expects a rejection."""
    from irp_b200 import ClassifierService

    logged = []

    class Logger:
        def error(self, msg, meta=None):
            logged.append((msg, meta))

        def debug(self, *a):
            pass

        def warn(self, *a):
            pass

    svc = ClassifierService(logger=Logger())
    with pytest.raises(Exception):
        asyncio.run(svc.analyze(bytes(100)))
    assert logged and logged[0][0] == "[classifier] Analysis failed"  # classifier.js:94


def test_preprocess_missing_file_is_problem_400():
    from irp_b200.preprocess import make_request, preprocess_image

    calls = []
    preprocess_image(make_request(None), None, lambda *a: calls.append(a))
    (problem,) = calls[0]
    assert problem.status == 400 and problem.title == "Image File Required"  # imagePreprocess.js:25-34


def test_preprocess_bad_image_is_problem_422():
    from irp_b200.preprocess import make_request, preprocess_image

    calls = []
    preprocess_image(make_request(b"not an image"), None, lambda *a: calls.append(a))
    (problem,) = calls[0]
    assert problem.status == 422 and problem.title == "Image Preprocessing Failed"  # imagePreprocess.js:81-90


def test_resize_dimension_helpers_follow_the_js():
    from irp_b200.preprocess import calculate_resize_dimensions, needs_resize

    assert not needs_resize(2048, 2048) and needs_resize(2049, 10) and not needs_resize(0, 5000)
    assert calculate_resize_dimensions(4000, 3000) == {"width": 2048, "height": 1536}
    assert calculate_resize_dimensions(3840, 2160) == {"width": 2048, "height": 1152}
    assert calculate_resize_dimensions(6000, 4000) == {"width": 2048, "height": 1365}
    assert calculate_resize_dimensions(800, 600) == {"width": 800, "height": 600}
    assert calculate_resize_dimensions(0, 600) == {}


def test_decode_image_reports_format_channels_orientation():
    from PIL import Image

    from irp_b200.classifier import decode_image

    a = np.random.default_rng(0).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    im = Image.fromarray(a)
    exif = im.getexif()
    exif[0x0112] = 6
    buf = io.BytesIO()
    im.save(buf, format="JPEG", quality=90, exif=exif.tobytes())
    px, fmt, o = decode_image(buf.getvalue())
    assert px.shape == (20, 30, 3) and fmt == "jpeg" and o == 6  # stored dims, no rotation applied
    buf = io.BytesIO()
    Image.fromarray(np.dstack([a, a[:, :, :1]])).save(buf, format="PNG")
    px, fmt, o = decode_image(buf.getvalue())
    assert px.shape == (20, 30, 4) and fmt == "png" and o == 1


def test_lpt_sharding_is_a_balanced_partition():
    from irp_b200.sharding import lpt_assign, shard_groups
    from irp_b200.synth import mixed_resolution_sizes

    sizes = mixed_resolution_sizes(512)
    costs = [w * h for w, h in sizes]
    assert 0.45e6 < min(costs) and max(costs) < 25e6
    for n in (1, 2, 4, 8):
        shards = lpt_assign(costs, n)
        assert sorted(i for s in shards for i in s) == list(range(512))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)  # LPT bound
    assert lpt_assign(costs, 4) == lpt_assign(costs, 4)  # deterministic: every rank derives the same split
    assert shard_groups(7, 2) == [[0, 2, 4, 6], [1, 3, 5]]


def test_pull_queue_without_a_process_group_is_a_local_counter():
    from irp_b200.sharding import PullQueue

    q = PullQueue(5, "k", PullQueue.default_store())
    assert [q.pull() for _ in range(7)] == [0, 1, 2, 3, 4, None, None]


def test_synthetic_generator_is_seeded_and_nontrivial():
    from irp_b200.synth import synth_batch, synth_image

    a, b = synth_image(320, 240, 3), synth_image(320, 240, 3)
    assert a.shape == (240, 320, 3) and a.dtype == np.uint8 and np.array_equal(a, b)
    assert not np.array_equal(a, synth_image(320, 240, 4))
    batch = synth_batch(64, 48, 5, distinct=2)
    assert len(batch) == 5 and not np.array_equal(batch[0], batch[2])


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import irp_b200  # noqa: F401  (loads the package; no GPU work in this test)
import time
from irp_b200.sharding import lpt_assign, gather_results, shard_groups, PullQueue, cleanup_queue_files
from irp_b200.synth import mixed_resolution_sizes
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sizes = mixed_resolution_sizes(40)
shards = lpt_assign([w * h for w, h in sizes], world)
mine = shards[rank]
local = [{"idx": i, "rank": rank, "pixels": sizes[i][0] * sizes[i][1]} for i in mine]   # stand-in for per-image results
allres = gather_results(local, mine, len(sizes))
assert [r["idx"] for r in allres] == list(range(len(sizes)))
assert {r["rank"] for r in allres} == set(range(world))
assert all(r["pixels"] == sizes[i][0] * sizes[i][1] for i, r in enumerate(allres))
# fusion triplets stay whole (configs[2]) ...
groups = shard_groups(7, world)
assert sorted(g for s in groups for g in s) == list(range(7))
# ... and the mixed-resolution queue (configs[4]) is pulled dynamically from ONE shared counter: every chunk goes to
# exactly one rank, a slow rank draws fewer
for step in range(2):
    q = PullQueue(37, f"test_queue_{step}", PullQueue.default_store())
    assert q.store is not None and q._fd is not None   # one box: the counter is a flock-guarded file in /dev/shm
    got = []
    while True:
        i = q.pull()
        if i is None:
            break
        got.append(i)
        if rank == 0:
            time.sleep(0.01)   # the slow rank
    parts = [None] * world
    dist.all_gather_object(parts, got)
    assert sorted(i for p in parts for i in p) == list(range(37)), parts
    assert len(parts[1]) > len(parts[0]), [len(p) for p in parts]
    q.close()
# the rendezvous store's atomic add is the fallback where ranks share no /dev/shm
q = PullQueue(11, "test_queue_store", PullQueue.default_store(), shm_dir="")
assert q._fd is None
got = []
while (i := q.pull()) is not None:
    got.append(i)
parts = [None] * world
dist.all_gather_object(parts, got)
assert sorted(i for p in parts for i in p) == list(range(11))
dist.barrier()
if rank == 0:
    cleanup_queue_files()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_world_size_2_gloo_shard_and_gather(tmp_path):
    """The N > 1 path on CPU: two processes derive the same LPT partition, each 'processes' its shard,
    results are gathered on the host in batch order — no data-path collective."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script), ROOT], capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2


def test_top_issues_follow_the_prompt_enhancer():
    """irp_result.issues: PromptEnhancerService._identifyTopIssues /
    _determineSeverity, promptEnhancer.js:121-145 — above 0.3, highest first, stable for ties, at most 3."""
    import ctypes as C

    from irp_b200 import _ffi
    from irp_b200.engine import SCORE_KEYS, issues_of

    lib = _ffi.load()

    def js(scores):   # the reference, literally
        issues = [{"type": t, "confidence": c, "severity": "high" if c >= 0.7 else ("medium" if c >= 0.5 else "low")}
                  for t, c in zip(SCORE_KEYS, scores) if c > 0.3]
        return sorted(issues, key=lambda i: -i["confidence"])[:3]   # sorted() is stable, like Array.prototype.sort

    rng = np.random.default_rng(7)
    cases = [[0.0] * 7, [0.3] * 7, [0.31] * 7, [0.7, 0.5, 0.49999, 0.3000001, 1.0, 0.69999, 0.2], [0.9, 0.1, 0.9, 0.1, 0.9, 0.9, 0.1]]
    cases += [list(np.round(rng.random(7), 2)) for _ in range(200)]
    for sc in cases:
        r = _ffi.Result()
        for k, v in enumerate(sc):
            r.score[k] = v
        assert lib.irp_top_issues(C.byref(r)) == 0
        assert issues_of(r) == js(sc), sc
