"""bench_configs.py — the BASELINE.json configs bench.py's default line does not cover (`bench.py --config c1|c3|c4|c5`;
c2 = configs[1] is bench.py's own path).  Each prints ONE JSON line in bench.py's format.

  c1  configs[0]: one 1024x1024 image — the CPU path's latency (oracle port, 1 thread and T threads) next to the GPU's
      single-image call.
  c3  configs[2]: 32 fusion triplets (12 MP each, mixed aspect) -> 2048x2048 canvases; under N GPUs whole triplets are
      dealt round-robin (sharding.shard_groups).
  c4  configs[3]: ONE batch of 256 4K images split over N GPUs by sharding.lpt_assign (strong scaling); the per-image
      results are gathered on the host and rank 0 checks them against its own single-GPU run of the whole batch.
  c5  configs[4]: the 512-image 0.5-24 MP queue, pulled dynamically: ranks fetch chunks (largest images first) from a
      shared counter in the rendezvous store (host side; no collective on the data path), so a rank that drew small
      images simply comes back sooner.

Timing as in bench.py: W >= 3 warm-up steps, barrier + synchronize on both sides, CUDA events on the launch stream, max
over ranks; inputs are device-resident and far larger than the 126 MB L2 (no flush needed)."""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
import time

import numpy as np

CONFIG_NAMES = {
    "c1": "configs[0]: degradation classifier on one synthetic 1024x1024 sRGB image (CPU path timed beside the GPU call)",
    "c3": "configs[2]: 3-image fusion preprocessing, batch of 32 triplets (4000x3000, 3000x4000, 3840x2160) -> 2048^2 canvases",
    "c4": "configs[3]: one batch of 256 4K (3840x2160) images, classify+preprocess, sharded across the GPUs (LPT)",
    "c5": "configs[4]: mixed-resolution queue, 512 images of 0.5-24 MP, classify+preprocess, chunks pulled dynamically",
}


def _dist_setup(local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # whatever NCCL logs goes to stderr: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return torch, dist


def _timed(torch, dist, world, stream, steps, warmup, fn, sampler):
    """fn() once per step; returns ms per step (max over ranks)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        fn()
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms / steps


def _line(a, world, value, ms_step, scaling, workload, extra_config, launches, clocks, **more):
    from bench import METRIC, UNIT

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": dict({"workload": workload, "l2": "inputs far larger than the 126 MB L2 (no flush needed)"}, **extra_config),
            "gpu_launches": launches, "clocks": clocks}
    line.update(more)
    return line


def _scores_equal(a, b) -> bool:
    keys = ("sum", "sumsq", "e_sum", "e_sumsq", "b_sum", "b_sumsq", "scratch_v", "scratch_h", "luma_hist")
    return all(x[k] == y[k] for x, y in zip(a, b) for k in keys) and all(x["scores"] == y["scores"] for x, y in zip(a, b))


# ------------------------------------------------------------------------------------------------------------------
def run_c1(a, rank, local_rank, world):
    """configs[0]: the reference's own CPU-runnable case."""
    if rank != 0:
        return 0
    import torch

    import irp_b200
    from bench import ClockSampler, host_threads
    from irp_b200.synth import synth_image
    from oracle import oracle

    W = H = 1024
    img = synth_image(W, H, 1)
    oracle.build()
    T = host_threads()
    reps = max(3, min(a.steps, 10))
    t1 = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.analyze_batch([img], threads=1, with_preprocess=True)
        t1.append(time.perf_counter() - t0)
    tT = []
    for _ in range(max(2, reps // 2)):
        t0 = time.perf_counter()
        oracle.analyze_batch([img] * T, threads=T, with_preprocess=True)
        tT.append(time.perf_counter() - t0)
    cpu = {"value": W * H / 1e6 / min(t1), "unit": "MPix/s", "cores": 1, "kind": "port", "latency_ms_1_thread": 1e3 * min(t1),
           "throughput_mpix_s_T_threads": T * W * H / 1e6 / min(tT), "threads_T": T,
           "sample": f"one 1024x1024 image, best of {reps} (1 thread); {T} copies on {T} threads for the throughput figure",
           "note": "oracle port of the reference arithmetic; sharp itself cannot run offline (SURVEY.md section 8c)"}
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    with irp_b200.Engine(local_rank) as eng:
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        d = eng.upload(img)
        launches = [0]

        def dev():
            eng.analyze_batch([d], raw=True)
            launches[0] += eng.timing()["kernel_launches"]

        ms_dev = _timed(torch, None, 1, stream, a.steps, a.warmup, dev, sampler)
        n_l = launches[0] // (a.steps + max(a.warmup, 3)) * a.steps
        pin = eng.pinned_empty(img.shape)
        pin[...] = img
        th = []
        for _ in range(max(5, a.steps)):
            t0 = time.perf_counter()
            eng.analyze_batch([pin])
            th.append(time.perf_counter() - t0)
        res = eng.classify_batch([d])[0]
    ref = oracle.classify(img)
    line = _line(a, 1, W * H / 1e6 / (ms_dev / 1e3), ms_dev, "weak", CONFIG_NAMES["c1"], {"width": W, "height": H, "batch": 1}, n_l, sampler.stop(),
                 cpu_baseline=cpu, e2e={"value": W * H / 1e6 / min(th), "unit": "MPix/s", "ms_per_step": 1e3 * min(th), "h2d_bytes_per_step": W * H * 3,
                                        "d2h_bytes_per_step": W * H * 3 + C.sizeof(irp_b200._ffi.Result), "note": "one pinned host image per call, wall clock"},
                 parity={"integer_fields_equal_oracle": all(res[k] == ref[k] for k in ("e_sum", "e_sumsq", "b_sum", "b_sumsq", "scratch_v", "scratch_h", "luma_hist"))})
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------
def run_c3(a, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world)
    import irp_b200
    from bench import ClockSampler
    from irp_b200.sharding import shard_groups
    from irp_b200.synth import synth_image

    shapes = [(4000, 3000), (3000, 4000), (3840, 2160)]
    n_groups = 32
    mine = shard_groups(n_groups, world)[rank]
    base = [synth_image(w, h, i) for i, (w, h) in enumerate(shapes)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    with irp_b200.Engine(local_rank) as eng:
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        groups = [[eng.upload(np.roll(base[k], 13 * g, axis=0)) for k in range(3)] for g in mine]
        canv = [eng.alloc_device(2048, 2048, 3) for _ in range(3 * len(mine))]
        launches = [0]

        def step():
            if groups:
                eng.fusion_prepare_batch(groups, device_outputs=canv)
                launches[0] += eng.timing()["kernel_launches"]

        ms = _timed(torch, dist, world, stream, a.steps, a.warmup, step, sampler)
        ok = True
        if rank == 0 and groups:   # one canvas of every aspect against the oracle
            from oracle import oracle
            for k in range(3):
                ok &= bool(np.array_equal(eng.download(canv[k]), oracle.fusion_canvas(np.roll(base[k], 13 * mine[0], axis=0), 1)))
    px = n_groups * sum(w * h for w, h in shapes)
    if rank == 0:
        n_l = launches[0] // (a.steps + max(a.warmup, 3)) * a.steps
        line = _line(a, world, px / 1e6 / (ms / 1e3), ms, "strong", CONFIG_NAMES["c3"],
                     {"triplets": n_groups, "triplets_per_s": n_groups / (ms / 1e3), "parallelism": f"{world} GPUs, whole triplets round-robin, no collective",
                      "algorithmic_bytes": px * 3 + n_groups * 3 * 2048 * 2048 * 3}, n_l, sampler.stop(), parity={"first_triplet_equals_oracle": ok})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------------------------
def run_c4(a, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world)
    import irp_b200
    from bench import ClockSampler
    from irp_b200.sharding import gather_results, lpt_assign
    from irp_b200.synth import synth_batch

    W, H, B = 3840, 2160, 256
    imgs = synth_batch(W, H, B, distinct=4)
    shards = lpt_assign([W * H] * B, world)
    mine = shards[rank]
    sampler = ClockSampler(local_rank)
    sampler.start()
    with irp_b200.Engine(local_rank) as eng:
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        ow, oh = eng.preprocess_dims(W, H)
        d_in = [eng.upload(imgs[i]) for i in mine]
        d_out = [eng.alloc_device(ow, oh, 3) for _ in mine]
        launches = [0]

        def step():
            eng.analyze_batch(d_in, device_outputs=d_out, raw=True)
            launches[0] += eng.timing()["kernel_launches"]

        ms = _timed(torch, dist, world, stream, a.steps, a.warmup, step, sampler)
        t = eng.timing()
        local = eng.analyze_batch(d_in, device_outputs=d_out)[0]   # once more, untimed, as dictionaries for the gather
        gathered = gather_results(local, mine, B)
        check = None
        if rank == 0 and world > 1:   # the whole batch on this GPU alone: the sharded results must be the same numbers
            for d in d_in + d_out:
                eng.free(d)
            full_in = [eng.upload(im) for im in imgs]
            full_out = [eng.alloc_device(ow, oh, 3) for _ in imgs]
            single = eng.analyze_batch(full_in, device_outputs=full_out)[0]
            check = _scores_equal(gathered, single)
    if rank == 0:
        n_l = launches[0] // (a.steps + max(a.warmup, 3)) * a.steps
        line = _line(a, world, B * W * H / 1e6 / (ms / 1e3), ms, "strong", CONFIG_NAMES["c4"],
                     {"batch_total": B, "images_per_gpu": [len(s) for s in shards], "width": W, "height": H, "out_width": ow, "out_height": oh,
                      "parallelism": f"{world} GPUs, sharding.lpt_assign over pixel counts, results gathered on the host, no collective",
                      "algorithmic_bytes": B * (W * H * 3 + ow * oh * 3)}, n_l, sampler.stop(),
                     kernels={"classify_ms": t["classify_ms"], "preprocess_ms": t["preprocess_ms"]},
                     parity={"gathered_results": len([g for g in gathered if g is not None]), "equal_to_single_gpu_run": check})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------------------------
def run_c5(a, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world)
    import irp_b200
    from bench import ClockSampler
    from irp_b200.sharding import PullQueue, cleanup_queue_files
    from irp_b200.synth import mixed_resolution_sizes, synth_image

    sizes = mixed_resolution_sizes(512)
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i][0] * sizes[i][1], i))   # largest first: the tail of the queue is small change
    chunk = 16
    chunks = [order[k:k + chunk] for k in range(0, len(order), chunk)]
    big = synth_image(6600, 6100, 11)   # covers 24 MP at every aspect of the set; every image is a window of it
    store = PullQueue.default_store()
    sampler = ClockSampler(local_rank)
    sampler.start()
    with irp_b200.Engine(local_rank) as eng:
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        d_in, d_out, px = [], [], 0
        for i, (w, h) in enumerate(sizes):
            x0, y0 = (37 * i) % max(1, 6600 - w), (91 * i) % max(1, 6100 - h)
            d_in.append(eng.upload(np.ascontiguousarray(big[y0:y0 + h, x0:x0 + w])))
            o_w, o_h = eng.preprocess_dims(w, h)
            d_out.append(eng.alloc_device(o_w, o_h, 3))
            px += w * h
        launches, step_no, done = [0], [0], []

        def step():
            q = PullQueue(len(chunks), f"irp_c5_q{step_no[0]}", store)   # a fresh counter per pass over the queue
            step_no[0] += 1
            mine = []
            nxt = q.pull()
            while nxt is not None:
                cur = chunks[nxt]
                got = [None]
                th = threading.Thread(target=lambda: got.__setitem__(0, q.pull()))   # the next index while the GPU works
                th.start()
                eng.analyze_batch([d_in[i] for i in cur], device_outputs=[d_out[i] for i in cur], raw=True)
                launches[0] += eng.timing()["kernel_launches"]
                mine.append(nxt)
                th.join()
                nxt = got[0]
            q.close(unlink=False)
            done.append(mine)

        # geometry plans (tap tables, the tensor-core kernel's operand matrices) are cached per context and cost ~1 ms to
        # build; with a dynamic queue a rank meets different images every step, so every rank walks the whole queue once
        # before the clock starts — what a long-running worker's cache looks like
        for cur in chunks:
            eng.analyze_batch([d_in[i] for i in cur], device_outputs=[d_out[i] for i in cur], raw=True)
        ms = _timed(torch, dist, world, stream, a.steps, a.warmup, step, sampler)
        mine_last = done[-1]
        counts = [len(mine_last)]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine_last)
            counts = [len(p) for p in parts]
            covered = sorted(c for p in parts for c in p) == list(range(len(chunks)))
        else:
            covered = sorted(mine_last) == list(range(len(chunks)))
        ok = None
        if rank == 0:   # a few images of the last step against the oracle
            from oracle import oracle
            ok = True
            for ci in mine_last[-2:]:
                i = chunks[ci][-1]
                w, h = sizes[i]
                x0, y0 = (37 * i) % max(1, 6600 - w), (91 * i) % max(1, 6100 - h)
                ok &= bool(np.array_equal(eng.download(d_out[i]), oracle.preprocess(np.ascontiguousarray(big[y0:y0 + h, x0:x0 + w]), 1)))
    if world > 1:
        dist.barrier()
        if rank == 0:
            cleanup_queue_files()
    if rank == 0:
        n_l = launches[0] // (a.steps + max(a.warmup, 3)) * a.steps
        line = _line(a, world, px / 1e6 / (ms / 1e3), ms, "strong", CONFIG_NAMES["c5"],
                     {"images": len(sizes), "total_mpix": px / 1e6, "chunk_images": chunk, "chunks": len(chunks), "chunks_per_gpu_last_step": counts,
                      "parallelism": f"{world} GPUs pulling chunks (largest images first) from one shared counter (a flock-guarded file in /dev/shm); no collective on the data path"},
                     n_l, sampler.stop(), parity={"every_chunk_processed_exactly_once": covered, "sampled_outputs_equal_oracle": ok})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


RUNNERS = {"c1": run_c1, "c3": run_c3, "c4": run_c4, "c5": run_c5}
