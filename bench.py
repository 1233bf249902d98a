#!/usr/bin/env python
"""Benchmark of the match-scoring hot path (BASELINE.json configs[1], weak-scaled over N GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU path (port)

A step = one single-query pass of the hot path over every rank's resident shard of 1M synthetic
clips (2 streams x 1024-d fp32 = 8192 B/clip, 8.19 GB per GPU > 126 MB L2, so no flush is needed):
fused scan K1 (dots, fusion, score) + selection K2 (match / near-miss / tie lists in database
order, exact top-100) and, for N > 1, the peer-memory exchange kernel that merges the ranks' counts and top-k.
`e2e` is the public call with host buffers: FeatureStore.scan (+ lists) — for N > 1 from ONE process that drives all
N GPUs (the broker's arrangement, what compute_matches calls); the one-process-per-GPU arrangement is `e2e_ranks`.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")          # reference Dockerfile:15
os.environ.setdefault("RANDOM_SEED", "73459912436")

STREAMS = ("rgb", "warped_optical_flow")
DIM = 1024
ROW_BYTES = len(STREAMS) * DIM * 4                       # algorithmic bytes per clip per scan
WEIGHTS = (1.0, 1.5)                                     # broker.py:36-41
THRESHOLD, NEAR_MISS, EPS, TOPK = 0.8, 0.35, 3e-6, 100
DATA_SEED = 20261018
REF_ROW = 18120                                          # VQSYN-1 row with alpha ~ 0.9: ~9 % of clips match
METRIC = "clips scored/sec per query"
KERNELS_PER_STEP = 4                                     # scan_rows, select_count, select_finish, select_compact


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config4", "config5"],
                    help="config2 (default, the metric's configuration): single-query scan of 1M clips per GPU; config4: "
                         "256-query batch against 10M clips per GPU on the tensor cores; config5: 1000 bootstrap replicates "
                         "of the weight update over 5000 labelled clips of a 1M-clip DB")
    ap.add_argument("--clips-per-gpu", type=int, default=None, help="default 1M (config2, config5) / 10M (config4)")
    ap.add_argument("--queries", type=int, default=256, help="config4: queries per batch")
    ap.add_argument("--labelled", type=int, default=5000, help="config5: labelled clips")
    ap.add_argument("--replicates", type=int, default=1000, help="config5: bootstrap replicates per step")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold end-to-end leg (8 GB upload per step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config3", action="store_true", help="skip the 12.5M-clips-per-GPU leg (config3_shard)")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the short config4 / config5 legs (other_workloads)")
    ap.add_argument("--exchange", default="p2p-lagged", choices=["p2p-lagged", "p2p", "nccl"],
                    help="N > 1: peer-memory push + fused merge kernel, merging the previous step's payloads while this "
                         "step's are in flight (default; the last step is flushed inside the timed region), the same "
                         "merging in-step, or NCCL allgather + merge kernel")
    a = ap.parse_args()
    if a.clips_per_gpu is None:
        a.clips_per_gpu = 10_000_000 if a.workload == "config4" else 1_000_000
    return a


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag = index, [], 0, False
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                            # pragma: no cover
            self.err = repr(e)

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b], "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU arms
def cpu_port_throughput(n_sample, steps=1, warmup=0):
    """The reference as written, restated (oracle/loop_port.py): dicts of Python lists, one
    np.dot(list, list) per (clip, stream, split), single thread (the reference is GIL-bound).
    Payload construction is outside the timed region, like the HTTP fetch it stands for."""
    from oracle import loop_port, scoring, synth
    X = synth.rows(DATA_SEED, np.arange(n_sample)).astype(np.float64)[:, :, None, :]
    ref = synth.rows(DATA_SEED, [REF_ROW]).astype(np.float64)[0][:, None, :]
    T = scoring.scale_target(ref)
    cand = loop_port.make_candidates(X, np.arange(n_sample), STREAMS, [1])
    tf = loop_port.make_target(T, STREAMS, [1])
    w = dict(zip(STREAMS, WEIGHTS))
    for _ in range(warmup):
        loop_port.scoring_step(tf, cand, w, THRESHOLD, NEAR_MISS)
    t0 = time.perf_counter()
    for _ in range(steps):
        loop_port.scoring_step(tf, cand, w, THRESHOLD, NEAR_MISS)
    dt = time.perf_counter() - t0
    return n_sample * steps / dt, dt


def cpu_vectorised_throughput(n_sample):
    """'Fair CPU' row: float64 numpy restatement (BLAS, all host cores) on fp32-stored rows."""
    from oracle import scoring, synth
    X = synth.database(DATA_SEED, n_sample)[:, :, None, :]
    T = scoring.scale_target(synth.rows(DATA_SEED, [REF_ROW]).astype(np.float64)[0][:, None, :])
    t0 = time.perf_counter()
    sims, _ = scoring.similarities(X, T)
    sc = scoring.scores(sims, WEIGHTS)
    scoring.classify(sc, THRESHOLD, NEAR_MISS)
    scoring.topk_stable(sc, TOPK)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    total = max(args.steps + args.warmup, 1)
    n_sample = int(min(20000, max(200, 120.0 * 3500 / total)))
    v, dt = cpu_port_throughput(n_sample, steps=args.steps, warmup=args.warmup)
    sample = ("%d-clip slice of the workload per step (loop port of ticket.py:120-180,325-327; features "
              "as Python lists already in memory, HTTP/payload construction excluded)" % n_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (VQSYN-1)",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": 1, "cores_available": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is single-threaded Python (GIL-bound loops); it cannot use more host threads",
    }
    if not args.no_cpu:                                   # the 'fair CPU' row beside it: the same arithmetic vectorised, all cores
        vv, vdt = cpu_vectorised_throughput(200000)
        line["cpu_baseline"]["vectorised_numpy_f64"] = {
            "value": vv, "unit": "clips/s", "cores": cores,
            "sample": "200000-clip slice, one pass (%.1f s), float64 numpy restatement, BLAS on all cores" % vdt}
    emit(line)


def workload_config(args, world):
    return {"workload": "configs[1]: %d-clip synthetic DB per GPU, single-query scan + threshold/near-miss lists "
                        "+ top-%d" % (args.clips_per_gpu, TOPK),
            "clips_per_gpu": args.clips_per_gpu, "global_clips": args.clips_per_gpu * world, "streams": 2,
            "dim": DIM, "bytes_per_clip": ROW_BYTES, "topk": TOPK, "threshold": THRESHOLD, "near_miss": NEAR_MISS,
            "weights": list(WEIGHTS), "parallelism": "clip-range shards x%d" % world,
            "l2": "input %.2f GB per GPU > 126 MB L2, no flush" % (args.clips_per_gpu * ROW_BYTES / 1e9)}


# ------------------------------------------------------------------------------------ GPU arm
def emit(line):
    """Exactly one JSON line on the real stdout (libraries such as NCCL print banners to fd 1)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)                      # everything else that writes to stdout goes to stderr


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.workload == "config4":
        line = run_batched(args, rank, world, local_rank)
        if line is not None:
            emit(line)
        return
    if args.workload == "config5":
        line = run_bootstrap(args, rank, world, local_rank)
        if line is not None:
            emit(line)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import video_query_algorithms_b200 as vq
    from video_query_algorithms_b200 import _ffi
    from video_query_algorithms_b200.store import make_params
    lib = _ffi.lib()

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # host-only barrier: ranks that wait leave their GPU idle
    n = args.clips_per_gpu
    st = vq.FeatureStore(n, STREAMS, [1], DIM, devices=[local_rank], first_global_row=rank * n)
    st.fill_synthetic(DATA_SEED)
    handle = st.shards[0].handle

    # target = reference clip scaled by its squared norm (target_clip.py:311-313); the clip lives on rank 0
    T = torch.zeros(2 * DIM, dtype=torch.float64, device=dev)
    if rank == 0:
        f = st.download(REF_ROW, 1)[0].astype(np.float64)            # [2, 1, 1024]
        t = np.stack([vq.TargetClip._scale_feature(f[s, 0]) for s in range(2)])
        T.copy_(torch.from_numpy(t.reshape(-1)))
    if world > 1:
        dist.broadcast(T, src=0)
    T_host = T.cpu().numpy().reshape(2, 1, DIM)
    target_dev = T.to(torch.float32).contiguous()
    tdict = {s: {1: T_host[i, 0]} for i, s in enumerate(STREAMS)}
    lower = THRESHOLD - NEAR_MISS * (1 - THRESHOLD)
    params = make_params(WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK)

    stream = torch.cuda.Stream(device=dev)       # explicit stream: handle 0 would mean "the store's own stream"
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    from video_query_algorithms_b200.sharded import RankScan

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t_ = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def device_leg(handle_, n_local, steps, warmup, sampler=None):
        """K1 + K2 (+ exchange) per step on this rank's resident shard, timed on the device: W warm-up steps, a barrier,
        then exactly `steps` steps between two events on the launching stream; max over ranks.  Returns the whole-job
        ms per step and the per-phase kernel times (max over ranks of each rank's mean)."""
        rs = RankScan(handle_, TOPK, local_rank, dist if world > 1 else None, torch, exchange=args.exchange)
        for _ in range(warmup):
            rs.enqueue(target_dev.data_ptr(), params, stream.cuda_stream)
        rs.flush(stream.cuda_stream)
        barrier()
        tmp = np.empty(1024, np.float32)
        cnt = C.c_int32()
        lib.vq_scan_kernel_times(handle_, 1024, _ffi.ptr(tmp), C.byref(cnt))       # drop the warm-up timings
        rs.exchange_times()
        if sampler is not None:
            sampler.start()                          # NVML was initialised when the sampler was built: nothing slow from here on
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()                                    # every rank starts its timed region together
        ev0.record(stream)
        for _ in range(steps):
            rs.enqueue(target_dev.data_ptr(), params, stream.cuda_stream)
        rs.flush(stream.cuda_stream)                 # the last step's merge (and the exchange stream) belong to the timed region
        ev1.record(stream)
        barrier()
        if sampler is not None:
            sampler.stop_flag = True
            sampler.join()
        ms_total = max_over_ranks(ev0.elapsed_time(ev1))
        k1, sel = np.empty(1024, np.float32), np.empty(1024, np.float32)
        _ffi.check(lib.vq_scan_phase_times(handle_, 1024, _ffi.ptr(k1), _ffi.ptr(sel), C.byref(cnt)), "vq_scan_phase_times")
        xt = rs.exchange_times()
        k1_mine = float(np.mean(k1[:cnt.value])) if cnt.value else float("nan")
        phases = {"k1_ms": max_over_ranks(k1_mine), "k1_ms_rank0": k1_mine,
                  "select_ms": max_over_ranks(float(np.mean(sel[:cnt.value])) if cnt.value else float("nan")),
                  "exchange_ms_overlapped": max_over_ranks(float(np.mean(xt))) if len(xt) else 0.0,
                  "launches_timed": int(cnt.value)}
        if world > 1 and rs._x is not None:
            # the exchange kernel by itself on an otherwise idle GPU (in the timed region it runs beside the next step's
            # scan, where every one of its dependent memory round trips queues behind a saturated HBM)
            fn = lib.vq_scan_exchange_enqueue_lagged if args.exchange == "p2p-lagged" else lib.vq_scan_exchange_enqueue
            for _ in range(24):
                _ffi.check(fn(handle_, rs._x, sptr), "vq_scan_exchange_enqueue")
            rs.flush(stream.cuda_stream)
            barrier()
            xi, xp = rs.exchange_times(parts=True)
            phases["exchange_ms_alone"] = max_over_ranks(float(np.mean(xi[4:]))) if len(xi) > 4 else None
            if len(xi) > 4:                                  # ... split: NVLink pushes | waiting for the slowest peer | the merge itself
                for j, name in enumerate(("push", "wait_for_peers", "merge")):
                    phases["exchange_ms_alone_" + name] = max_over_ranks(float(np.mean(xp[4:, j])))
        return ms_total / steps, phases, rs

    sampler = ClockSampler(local_rank)               # NVML init happens here, outside every timed region
    warm = max(args.warmup, 3)
    ms_step, phases, rank_scan = device_leg(handle, n, args.steps, warm, sampler)
    k1_ms = phases["k1_ms_rank0"]
    value = world * n * args.steps / (ms_step * args.steps * 1e-3)

    # counts of the last step (sanity: the work is real)
    sc_counts = _ffi.ScanCounts()
    _ffi.check(lib.vq_scan_wait(handle, sptr, C.byref(sc_counts)), "vq_scan_wait")
    g_counts, g_rows, g_scores = rank_scan.result()
    kernels_per_step = rank_scan.kernels_per_step()
    exchange_mode = rank_scan.exchange
    rank_scan.close()

    # ---- end to end through the public API, host buffers: target H2D, counts + lists + top-k D2H.
    # N > 1 (one process per GPU): RankStore — every rank ends each step with the search set's counts and merged top-k and
    # with ITS segment of the ordered lists in pinned host memory (the segments in rank order are the lists; nothing is
    # replicated, no consumer of the path needs that).  e2e_root: the whole lists on rank 0 only.
    e2e_steps = max(3, min(args.steps, 100))
    rstore = None
    if world > 1:
        from video_query_algorithms_b200.sharded import RankStore
        rstore = RankStore(st, dist, torch, dev)

    def timed(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = fn()
        barrier()
        return max_over_ranks(time.perf_counter() - t0), out

    def e2e_step():
        if rstore is not None:
            g_cnt, sh, g_top = rstore.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK, copy=False)
            mine = sum(len(sh.local(w)[0]) for w in ("matches", "near_misses", "ties"))
            return st.last, mine
        res = st.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK)
        k_rows, k_sc = st.topk()
        m_rows, m_sc = st.matches(copy=False)          # as Ticket.select_clips_to_review reads them: views of the
        nm_rows, nm_sc = st.near_misses(copy=False)    # pinned host mirror the scan's publish kernel wrote
        return res, res.n_match + res.n_near + res.n_tie

    e2e_s, (res, n_listed) = timed(e2e_step, e2e_steps)
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = ROW_BYTES + C.sizeof(_ffi.ScanParams)
    d2h = 40 + TOPK * 12 + n_listed * 12       # per rank: its shard's lists
    e2e_what = ("FeatureStore.scan + topk + matches + near_misses through the C ABI: target from a "
                "host buffer, counts / top-k / ordered lists (int64 rows + fp32 scores) published into pinned host "
                "memory inside the call; the shard stays resident in HBM between queries (the store outlives "
                "broker ticks)")
    sel_what = ("the review round as Ticket.select_clips_to_review runs it: FeatureStore.scan(lists=False) "
                "+ topk + tie band + gather of 10 sampled matches and 9 sampled near misses + best near miss; "
                "the match / near-miss lists stay on the device")
    e2e_root = None
    if world > 1:
        small_how = ("through the shared-memory host mailbox, vq_hostx_allgather" if rstore.mailbox is not None
                     else "NCCL collectives staged through the devices")
        e2e_what = ("RankStore.scan on every rank (one process per GPU): the same call as N = 1 on the rank's shard — target "
                    "from a host buffer, the rank's segment of the ordered lists (int64 global rows + fp32 scores, %d entries on "
                    "rank 0) published into its pinned host memory — plus ONE exchange of a summary record per rank (%s) that "
                    "gives every rank the search set's counts, the merged top-k, the tie band and every segment's length; "
                    "the segments in rank order are the search set's lists in database order, nothing is replicated; "
                    "h2d / d2h bytes are per rank" % (n_listed, small_how))
        sel_what = ("RankStore.scan_select + gather_many on every rank: lists stay on each rank's device; two exchanges (%s) — "
                    "one summary record per rank (counts, top-k, tie band, best near miss) and 16 B per position to fetch "
                    "the 19 sampled entries from the ranks that own them" % small_how)

        def root_step():
            g_cnt, lists, g_top = rstore.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK, copy=False, lists="root")
            return int(g_cnt[0]) + int(g_cnt[1]) + int(g_cnt[2])
        root_s, n_all = timed(root_step, e2e_steps)
        e2e_root = {"value": world * n * e2e_steps / root_s, "unit": "clips/s", "steps": e2e_steps,
                    "d2h_bytes_per_step_rank0": int(40 + TOPK * 12 + n_all * 12),
                    "what": "the same with the WHOLE lists (%d entries) delivered to rank 0 only (the rank that talks to the API): "
                            "unpadded NCCL send / recv of each rank's segment device to device, one copy into rank 0's pinned "
                            "memory (RankStore.scan(lists='root'))" % n_all}

    # ---- the review round's call (Ticket.select_clips_to_review): lists stay on the device, the host draws 20 list
    # positions with the reference's RNG and gathers just those entries + the best near miss
    import random as _random
    _random.seed(a=os.environ["RANDOM_SEED"])
    def select_step():
        if rstore is not None:                         # same seed on every rank -> same positions on every rank
            cnt, top, ties, nb = rstore.scan_select(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK)
            n_m, n_nm = int(cnt[0]), int(cnt[1])
        else:
            r2 = st.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK, lists=False)
            st.topk()
            st.ties(copy=False)
            nb = st.near_best()
            n_m, n_nm = r2.n_match, r2.n_near
        pos_m = _random.sample(range(n_m), min(10, n_m))
        pos_n = _random.sample(range(max(n_nm - 1, 0)), min(9, max(n_nm - 1, 0)))
        if rstore is not None:
            rstore.gather_many([("matches", pos_m), ("near_misses", pos_n)])
        else:
            st.gather_many([("matches", pos_m), ("near_misses", pos_n)])

    sel_s, _ = timed(select_step, e2e_steps)
    e2e_select_value = world * n * e2e_steps / sel_s

    # ---- cold end to end: the shard itself is uploaded from pinned host memory every step
    cold = None
    if world == 1 and not args.no_cold:
        try:
            pinned = torch.empty(n * 2 * DIM, dtype=torch.float32, pin_memory=True)
            host = pinned.numpy()
            _ffi.check(lib.vq_store_download(handle, 0, n, _ffi.ptr(host)), "vq_store_download")
            st.upload(0, host)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cold_steps = 3
            for _ in range(cold_steps):
                st.upload(0, host)
                st.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK)
                st.topk(); st.matches(); st.near_misses()
            dt = time.perf_counter() - t0
            cold = {"value": n * cold_steps / dt, "unit": "clips/s", "h2d_bytes_per_step": n * ROW_BYTES + h2d,
                    "d2h_bytes_per_step": d2h, "steps": cold_steps,
                    "what": "feature DB re-uploaded from pinned host memory every step (the reference re-fetches "
                            "it over HTTP every job); PCIe-bound"}
            del pinned, host
        except Exception as e:                            # pragma: no cover
            cold = {"value": None, "error": repr(e)[:200]}

    if rstore is not None:
        rstore.close()
    st.close()

    # ---- the broker's arrangement: ONE process drives every GPU of the box (reference src/broker.py:62-92; this is what
    # compute_matches -> Ticket -> FeatureStore.scan does): rank 0 holds a store of N x clips_per_gpu clips sharded over
    # all N devices and scans it with one library call per query (vq_scan_multi); the other ranks wait on a host barrier
    single = None
    if world > 1:
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                big = vq.FeatureStore(world * n, STREAMS, [1], DIM, devices=list(range(world)))
                big.fill_synthetic(DATA_SEED)

                def sp_step():
                    r_ = big.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK)
                    big.topk()
                    big.matches(copy=False)
                    big.near_misses(copy=False)
                    return r_
                for _ in range(3):
                    sp_step()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    r_ = sp_step()
                dt = time.perf_counter() - t0

                def sp_select():
                    r2 = big.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=TOPK, lists=False)
                    big.topk(); big.ties(copy=False); big.near_best()
                    big.gather_many([("matches", _random.sample(range(r2.n_match), 10)),
                                     ("near_misses", _random.sample(range(r2.n_near - 1), 9))])
                for _ in range(3):
                    sp_select()
                t1 = time.perf_counter()
                for _ in range(e2e_steps):
                    sp_select()
                dt_sel = time.perf_counter() - t1
                single = {"value": world * n * e2e_steps / dt, "unit": "clips/s", "steps": e2e_steps,
                          "ms_per_query": 1e3 * dt / e2e_steps, "select_value": world * n * e2e_steps / dt_sel,
                          "select_ms_per_query": 1e3 * dt_sel / e2e_steps,
                          "k1_ms_max_over_shards": float(r_.scan_ms), "n_match": int(r_.n_match), "n_near": int(r_.n_near),
                          "d2h_bytes_per_step": int(40 * world + TOPK * 12 * world + (r_.n_match + r_.n_near + r_.n_tie) * 12),
                          "what": "ONE process, one thread, %d GPUs: FeatureStore(devices=[0..%d]).scan + topk + matches + "
                                  "near_misses — the call compute_matches makes — through vq_scan_multi (target copy, K1, K2 and "
                                  "the publish kernel enqueued on every shard's stream, then one wait per stream; top-k merged "
                                  "in the call); select_value: the review round's variant (lists stay on the devices, 19 "
                                  "sampled entries gathered)" % (world, world - 1)}
                big.close()
            except Exception as e:                        # pragma: no cover
                single = {"value": None, "error": repr(e)[:300]}
        dist.barrier(group=host_group)

    # ---- BASELINE configs[2], one shard of it: 12.5M clips per GPU (102.4 GB; 8 GPUs = the 100M-clip database), a few
    # device-timed steps with the same kernels and exchange
    shard3 = None
    n3 = 12_500_000
    free_b, total_b = torch.cuda.mem_get_info(dev)
    fits = torch.tensor([1 if (not args.no_config3 and free_b > n3 * ROW_BYTES * 1.04 + (2 << 30)) else 0], device=dev)
    if world > 1:
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    if int(fits.item()) == 1:
        st3 = vq.FeatureStore(n3, STREAMS, [1], DIM, devices=[local_rank], first_global_row=rank * n3)
        st3.fill_synthetic(DATA_SEED)
        steps3 = 10
        sampler3 = ClockSampler(local_rank)
        ms3, ph3, rs3 = device_leg(st3.shards[0].handle, n3, steps3, 3, sampler3)
        c3, _, _ = rs3.result()
        rs3.close()
        st3.close()
        shard3 = {"clips_per_gpu": n3, "global_clips": n3 * world, "steps": steps3, "warmup": 3, "ms_per_step": ms3,
                  "value": world * n3 / (ms3 * 1e-3), "unit": "clips/s", "hbm_gbs_aggregate": world * n3 * ROW_BYTES / (ms3 * 1e-3) / 1e9,
                  "frac_of_nominal_8TBs_per_gpu": n3 * ROW_BYTES / (ms3 * 1e-3) / 1e9 / 8000.0, "kernel_times": ph3,
                  "global_counts": {"n_match": int(c3[0]), "n_near": int(c3[1]), "n_tie": int(c3[2])},
                  "clocks": sampler3.result(),
                  "what": "weak-scaled shard of BASELINE configs[2]: %d clips (%.1f GB) resident per GPU; N = 8 is the 100M-clip, "
                          "819 GB database of the north star (target: <= 16.0 ms per query = 80 %% of 8 x 8 TB/s)"
                          % (n3, n3 * ROW_BYTES / 1e9)}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = n * ROW_BYTES / (k1_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
            if int(tj["clips_per_launch"]) == n:          # one ncu --set full capture of this very launch shape; never rescaled
                traffic, traffic_src = int(tj["dram_bytes_per_launch"]), tj.get("source")
        except Exception:
            pass
        # N = 1: the one call there is.  N > 1: the headline `e2e` is the broker's arrangement — ONE process drives all N GPUs
        # through FeatureStore.scan -> vq_scan_multi, the call compute_matches makes, and ends with the WHOLE result (all
        # lists) in one place like the N = 1 call; the one-process-per-GPU arrangement (RankStore; every rank keeps its
        # segment of the lists) is reported beside it as e2e_ranks / e2e_ranks_select / e2e_root
        ranks_e2e = {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "steps": e2e_steps, "what": e2e_what}
        ranks_sel = {"value": e2e_select_value, "unit": "clips/s", "h2d_bytes_per_step": int(h2d + 19 * 8),
                     "d2h_bytes_per_step": int(64 + TOPK * 12 + res.n_tie * 12 + 19 * 12), "steps": e2e_steps, "what": sel_what}
        e2e_main, e2e_sel_main, e2e_ranks, e2e_ranks_sel = ranks_e2e, ranks_sel, None, None
        if world > 1 and single is not None and single.get("value"):
            e2e_main = {"value": single["value"], "unit": "clips/s", "h2d_bytes_per_step": int(h2d * world),
                        "d2h_bytes_per_step": int(single["d2h_bytes_per_step"]), "steps": single["steps"],
                        "ms_per_query": single["ms_per_query"], "what": single["what"]}
            e2e_sel_main = {"value": single["select_value"], "unit": "clips/s", "h2d_bytes_per_step": int((h2d + 19 * 8) * world),
                            "d2h_bytes_per_step": int((64 + TOPK * 12) * world + res.n_tie * 12 * world + 19 * 12), "steps": single["steps"],
                            "ms_per_query": single["select_ms_per_query"],
                            "what": "the review round in the same arrangement (one process, all GPUs): FeatureStore.scan(lists=False) + "
                                    "topk + tie band + best near miss + ONE vq_gather_list_multi for the 19 sampled entries"}
            e2e_ranks, e2e_ranks_sel = ranks_e2e, ranks_sel
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (VQSYN-1 counter-based generator, on device)",
            "config": workload_config(args, world),
            "hbm_gbs_whole_step": value * ROW_BYTES / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "scan_rows_reg<2,8> (K1)", "kernel_ms": k1_ms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": n * ROW_BYTES},
            "kernel_times": dict(phases, what="CUDA events inside the timed region, mean per launch, max over ranks: K1 scan, "
                                 "K2a-c selection; exchange kernel timed by itself on the global timer — `overlapped`: inside the timed "
                                 "region, on its own stream beside the next step's K1 (off the scan stream's critical path); "
                                 "`alone`: the same kernel back to back on an idle GPU"),
            "e2e": e2e_main,
            "e2e_select": e2e_sel_main,
            "e2e_ranks": e2e_ranks,
            "e2e_ranks_select": e2e_ranks_sel,
            "e2e_root": e2e_root,
            "e2e_single_process": single,
            "e2e_cold": cold,
            "config3_shard": shard3,
            "gpu_launches": kernels_per_step * args.steps,
            "exchange": exchange_mode,
            "clocks": sampler.result(),
            "last_step_counts": {"n_match": int(sc_counts.n_match), "n_near": int(sc_counts.n_near),
                                 "n_tie": int(sc_counts.n_tie), "n_topk": int(sc_counts.n_topk)},
            "last_step_global": {"n_match": int(g_counts[0]), "n_near": int(g_counts[1]), "n_topk": int(g_counts[3]),
                                 "top1_row": int(g_rows[0]) if len(g_rows) else None,
                                 "top1_score": float(g_scores[0]) if len(g_scores) else None},
        }
        if world == 1 and not args.no_extra:
            # the other two GPU configurations of BASELINE.json, a few steps each, so that every round has driver-timed numbers
            extra = {}
            for name, fn, kw in (("config4", run_batched, {"steps": 5, "clips_per_gpu": 10_000_000}),
                                 ("config5", run_bootstrap, {"steps": 5, "clips_per_gpu": 1_000_000})):
                try:
                    a2 = argparse.Namespace(**dict(vars(args), no_cpu=True, **kw))
                    sub = fn(a2, 0, 1, local_rank, steps=kw["steps"])
                    extra[name] = {k_: sub[k_] for k_ in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "dtype",
                                                          "config", "roofline", "e2e", "clocks", "last_step") if k_ in sub}
                except Exception as e:                    # pragma: no cover
                    extra[name] = {"error": repr(e)[:300]}
            line["other_workloads"] = extra
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            v, dt = cpu_port_throughput(40000)
            vv, vdt = cpu_vectorised_throughput(200000)
            line["cpu_baseline"] = {
                "value": v, "unit": "clips/s", "cores": 1, "cores_available": cores, "kind": "port",
                "sample": "40000-clip slice of the workload, one pass (%.1f s): loop port of the reference's "
                          "compute_similarities/compute_scores/candidate scans, Python lists in memory" % dt,
                "vectorised_numpy_f64": {"value": vv, "unit": "clips/s", "cores": cores,
                                         "sample": "200000-clip slice, one pass (%.1f s), BLAS on all cores" % vdt}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ the other GPU configs
def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _setup(local_rank, world):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import video_query_algorithms_b200 as vq
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, vq, dev


def run_batched(args, rank, world, local_rank, steps=None):
    """BASELINE configs[3]: Q targets against every rank's resident shard in one pass on the tensor cores (K3), per-query
    counts + top-k; N > 1: weak scaling, per-query top-k merged across ranks (RankStore.scan_batch).  A step = one
    scan_batch call with host buffers in and out; `value` uses the device time of its kernels (CUDA events on the
    library's stream, max over ranks), `e2e` the wall time of the call between barriers."""
    torch, dist, vq, dev = _setup(local_rank, world)
    n, Q = args.clips_per_gpu, args.queries
    st = vq.FeatureStore(n, STREAMS, [1], DIM, devices=[local_rank], first_global_row=rank * n)
    st.fill_synthetic(DATA_SEED)
    T = torch.zeros(Q * 2 * DIM, dtype=torch.float32, device=dev)
    if rank == 0:                                            # the query clips live in rank 0's shard
        t = np.empty((Q, 2, 1, DIM), np.float32)
        for q in range(Q):
            f = st.download(REF_ROW + 37 * q, 1)[0].astype(np.float64)
            t[q, :, 0] = [vq.TargetClip._scale_feature(f[s_, 0]) for s_ in range(2)]
        T.copy_(torch.from_numpy(t.reshape(-1)))
    if world > 1:
        dist.broadcast(T, src=0)
    targets = T.cpu().numpy().reshape(Q, 2, 1, DIM)
    lower = THRESHOLD - NEAR_MISS * (1 - THRESHOLD)
    rstore = None
    if world > 1:
        from video_query_algorithms_b200.sharded import RankStore
        rstore = RankStore(st, dist, torch, dev)
    call = (rstore or st).scan_batch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        call(targets, WEIGHTS, THRESHOLD, lower, topk=TOPK)
    if steps is None:
        steps = args.steps if args.steps != 300 else 20      # the default K is config2's; a step here is ~30 ms
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    k_ms = []
    t0 = time.perf_counter()
    for _ in range(steps):
        counts, rows, scores, ms = call(targets, WEIGHTS, THRESHOLD, lower, topk=TOPK)
        k_ms.append(ms)
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    tt = torch.tensor(k_ms + [wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tt = tt.tolist()
    k_ms, wall = tt[:-1], tt[-1]
    line = None
    if rank == 0:
        ms_step = float(np.mean(k_ms))
        pairs = float(world) * n * Q
        flops = 2.0 * pairs * 2 * DIM                        # algorithmic: 2 * Q * N * S * D
        pk = _peaks()
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)))
        tensor_pipe = None                                   # from ONE ncu --set full capture of a steady-state launch (profiles/)
        try:
            with open(os.path.join(ROOT, "profiles", "k3_tensor_pipe.json")) as f:
                tj = json.load(f)
            if int(tj["queries"]) == Q:
                tensor_pipe = {"active_pct_of_elapsed": tj["tensor_pipe_active_pct_of_elapsed"], "source": tj["source"]}
        except Exception:
            pass
        achieved = flops / world / (ms_step * 1e-3) / 1e12   # per GPU, like the peak
        line = {
            "metric": "clip-query pairs scored/sec (batched queries)", "value": pairs / (ms_step * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x2 (fp32 split into two bf16 terms, fp32 accumulate)",
            "data": "synthetic (VQSYN-1 counter-based generator, on device)",
            "config": {"workload": "configs[3]: batched %d-query scoring vs %d clips per GPU as tcgen05 GEMM + fused "
                                   "per-query counts and top-%d" % (Q, n, TOPK), "clips_per_gpu": n, "global_clips": n * world,
                       "queries": Q, "streams": 2, "dim": DIM, "topk": TOPK, "threshold": THRESHOLD, "near_miss": NEAR_MISS,
                       "weights": list(WEIGHTS), "parallelism": "clip-range shards x%d" % world,
                       "l2": "input %.1f GB per GPU > 126 MB L2, no flush" % (n * ROW_BYTES / 1e9)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "batch_scan_bf16 (K3) + batch_compact", "kernel_ms": ms_step,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, back to back)" if pk
                         else "fallback 1590", "algorithmic_flops_per_step_per_gpu": flops / world,
                         "executed_tflops_bf16x2": 3 * achieved, "frac_executed": 3 * achieved / peak,
                         "tensor_pipe_ncu": tensor_pipe,
                         "note": "three bf16 MMAs per fp32 product (x1*t1 + x2*t1 + x1*t2): the tensor pipe executes 3x "
                                 "the algorithmic flops"},
            "e2e": {"value": pairs * steps / wall, "unit": "pairs/s", "h2d_bytes_per_step": int(Q * ROW_BYTES + 64),
                    "d2h_bytes_per_step": int(Q * (16 + TOPK * 12)), "steps": steps,
                    "what": "%s.scan_batch with host buffers: targets H2D, per-query counts and top-%d D2H%s"
                            % ("RankStore" if world > 1 else "FeatureStore", TOPK,
                               ", per-query merge across ranks (allgather + vq_merge_topk_batch)" if world > 1 else "")},
            "gpu_launches": None, "clocks": sampler.result(),
            "last_step": {"query0_counts": [int(x) for x in counts[0]], "query0_top1_row": int(rows[0][0]),
                          "query0_top1_score": float(scores[0][0])},
        }
        if world == 1 and not args.no_cpu:
            v, dt = cpu_port_throughput(20000)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": 1,
                                    "cores_available": len(os.sched_getaffinity(0)), "kind": "port",
                                    "sample": "one query against a 20000-clip slice (%.1f s): the reference scores one "
                                              "query per job, so its pairs/s is its clips/s" % dt}
    if rstore is not None:
        rstore.close()
    st.close()
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


def run_bootstrap(args, rank, world, local_rank, steps=None):
    """BASELINE configs[4]: the weight update (hyperparameter.py:29-76) for R bootstrap replicates at once over L labelled
    clips of a 1M-clip DB; replicate index sets drawn host-side from Python's RANDOM_SEED-driven generator exactly as
    the reference's bagging draws them.  The labelled rows live on one GPU: replicas only (rank 0 runs, N is ignored)."""
    if rank != 0:
        return None
    import random
    torch, dist, vq, dev = _setup(local_rank, 1)
    n, L, R = args.clips_per_gpu, args.labelled, args.replicates
    st = vq.FeatureStore(n, STREAMS, [1], DIM, devices=[local_rank])
    st.fill_synthetic(DATA_SEED)
    st.set_clip_ids(np.arange(n))
    f = st.download(REF_ROW, 1)[0].astype(np.float64)
    tdict = {s_: {1: vq.TargetClip._scale_feature(f[i, 0])} for i, s_ in enumerate(STREAMS)}
    lower = THRESHOLD - NEAR_MISS * (1 - THRESHOLD)
    st.scan(tdict, WEIGHTS, THRESHOLD, lower, EPS, topk=0)
    m_rows, m_sc = st.matches()
    n_rows, n_sc = st.near_misses()
    rng = np.random.default_rng(7)                           # labelled set: half matches, half near misses; label = score >= 0.82
    im, inm = rng.choice(len(m_rows), L // 2, replace=False), rng.choice(len(n_rows), L - L // 2, replace=False)
    pick = np.concatenate([m_rows[im], n_rows[inm]])
    sc_ = np.concatenate([m_sc[im], n_sc[inm]])
    order = np.argsort(pick)
    matches = [{"video_clip": int(c), "user_match": bool(v >= 0.82), "is_match": bool(v >= THRESHOLD)}
               for c, v in zip(pick[order], sc_[order])]

    class _Ticket:                                           # what optimize_weights reads from a ticket
        pass
    t = _Ticket()
    t.matches, t.target = matches, _Ticket()
    t.target.target_features = tdict
    t.feature_store = lambda optional=False: st
    hp = vq.Hyperparameter(dict(zip(STREAMS, WEIGHTS)), ballast=0.0)
    if steps is None:
        steps = args.steps if args.steps != 300 else 10
    random.seed(a=os.environ["RANDOM_SEED"])

    def step():
        t0 = time.perf_counter()
        reps = vq.resample_labelled(L, R, random)
        t1 = time.perf_counter()
        w, th = hp.optimize_weights_replicates(t, reps)
        return t1 - t0, time.perf_counter() - t1, w, th

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    parts = [step() for _ in range(steps)]
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    draw_s, upd_s = float(np.mean([p_[0] for p_ in parts])), float(np.mean([p_[1] for p_ in parts]))
    w, th = parts[-1][2], parts[-1][3]
    line = {
        "metric": "bootstrap replicates of the weight update per second", "value": R / upd_s, "unit": "replicates/s",
        "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * upd_s, "higher_is_better": True,
        "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic (VQSYN-1 counter-based generator, on device)",
        "config": {"workload": "configs[4]: bootstrap weight update, %d seeded replicates over %d labelled clips on a "
                               "%d-clip DB" % (R, L, n), "replicates": R, "labelled": L, "clips": n,
                   "grid": "40 weights x 31 thresholds", "mean_replicate_size": float(np.mean([len(r_) for r_ in
                                                                                              vq.resample_labelled(L, 8, random)]))},
        "e2e": {"value": R * steps / wall, "unit": "replicates/s", "h2d_bytes_per_step": int(L * 17 + R * L * 0.632 * 4),
                "d2h_bytes_per_step": int(R * 40 * 31 * 8), "steps": steps,
                "what": "resample_labelled (host, Python's own Mersenne Twister stream: %.1f ms) + Hyperparameter."
                        "optimize_weights_replicates (labelled fp64 similarities K4, loss grid K5 for all replicates, argmin + "
                        "parabola fit on the host: %.1f ms)" % (1e3 * draw_s, 1e3 * upd_s)},
        "gpu_launches": None, "clocks": sampler.result(),
        "last_step": {"weight_mean_std": [float(w.mean()), float(w.std())], "threshold_mean_std": [float(th.mean()), float(th.std())]},
        "cpu_baseline": {"value": 1.0 / (40 * n * 3.1e-6 + 40 * 31 * 0.632 * L * 1.7e-6), "unit": "replicates/s", "cores": 1,
                         "kind": "port", "sample": "projection, not a run: the reference rescans the whole DB for each of its 40 "
                                                   "weights (3.1 us per clip) and walks 40 x 31 x the replicate's labelled clips "
                                                   "(1.7 us each), per-item costs measured on the loop port (SURVEY.md §8 A8)"},
    }
    st.close()
    return line


if __name__ == "__main__":
    main()
