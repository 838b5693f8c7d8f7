#!/usr/bin/env python3
"""bench.py — per-frame stereo front-end throughput on B200 (BASELINE.json metric).

One "step" = one batch of `--batch` stereo frames through the whole hot path: extract(L) +
extract(R) + sparse stereo/SAD + BFMatcher(cur->prev) + greedy pass 1 (previous frame's map
points) + greedy pass 2 (5k-row local map).  Workload = BASELINE.json configs[1]: a synthetic
KITTI-shape (1241x376) stereo sequence with KITTI04-12 intrinsics, 2000 ORB features, 8 levels.

  value : frames/s with every input already resident in HBM (device pointers into the C ABI)
  e2e   : frames/s through the same C-ABI call with pinned HOST buffers — H2D of both images,
          the previous-frame descriptors and the 5k-row map, and D2H of every result, all inside
          the timed region.

Multi-GPU (`torchrun ... bench.py --gpus N`): independent sequences, one per rank, no data-path
collective ("replicas only"); torch.distributed is used for the barrier and the max-over-ranks.

`--impl reference` times the reference's CPU implementation of the same path on all host cores
(cv2.ORB, which is what frame.cc:77 calls, + the C port of the pnpmatch.cc loops and of the
sparse-stereo stage from oracle/).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stereo-semantic-vo_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import replicas  # noqa: E402
import synth  # noqa: E402

W_IMG, H_IMG, NFEAT, NLEVELS, MAP_ROWS = 1241, 376, 2000, 8, 5000
CAL = synth.KITTI_04_12
BF = float(np.float32(CAL["bf"]))
BASELINE = float(np.float32(CAL["bf"] / CAL["fx"]))
DISTRIBUTION = 0          # 1: the opt-in octree distribution (--distribution octree)
METRIC = "stereo_frames_per_sec"
UNIT = "frames/s"


def kp_cap(nf):
    return (nf + nf // 8 + 64 + 63) // 64 * 64


def level_sizes():
    # cv::ORB geometry, scale 1.2: cvRound(size / (float)pow(1.2, l)); 1241x376 gives the SURVEY.md section 8 table
    out = []
    for l in range(NLEVELS):
        s = float(np.float32(pow(float(np.float32(1.2)), l)))
        out.append((int(np.rint(np.float32(W_IMG) / np.float32(s))), int(np.rint(np.float32(H_IMG) / np.float32(s)))))
    return out


def algorithmic_bytes_per_image():
    """algorithmic bytes of each extraction kernel per image."""
    px = [w * h for w, h in level_sizes()]
    return {
        "pyramid": sum(px[l - 1] + px[l] for l in range(1, 8)),
        "fast": sum(px),
        "blur": 2 * sum(px),
        "harris": 2 * NFEAT * 81,
        "describe": NFEAT * (512 + 32) + NFEAT * 749,
    }


def build_map(descs, live_prev, rng):
    """5k-row local map for a frame from the 4 previous frames' left descriptors, padded with
    bit-flipped copies; rows taken from the immediately previous frame carry map_prev_row."""
    rows, prow = [], []
    per = MAP_ROWS // 4
    for age, d in enumerate(descs):            # descs[0] = frame t-1, descs[1] = t-2, ...
        take = rng.permutation(len(d))[:per]
        rows.append(d[take])
        prow.append(take.astype(np.int32) if age == 0 else np.full(len(take), -1, np.int32))
    rows = np.concatenate(rows, 0); prow = np.concatenate(prow, 0)
    k = 0
    while len(rows) < MAP_ROWS:
        need = MAP_ROWS - len(rows)
        src = rows[rng.permutation(len(rows))[:need]]
        flips = np.packbits(rng.random((len(src), 256)) < (0.02, 0.1, 0.5)[k % 3], axis=1, bitorder="little")
        rows = np.concatenate([rows, src ^ flips], 0); prow = np.concatenate([prow, np.full(len(src), -1, np.int32)], 0)
        k += 1
    order = rng.permutation(MAP_ROWS)          # explicit, fixed scan order
    return np.ascontiguousarray(rows[order]), np.ascontiguousarray(prow[order])


def workload_string(distribution="retainbest"):
    head = ("configs[1]" if (W_IMG, H_IMG, NFEAT, MAP_ROWS, distribution) == (1241, 376, 2000, 5000, "retainbest")
            else "non-headline configuration" + (" (opt-in octree distribution)" if distribution == "octree" else ""))
    return ("%s: synthetic %s %dx%d stereo sequence, KITTI04-12 intrinsics, %d ORB features / 8 levels / 1.2, full front-end: "
            "extract L+R, sparse stereo + SAD, BF match vs previous frame, greedy pass 1 + pass 2 vs %d-row local map"
            % (head, "KITTI-shape" if (W_IMG, H_IMG) == (1241, 376) else "high-res", W_IMG, H_IMG, NFEAT, MAP_ROWS))


def verify_extract(r, left, right):
    """Checker (not timed): one frame's keypoints and descriptors bit for bit, stereo validity / match-level agreement
    within 1e-3, against the CPU oracle.  Returns the oracle's left descriptors."""
    from oracle import oracle as O
    kl, dl, pl = O.orb(left, NFEAT, with_pyramid=True, distribution=DISTRIBUTION)
    kr, dr, pr = O.orb(right, NFEAT, with_pyramid=True, distribution=DISTRIBUTION)
    ur, dep, mr, sad = O.stereo_sparse(kl, dl, pl, kr, dr, pr, BF, BASELINE)
    O.pyramid_free(pl); O.pyramid_free(pr)
    assert r["status"] == 0 and r["n_left"] == len(kl) and r["n_right"] == len(kr), "keypoint counts"
    for f in ("x", "y", "size", "angle", "response"):
        assert (r["kp_left"][f].view(np.uint32) == kl[f].view(np.uint32)).all(), "left keypoints: " + f
    assert (r["kp_left"]["octave"] == kl["octave"]).all() and (r["desc_left"] == dl).all(), "descriptors"
    if r["kp_right"] is not None:      # SVO_OUT_NO_RIGHT leaves them on the device: the stereo results below still depend on them
        for f in ("x", "y", "size", "angle", "response"):
            assert (r["kp_right"][f].view(np.uint32) == kr[f].view(np.uint32)).all(), "right keypoints: " + f
        assert (r["desc_right"] == dr).all(), "right descriptors"
    valid = dep > 0
    assert ((r["depth"] > 0) == valid).all(), "stereo validity"
    if valid.any():
        assert np.abs(r["u_right"][valid] - ur[valid]).max() <= 1e-3, "u_right"
        assert (np.abs(r["depth"][valid] - dep[valid]) <= 1e-3 * np.abs(dep[valid])).all(), "depth"
    return dl


def verify_frame(r, job):
    """Checker for the bench's own configuration (not timed): one frame's results against the CPU oracle —
    extraction and stereo (verify_extract), BF, pass 1, pass 2."""
    from oracle import oracle as O
    dl = verify_extract(r, job["left"], job["right"])
    oi, od, ok = O.match_bf(dl, job["prev_desc"])
    assert (r["bf_idx"] == oi).all() and (r["bf_dist"] == od).all() and (r["bf_keep"] == ok).all(), "BF"
    p1 = O.match_greedy(job["prev_desc"], dl, 0, row_live=job["prev_live"])
    lv = job["prev_live"].astype(bool)
    assert (r["p1_row_claimed"] == p1["row_claimed"]).all(), "pass 1 claims"
    for k in ("best_idx", "best", "second"):
        assert (r["p1_" + k][lv] == p1[k][lv]).all(), "pass 1 " + k
    live2 = np.ones(len(job["map_desc"]), np.uint8)
    m = job["map_prev_row"] >= 0
    live2[m] = 1 - p1["row_claimed"][job["map_prev_row"][m]]
    p2 = O.match_greedy(job["map_desc"], dl, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                        row_base=len(job["prev_desc"]))
    assert (r["p2_row_claimed"] == p2["row_claimed"]).all() and (r["claim_row"] == p2["claim_row"]).all(), "pass 2"
    return int(p1["row_claimed"].sum()), int(p2["row_claimed"].sum())


def verify_tracked_frame(r, left, right, pre, post, frame_id, map_cap, K4):
    """Checker for a TRACKED frame (not timed): extraction and stereo against the oracle, then oracle/track.py continued
    from the state the frame read (`pre`) must give the frame's matches and the state it left (`post`), bit for bit."""
    from oracle import track as T
    verify_extract(r, left, right)
    trk = T.Tracker.from_state(pre, window=4, map_cap=map_cap)
    xy = np.stack([r["kp_left"]["x"], r["kp_left"]["y"]], 1)
    o = trk.step(xy, r["desc_left"], r["depth"], frame_id, K4=K4)
    assert r["n_prev"] == o["n_prev"] and r["n_map"] == o["n_map"], "tracked row counts"
    keys = ["claim_row", "mp_create"]
    if o["n_prev"]:
        keys += ["bf_idx", "bf_dist", "bf_keep", "p1_best_idx", "p1_best", "p1_second", "p1_row_claimed"]
    if o["n_map"]:
        keys.append("p2_row_claimed")
    for k in keys:
        assert np.array_equal(r[k], o[k]), "tracked " + k
    assert (r["mp_xyz"].view(np.uint32) == o["mp_xyz"].view(np.uint32)).all(), "tracked mp_xyz"
    for k in ("last_desc", "prev_desc", "prev_live", "prev_map_row", "prev_create", "map_desc", "map_create", "map_link"):
        assert np.array_equal(post[k], getattr(trk, k)), "tracked state " + k
    return int(o["p1_row_claimed"].sum()) if o["n_prev"] else 0, int(o["p2_row_claimed"].sum()) if o["n_map"] else 0


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {getattr(nv, n): n[len("nvmlClocksEventReason"):] for n in dir(nv) if n.startswith("nvmlClocksEventReason")
                 and isinstance(getattr(nv, n), int)}
        if not names:
            names = {getattr(nv, n): n[len("nvmlClocksThrottleReason"):] for n in dir(nv)
                     if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if bit and (r & bit) and n not in ("None", "GpuIdle", "All"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import svo
    dev = local_rank
    torch.cuda.set_device(dev)
    bus = None
    try:
        pr = torch.cuda.get_device_properties(dev)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        pass
    near = replicas.bind_near_gpu(dev, bus)   # before any pinned allocation: host buffers land on the GPU's NUMA node
    grp = replicas.Group(backend="nccl", device=dev)   # barrier + max-over-ranks only; no data-path collective
    B, P = args.batch, args.pool
    ctx = svo.Context(W_IMG, H_IMG, nfeatures=NFEAT, nlevels=NLEVELS, max_batch=B, lanes=args.lanes,
                      max_rows=max(MAP_ROWS, kp_cap(NFEAT)), device=dev,
                      distribution=svo.DIST_OCTREE if args.distribution == "octree" else svo.DIST_RETAIN_BEST)
    # ---- synthetic sequence (one per rank): pool of P distinct stereo frames in pinned memory
    seq_id = replicas.assign_sequences(world, world, rank)[0]
    seq = synth.Sequence((H_IMG, W_IMG), seed=seq_id)
    rng = np.random.default_rng(1000 + rank)
    pitch = W_IMG
    hl = ctx.pinned_array((P, H_IMG, pitch)); hr = ctx.pinned_array((P, H_IMG, pitch))
    for t in range(P):
        L, R = seq.frame(t)
        hl[t] = L; hr[t] = R
    # ---- untimed setup pass: extract the pool once to obtain the previous-frame descriptors and maps
    K = kp_cap(NFEAT)
    descs, lives = [], []
    for t0 in range(0, P, B):
        n = min(B, P - t0)
        ctx.batch_submit(0, [dict(left=hl[t0 + i], right=hr[t0 + i], bf=BF, baseline=BASELINE) for i in range(n)])
        ctx.batch_wait(0)
        for i in range(n):
            r = ctx.batch_result(0, i)
            assert r["status"] == 0
            descs.append(r["desc_left"]); lives.append((r["depth"] > 0).astype(np.uint8))
    h_prev = ctx.pinned_array((P, K, 32)); h_live = ctx.pinned_array((P, K))
    h_map = ctx.pinned_array((P, MAP_ROWS, 32)); h_mpr = ctx.pinned_array((P, MAP_ROWS), np.int32)
    n_prev = np.zeros(P, np.int32)
    for t in range(P):
        prev = [descs[(t - a) % P] for a in range(1, 5)]
        n_prev[t] = len(prev[0])
        h_prev[t, :n_prev[t]] = prev[0]; h_live[t, :n_prev[t]] = lives[(t - 1) % P]
        h_map[t], h_mpr[t] = build_map(prev, lives[(t - 1) % P], rng)
    # device-resident copies for the `value` measurement
    d_l, d_r = ctx.to_device(hl), ctx.to_device(hr)
    d_prev, d_live, d_map, d_mpr = ctx.to_device(h_prev), ctx.to_device(h_live), ctx.to_device(h_map), ctx.to_device(h_mpr)
    img_b = H_IMG * pitch

    def frame_host(t):
        return dict(left=hl[t], right=hr[t], bf=BF, baseline=BASELINE, prev_desc=h_prev[t, :n_prev[t]],
                    prev_live=h_live[t, :n_prev[t]], map_desc=h_map[t], map_prev_row=h_mpr[t])

    def frame_dev(t):
        return dict(left=d_l + t * img_b, right=d_r + t * img_b, stride=pitch, bf=BF, baseline=BASELINE,
                    prev_desc=d_prev + t * K * 32, n_prev=int(n_prev[t]), prev_live=d_live + t * K,
                    map_desc=d_map + t * MAP_ROWS * 32, n_map=MAP_ROWS, map_prev_row=d_mpr + t * MAP_ROWS * 4)

    # untimed: columns pass 1 leaves free in this workload (pass 2 scans only those; roofline work count below)
    free_cols = []
    for t0 in range(0, min(P, 2 * B), B):
        n = min(B, P - t0)
        ctx.batch_submit(0, [frame_host(t0 + i) for i in range(n)])
        ctx.batch_wait(0)
        for i in range(n):
            r = ctx.batch_result(0, i)
            free_cols.append(int(r["n_left"]) - int(np.asarray(r["p1_row_claimed"]).sum()))
    free_cols = float(np.mean(free_cols))

    streams = [torch.cuda.ExternalStream(ctx.lane_stream(l), device=dev) for l in range(args.lanes)]

    barrier = grp.barrier
    last_ts = {}

    def verify_last_batches(per_lane):
        """After a timed region: the results still sitting in each lane's arena (the LAST timed batch of every lane, at
        the bench's own batch size / lane count / graph replay) are checked against the oracle."""
        nv = claims = 0
        for lane in range(args.lanes):
            if lane not in last_ts:
                continue
            for i in sorted(set(np.linspace(0, B - 1, per_lane).astype(int).tolist())):
                c1, c2 = verify_frame(ctx.batch_result(lane, i), frame_host(last_ts[lane][i]))
                nv += 1; claims += c1 + c2
        return nv, claims

    cursor = [0]

    def pool_batch(make_frame):
        def make(lane):
            ts = [(cursor[0] + i) % P for i in range(B)]
            cursor[0] = (cursor[0] + B) % P
            last_ts[lane] = ts
            return [make_frame(t) for t in ts]
        return make

    def timed(make_batch, steps, warmup, profile):
        def submit(lane):
            ctx.batch_submit(lane, make_batch(lane))

        for s in range(warmup):
            lane = s % args.lanes
            ctx.batch_wait(lane)
            submit(lane)
        for lane in range(args.lanes):
            ctx.batch_wait(lane)
        ctx.set_profiling(profile)
        stage = {}
        barrier()
        launches0 = ctx.launch_count()
        start = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.lanes)]
        start.record(streams[0])
        for l in range(1, args.lanes):
            streams[l].wait_event(start)
        t0 = time.perf_counter()
        for s in range(steps):
            lane = s % args.lanes
            if s >= args.lanes:
                ctx.batch_wait(lane)
                if profile:
                    for k, v in ctx.stage_ms(lane).items():
                        stage[k] = stage.get(k, 0.0) + v
            submit(lane)
        for lane in range(args.lanes):
            ends[lane].record(streams[lane])
        for lane in range(args.lanes):
            ctx.batch_wait(lane)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = max(start.elapsed_time(e) for e in ends)
        launches = ctx.launch_count() - launches0
        ctx.set_profiling(0)
        if profile:
            nprof = max(steps - args.lanes, 1)
            stage = {k: v / nprof for k, v in stage.items()}
        return ms, wall * 1e3, launches, stage

    sampler = ClockSampler(dev)
    sampler.start()
    # headline numbers: graph replay, no per-stage events
    ms_dev, wall_dev, launches, _ = timed(pool_batch(frame_dev), args.steps, args.warmup, 0)
    verified = claims = 0
    if args.verify and rank == 0:
        verified, claims = verify_last_batches(args.verify)      # resident-input pass: last timed batch of every lane
    ms_hc, wall_hc, _, _ = timed(pool_batch(frame_host), args.steps, args.warmup, 0)
    if args.verify and rank == 0:
        v2, c2 = verify_last_batches(args.verify)                # host-input pass
        verified += v2; claims += c2
    # ---- tracked end-to-end pass: the tracker state (last frame's descriptors and map points, local map) stays in HBM
    # (svo_track_*), so only the images cross PCIe.  Every lane owns B sequences; a step advances each by one frame.
    trk = None
    ms_trk = wall_trk = ms_pose = wall_pose = None
    stage_trk = {}
    if not args.no_tracked:
        K4 = (float(CAL["fx"]), float(CAL["fy"]), float(CAL["cx"]), float(CAL["cy"]))
        NS = args.lanes * B
        ctx.track_create(NS, MAP_ROWS, 4)
        step_of = [0] * args.lanes
        last_k = {}
        seen = {"n_prev": [], "n_map": []}

        def tracked_batch(lane):
            k = step_of[lane]; step_of[lane] += 1
            ts = [((lane * B + i) * 5 + k) % P for i in range(B)]      # sequence s starts at pool frame 5 s
            last_ts[lane] = ts; last_k[lane] = k
            return [dict(left=hl[t], right=hr[t], bf=BF, baseline=BASELINE, track_seq=lane * B + i, frame_id=k, K=K4)
                    for i, t in enumerate(ts)]

        def run_untimed(steps_per_lane, record):
            for k in range(steps_per_lane):
                for lane in range(args.lanes):
                    ctx.batch_submit(lane, tracked_batch(lane))
                for lane in range(args.lanes):
                    ctx.batch_wait(lane)
                    if record:
                        for i in range(0, B, 4):
                            r = ctx.batch_result(lane, i, copy=False)
                            seen["n_prev"].append(r["n_prev"]); seen["n_map"].append(r["n_map"])

        # natural size of the local map in this workload (4 frames of new points), then ballast up to the configured rows
        for sq in range(NS):
            ctx.track_reset(sq)
        run_untimed(6, False); run_untimed(2, True)
        natural = float(np.mean(seen["n_map"]))
        n_ballast = int(max(0, min(MAP_ROWS, round(MAP_ROWS - natural))))
        ballast = rng.integers(0, 256, (max(n_ballast, 1), 32), dtype=np.uint8)
        for sq in range(NS):
            ctx.track_reset(sq, ballast[:n_ballast] if n_ballast else None)
        for lane in range(args.lanes):
            step_of[lane] = 0
        run_untimed(6, False)                                           # steady state: the 4-frame window is full
        # results: nfeatures + 32 rows per array instead of the capacities, and no right-image features (the reference's
        # frame keeps none); every other output of the path comes back
        ctx.set_outputs(svo.OUT_COMPACT | svo.OUT_NO_RIGHT)
        ms_trk, wall_trk, _, _ = timed(tracked_batch, args.steps, args.warmup, 0)
        tv = tc = 0
        if args.verify and rank == 0:
            for lane in range(args.lanes):
                for i in sorted(set(np.linspace(0, B - 1, args.verify).astype(int).tolist())):
                    sq, t = lane * B + i, last_ts[lane][i]
                    c1, c2 = verify_tracked_frame(ctx.batch_result(lane, i), hl[t], hr[t], ctx.track_state(sq, previous=True),
                                                  ctx.track_state(sq), last_k[lane], MAP_ROWS, K4)
                    tv += 1; tc += c1 + c2
        seen = {"n_prev": [], "n_map": []}
        run_untimed(2, True)
        trk = {"sequences": NS, "natural_map_rows": natural, "ballast_rows": n_ballast, "mean_map_rows": float(np.mean(seen["n_map"])),
               "mean_prev_rows": float(np.mean(seen["n_prev"])), "verified_frames": tv, "verified_claims": tc}
        verified += tv; claims += tc
        # the same pass returning only what the steps after the matchers read (SVO_OUT_POSE_INPUTS: left keypoints, depth,
        # claims and the owned points' names / positions); the subset equals the full outputs (tests/test_gpu_track.py)
        ctx.set_outputs(svo.OUT_COMPACT | svo.OUT_POSE_INPUTS)
        ms_pose, wall_pose, _, _ = timed(tracked_batch, args.steps, args.warmup, 0)
        ctx.set_outputs(0)
    # per-stage / per-kernel durations: the same steps again with CUDA events on the lanes' own streams (the events
    # split the captured graph into plain launches, so this pass is a few percent slower than the headline)
    psteps = max(args.lanes + 2, min(args.steps, 24))
    ms_prof, _, _, stage = timed(pool_batch(frame_dev), psteps, 3, 1)
    _, _, _, stage_hc = timed(pool_batch(frame_host), psteps, 3, 1)
    if trk is not None:
        ctx.set_outputs(svo.OUT_COMPACT | svo.OUT_NO_RIGHT)
        _, _, _, stage_trk = timed(tracked_batch, psteps, 3, 1)
        ctx.set_outputs(0)
    sampler.stop_flag = True
    sampler.join(timeout=1)

    # single-frame latency (p50 of 30 frames): one frame per call, submit -> wait, host buffers in and out.  Measured on
    # the batch context above and on a LATENCY context (max_batch = 1, one lane: what a deployment that serves one
    # sequence per GPU creates; svo_create then picks the low-latency kernel shapes, e.g. 8-row FAST bands)
    def p50_single(c, tracked_mode):
        lt = []
        if tracked_mode:
            c.set_outputs(svo.OUT_COMPACT | svo.OUT_NO_RIGHT)
        for t in range(min(34, P)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if tracked_mode:   # consecutive frames of ONE sequence through the tracker state
                c.batch_submit(0, [dict(left=hl[t], right=hr[t], bf=BF, baseline=BASELINE, track_seq=0, frame_id=1000 + t, K=K4)])
            else:
                c.batch_submit(0, [frame_host(t)])
            c.batch_wait(0)
            lt.append((time.perf_counter() - t0) * 1e3)
        c.set_outputs(0)
        return float(np.median(lt[4:]))

    K4 = (float(CAL["fx"]), float(CAL["fy"]), float(CAL["cx"]), float(CAL["cy"]))
    p50_batch_ctx = p50_single(ctx, False)
    lctx = svo.Context(W_IMG, H_IMG, nfeatures=NFEAT, nlevels=NLEVELS, max_batch=1, lanes=1, max_rows=max(MAP_ROWS, kp_cap(NFEAT)),
                       device=dev, distribution=svo.DIST_OCTREE if args.distribution == "octree" else svo.DIST_RETAIN_BEST)
    p50 = p50_single(lctx, False)
    p50_trk = None
    if trk is not None:
        lctx.track_create(1, MAP_ROWS, 4)
        lctx.track_reset(0, ballast[:n_ballast] if n_ballast else None)
        p50_trk = p50_single(lctx, True)
    lctx.close()

    # pose stage (not part of the headline metric): PnP RANSAC + pose-only LM for a batch of
    # B frames with ~1000 matched map points each, through the synchronous C-ABI calls (host buffers, copies included)
    pose = None
    if rank == 0:
        probs = []
        for i in range(B):
            Xw, obs, K4, Rt, tt, _ = synth.pose_problem(1000, 500 + i, 0.3, 0.5, cal=CAL)
            probs.append(dict(pts3d=Xw, pts2d=obs, K=K4))
        for _ in range(3):
            rr = ctx.pnp_ransac(probs); ctx.pose_optimize(probs)
        t_r, t_l = [], []
        for _ in range(10):
            t0 = time.perf_counter(); rr = ctx.pnp_ransac(probs); t1 = time.perf_counter()
            for q, r in zip(probs, rr):
                T = np.eye(4, dtype=np.float32); T[:3, :3] = r["R"]; T[:3, 3] = r["t"]; q["Tcw"] = T
            t2 = time.perf_counter(); ctx.pose_optimize(probs); t3 = time.perf_counter()
            t_r.append((t1 - t0) * 1e3); t_l.append((t3 - t2) * 1e3)
        pose = {"frames_per_call": B, "points_per_frame": 1000, "outlier_fraction": 0.3,
                "pnp_ransac_ms_per_call": float(np.median(t_r)), "pose_optimize_ms_per_call": float(np.median(t_l)),
                "frames_per_s": B / ((np.median(t_r) + np.median(t_l)) * 1e-3),
                "note": "wall clock around svo_pnp_ransac (100 samples, 8 px, refit) and svo_pose_optimize (g2o LM, 10 iterations), "
                        "host buffers in and out; includes the Python binding's packing"}

    # input staging beside the path: main.cpp:160-162 reads PNG files; svo_png_decode (host, zlib)
    # turns them into the pinned buffers the calls above take.  Reported so that the host side can be read against the GPU
    # rate: at these frame rates frames must arrive raw (or be decoded by many cores).
    staging = None
    if rank == 0:
        try:
            import cv2
            files = [cv2.imencode(".png", np.ascontiguousarray(hl[t]))[1].tobytes() for t in range(min(8, P))]
            dst = ctx.pinned_array((H_IMG, W_IMG))
            for f in files[:2]:
                svo.png_decode(f, out=dst)
            t0 = time.perf_counter()
            for _ in range(3):
                for f in files:
                    svo.png_decode(f, out=dst)
            ms_img = (time.perf_counter() - t0) * 1e3 / (3 * len(files))
            staging = {"png_decode_ms_per_image_1_core": ms_img, "png_bytes_per_image": int(np.mean([len(f) for f in files])),
                       "stereo_frames_per_s_per_core": 1e3 / (2 * ms_img),
                       "note": "svo_png_decode (host, zlib) of %dx%d 8-bit gray frames encoded by cv2.imencode, into pinned memory; "
                               "a stereo frame is two images" % (W_IMG, H_IMG)}
        except Exception as ex:   # no cv2 on this host: nothing to encode the sample files with
            staging = {"unavailable": str(ex)[:100]}

    ms_dev, ms_hc, ms_trk_m, ms_pose_m = grp.max_over_ranks([ms_dev, ms_hc, ms_trk if ms_trk is not None else 0.0,
                                                             ms_pose if ms_pose is not None else 0.0])
    frames_total = int(grp.sum_over_ranks([args.steps * B])[0])
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        ab = algorithmic_bytes_per_image()
        kernels = {k: stage.get(k, 0.0) for k in ("pyramid", "fast", "select1", "harris", "select2", "blur", "describe", "stereo", "match")}
        nimg = 2 * B
        npv = float(n_prev.mean())
        # pass-2 rows k_shortlist really scans: rows linked to a pass-1 row are skipped or served from the pass-1
        # distance matrix (k_reuse); in this workload those are exactly the rows with map_prev_row >= 0
        need_rows = float((h_mpr < 0).sum()) / P
        # single kernels bracketed by their own events on the lane's stream (DESIGN.md section 4 lists the bytes); the
        # second name is the kernel's key in profiles/traffic.json (ncu: DRAM bytes and warp instructions per launch)
        single = {"k_fast": ("fast", ab["fast"] * nimg, "k_fast"), "k_blur": ("blur", ab["blur"] * nimg, "k_blur"),
                  "k_describe": ("describe", ab["describe"] * nimg, "k_describe"), "k_harris": ("harris", ab["harris"] * nimg, "k_harris"),
                  # tensor-core Hamming tiles: descriptors in as operand images (256 B per descriptor), lists / minima out
                  "k_tc_hamming<PAIRS> (BF + pass-1 distances)": ("k_pairs", B * (npv + NFEAT) * 256, "k_tc_hamming<0>"),
                  "k_tc_hamming<SHORT> (pass-2 short lists)": ("k_shortlist2", B * (MAP_ROWS + free_cols) * 256, "k_tc_hamming<2>")}
        counts, ceil = {}, {}
        try:
            counts = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ceil = json.load(open(os.path.join(ROOT, "profiles", "r2_ceilings.json")))
        except Exception:
            pass
        headline_cfg = (W_IMG, H_IMG, NFEAT, MAP_ROWS, B, args.distribution) == (1241, 376, 2000, 5000, 32, "retainbest")
        # the dominant kernel = the largest share of the step when every kernel runs alone (the ncu launch list); the live
        # event brackets are taken while other lanes' kernels share the SMs and the tensor-core brackets include the
        # operand expansion, so they rank kernels less reliably
        def kcount(name):
            # traffic.json keys are ncu's kernel names: an exact key, else the first one that extends it (k_harris -> k_harris4,
            # k_blur -> k_blur<(bool)1>)
            ks = counts.get("kernels", {})
            if name in ks:
                return ks[name]
            return next((v for k, v in ks.items() if k.startswith(name) and "#" not in k), {})
        alone = {k: kcount(v[2]).get("ncu_duration_us", 0.0) for k, v in single.items()} if headline_cfg else {}
        dom = max(single, key=lambda k: (alone.get(k, 0.0), stage.get(single[k][0], 0.0)))
        dur_ms = stage.get(single[dom][0], 0.0)
        alg_bytes = float(single[dom][1])
        achieved = alg_bytes / (dur_ms * 1e-3) / 1e9 if dur_ms > 0 else None
        traffic = None
        kc = kcount(single[dom][2]) if headline_cfg else {}
        traffic = kc.get("dram_bytes_per_launch")
        # What binds the path is instruction issue, not bytes: the kernel's warp instructions per launch
        # (ncu, a property of the workload) over its live launch time against the MEASURED issue ceiling of this GPU
        # (tools/ceilings.cu: 8 independent integer chains per thread, every SM full)
        issue = None
        peak_issue = ceil.get("issue_warp_inst_per_s")
        if kc.get("warp_inst_per_launch") and peak_issue and dur_ms > 0:
            a_i = kc["warp_inst_per_launch"] / (dur_ms * 1e-3)
            issue = {"bound": "issue", "warp_inst_per_launch": kc["warp_inst_per_launch"], "achieved_ginst_s": a_i / 1e9,
                     "peak_ginst_s": peak_issue / 1e9, "frac": a_i / peak_issue,
                     "peak_source": "measured on this GPU model (tools/ceilings.cu -> profiles/r2_ceilings.json), %.2f warp instructions/clk/SM"
                                    % ceil.get("issue_per_sm_per_clk_at_max_clock", 0.0),
                     "alone_under_ncu": {"duration_us": kc.get("ncu_duration_us"), "issue_active_pct": kc.get("issue_active_pct"),
                                         "pipe_alu_pct": kc.get("pipe_alu_pct"), "pipe_fma_pct": kc.get("pipe_fma_pct"),
                                         "pipe_xu_pct": kc.get("pipe_xu_pct"), "pipe_lsu_pct": kc.get("pipe_lsu_pct")},
                     "note": "launch_ms is the kernel's time while the other lanes' kernels share the SMs; alone_under_ncu is the same "
                             "kernel alone (cold cache, serialised)"}
        step_issue = None
        if headline_cfg and counts.get("step_warp_inst") and peak_issue:
            a_s = counts["step_warp_inst"] / (ms_dev / args.steps * 1e-3)
            step_issue = {"warp_inst_per_step": counts["step_warp_inst"], "achieved_ginst_s": a_s / 1e9, "peak_ginst_s": peak_issue / 1e9,
                          "frac": a_s / peak_issue, "dram_bytes_per_step": counts.get("step_dram_bytes"),
                          "dram_gbs": counts.get("step_dram_bytes", 0) / (ms_dev / args.steps * 1e-3) / 1e9, "hbm_frac":
                          counts.get("step_dram_bytes", 0) / (ms_dev / args.steps * 1e-3) / 1e9 / peak,
                          "note": "all kernels of a 32-frame step (ncu: warp instructions and DRAM bytes) over the measured step time"}
        h2d = B * (2 * W_IMG * H_IMG + int(n_prev.mean()) * 33 + MAP_ROWS * 36 + 16)
        d2h = 2 * B * (K * 56 + 8) + B * (K * 21 + 4) + B * max(MAP_ROWS, K) * 14
        e2e_hc = {"value": frames_total / (ms_hc * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": ms_hc / args.steps, "wall_ms_per_step": wall_hc / args.steps,
                  "h2d_ms_per_step": stage_hc.get("h2d"), "d2h_ms_per_step": stage_hc.get("d2h"),
                  "lane_total_ms_per_step": stage_hc.get("total"),
                  "mode": "the caller chains frames through the host: previous-frame descriptors, liveness, the 5k-row map and its links "
                          "are uploaded with every frame (svo_frame_in.prev_desc / map_desc)"}
        e2e_main = e2e_hc
        if trk is not None:
            e2e_main = {"value": frames_total / (ms_trk_m * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": B * (2 * W_IMG * H_IMG + 16) + B * 200,
                        "d2h_bytes_per_step": B * ((NFEAT + 32) * 106 + MAP_ROWS + 40),
                        "ms_per_step": ms_trk_m / args.steps, "wall_ms_per_step": wall_trk / args.steps,
                        "h2d_ms_per_step": stage_trk.get("h2d"), "d2h_ms_per_step": stage_trk.get("d2h"),
                        "lane_total_ms_per_step": stage_trk.get("total"),
                        "mode": "device-resident tracker state (svo_track_*, svo_frame_in.track_seq): pinned host images in, every result "
                                "of the reference's frame / pnpmatch out (compact copies, right-image features stay on the device: "
                                "svo_set_outputs); the last frame's descriptors and map points and the local map are advanced in HBM as "
                                "src/Tracking.cc:237-250 does on the host (%d sequences, one frame of each per step; mean %.0f pass-1 rows, "
                                "%.0f local-map rows of which %d ballast)" % (trk["sequences"], trk["mean_prev_rows"], trk["mean_map_rows"],
                                                                              trk["ballast_rows"])}
        out = {
            "metric": METRIC, "value": frames_total / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (f32 for response, angle, sub-pixel)",
            "data": "synthetic",
            "config": {"workload": workload_string(args.distribution),
                       "frames_per_step": B, "lanes": args.lanes, "pool_frames": P,
                       "l2": "inputs larger than L2: %d-frame pool = %.0f MB of images + %.0f MB of descriptors cycled"
                             % (P, 2 * P * img_b / 1e6, P * (K * 33 + MAP_ROWS * 36) / 1e6),
                       "parallelism": "replicas only: one independent sequence per GPU, no collective",
                       "host_placement": ("process bound to the %d cores NVML reports local to its GPU" % len(near)) if near else "unbound"},
            "e2e": e2e_main,
            "e2e_host_chained": e2e_hc,
            "e2e_pose_inputs": None if trk is None else {
                "value": frames_total / (ms_pose_m * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * (2 * W_IMG * H_IMG + 16) + B * 200,
                "d2h_bytes_per_step": B * ((NFEAT + 32) * 48 + 40), "ms_per_step": ms_pose_m / args.steps,
                "mode": "tracked pass returning only what solvePnPRansac / PoseOptimization read (svo_set_outputs: SVO_OUT_POSE_INPUTS): "
                        "left keypoints, depth, claim_row, mp_create, mp_xyz; descriptors, BF matches, match_score and row flags stay in HBM"},
            "tracked": trk,
            "gpu_launches": launches,
            "verified_frames": verified,
            "verified_note": ("frames of the LAST timed batch of every lane (resident-input, host-chained and tracked passes, batch %d, "
                              "%d lanes, graph replay) checked bit for bit against the CPU oracle after the timed regions: keypoints, "
                              "descriptors, stereo, BF, pass 1, pass 2, and for tracked frames the state they left (%d claims among them)"
                              % (B, args.lanes, claims)) if verified else "verification off",
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "launch_ms": dur_ms, "algorithmic_bytes_per_launch": alg_bytes, "issue": issue, "step": step_issue,
                         "note": "this path is not HBM-bound: per-frame working sets are small and the dominant kernels are bound by "
                                 "instruction issue / the integer ALU pipe, so the HBM fraction is small by construction; the operative "
                                 "ceiling is roofline.issue (measured issue rate of the GPU), roofline.step gives both figures for the "
                                 "whole step (DESIGN.md section 4)"},
            "profiled_pass": {"steps": psteps, "ms_per_step": ms_prof / psteps,
                              "note": "stage_ms_per_step, kernel_ms_per_launch and roofline.launch_ms come from this pass"},
            "stage_ms_per_step": kernels, "kernel_ms_per_launch": {k: stage.get(v[0], 0.0) for k, v in single.items()}, "p50_ms_per_frame_single": p50, "p50_ms_per_frame_single_tracked": p50_trk,
            "p50_ms_per_frame_single_on_the_batch_context": p50_batch_ctx, "wall_ms_per_step": wall_dev / args.steps,
        }
        out["pose_stage"] = pose
        out["input_staging"] = staging
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(seq, cores=1, budget_s=args.cpu_seconds)
    ctx.close()
    grp.close()
    return out


# ------------------------------------------------------------------------------------------
def _cpu_worker_init():
    global _O, _cv2, _NAT, _GEOM
    from oracle import oracle as _O  # noqa: F401  (bench's cpu legs may execute oracle/)
    _NAT = _O.native_matchers()      # svo_matchers.c built on this host with -O3 -march=native
    _GEOM = _O.geometry(W_IMG, H_IMG, NLEVELS, 1.2, NFEAT)
    try:
        import cv2 as _cv2
        _cv2.setNumThreads(1)
    except Exception:
        _cv2 = None


CPU_STAGES = ("orb_x2", "pyramids_x2", "stereo", "bf", "pass1", "pass2")


def _cpu_pyramid(img):
    """Un-blurred levels for the sparse-stereo stage: chained INTER_LINEAR_EXACT resizes, byte-identical to the
    oracle's (tests/test_oracle_vs_cv2.py) — cv::ORB builds the same levels internally but does not hand them out."""
    lw, lh, ls, _ = _GEOM
    levels = [img]
    for l in range(1, NLEVELS):
        levels.append(_cv2.resize(levels[-1], (int(lw[l]), int(lh[l])), interpolation=_cv2.INTER_LINEAR_EXACT))
    return _O.pyramid_from_levels(levels, ls)


def _cpu_frame(job):
    """The reference's per-frame CPU front-end on one core (GUI, drawing, sleeps and per-row vector copies of
    pnpmatch.cc removed): cv::ORB on L and R (src/frame.cc:77), sparse stereo, BFMatcher +
    greedy pass 1 + pass 2 (src/pnpmatch.cc, -O3 -march=native).  Returns the per-stage seconds."""
    L, R, prev_desc, prev_live, map_desc = job
    t = [time.perf_counter()]
    if _cv2 is not None:
        orb = _cv2.ORB_create(nfeatures=NFEAT, scaleFactor=1.2, nlevels=NLEVELS)
        out = []
        for img in (L, R):
            kp, desc = orb.detectAndCompute(img, None)
            out.append((np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=_O.KP_DTYPE), desc))
        (kl, dl), (kr, dr) = out
        t.append(time.perf_counter())
        pl, pr = _cpu_pyramid(L), _cpu_pyramid(R)
        t.append(time.perf_counter())
        _O.stereo_sparse(kl, dl, pl, kr, dr, pr, BF, BASELINE)
    else:   # no cv2 on this host: the oracle's own extractor (slower; says so in `sample`)
        kl, dl, pl = _O.orb(L, NFEAT, with_pyramid=True)
        kr, dr, pr = _O.orb(R, NFEAT, with_pyramid=True)
        t.append(time.perf_counter()); t.append(time.perf_counter())
        _O.stereo_sparse(kl, dl, pl, kr, dr, pr, BF, BASELINE)
        _O.pyramid_free(pl); _O.pyramid_free(pr)
    t.append(time.perf_counter())
    _O.match_bf(dl, prev_desc, L=_NAT)
    t.append(time.perf_counter())
    p1 = _O.match_greedy(prev_desc, dl, 0, row_live=prev_live, L=_NAT)
    t.append(time.perf_counter())
    _O.match_greedy(map_desc, dl, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_base=len(prev_desc), L=_NAT)
    t.append(time.perf_counter())
    return [t[i + 1] - t[i] for i in range(6)]


def _cpu_jobs(seq, n):
    _cpu_worker_init()
    rng = np.random.default_rng(77)
    frames = [seq.frame(t) for t in range(n + 1)]
    descs = []
    for f in frames:
        if _cv2 is not None:
            descs.append(_cv2.ORB_create(nfeatures=NFEAT, scaleFactor=1.2, nlevels=NLEVELS).detectAndCompute(f[0], None)[1])
        else:
            descs.append(_O.orb(f[0], NFEAT)[1])
    jobs = []
    for t in range(1, n + 1):
        prev = [descs[max(t - a, 0)] for a in range(1, 5)]
        mp, _ = build_map(prev, None, rng)
        jobs.append((frames[t][0], frames[t][1], prev[0], np.ones(len(prev[0]), np.uint8), mp))
    return jobs


def _cpu_sample_text():
    return ("%s ORB x2 (what src/frame.cc:77 calls) + chained cv2.resize pyramids + sparse stereo (oracle/, -O2) + BFMatcher / greedy "
            "pass 1 / pass 2 (oracle/svo_matchers.c, -O3 -march=native, popcnt)" % ("cv2 %s" % _cv2.__version__ if _cv2 else "oracle C"))


def cpu_baseline(seq, cores, budget_s):
    """Bounded sample of the same workload on the host (reported beside the GPU number; not the target)."""
    jobs = _cpu_jobs(seq, 4)
    t1 = sum(_cpu_frame(jobs[0]))                            # warm-up + per-frame cost estimate
    n = int(max(4, min(400, budget_s / max(t1, 1e-3))))
    acc = np.zeros(6)
    t0 = time.perf_counter()
    for i in range(n):
        acc += _cpu_frame(jobs[i % len(jobs)])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d frames of the same synthetic sequence on 1 core: %s" % (n, _cpu_sample_text()),
            "ms_per_frame": dt / n * 1e3, "stage_ms_per_frame": dict(zip(CPU_STAGES, (acc / n * 1e3).round(3).tolist()))}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU front-end on all host cores of the box (rank 0 only)."""
    if rank != 0:
        return None
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    seq = synth.Sequence((H_IMG, W_IMG), seed=0)
    jobs = _cpu_jobs(seq, 4)
    per_step = args.ref_frames or cores                        # a bounded sample: one frame per core per step
    cores = min(cores, per_step)
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init)
    work = [jobs[i % len(jobs)] for i in range(per_step)]
    steps, warm = args.steps, args.warmup                      # the driver's --steps / --warmup, as given
    for _ in range(warm):
        pool.map(_cpu_frame, work)
    acc = np.zeros(6)
    t0 = time.perf_counter()
    for _ in range(steps):
        for st in pool.map(_cpu_frame, work):
            acc += st
    dt = time.perf_counter() - t0
    pool.close()
    v = steps * per_step / dt
    sample = "%d steps x %d frames (one per host core) of the same synthetic sequence; %s" % (steps, per_step, _cpu_sample_text())
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32 (f32 for response, angle, sub-pixel)", "data": "synthetic",
            "config": {"workload": workload_string(), "frames_per_step": per_step,
                       "note": "each step is a bounded sample of the workload (one frame per host core), CPU only"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "stage_ms_per_frame": dict(zip(CPU_STAGES, (acc / (steps * per_step) * 1e3).round(3).tolist()))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="stereo frames per step")
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--pool", type=int, default=160, help="distinct synthetic frames cycled (must exceed L2)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tracked", action="store_true", help="skip the tracked end-to-end pass (e2e is then the host-chained one)")
    ap.add_argument("--verify", type=int, default=2, help="frames per lane of the last timed batch checked against the oracle (0: off)")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per step (default: one per host core)")
    # other BASELINE.json configurations (parity-test / breakdown cases, not the headline): e.g. configs[2]
    # `--features 4000`, configs[3] `--width 2560 --height 720 --features 8000 --batch 8 --pool 48`
    ap.add_argument("--width", type=int, default=1241)
    ap.add_argument("--height", type=int, default=376)
    ap.add_argument("--features", type=int, default=2000)
    ap.add_argument("--map-rows", type=int, default=5000)
    ap.add_argument("--distribution", default="retainbest", choices=["retainbest", "octree"],
                    help="keypoint selection: cv::ORB retainBest (headline, parity) or the opt-in quadtree distribution")
    args = ap.parse_args()
    globals().update(W_IMG=args.width, H_IMG=args.height, NFEAT=args.features, MAP_ROWS=args.map_rows,
                     DISTRIBUTION=1 if args.distribution == "octree" else 0)
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        _cpu_worker_init()
        out = run_reference(args, rank, world)
    else:
        out = run_gpu(args, rank, world, local_rank)
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
