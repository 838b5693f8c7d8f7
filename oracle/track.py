"""TEST INFRASTRUCTURE — CPU restatement of what Tracking::Track carries from one frame to the next
(/root/reference/src/Tracking.cc:225-250), as the oracle of the device-resident tracker state (csrc/track.cu):

    pnpmatch::poseEstimationPnP(currentframe, lastframe, LocalMapPoints, ...)   src/Tracking.cc:114 -> src/pnpmatch.cc:33-251
        find_feature_matches (BF, last frame's OWN descriptors)                 src/pnpmatch.cc:253-300
        pass 1: LastFrame.MapPoints[i]->m_descriptor (frozen, src/mappoint.cc:12) with the YOLO/epipolar veto   :61-156
        pass 2: every point of LocalMapPoints not bad and not observed by the current frame                     :160-199
    lastframe = frame(currentframe); lastframe.createmappoint(LocalMapPoints)   src/Tracking.cc:237-238, src/frame.cc:182-238
    erase every point with create_id <= frame_num - 4 (when frame_num >= 4)     src/Tracking.cc:239-250

The scans themselves are oracle/svo_matchers.c (O.match_bf / O.match_greedy, pinned to the reference's own code by
tests/test_ref_pin.py); this file only restates the bookkeeping between them.  Pinned against the reference's own
classes driven for several frames (oracle/ref.py:run_sequence) by tests/test_oracle_track.py.

Scan order of pass 2.  The reference walks a std::set<mappoint*> in pointer order, which is whatever the allocator
produced.  The device keeps "survivors in their previous order, then the new points in keypoint order"; `reorder()`
lets the pin test impose the reference's own order before each frame so that everything else can be compared exactly.
"""
import numpy as np

from . import oracle as O

BALLAST = np.iinfo(np.int32).max


class Tracker:
    def __init__(self, window=4, ballast=None, map_cap=None):
        self.window = window
        self.map_cap = map_cap
        self.last_desc = np.zeros((0, 32), np.uint8)       # last frame's own descriptors
        self.prev_desc = np.zeros((0, 32), np.uint8)       # frozen descriptor of the point each last-frame keypoint owns
        self.prev_live = np.zeros(0, np.uint8)
        self.prev_map_row = np.zeros(0, np.int32)
        self.prev_create = np.zeros(0, np.int32)           # name of the owned point: (create_id, keypoint index there)
        self.prev_idx = np.zeros(0, np.int32)
        self.prev_xyz = np.zeros((0, 3), np.float32)
        self.prev_xy = np.zeros((0, 2), np.float32)
        nb = 0 if ballast is None else len(ballast)
        self.map_desc = np.zeros((nb, 32), np.uint8) if ballast is None else np.ascontiguousarray(ballast, np.uint8).reshape(-1, 32).copy()
        self.map_create = np.full(nb, BALLAST, np.int32)
        self.map_idx = np.arange(nb, dtype=np.int32)
        self.map_link = np.full(nb, -1, np.int32)
        self.map_xyz = np.zeros((nb, 3), np.float32)

    @classmethod
    def from_state(cls, st, window=4, map_cap=None):
        """A tracker continuing from a state dump (svo.Context.track_state); the points' keypoint-index names are unknown (-1)."""
        t = cls(window=window, map_cap=map_cap)
        for k in ("last_desc", "prev_desc", "prev_live", "prev_map_row", "prev_create", "prev_xyz", "prev_xy", "map_desc", "map_create",
                  "map_link", "map_xyz"):
            setattr(t, k, np.array(st[k]))
        t.prev_idx = np.full(len(t.prev_desc), -1, np.int32); t.map_idx = np.full(len(t.map_desc), -1, np.int32)
        return t

    # ------------------------------------------------------------------------------------------------------------
    def names(self):
        return list(zip(self.map_create.tolist(), self.map_idx.tolist()))

    def reorder(self, names):
        """Permute the local map into the given order of (create_id, idx) names (the reference's set order)."""
        pos = {nm: i for i, nm in enumerate(self.names())}
        assert len(pos) == len(self.map_create) and sorted(pos) == sorted(names), "not a permutation of the local map"
        perm = np.array([pos[nm] for nm in names], np.int64)
        inv = np.empty(len(perm), np.int32); inv[perm] = np.arange(len(perm), dtype=np.int32)
        self.map_desc = self.map_desc[perm]; self.map_create = self.map_create[perm]; self.map_idx = self.map_idx[perm]
        self.map_link = self.map_link[perm]; self.map_xyz = self.map_xyz[perm]
        has = self.prev_map_row >= 0
        self.prev_map_row = np.where(has, inv[np.where(has, self.prev_map_row, 0)], -1).astype(np.int32)

    # ------------------------------------------------------------------------------------------------------------
    def step(self, kp_xy, desc, depth, frame_id, boxes=None, F=None, K4=None):
        """One frame.  kp_xy (N,2) f32, desc (N,32) u8, depth (N,) f32 = depth at the keypoints; boxes (n,4) int
        left/right/top/bottom; F (3,3) f64 enables the pass-1 veto (with boxes); K4 = (fx, fy, cx, cy)."""
        kp_xy = np.ascontiguousarray(kp_xy, np.float32).reshape(-1, 2)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        depth = np.ascontiguousarray(depth, np.float32)
        N, npv, nm = len(desc), len(self.prev_desc), len(self.map_desc)
        boxes = np.zeros((0, 4), np.int32) if boxes is None else np.asarray(boxes, np.int32).reshape(-1, 4)
        out = dict(n_prev=npv, n_map=nm)
        # ---- find_feature_matches: query = current frame, train = the last frame's own descriptors
        if npv:
            out["bf_idx"], out["bf_dist"], out["bf_keep"] = O.match_bf(desc, self.last_desc)
        # ---- pass 1
        claimed = np.zeros(N, np.uint8); claim_row = np.full(N, -1, np.int32)
        row_claimed1 = np.zeros(npv, np.uint8); row_bad = np.zeros(npv, np.uint8)
        if npv:
            veto = None
            if len(boxes) and F is not None:
                veto = dict(boxes=boxes, F=F, row_xy=self.prev_xy, cur_xy=kp_xy)
            p1 = O.match_greedy(self.prev_desc, desc, 0, row_live=self.prev_live, veto=veto)
            claimed, claim_row, row_claimed1, row_bad = p1["claimed"], p1["claim_row"], p1["row_claimed"], p1["row_bad"]
            out.update(p1_best_idx=p1["best_idx"], p1_best=p1["best"], p1_second=p1["second"], p1_row_claimed=row_claimed1,
                       p1_row_bad=row_bad)
        # ---- pass 2: skip points the current frame already observes (pass-1 matches) and bad ones
        row_claimed2 = np.zeros(nm, np.uint8)
        if nm:
            live2 = np.ones(nm, np.uint8)
            lk = self.map_link
            has = (lk >= 0) & (lk < npv)
            if npv:
                l0 = np.where(has, lk, 0)
                live2[has & ((row_claimed1[l0] == 1) | (row_bad[l0] == 1))] = 0
            p2 = O.match_greedy(self.map_desc, desc, 1, claimed=claimed, claim_row=claim_row, row_live=live2, row_base=npv)
            claimed, claim_row, row_claimed2 = p2["claimed"], p2["claim_row"], p2["row_claimed"]
            out["p2_row_claimed"] = row_claimed2
        out["claim_row"] = claim_row
        # ---- which old map rows survive: not aged out, not turned bad by this frame's veto
        keep = self.map_create > frame_id - self.window
        bad_rows = self.prev_map_row[(row_bad == 1) & (self.prev_map_row >= 0)] if npv else np.zeros(0, np.int64)
        keep[bad_rows] = False
        remap = np.full(nm, -1, np.int32)
        remap[keep] = np.arange(int(keep.sum()), dtype=np.int32)
        nsurv = int(keep.sum())
        # ---- CurrentFrame->MapPoints after the matching and createmappoint
        by_p1 = (claim_row >= 0) & (claim_row < npv)
        by_p2 = claim_row >= npv
        inside = np.zeros(N, bool)
        for b in boxes:                                         # src/frame.cc:196-207: box grown by 5 px
            inside |= (kp_xy[:, 0] > b[0] - 5) & (kp_xy[:, 0] < b[1] + 5) & (kp_xy[:, 1] > b[2] - 5) & (kp_xy[:, 1] < b[3] + 5)
        create = ~by_p1 & ~by_p2 & (depth > 0) & ~inside
        new_row = nsurv + np.cumsum(create) - 1
        if self.map_cap is not None:
            new_row = np.where(new_row < self.map_cap, new_row, -1)
        mp_create = np.full(N, -1, np.int32); mp_idx = np.full(N, -1, np.int32); mp_xyz = np.zeros((N, 3), np.float32)
        owner = np.full(N, -1, np.int32); frozen = desc.copy()
        r1 = np.where(by_p1, claim_row, 0)
        if npv:
            mp_create[by_p1] = self.prev_create[r1[by_p1]]; mp_idx[by_p1] = self.prev_idx[r1[by_p1]]
            mp_xyz[by_p1] = self.prev_xyz[r1[by_p1]]; frozen[by_p1] = self.prev_desc[r1[by_p1]]
            pm = self.prev_map_row[r1]
            owner[by_p1] = np.where(pm[by_p1] >= 0, remap[np.where(pm[by_p1] >= 0, pm[by_p1], 0)], -1)
        r2 = np.where(by_p2, claim_row - npv, 0)
        if nm:
            mp_create[by_p2] = self.map_create[r2[by_p2]]; mp_idx[by_p2] = self.map_idx[r2[by_p2]]
            mp_xyz[by_p2] = self.map_xyz[r2[by_p2]]; frozen[by_p2] = self.map_desc[r2[by_p2]]
            owner[by_p2] = remap[r2[by_p2]]
        mp_create[create] = frame_id; mp_idx[create] = np.nonzero(create)[0]
        xyz_new = np.zeros((N, 3), np.float32)
        xyz_new[:, 2] = depth
        if K4 is not None:
            fx, fy, cx, cy = [np.float32(v) for v in K4]
            xyz_new[:, 0] = (kp_xy[:, 0] - cx) * depth * (np.float32(1) / fx)       # src/frame.cc:171-172
            xyz_new[:, 1] = (kp_xy[:, 1] - cy) * depth * (np.float32(1) / fy)
        mp_xyz[create] = xyz_new[create]
        owner[create] = new_row[create]
        out.update(mp_create=mp_create, mp_idx=mp_idx, mp_xyz=mp_xyz, created=int(create.sum()))
        # ---- the next frame's state
        n_new = int((create & (owner >= 0)).sum())
        total = nsurv + n_new
        md = np.zeros((total, 32), np.uint8); mc = np.zeros(total, np.int32); mi = np.zeros(total, np.int32)
        ml = np.full(total, -1, np.int32); mx = np.zeros((total, 3), np.float32)
        md[:nsurv] = self.map_desc[keep]; mc[:nsurv] = self.map_create[keep]; mi[:nsurv] = self.map_idx[keep]; mx[:nsurv] = self.map_xyz[keep]
        cn = create & (owner >= 0)
        md[owner[cn]] = desc[cn]; mc[owner[cn]] = frame_id; mi[owner[cn]] = np.nonzero(cn)[0]; mx[owner[cn]] = mp_xyz[cn]
        own = owner >= 0
        ml[owner[own]] = np.nonzero(own)[0]
        self.map_desc, self.map_create, self.map_idx, self.map_link, self.map_xyz = md, mc, mi, ml, mx
        self.last_desc = desc.copy(); self.prev_desc = frozen
        self.prev_live = (by_p1 | by_p2 | create).astype(np.uint8)
        self.prev_map_row = owner; self.prev_create = mp_create; self.prev_idx = mp_idx; self.prev_xyz = mp_xyz
        self.prev_xy = kp_xy.copy()
        return out
