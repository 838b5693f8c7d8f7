// track_loop_test.cc — the host loop of INTEGRATION.md section 3a, compiled: what Tracking::Track does per stereo pair
// (src/Tracking.cc:184-250) when lastframe / LocalMapPoints live in HBM.  Plain C++ over the C ABI (no cv:: types):
// per frame, upload the pair with svo_frame_in.track_seq, read the matches and the points the keypoints own, keep the
// poses on the host.  tests/test_gpu_adapter.py runs it on a synthetic sequence and compares the dump with
// oracle/track.py (pinned to the reference's own code by tests/test_oracle_track.py).
//   track_loop_test W H nfeatures nframes frames.raw boxes.txt out.txt
// frames.raw: nframes x (left, right) gray images back to back; boxes.txt: "frame left right top bottom" per line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "svo_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, svo_last_error(ctx)); return 1; } } while (0)

int main(int argc, char **argv)
{
    if (argc != 8) { fprintf(stderr, "usage: %s W H nfeatures nframes frames.raw boxes.txt out.txt\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]), nf = atoi(argv[3]), T = atoi(argv[4]);
    std::vector<uint8_t> img((size_t)T * 2 * W * H);
    {
        FILE *f = fopen(argv[5], "rb");
        if (!f || fread(img.data(), 1, img.size(), f) != img.size()) { fprintf(stderr, "cannot read %s\n", argv[5]); return 2; }
        fclose(f);
    }
    std::vector<std::vector<int32_t>> boxes((size_t)T);
    {
        FILE *f = fopen(argv[6], "r");
        int t, b[4];
        while (f && fscanf(f, "%d %d %d %d %d", &t, &b[0], &b[1], &b[2], &b[3]) == 5)
            if (t >= 0 && t < T) boxes[(size_t)t].insert(boxes[(size_t)t].end(), b, b + 4);
        if (f) fclose(f);
    }
    svo_ctx *ctx = nullptr;
    svo_config cfg;
    svo_default_config(&cfg);
    cfg.width = W; cfg.height = H; cfg.nfeatures = nf; cfg.max_batch = 1; cfg.lanes = 1; cfg.max_rows = 3000;
    if (svo_create(&cfg, &ctx) != SVO_OK) { fprintf(stderr, "svo_create: %s\n", ctx ? svo_last_error(ctx) : "no context"); return 1; }
    CHECK(svo_track_create(ctx, 1, 3000, 4));                 // one sequence, the 4-frame window of src/Tracking.cc:239-250
    CHECK(svo_track_reset(ctx, 0, nullptr, 0));
    CHECK(svo_set_outputs(ctx, SVO_OUT_COMPACT | SVO_OUT_NO_RIGHT));
    const float fx = 707.0912f, fy = 707.0912f, cx = 601.8873f, cy = 183.1104f, bf = 379.8145f;
    // an F whose epipolar lines are (almost) the image rows; a real caller passes findFundamentalMat's (src/pnpmatch.cc:336)
    alignas(8) const double F[9] = {1.1e-9, 2.3e-7, -3.1e-4, -2.2e-7, 0.9e-9, 0.8312, 2.9e-4, -0.8297, 1.0};
    FILE *o = fopen(argv[7], "w");
    if (!o) return 2;
    for (int t = 0; t < T; ++t) {
        svo_frame_in in;
        memset(&in, 0, sizeof(in));
        in.left = img.data() + (size_t)(2 * t) * W * H; in.right = img.data() + (size_t)(2 * t + 1) * W * H;
        in.stride = W; in.channels = 1;
        in.bf = bf; in.baseline = bf / fx; in.fx = fx; in.fy = fy; in.cx = cx; in.cy = cy;
        in.track_seq = 1 + 0; in.frame_id = t;                              // Tracking::frame_num
        if (!boxes[(size_t)t].empty()) { in.boxes = boxes[(size_t)t].data(); in.n_boxes = (int)boxes[(size_t)t].size() / 4; in.F = F; }
        CHECK(svo_batch_submit(ctx, 0, &in, 1));
        CHECK(svo_batch_wait(ctx, 0));
        svo_frame_out r;
        CHECK(svo_batch_result(ctx, 0, 0, &r));
        if (r.status != SVO_OK) { fprintf(stderr, "frame %d: status %d\n", t, r.status); return 1; }
        fprintf(o, "frame %d %d %d %d %d\n", t, r.n_left, r.n_prev, r.n_map, r.n_stereo);
        int matched = 0;
        for (int j = 0; j < r.n_left; ++j) {
            // a matched keypoint is a 3D-2D pair for solvePnPRansac (src/pnpmatch.cc:215-227): world point =
            // Twc[mp_create] * mp_xyz, pixel = kp_left[j]
            matched += r.claim_row[j] >= 0;
            fprintf(o, "kp %d %d %.9g %.9g %.9g %d %d %.9g %.9g %.9g\n", t, j, r.kp_left[j].x, r.kp_left[j].y, r.depth[j], r.claim_row[j],
                    r.mp_create[j], r.mp_xyz[3 * j], r.mp_xyz[3 * j + 1], r.mp_xyz[3 * j + 2]);
        }
        for (int i = 0; i < r.n_prev; ++i)
            if (r.p1_row_bad && r.p1_row_bad[i]) fprintf(o, "bad %d %d\n", t, i);
        fprintf(o, "matched %d %d\n", t, matched);
    }
    fclose(o);
    svo_destroy(ctx);
    return 0;
}
