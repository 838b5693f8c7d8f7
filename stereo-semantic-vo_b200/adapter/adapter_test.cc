// adapter_test.cc — exercises the frame / pnpmatch drop-in exactly as Tracking::Track does
// (src/Tracking.cc:225-238) on two stereo pairs given as raw 8-bit files, and dumps what the
// tracker would read back so tests/test_gpu_adapter.py can compare it with the oracle.
//   adapter_test W H nfeatures L0.raw R0.raw L1.raw R1.raw out.txt
#include <cstdio>
#include <cstdlib>
#include <map>
#include "pnpmatch.h"
#include "Optimizer.h"

static cv::Mat load_raw(const char *path, int w, int h)
{
    cv::Mat m(h, w, CV_8UC1);
    FILE *f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

static void run_frontend(frame &f, cv::Mat &L, cv::Mat &R)
{
    f.featuredetect(L);                 // Tracking.cc:225
    f.dispimg = f.MB(L, R);             // :226
    f.computekeypoint_r();              // :227
    f.disp2Depth(f.bf);                 // :228
}

int main(int argc, char **argv)
{
    if (argc != 9) { fprintf(stderr, "usage: %s W H nfeatures L0 R0 L1 R1 out\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]), nf = atoi(argv[3]);
    try {
        frame::configure(nf);
        cv::Mat K(3, 3, CV_32F, 0.0);
        K.at<float>(0, 0) = 707.0912f; K.at<float>(1, 1) = 707.0912f; K.at<float>(0, 2) = 601.8873f; K.at<float>(1, 2) = 183.1104f;
        K.at<float>(2, 2) = 1.f;
        float bf = 379.8145f;
        std::vector<std::vector<int>> boxes;
        cv::Mat L0 = load_raw(argv[4], W, H), R0 = load_raw(argv[5], W, H), L1 = load_raw(argv[6], W, H), R1 = load_raw(argv[7], W, H);
        cv::Mat none;
        double t0 = 0.0, t1 = 0.1;
        std::set<mappoint *> localmap;
        frame *f0 = new frame(L0, R0, none, L0, t0, K, bf, boxes);
        run_frontend(*f0, L0, R0);
        f0->id = 0;
        frame last(f0);                                   // Tracking.cc:237
        last.createmappoint(localmap);                    // :238
        frame *f1 = new frame(L1, R1, none, L1, t1, K, bf, boxes);
        run_frontend(*f1, L1, R1);
        f1->id = 1;
        cv::Mat vel, F;
        const int n1 = pnpmatch::match_last_frame(f1, last, F);
        // pass 2 against every map point of the last frame, in the set's own order
        std::map<mappoint *, int> src;
        for (int i = 0; i < (int)last.MapPoints.size(); ++i) if (last.MapPoints[(size_t)i]) src[last.MapPoints[(size_t)i]] = i;
        const int n2 = pnpmatch::match_local_map(f1, localmap);
        FILE *o = fopen(argv[8], "w");
        fprintf(o, "counts %d %d %d %d\n", (int)f0->keypoints_l.size(), (int)f1->keypoints_l.size(), n1, n2);
        for (size_t i = 0; i < f1->keypoints_l.size(); ++i) {
            const cv::KeyPoint &k = f1->keypoints_l[i];
            const float z = f1->depthimg.at<float>((int)k.pt.y, (int)k.pt.x);
            fprintf(o, "kp %zu %.9g %.9g %.9g %.9g %d %.9g %.9g %d", i, k.pt.x, k.pt.y, k.angle, k.response, k.octave,
                    f1->keypoints_r[i].x, z, f1->MapPoints[i] ? src[f1->MapPoints[i]] : -1);
            for (int b = 0; b < 32; ++b) fprintf(o, " %d", f1->f_descriptor.at<uint8_t>((int)i, b));
            fprintf(o, "\n");
        }
        int rank = 0;
        for (mappoint *mp : localmap) fprintf(o, "map %d %d\n", rank++, src[mp]);
        for (size_t i = 0; i < last.MapPoints.size(); ++i)
            if (last.MapPoints[i]) fprintf(o, "score %zu %.9g\n", i, f1->match_score[i]);
        // the pose stage as Tracking::Tracklastframe runs it (src/Tracking.cc:114-120): PnP RANSAC on the matched
        // 3D-2D pairs (src/pnpmatch.cc:211-247), SetPose(Tcl * Tcw), then the pose-only optimisation
        std::vector<cv::Mat> pts3d;
        std::vector<cv::Point2f> pts2d;
        for (size_t j = 0; j < f1->keypoints_l.size(); ++j)
            if (mappoint *mp = f1->MapPoints[j]) { pts2d.push_back(f1->keypoints_l[j].pt); pts3d.push_back(mp->worldpos); }
        for (size_t k = 0; k < pts2d.size(); ++k)
            fprintf(o, "pair %zu %.9g %.9g %.9g %.9g %.9g\n", k, pts3d[k].at<float>(0, 0), pts3d[k].at<float>(1, 0),
                    pts3d[k].at<float>(2, 0), pts2d[k].x, pts2d[k].y);
        cv::Mat Tcl;
        int inl = 0;
        const bool ok = pnpmatch::pnp_solver(pts3d, pts2d, K, Tcl, inl);
        fprintf(o, "pnp %d %d", ok ? 1 : 0, inl);
        if (ok) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) fprintf(o, " %.9g", Tcl.at<float>(r, c));
        fprintf(o, "\n");
        if (ok) f1->SetPose(Tcl);                           // f1->Tcw is the identity here, so Tcl * Tcw = Tcl
        const int nc = Optimizer::PoseOptimization(f1);
        fprintf(o, "lm %d", nc);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) fprintf(o, " %.9g", f1->Tcw.at<float>(r, c));
        fprintf(o, "\n");
        fclose(o);
        frame::shutdown();
    } catch (const std::exception &e) {
        fprintf(stderr, "adapter_test: %s\n", e.what());
        return 1;
    }
    return 0;
}
