// adapter_track_test.cc — drives the frame / pnpmatch drop-in through the SAME calls, in the same order, as
// Tracking::Track and Tracking::Tracklastframe make on the reference (src/Tracking.cc:184-250, :114), with non-empty
// offline YOLO boxes, and dumps everything the tracker reads back.  tests/test_gpu_adapter.py compares the dump with
// tests/golden/ref_track_*.npz, recorded from the reference's own compiled code (oracle/_ref), and with the oracle.
//   adapter_track_test W H nfeatures L0 R0 L1 R1 disp0.f32 disp1.f32 boxes.txt F.txt out.txt
// disp*.f32: the dense CV_32F disparity image standing in for frame::MB's output (:226); boxes.txt: "left right top
// bottom" per line; F.txt: the 9 doubles findFundamentalMat returned in the reference run (the hook below hands them
// back after dumping the point lists it was given, so those are compared too).
#include <cstdio>
#include <cstdlib>
#include <map>
#include "pnpmatch.h"

static cv::Mat load_raw(const char *path, int w, int h, int type, size_t esz)
{
    cv::Mat m(h, w, type);
    FILE *f = fopen(path, "rb");
    if (!f || fread(m.data, esz, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

int main(int argc, char **argv)
{
    if (argc != 13) { fprintf(stderr, "usage: %s W H nfeatures L0 R0 L1 R1 disp0 disp1 boxes F out\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]), nf = atoi(argv[3]);
    try {
        frame::configure(nf);
        cv::Mat K(3, 3, CV_32F, 0.0);
        K.at<float>(0, 0) = 707.0912f; K.at<float>(1, 1) = 707.0912f; K.at<float>(0, 2) = 601.8873f; K.at<float>(1, 2) = 183.1104f;
        K.at<float>(2, 2) = 1.f;
        float bf = 379.8145f;
        std::vector<std::vector<int>> boxes;
        {
            FILE *f = fopen(argv[10], "r");
            int b[4];
            while (f && fscanf(f, "%d %d %d %d", &b[0], &b[1], &b[2], &b[3]) == 4) boxes.push_back(std::vector<int>(b, b + 4));
            if (f) fclose(f);
        }
        double Fv[9];
        bool haveF = false;
        {
            FILE *f = fopen(argv[11], "r");
            haveF = f && fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", &Fv[0], &Fv[1], &Fv[2], &Fv[3], &Fv[4], &Fv[5], &Fv[6], &Fv[7], &Fv[8]) == 9;
            if (f) fclose(f);
        }
        FILE *o = fopen(argv[12], "w");
        pnpmatch::fundamental_solver = [&](const std::vector<cv::Point2f> &p1, const std::vector<cv::Point2f> &p2) {
            for (size_t i = 0; i < p1.size(); ++i) fprintf(o, "fpt %zu %.9g %.9g %.9g %.9g\n", i, p1[i].x, p1[i].y, p2[i].x, p2[i].y);
            cv::Mat F;
            if (haveF) { F = cv::Mat(3, 3, CV_64F); for (int i = 0; i < 9; ++i) F.at<double>(i / 3, i % 3) = Fv[i]; }
            return F;
        };
        cv::Mat L0 = load_raw(argv[4], W, H, CV_8UC1, 1), R0 = load_raw(argv[5], W, H, CV_8UC1, 1);
        cv::Mat L1 = load_raw(argv[6], W, H, CV_8UC1, 1), R1 = load_raw(argv[7], W, H, CV_8UC1, 1);
        cv::Mat D0 = load_raw(argv[8], W, H, CV_32F, 4), D1 = load_raw(argv[9], W, H, CV_32F, 4);
        cv::Mat none;
        double t0 = 0.0, t1 = 0.1;
        std::set<mappoint *> localmap;
        // Tracking::Track, first frame (src/Tracking.cc:184, :225-238)
        frame *f0 = new frame(L0, R0, none, L0, t0, K, bf, boxes);
        f0->id = 0;
        f0->featuredetect(f0->leftimg);
        f0->dispimg = D0;
        f0->computekeypoint_r();
        f0->disp2Depth(f0->bf);
        for (size_t i = 0; i < f0->keypoints_l.size(); ++i) {
            const cv::KeyPoint &k = f0->keypoints_l[i];
            fprintf(o, "f0 %zu %.9g %.9g %.9g %.9g\n", i, k.pt.x, k.pt.y, f0->keypoints_r[i].x, f0->depthimg.at<float>((int)k.pt.y, (int)k.pt.x));
        }
        frame last(f0);
        last.createmappoint(localmap);
        std::map<mappoint *, int> src;
        for (int i = 0; i < (int)last.MapPoints.size(); ++i) if (last.MapPoints[(size_t)i]) src[last.MapPoints[(size_t)i]] = i;
        int rank = 0;
        for (mappoint *mp : localmap)     // the set's own order: what pass 2 will walk
            fprintf(o, "map %d %d %.9g %.9g %.9g\n", rank++, src[mp], mp->worldpos.at<float>(0, 0), mp->worldpos.at<float>(1, 0), mp->worldpos.at<float>(2, 0));
        // second frame
        frame *f1 = new frame(L1, R1, none, L1, t1, K, bf, boxes);
        f1->id = 1;
        f1->featuredetect(f1->leftimg);
        f1->dispimg = D1;
        f1->computekeypoint_r();
        f1->disp2Depth(f1->bf);
        cv::Mat vel;
        const int ret = pnpmatch::poseEstimationPnP(f1, last, localmap, vel, K);      // src/Tracking.cc:114
        fprintf(o, "counts %d %d %d %d\n", (int)f0->keypoints_l.size(), (int)f1->keypoints_l.size(), (int)localmap.size(), ret);
        for (size_t j = 0; j < f1->keypoints_l.size(); ++j) {
            const cv::KeyPoint &k = f1->keypoints_l[j];
            fprintf(o, "kp %zu %.9g %.9g %d", j, k.pt.x, k.pt.y, f1->MapPoints[j] ? src[f1->MapPoints[j]] : -1);
            for (int b = 0; b < 32; ++b) fprintf(o, " %d", f1->f_descriptor.at<uint8_t>((int)j, b));
            fprintf(o, "\n");
        }
        for (size_t i = 0; i < last.MapPoints.size(); ++i)
            if (last.MapPoints[i]) fprintf(o, "row %zu %.9g %d\n", i, f1->match_score[i], last.MapPoints[i]->bad ? 1 : 0);
        fprintf(o, "pose");
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) fprintf(o, " %.9g", f1->Tcw.at<float>(r, c));
        fprintf(o, "\n");
        fclose(o);
        frame::shutdown();
    } catch (const std::exception &e) {
        fprintf(stderr, "adapter_track_test: %s\n", e.what());
        return 1;
    }
    return 0;
}
