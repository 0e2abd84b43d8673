/*
 * nms_ref_shim.cpp -- C entry point around the reference's OWN merge (nms/nms.h + its vendored Clipper), compiled
 * from the sources where they lie under /root/reference into oracle/_ref/libref_nms.so by oracle/Makefile.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for fots_b200_merge_candidates_host.  Nothing of the
 * reference is copied here; this file only includes nms.h, builds nms::Polygon objects from the candidate rows
 * exactly as nms/adaptor.cpp:101-113 does from its local variables, and calls nms::merge_iou (nms/adaptor.cpp:118),
 * then flattens the result like polys2floats (nms/adaptor.cpp:13-29).
 */
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "nms.h"      /* -I/root/reference/nms */

extern "C" int ref_merge_candidates(const int32_t* cand, int num_cand, int w, int h, float thr1, float thr2,
                                    float* boxes, int max_boxes) {
    std::vector<nms::Polygon> polys;
    polys.reserve(num_cand);
    for (int i = 0; i < num_cand; ++i) {
        const int32_t* r = cand + (size_t)i * 16;
        float score, probs[4];
        std::memcpy(&score, r + 8, 4);
        std::memcpy(probs, r + 9, 16);
        nms::Polygon p{{{r[0], r[1]}, {r[2], r[3]}, {r[4], r[5]}, {r[6], r[7]}}, score,
                       {probs[0], probs[1], probs[2], probs[3]}, r[13], r[14]};
        polys.push_back(p);
    }
    std::vector<int> poly_ptr((size_t)w * h, -1);           /* nms/__init__.py:26-27 */
    std::vector<nms::Polygon> out = nms::merge_iou(polys, poly_ptr.data(), w, h, thr1, thr2);
    int n = 0;
    for (auto& p : out) {
        if (n < max_boxes) {
            float* o = boxes + (size_t)n * 9;
            for (int k = 0; k < 4; ++k) { o[2 * k] = float(p.poly[k].X); o[2 * k + 1] = float(p.poly[k].Y); }
            o[8] = float(p.score);
        }
        ++n;
    }
    return n;
}
