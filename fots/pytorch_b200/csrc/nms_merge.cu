// nms_merge.cu -- HOST code (no kernels): the locality-aware merge + standard NMS that follows the per-pixel
// quadrangle decode in the reference's detector post-processing (nms/nms.h:149-213 merge_iou, :112-146
// standard_nms, :49-109 PolyMerger; called from nms/adaptor.cpp:118).  The reference runs it on the CPU over
// quadrangles it decoded on the CPU from three full maps; here the decode + compaction already happened on the GPU
// (fots_b200_decode_candidates) and this routine consumes the compact raster-ordered candidate rows after ONE
// device-to-host copy.  The algorithm is sequential by construction (every candidate is compared with the running
// result of the previous ones), so it stays host code, as SURVEY.md section 8f-2 prescribes.
//
// Written from the algorithm, not from the reference's text: plain arrays, no Clipper.  Polygon IoU: each
// quadrangle's even-odd region is split into two triangles (inner diagonal, or the two lobes of a self-intersecting
// one) and the intersection is the sum of four triangle-triangle clips (Sutherland-Hodgman in double precision on
// the x10000 fixed-point integer coordinates); the reference asks Clipper for the even-odd intersection and union,
// which is the same number up to Clipper's rounding of intersection vertices to integers (relative 1e-7).  Behaviours that are part of the observable result and are kept:
//   * merged coordinates are accumulated as int64 += (float)coordinate * weight, evaluated in fp32 and truncated
//     back to int64 on every addition (nms.h:62-75), and divided in fp32 (nms.h:90-97);
//   * the score of a merged polygon is the SUM of the scores (nms.h:77,100);
//   * every candidate after the first that starts a new polygon is appended twice (nms.h:199,203 -- both
//     emplace_back calls run); the duplicate is folded back by a later comparison and doubles that score;
//   * second stage: repeatedly take the highest-score polygon, fold every remaining polygon whose IoU with it
//     exceeds the threshold into it, keep the rest in order (nms.h:122-139).
#include "../../../include/fots_b200_pipeline.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <atomic>
#include <thread>
#include <vector>

namespace {

struct Quad {
    int64_t x[4], y[4];
    float score;
    float w[4];           // edge confidences p_left*p_bt, p_left*p_top, p_right*p_top, p_right*p_bt
    int px, py;           // pixel that produced the candidate
};

double signed_area(const double* px, const double* py, int n) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        a += px[i] * py[j] - px[j] * py[i];
    }
    return 0.5 * a;
}

struct Tri { double x[3], y[3]; };

inline double cross(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// proper crossing of segments (a0,a1) and (b0,b1); writes the crossing point
bool segments_cross(double a0x, double a0y, double a1x, double a1y, double b0x, double b0y, double b1x, double b1y,
                    double* ox, double* oy) {
    const double d = cross(a1x - a0x, a1y - a0y, b1x - b0x, b1y - b0y);
    if (d == 0.0) return false;
    const double t = cross(b0x - a0x, b0y - a0y, b1x - b0x, b1y - b0y) / d;
    const double u = cross(b0x - a0x, b0y - a0y, a1x - a0x, a1y - a0y) / d;
    if (!(t > 0.0 && t < 1.0 && u > 0.0 && u < 1.0)) return false;
    *ox = a0x + t * (a1x - a0x);
    *oy = a0y + t * (a1y - a0y);
    return true;
}

// The region a quadrangle covers under the even-odd rule (what Clipper is asked for, nms.h:30-31) as two triangles
// with disjoint interiors: a convex or concave quadrangle is split along the diagonal that lies inside it, a
// self-intersecting one (merged polygons can be) into its two lobes.
void split(const Quad& q, Tri out[2]) {
    double x[4], y[4];
    for (int i = 0; i < 4; ++i) { x[i] = (double)q.x[i]; y[i] = (double)q.y[i]; }
    double cx, cy;
    if (segments_cross(x[0], y[0], x[1], y[1], x[2], y[2], x[3], y[3], &cx, &cy)) {          // edge 0 x edge 2
        out[0] = {{cx, x[1], x[2]}, {cy, y[1], y[2]}};
        out[1] = {{cx, x[3], x[0]}, {cy, y[3], y[0]}};
    } else if (segments_cross(x[1], y[1], x[2], y[2], x[3], y[3], x[0], y[0], &cx, &cy)) {   // edge 1 x edge 3
        out[0] = {{x[0], x[1], cx}, {y[0], y[1], cy}};
        out[1] = {{cx, x[2], x[3]}, {cy, y[2], y[3]}};
    } else {
        const double s1 = cross(x[2] - x[0], y[2] - y[0], x[1] - x[0], y[1] - y[0]);
        const double s3 = cross(x[2] - x[0], y[2] - y[0], x[3] - x[0], y[3] - y[0]);
        if ((s1 > 0.0) != (s3 > 0.0) || s1 == 0.0 || s3 == 0.0) {      // p1, p3 on opposite sides: diagonal p0-p2 is inside
            out[0] = {{x[0], x[1], x[2]}, {y[0], y[1], y[2]}};
            out[1] = {{x[0], x[2], x[3]}, {y[0], y[2], y[3]}};
        } else {                                                         // reflex corner at p1 or p3: diagonal p1-p3
            out[0] = {{x[0], x[1], x[3]}, {y[0], y[1], y[3]}};
            out[1] = {{x[1], x[2], x[3]}, {y[1], y[2], y[3]}};
        }
    }
}

// area of the intersection of two triangles (Sutherland-Hodgman against the convex clip triangle)
double tri_intersection_area(const Tri& s, const Tri& c) {
    if (signed_area(c.x, c.y, 3) == 0.0) return 0.0;
    const double orient = signed_area(c.x, c.y, 3) > 0.0 ? 1.0 : -1.0;
    double ax[12], ay[12], bx[12], by[12];
    int n = 3;
    for (int i = 0; i < 3; ++i) { ax[i] = s.x[i]; ay[i] = s.y[i]; }
    for (int e = 0; e < 3 && n > 0; ++e) {
        const int e2 = (e + 1) % 3;
        const double ex = c.x[e2] - c.x[e], ey = c.y[e2] - c.y[e];
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double di = orient * (ex * (ay[i] - c.y[e]) - ey * (ax[i] - c.x[e]));   // >= 0: inside
            const double dj = orient * (ex * (ay[j] - c.y[e]) - ey * (ax[j] - c.x[e]));
            if (di >= 0.0) { bx[m] = ax[i]; by[m] = ay[i]; ++m; }
            if ((di >= 0.0) != (dj >= 0.0)) {
                const double t = di / (di - dj);
                bx[m] = ax[i] + t * (ax[j] - ax[i]);
                by[m] = ay[i] + t * (ay[j] - ay[i]);
                ++m;
            }
        }
        n = m;
        std::memcpy(ax, bx, sizeof(double) * n);
        std::memcpy(ay, by, sizeof(double) * n);
    }
    return n >= 3 ? std::fabs(signed_area(ax, ay, n)) : 0.0;
}

// Strictly convex quadrangle (all four turns the same way, none degenerate)?  Returns the orientation sign, 0 if not.
int convex_orientation(const double* x, const double* y) {
    int pos = 0, neg = 0;
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        const double c = cross(x[j] - x[i], y[j] - y[i], x[k] - x[j], y[k] - y[j]);
        if (c > 0.0) ++pos; else if (c < 0.0) ++neg; else return 0;
    }
    return pos == 4 ? 1 : neg == 4 ? -1 : 0;
}

// Area of the intersection of two strictly convex quadrangles: Sutherland-Hodgman of `s` against the four edges of `c`.
double convex_quad_intersection_area(const double* sx, const double* sy, const double* cx, const double* cy, int c_orient) {
    double bufx[2][16], bufy[2][16];                 // ping-pong: the clipped polygon of edge e is the input of edge e + 1
    int n = 4, cur = 0;
    for (int i = 0; i < 4; ++i) { bufx[0][i] = sx[i]; bufy[0][i] = sy[i]; }
    const double orient = (double)c_orient;
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double* ax = bufx[cur];
        const double* ay = bufy[cur];
        double* bx = bufx[cur ^ 1];
        double* by = bufy[cur ^ 1];
        const int e2 = (e + 1) & 3;
        const double ex = cx[e2] - cx[e], ey = cy[e2] - cy[e];
        int m = 0;
        double di = orient * (ex * (ay[0] - cy[e]) - ey * (ax[0] - cx[e]));
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double dj = orient * (ex * (ay[j] - cy[e]) - ey * (ax[j] - cx[e]));
            if (di >= 0.0) { bx[m] = ax[i]; by[m] = ay[i]; ++m; }
            if ((di >= 0.0) != (dj >= 0.0)) {
                const double t = di / (di - dj);
                bx[m] = ax[i] + t * (ax[j] - ax[i]);
                by[m] = ay[i] + t * (ay[j] - ay[i]);
                ++m;
            }
            di = dj;
        }
        n = m;
        cur ^= 1;
    }
    return n >= 3 ? std::fabs(signed_area(bufx[cur], bufy[cur], n)) : 0.0;
}

// Axis-aligned bounding box of a quadrangle, cached next to the running polygons: most comparisons of both stages (the
// polygon appended last in raster order; every remaining polygon against the current maximum in the second stage) are
// between far-apart boxes, and recomputing two bounding boxes per comparison was a quarter of the merge.
struct Box { int64_t lx, hx, ly, hy; };
inline Box box_of(const Quad& q) {
    Box b = {q.x[0], q.x[0], q.y[0], q.y[0]};
    for (int i = 1; i < 4; ++i) {
        b.lx = std::min(b.lx, q.x[i]); b.hx = std::max(b.hx, q.x[i]); b.ly = std::min(b.ly, q.y[i]); b.hy = std::max(b.hy, q.y[i]);
    }
    return b;
}
inline bool disjoint(const Box& a, const Box& b) { return a.hx < b.lx || b.hx < a.lx || a.hy < b.ly || b.hy < a.ly; }

float quad_iou_overlapping(const Quad& a, const Quad& b);

// Disjoint bounding boxes: the intersection is empty and the quotient is exactly 0 (never above a threshold).
inline float quad_iou(const Quad& a, const Box& ba, const Quad& b, const Box& bb) {
    return disjoint(ba, bb) ? 0.0f : quad_iou_overlapping(a, b);
}

float quad_iou_overlapping(const Quad& a, const Quad& b) {
    // Both strictly convex (decoded rectangles and almost all of their weighted means are): one quad-quad clip instead
    // of two splits and four triangle clips.  Same area up to double rounding.
    {
        double ax[4], ay[4], bx[4], by[4];
        for (int i = 0; i < 4; ++i) { ax[i] = (double)a.x[i]; ay[i] = (double)a.y[i]; bx[i] = (double)b.x[i]; by[i] = (double)b.y[i]; }
        const int oa = convex_orientation(ax, ay), ob = convex_orientation(bx, by);
        if (oa != 0 && ob != 0) {
            const double inter = convex_quad_intersection_area(ax, ay, bx, by, ob);
            const double uni = std::fabs(signed_area(ax, ay, 4)) + std::fabs(signed_area(bx, by, 4)) - inter;
            return std::fabs((float)inter) / std::max(std::fabs((float)uni), 1.0f);
        }
    }
    Tri ta[2], tb[2];
    split(a, ta);
    split(b, tb);
    double inter = 0.0, area_a = 0.0, area_b = 0.0;
    for (int i = 0; i < 2; ++i) {
        area_a += std::fabs(signed_area(ta[i].x, ta[i].y, 3));
        area_b += std::fabs(signed_area(tb[i].x, tb[i].y, 3));
        for (int j = 0; j < 2; ++j) inter += tri_intersection_area(ta[i], tb[j]);
    }
    const double uni = area_a + area_b - inter;
    // nms.h:21-36: areas are summed into floats; union clamped to >= 1
    return std::fabs((float)inter) / std::max(std::fabs((float)uni), 1.0f);
}

// fold `first` then `second` (nms.h:62-105); every += on the int64 sums goes through fp32, like the reference's
// `std::int64_t += cInt * float`
Quad fold(const Quad& first, const Quad& second) {
    static const int xw[4] = {0, 0, 2, 2};    // weight index used by x of corner k
    static const int yw[4] = {3, 1, 1, 3};    // ... and by y
    int64_t sx[4] = {0, 0, 0, 0}, sy[4] = {0, 0, 0, 0};
    float score = 0.0f, w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const Quad* both[2] = {&first, &second};
    for (const Quad* q : both) {
        for (int k = 0; k < 4; ++k) {
            const volatile float tx = (float)q->x[k] * q->w[xw[k]];
            const volatile float ty = (float)q->y[k] * q->w[yw[k]];
            sx[k] = (int64_t)((float)sx[k] + tx);
            sy[k] = (int64_t)((float)sy[k] + ty);
        }
        score += q->score;
        for (int k = 0; k < 4; ++k) w[k] += q->w[k];
    }
    Quad r;
    for (int k = 0; k < 4; ++k) {
        r.x[k] = (int64_t)((float)sx[k] / w[xw[k]]);
        r.y[k] = (int64_t)((float)sy[k] / w[yw[k]]);
        r.w[k] = w[k];
    }
    r.score = score;
    r.px = 0; r.py = 0;
    return r;
}

}  // namespace

extern "C" int fots_b200_merge_candidates_host(const int* cand, int num_cand, int w, int h, float iou_threshold1,
                                               float iou_threshold2, float* boxes, int max_boxes, int* num_boxes) {
    if (num_cand < 0 || w <= 0 || h <= 0 || !num_boxes || (num_cand > 0 && !cand) || max_boxes < 0 || (max_boxes > 0 && !boxes))
        return RROI_B200_ERR_INVALID_ARG;
    std::vector<Quad> acc;                       // first stage: running polygons
    std::vector<Box> abox;                       // ... and their bounding boxes
    std::vector<int> owner((size_t)w * h, -1);   // pixel -> index of the polygon it was folded into
    acc.reserve(256);
    abox.reserve(256);
    for (int i = 0; i < num_cand; ++i) {
        const int* row = cand + (size_t)i * 16;
        Quad q;
        for (int k = 0; k < 4; ++k) { q.x[k] = row[2 * k]; q.y[k] = row[2 * k + 1]; }
        std::memcpy(&q.score, row + 8, 4);
        std::memcpy(q.w, row + 9, 16);
        q.px = row[13]; q.py = row[14];
        if (q.px < 0 || q.px >= w || q.py < 0 || q.py >= h) return RROI_B200_ERR_INVALID_ARG;
        const size_t pix = (size_t)q.py * w + q.px;
        const Box qb = box_of(q);
        if (acc.empty()) {
            acc.push_back(q);
            abox.push_back(qb);
            owner[pix] = 0;
            continue;
        }
        // 1. the polygon touched last (the left neighbour in raster order, usually)
        int target = -1;
        if (quad_iou(q, qb, acc.back(), abox.back()) > iou_threshold1) {
            target = (int)acc.size() - 1;
        } else {
            if (q.py > 0) {
                // 2. the polygons of the pixels above.  The chain is nested exactly like the reference's: up-left and
                //    up-right are only looked at when the pixel straight above belongs to a polygon.
                const int up = owner[pix - w];
                if (up >= 0) {
                    if (quad_iou(q, qb, acc[up], abox[up]) > iou_threshold1) target = up;
                    if (target < 0 && q.px > 0) {
                        const int ul = owner[pix - w - 1];
                        if (ul >= 0 && quad_iou(q, qb, acc[ul], abox[ul]) > iou_threshold1) target = ul;
                    }
                    if (target < 0) {
                        // the reference reads poly_ptr[(y-1)*w + x + 1] without a bound check on x + 1; at the right
                        // edge that is the first pixel of the current row (same flat index) -- kept as is
                        const int ur = owner[pix - w + 1];
                        if (ur >= 0 && quad_iou(q, qb, acc[ur], abox[ur]) > iou_threshold1) target = ur;
                    }
                }
            }
            if (target < 0) { acc.push_back(q); abox.push_back(qb); }        // the reference's first (duplicate) append, nms.h:199
        }
        if (target >= 0) {
            acc[target] = fold(acc[target], q);
            abox[target] = box_of(acc[target]);
            owner[pix] = target;
        } else {
            acc.push_back(q);
            abox.push_back(qb);
            owner[pix] = (int)acc.size() - 1;
        }
    }

    // second stage (nms.h:112-146)
    const size_t n = acc.size();
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), (size_t)0);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return acc[a].score > acc[b].score; });
    int out = 0, total = 0;
    std::vector<size_t> kept;
    while (!order.empty()) {
        const size_t cur = order[0];
        size_t p = 0;
        for (size_t i = 1; i < order.size(); ++i) {
            if (quad_iou(acc[cur], abox[cur], acc[order[i]], abox[order[i]]) > iou_threshold2) {
                acc[cur] = fold(acc[order[i]], acc[cur]);
                abox[cur] = box_of(acc[cur]);
            } else {
                order[p++] = order[i];
            }
        }
        order.resize(p);
        kept.push_back(cur);
    }
    for (size_t idx : kept) {
        if (out < max_boxes) {
            float* o = boxes + (size_t)out * 9;
            for (int k = 0; k < 4; ++k) { o[2 * k] = (float)acc[idx].x[k]; o[2 * k + 1] = (float)acc[idx].y[k]; }
            o[8] = acc[idx].score;
            ++out;
        }
        ++total;
    }
    *num_boxes = total;
    return RROI_B200_OK;
}

// One call for a whole micro-batch: image b's candidates are cand[b * cap .. + counts[b]) (the layout
// fots_b200_decode_candidates writes); images are independent, so they are dealt to `threads` host threads (the per-rank
// worker pool of SURVEY.md section 8e).  boxes [B, max_boxes, 9], num_boxes [B].  Returns the first non-OK status.
extern "C" int fots_b200_merge_candidates_host_batch(const int* cand, const int* counts, int B, int cap, int w, int h,
                                                     float iou_threshold1, float iou_threshold2, float* boxes, int max_boxes,
                                                     int* num_boxes, int threads) {
    if (B < 0 || cap < 0 || !counts || !num_boxes || (B > 0 && max_boxes > 0 && !boxes)) return RROI_B200_ERR_INVALID_ARG;
    if (threads < 1) threads = 1;
    if (threads > B) threads = B;
    std::atomic<int> next(0), status(RROI_B200_OK);
    auto work = [&]() {
        for (int b = next.fetch_add(1); b < B; b = next.fetch_add(1)) {
            const int n = counts[b] < cap ? counts[b] : cap;
            const int st = fots_b200_merge_candidates_host(cand + (size_t)b * cap * 16, n < 0 ? 0 : n, w, h, iou_threshold1, iou_threshold2,
                                                           boxes + (size_t)b * max_boxes * 9, max_boxes, num_boxes + b);
            if (st != RROI_B200_OK) { int ok = RROI_B200_OK; status.compare_exchange_strong(ok, st); }
        }
    };
    if (threads <= 1) { work(); return status.load(); }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return status.load();
}
