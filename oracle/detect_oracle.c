/*
 * detect_oracle.c -- CPU restatement of the per-pixel quadrangle decode of the reference's NMS front end,
 * /root/reference/nms/adaptor.cpp:76-117 (threshold the score map; one quadrangle per positive pixel from the
 * four distances and (sin, cos); x10000 fixed point; raster order).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Checker for fots_b200_decode_candidates.
 *
 * The reference builds adaptor.cpp with `-O3` on baseline x86-64 (nms/Makefile:1): no FMA instructions, every
 * float operation rounded on its own -- this file is compiled with -ffp-contract=off and keeps the reference's
 * operation order.  Inputs here are the network's NCHW planes ([4,h,w] distances, [2,h,w] (sin,cos)); the reference
 * reads the same numbers from HWC numpy views (test.py:86-93).  The merge (nms/nms.h, Clipper) is out of scope.
 * Pinning: unpinned by reference fixtures (the reference has none for NMS); pinned only by construction against
 * adaptor.cpp's text.  Said so in DESIGN.md.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

/* cand: [max_rows][16] int32 as documented in include/fots_b200_pipeline.h.  Returns the number of positive pixels. */
int detect_oracle_decode(const float *segm, const float *rbox, const float *angle, int h, int w,
                         float segm_threshold, int max_rows, int32_t *cand) {
    const int hw = h * w;
    const float scale_factor = 4, precision = 10000;
    int n = 0;
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const int i = y * w + x;
            const float p = segm[i];
            if (!(p > segm_threshold)) continue;                       /* adaptor.cpp:81 */
            if (n < max_rows) {
                const float r0 = rbox[i], r1 = rbox[hw + i], r2 = rbox[2 * hw + i], r3 = rbox[3 * hw + i];
                const float angle_sin = angle[i], angle_cos = angle[hw + i];      /* :82-83 a[0], a[1] */
                const float xp = x + 0.25f, yp = y + 0.25f;                       /* :85-86 */
                const float pos_r_x = (xp - r2 * angle_cos) * scale_factor;       /* :88-91 */
                const float pos_r_y = (yp - r2 * angle_sin) * scale_factor;
                const float pos_r2_x = (xp + r3 * angle_cos) * scale_factor;
                const float pos_r2_y = (yp + r3 * angle_sin) * scale_factor;
                const float ph = 9, phx = 9;
                const float p_left = expf(-r2 / phx), p_top = expf(-r0 / ph);     /* :96-99 */
                const float p_right = expf(-r3 / phx), p_bt = expf(-r1 / ph);
                int32_t *o = cand + (long)n * 16;
                o[0] = (int32_t)(int64_t)roundf(precision * (pos_r_x - r1 * angle_sin * scale_factor));   /* :103-106 */
                o[1] = (int32_t)(int64_t)roundf(precision * (pos_r_y + r1 * angle_cos * scale_factor));
                o[2] = (int32_t)(int64_t)roundf(precision * (pos_r_x + r0 * angle_sin * scale_factor));
                o[3] = (int32_t)(int64_t)roundf(precision * (pos_r_y - r0 * angle_cos * scale_factor));
                o[4] = (int32_t)(int64_t)roundf(precision * (pos_r2_x + r0 * angle_sin * scale_factor));
                o[5] = (int32_t)(int64_t)roundf(precision * (pos_r2_y - r0 * angle_cos * scale_factor));
                o[6] = (int32_t)(int64_t)roundf(precision * (pos_r2_x - r1 * angle_sin * scale_factor));
                o[7] = (int32_t)(int64_t)roundf(precision * (pos_r2_y + r1 * angle_cos * scale_factor));
                o[8] = f2i(p);
                o[9] = f2i(p_left * p_bt); o[10] = f2i(p_left * p_top);           /* :108 */
                o[11] = f2i(p_right * p_top); o[12] = f2i(p_right * p_bt);
                o[13] = x; o[14] = y; o[15] = 0;
            }
            ++n;
        }
    }
    return n;
}
