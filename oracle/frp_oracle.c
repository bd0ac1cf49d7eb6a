/*
 * frp_oracle.c — plain-C CPU restatement of the recognition hot path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg; never by the product (libfrp_b200.so).
 *
 * What each function follows:
 *   orc_l2_distances   face_recognition.api.face_distance (face-recognition==1.3.0,
 *                      backend/requirements.txt:37) = np.linalg.norm(G - q, axis=1);
 *                      in-repo restatement backend/app/services/face_service.py:597.
 *   orc_topk           `distances <= tolerance` (face_service.py:411), stable sort by
 *                      distance (:431-432), `[:top_k]` (routes/face.py:692-693);
 *                      cosine: sim >= tau, descending (insightface compute_sim).
 *   orc_cos_scores     insightface ArcFaceONNX.compute_sim = dot/(|a||b|)  (third-party,
 *                      insightface==0.7.3, backend/requirements.txt:41).
 *   orc_warp_affine    cv::warpAffine 8UC3 INTER_LINEAR BORDER_CONSTANT(0) fixed-point
 *                      scheme (OpenCV imgwarp.cpp: AB_BITS=10, INTER_BITS=5,
 *                      INTER_REMAP_COEF_BITS=15) as called by insightface
 *                      face_align.norm_crop.
 *   orc_blob           cv2.dnn.blobFromImages(1/127.5, mean 127.5, swapRB) -> fp32 CHW.
 *
 * Pinning: l2/topk against vectors produced by the unmodified reference
 * (tests/golden/l2_c1.npz); warp/blob against the installed OpenCV
 * (tests/golden/align_small.npz); cosine: parity unpinned (no reference code).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API void orc_l2_distances(const double* G, int64_t n, int dim, const double* q, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        const double* g = G + i * dim;
        double acc = 0.0;
        for (int j = 0; j < dim; ++j) {
            const double d = g[j] - q[j];
            acc += d * d;
        }
        out[i] = sqrt(acc);
    }
}

ORC_API void orc_cos_scores(const float* G, int64_t n, int dim, const float* q, double* out) {
    double qn = 0.0;
    for (int j = 0; j < dim; ++j) qn += (double)q[j] * (double)q[j];
    qn = sqrt(qn);
    for (int64_t i = 0; i < n; ++i) {
        const float* g = G + i * dim;
        double dot = 0.0, gn = 0.0;
        for (int j = 0; j < dim; ++j) {
            dot += (double)g[j] * (double)q[j];
            gn += (double)g[j] * (double)g[j];
        }
        out[i] = dot / (sqrt(gn) * qn);
    }
}

/* threshold (optional) then top-k with ties by row index.  larger_is_better: 1 for
 * cosine, 0 for L2.  Returns the number of results written (<= k). */
ORC_API int orc_topk(const double* v, int64_t n, int larger_is_better, int thresholded, double tau,
                     int k, int64_t* out_idx, double* out_val) {
    int cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double x = v[i];
        if (x != x) continue;
        if (thresholded && (larger_is_better ? !(x >= tau) : !(x <= tau))) continue;
        /* insertion into the sorted prefix; strict comparison keeps earlier rows first on ties */
        int pos = cnt;
        while (pos > 0 && (larger_is_better ? (x > out_val[pos - 1]) : (x < out_val[pos - 1]))) --pos;
        if (pos >= k) continue;
        const int last = cnt < k ? cnt : k - 1;
        for (int j = last; j > pos; --j) { out_val[j] = out_val[j - 1]; out_idx[j] = out_idx[j - 1]; }
        out_val[pos] = x;
        out_idx[pos] = i;
        if (cnt < k) ++cnt;
    }
    return cnt;
}

static int64_t rint_i64(double x) { return (int64_t)llrint(x); } /* round half to even, like cvRound */

/* forward 2x3 matrix M (double) -> dst[size][size][3]; src is H x W x 3 with row stride `stride` */
ORC_API void orc_warp_affine(const uint8_t* src, int H, int W, int64_t stride, const double* M,
                             int size, uint8_t* dst) {
    double m[6];
    memcpy(m, M, sizeof(m));
    double D = m[0] * m[4] - m[1] * m[3];
    D = D != 0 ? 1.0 / D : 0.0;
    const double A11 = m[4] * D, A22 = m[0] * D;
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22;
    const double b1 = -m[0] * m[2] - m[1] * m[5];
    const double b2 = -m[3] * m[2] - m[4] * m[5];
    m[2] = b1; m[5] = b2;
    int64_t* adelta = (int64_t*)malloc(sizeof(int64_t) * size * 2);
    int64_t* bdelta = adelta + size;
    for (int x = 0; x < size; ++x) {
        adelta[x] = rint_i64(m[0] * x * 1024.0);
        bdelta[x] = rint_i64(m[3] * x * 1024.0);
    }
    for (int y = 0; y < size; ++y) {
        const int64_t X0 = rint_i64((m[1] * y + m[2]) * 1024.0) + 16;
        const int64_t Y0 = rint_i64((m[4] * y + m[5]) * 1024.0) + 16;
        for (int x = 0; x < size; ++x) {
            const int64_t X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            int64_t sx = X >> 5, sy = Y >> 5;
            if (sx < -32768) sx = -32768; if (sx > 32767) sx = 32767;
            if (sy < -32768) sy = -32768; if (sy > 32767) sy = 32767;
            const int fx = (int)(X & 31), fy = (int)(Y & 31);
            const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
            const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            for (int c = 0; c < 3; ++c) {
                int p00 = 0, p01 = 0, p10 = 0, p11 = 0;
                if (sy >= 0 && sy < H) {
                    if (sx >= 0 && sx < W) p00 = src[sy * stride + 3 * sx + c];
                    if (sx + 1 >= 0 && sx + 1 < W) p01 = src[sy * stride + 3 * (sx + 1) + c];
                }
                if (sy + 1 >= 0 && sy + 1 < H) {
                    if (sx >= 0 && sx < W) p10 = src[(sy + 1) * stride + 3 * sx + c];
                    if (sx + 1 >= 0 && sx + 1 < W) p11 = src[(sy + 1) * stride + 3 * (sx + 1) + c];
                }
                const int acc = p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11;
                dst[((int64_t)y * size + x) * 3 + c] = (uint8_t)((acc + 16384) >> 15);
            }
        }
    }
    free(adelta);
}

/* aligned HWC BGR u8 -> CHW RGB fp32: (v - 127.5) * (float)(1/127.5) */
ORC_API void orc_blob(const uint8_t* aligned, int size, float* out) {
    const float scale = (float)(1.0 / 127.5);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < size * size; ++i)
            out[(int64_t)c * size * size + i] = ((float)aligned[(int64_t)i * 3 + (2 - c)] - 127.5f) * scale;
}
