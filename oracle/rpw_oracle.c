/* oracle/rpw_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's per-scan ground segmentation,
 * written from the reference's behaviour (not copied from it).  It exists to CHECK the CUDA
 * path; the product never calls it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.
 *
 * What it follows (RP = /root/reference/src/recursive_patchwork):
 *   filterGroundPoints            RP/src/recursive_patchwork.cpp:310-426
 *   cleanPoints                   RP/src/recursive_patchwork.cpp:19-35
 *   fitPlanePCA                   RP/src/recursive_patchwork.cpp:77-107
 *   fitPlaneAndSplit              RP/src/recursive_patchwork.cpp:109-308
 *   computeCentroid / Covariance  RP/src/point_cloud_processor.cpp:58-86
 *   cuda::ops CPU branches        RP/cuda/cuda_interface.cu:582-654
 *   PatchworkConfig defaults      RP/include/recursive_patchwork.hpp:25-36
 * Third-party arithmetic that is not under /root/reference:
 *   Eigen 3.4.0 (un-vendored, version unpinned by the reference; 3.4.0 is what ROS2 Humble
 *   ships) SelfAdjointEigenSolver<Matrix3f> and the 3-coefficient dot product: restated in
 *   eig3_f32()/dot3() below and, independently, in oracle/eigen_standin/Eigen/Dense.
 *   glibc 2.39 libm atan2f/powf/sqrtf: called directly.
 *
 * PARITY PIN.  The reference's own tests hold no golden vectors for this path,
 * so this restatement is pinned against the reference ITSELF: the reference's translation units
 * are compiled unmodified into oracle/_ref/libref_strict.so (oracle/Makefile) and
 * tests/test_oracle.py requires the two to return bit-identical ground / non-ground clouds on
 * every fixture cloud; tests/golden/ holds label fixtures generated from libref_strict.so by
 * tests/golden/make_golden.py.  All arithmetic here is float32 in the reference's evaluation
 * order; build with -fno-fast-math -ffp-contract=off.
 */
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include "rpw_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NUM_RINGS 8 /* RP/src/recursive_patchwork.cpp:345 */

void rpwo_default_config(rpwo_config* c) {
    /* RP/include/recursive_patchwork.hpp:25-36 */
    c->sensor_height = 1.2f;
    c->max_range = 150.0f;
    c->num_sectors = 10;
    c->max_iter = 100;
    c->adaptive_seed_height = 1;
    c->th_seeds = 0.15f;
    c->th_dist = 0.2f;
    c->th_outlier = 0.08f;
    c->filtering_radius = 150.0f;
    c->max_split_depth = 1000;
}

void rpwo_zone_model(const rpwo_config* cfg, float ring_edges[9], float* sector_angle) {
    /* :344-352.  std::pow(float,float) is powf; 2.0f*M_PI/num_sectors is a double expression
     * rounded once on assignment to float. */
    const float r_min = 1.0f, r_max = cfg->filtering_radius;
    for (int i = 0; i <= NUM_RINGS; ++i) ring_edges[i] = r_min * powf(r_max / r_min, (float)i / NUM_RINGS);
    *sector_angle = (float)((double)2.0f * M_PI / (double)cfg->num_sectors);
}

/* ------------------------------------------------------------------------------------------
 * Eigen 3.4.0 pieces, restated.
 * ---------------------------------------------------------------------------------------- */
static inline float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    /* fixed-size redux, fully unrolled: c0 + (c1 + c2) */
    const float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
    return p0 + (p1 + p2);
}

static void make_givens(float p, float q, float* c, float* s) {
    if (q == 0.f) {
        *c = p < 0.f ? -1.f : 1.f;
        *s = 0.f;
    } else if (p == 0.f) {
        *c = 0.f;
        *s = q < 0.f ? 1.f : -1.f;
    } else if (fabsf(p) > fabsf(q)) {
        float t = q / p;
        float u = sqrtf(1.f + t * t);
        if (p < 0.f) u = -u;
        *c = 1.f / u;
        *s = -t * *c;
    } else {
        float t = p / q;
        float u = sqrtf(1.f + t * t);
        if (q < 0.f) u = -u;
        *s = -1.f / u;
        *c = -t * *s;
    }
}

static float eigen_hypot(float x, float y) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float p = ax > ay ? ax : ay;
    if (p == 0.f) return 0.f;
    const float qp = (ay < ax ? ay : ax) / p;
    return p * sqrtf(1.f + qp * qp);
}

static void qr_step(float* diag, float* sub, int start, int end, float q[3][3]) {
    const float td = (diag[end - 1] - diag[end]) * 0.5f;
    const float e = sub[end - 1];
    float mu = diag[end];
    if (td == 0.f) {
        mu -= fabsf(e);
    } else if (e != 0.f) {
        const float e2 = e * e;
        const float h = eigen_hypot(td, e);
        if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
        else mu -= e2 / (td + (td > 0.f ? h : -h));
    }
    float x = diag[start] - mu;
    float z = sub[start];
    for (int k = start; k < end && z != 0.f; ++k) {
        float c, s;
        make_givens(x, z, &c, &s);
        const float sdk = s * diag[k] + c * sub[k];
        const float dkp1 = s * sub[k] + c * diag[k + 1];
        diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
        diag[k + 1] = s * sdk + c * dkp1;
        sub[k] = c * sdk - s * dkp1;
        if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
        x = sub[k];
        if (k < end - 1) {
            z = -s * sub[k + 1];
            sub[k + 1] = c * sub[k + 1];
        }
        if (!(c == 1.f && s == 0.f)) {
            for (int i = 0; i < 3; ++i) {
                const float xi = q[i][k], yi = q[i][k + 1];
                q[i][k] = c * xi - s * yi;
                q[i][k + 1] = s * xi + c * yi;
            }
        }
    }
}

int rpwo_eig3_f32(const float a[9], float evals[3], float evecs[9]) {
    float m00 = a[0], m10 = a[3], m11 = a[4], m20 = a[6], m21 = a[7], m22 = a[8];
    float scale = fabsf(m00);
    if (fabsf(m10) > scale) scale = fabsf(m10);
    if (fabsf(m11) > scale) scale = fabsf(m11);
    if (fabsf(m20) > scale) scale = fabsf(m20);
    if (fabsf(m21) > scale) scale = fabsf(m21);
    if (fabsf(m22) > scale) scale = fabsf(m22);
    if (scale == 0.f) scale = 1.f;
    m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;

    float diag[3], sub[2];
    float q[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    diag[0] = m00;
    const float v1norm2 = m20 * m20;
    if (v1norm2 <= FLT_MIN) {
        diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
    } else {
        const float beta = sqrtf(m10 * m10 + v1norm2);
        const float invBeta = 1.f / beta;
        const float m01 = m10 * invBeta;
        const float m02 = m20 * invBeta;
        const float qq = 2.f * m01 * m21 + m02 * (m22 - m11);
        diag[1] = m11 + m02 * qq;
        diag[2] = m22 - m02 * qq;
        sub[0] = beta;
        sub[1] = m21 - m01 * qq;
        q[1][1] = m01; q[1][2] = m02; q[2][1] = m02; q[2][2] = -m01;
    }
    int end = 2, start = 0, iter = 0;
    const int max_total = 30 * 3;
    const float precision_inv = 1.f / FLT_EPSILON;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (fabsf(sub[i]) < FLT_MIN) {
                sub[i] = 0.f;
            } else {
                const float sc = precision_inv * sub[i];
                if (sc * sc <= (fabsf(diag[i]) + fabsf(diag[i + 1]))) sub[i] = 0.f;
            }
        }
        while (end > 0 && sub[end - 1] == 0.f) end--;
        if (end <= 0) break;
        iter++;
        if (iter > max_total) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.f) start--;
        qr_step(diag, sub, start, end, q);
    }
    const int ok = iter <= max_total;
    if (ok) {
        for (int i = 0; i < 2; ++i) {
            int k = 0;
            float best = diag[i];
            for (int j = 1; j < 3 - i; ++j)
                if (diag[i + j] < best) { best = diag[i + j]; k = j; }
            if (k > 0) {
                float t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
                for (int r = 0; r < 3; ++r) { t = q[r][i]; q[r][i] = q[r][k + i]; q[r][k + i] = t; }
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        evals[i] = diag[i] * scale;
        for (int j = 0; j < 3; ++j) evecs[i * 3 + j] = q[i][j];
    }
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * glibc 2.39 atan2f (sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c, the fdlibm float code),
 * restated with explicit IEEE float operations.  The product evaluates the same sequence on
 * the device so that sector indices are bit-identical to the reference's host libm; this copy
 * lets the CPU tests prove restated == libm over hundreds of millions of inputs.
 * ---------------------------------------------------------------------------------------- */
static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

static const float k_atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
static const float k_atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
static const float k_aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                               9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                               4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};

static float atanf_restated(float x) {
    float w, s1, s2, z;
    int32_t id;
    const int32_t hx = f2i(x), ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) {
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? k_atanhi[3] + k_atanlo[3] : -k_atanhi[3] - k_atanlo[3];
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (k_aT[0] + w * (k_aT[2] + w * (k_aT[4] + w * (k_aT[6] + w * (k_aT[8] + w * k_aT[10])))));
    s2 = w * (k_aT[1] + w * (k_aT[3] + w * (k_aT[5] + w * (k_aT[7] + w * k_aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = k_atanhi[id] - ((x * (s1 + s2) - k_atanlo[id]) - x);
    return hx < 0 ? -z : z;
}

float rpwo_atan2f_restated(float y, float x) {
    static const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    const int32_t hx = f2i(x), ix = hx & 0x7fffffff, hy = f2i(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_restated(y);
    const int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000 || iy == 0x7f800000) return atan2f(y, x); /* infinities never reach the binning stage */
    const int32_t k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_restated(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return i2f(f2i(z) ^ (int32_t)0x80000000);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

/* Counts inputs on which the restatement and the host libm disagree bitwise.  Inputs: xorshift64
 * stream from `seed`; two thirds uniform in [-range, range]^2 (the coordinates a scan has), one
 * third arbitrary finite bit patterns. */
uint64_t rpwo_atan2f_selfcheck(uint64_t n, uint64_t seed, float range) {
    uint64_t s = seed ? seed : 88172645463325252ULL, bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float x = (float)((int32_t)(s & 0xffffff) - 0x800000) * (range / 8388608.0f);
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float y = (float)((int32_t)(s & 0xffffff) - 0x800000) * (range / 8388608.0f);
        if (i % 3 == 1) {
            x = i2f((int32_t)(s >> 32));
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            y = i2f((int32_t)(s >> 32));
            if (!isfinite(x) || !isfinite(y)) continue;
        }
        const float a = atan2f(y, x), b = rpwo_atan2f_restated(y, x);
        if (f2i(a) != f2i(b)) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------
 * fitPlanePCA, RP/src/recursive_patchwork.cpp:77-107
 * ---------------------------------------------------------------------------------------- */
typedef struct { float c[3]; float nrm[3]; float residual; } plane_t;

/* px/py/pz: node points; sel: optional mask (only points with sel[i]!=0 take part). */
static plane_t fit_plane_pca(const float* px, const float* py, const float* pz, const uint8_t* sel, size_t n, size_t cnt) {
    plane_t r;
    if (cnt < 3) { /* :78-80 */
        r.c[0] = r.c[1] = r.c[2] = 0.f;
        r.nrm[0] = 0.f; r.nrm[1] = 0.f; r.nrm[2] = 1.f;
        r.residual = FLT_MAX;
        return r;
    }
    /* computeCentroid, point_cloud_processor.cpp:58-70: sequential float sums, then /= size */
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (size_t i = 0; i < n; ++i)
        if (!sel || sel[i]) { cx += px[i]; cy += py[i]; cz += pz[i]; }
    const float fn = (float)cnt;
    cx = cx / fn; cy = cy / fn; cz = cz / fn;
    /* computeCovariance, point_cloud_processor.cpp:72-86: sum of outer products, /= (size-1) */
    float cov[9] = {0};
    for (size_t i = 0; i < n; ++i) {
        if (sel && !sel[i]) continue;
        const float d0 = px[i] - cx, d1 = py[i] - cy, d2 = pz[i] - cz;
        cov[0] += d0 * d0; cov[1] += d0 * d1; cov[2] += d0 * d2;
        cov[3] += d1 * d0; cov[4] += d1 * d1; cov[5] += d1 * d2;
        cov[6] += d2 * d0; cov[7] += d2 * d1; cov[8] += d2 * d2;
    }
    const float fn1 = (float)(cnt - 1);
    for (int k = 0; k < 9; ++k) cov[k] = cov[k] / fn1;
    /* :89-95 smallest-eigenvalue eigenvector, z-up */
    float ev[3], V[9];
    rpwo_eig3_f32(cov, ev, V);
    float nx = V[0], ny = V[3], nz = V[6];
    if (nz < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
    /* :98-104 mean absolute point-plane distance */
    float res = 0.f;
    for (size_t i = 0; i < n; ++i) {
        if (sel && !sel[i]) continue;
        res += fabsf(dot3(px[i] - cx, py[i] - cy, pz[i] - cz, nx, ny, nz));
    }
    res = res / fn;
    r.c[0] = cx; r.c[1] = cy; r.c[2] = cz;
    r.nrm[0] = nx; r.nrm[1] = ny; r.nrm[2] = nz;
    r.residual = res;
    return r;
}

/* ------------------------------------------------------------------------------------------
 * libstdc++ std::partial_sort(first, first+3, last, comp) selection — which three indices end
 * up in front (RP/src/recursive_patchwork.cpp:175-176).  Only the SET matters to the caller.
 * Emulates __heap_select with __make_heap/__pop_heap/__adjust_heap/__push_heap on a 3-heap so
 * that ties in z resolve exactly as in the reference binary.
 * ---------------------------------------------------------------------------------------- */
static void adjust_heap3(size_t* h, size_t hole, size_t value, const float* z) {
    const size_t len = 3, top = hole;
    size_t second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (z[h[second]] < z[h[second - 1]]) second--;
        h[hole] = h[second];
        hole = second;
    }
    /* len is odd: no lone-child case */
    size_t parent = (hole - 1) / 2;
    while (hole > top && z[h[parent]] < z[value]) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}

static void lowest3(const float* z, size_t n, size_t out[3]) {
    size_t h[3] = {0, 1, 2};
    /* __make_heap: parent = (len-2)/2 = 0 only */
    adjust_heap3(h, 0, h[0], z);
    for (size_t i = 3; i < n; ++i) {
        if (z[i] < z[h[0]]) adjust_heap3(h, 0, i, z); /* __pop_heap: old top leaves, i enters */
    }
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
}

static int cmp_float(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------
 * fitPlaneAndSplit, RP/src/recursive_patchwork.cpp:109-308
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const rpwo_config* cfg;
    rpwo_node* nodes;
    size_t nodes_cap, n_nodes;
    rpwo_stats st;
    int root;
    int oom;
} ctx_t;

static rpwo_node* new_node(ctx_t* C, int depth, size_t start, size_t n) {
    static rpwo_node scratch;
    rpwo_node* nd = (C->nodes && C->n_nodes < C->nodes_cap) ? &C->nodes[C->n_nodes] : &scratch;
    C->n_nodes++;
    memset(nd, 0, sizeof(*nd));
    nd->root = C->root; nd->depth = depth; nd->start = (int32_t)start; nd->n = (int32_t)n;
    nd->split_axis = -1; nd->normal[2] = 1.f; nd->residual = FLT_MAX;
    C->st.n_nodes++;
    if (depth > C->st.max_depth) C->st.max_depth = depth;
    return nd;
}

/* Writes the returned vector<bool> into out[0..n).  `start` is bookkeeping only. */
static void fit_plane_and_split(ctx_t* C, const float* px, const float* py, const float* pz, size_t n,
                                float mean_dist, int depth, size_t start, uint8_t* out) {
    const rpwo_config* cfg = C->cfg;
    rpwo_node* nd = new_node(C, depth, start, n);
    /* :111-113 */
    if (n < 3 || depth > cfg->max_split_depth) {
        memset(out, 0, n);
        nd->outcome = RPWO_NODE_SMALL; C->st.n_leaves++;
        return;
    }
    /* :116-129 */
    float x_min = FLT_MAX, x_max = -FLT_MAX, y_min = FLT_MAX, y_max = -FLT_MAX;
    for (size_t i = 0; i < n; ++i) {
        if (px[i] < x_min) x_min = px[i];
        if (x_max < px[i]) x_max = px[i];
        if (py[i] < y_min) y_min = py[i];
        if (y_max < py[i]) y_max = py[i];
    }
    const float area = (x_max - x_min) * (y_max - y_min);
    if (area < 25.0f && depth > 0) {
        memset(out, 1, n);
        nd->outcome = RPWO_NODE_AREA; C->st.n_leaves++;
        return;
    }
    /* :132-140 */
    float z_min = FLT_MAX, z_max = -FLT_MAX;
    for (size_t i = 0; i < n; ++i) {
        if (pz[i] < z_min) z_min = pz[i];
        if (z_max < pz[i]) z_max = pz[i];
    }
    if ((z_max - z_min) < 0.05f && n > 10) {
        memset(out, 1, n);
        nd->outcome = RPWO_NODE_FLAT; C->st.n_leaves++;
        return;
    }
    /* :149-160 */
    const float rel_dist = mean_dist / cfg->filtering_radius;
    float z_th;
    if (cfg->adaptive_seed_height) {
        z_th = cfg->sensor_height + 0.2f * rel_dist;
    } else {
        float* sz = (float*)malloc(n * sizeof(float));
        if (!sz) { C->oom = 1; memset(out, 0, n); return; }
        memcpy(sz, pz, n * sizeof(float));
        qsort(sz, n, sizeof(float), cmp_float);
        const size_t idx = (size_t)(0.1f * (float)n);
        z_th = sz[idx] + cfg->th_seeds;
        free(sz);
    }
    /* :163-182 */
    uint8_t* mask = (uint8_t*)malloc(n);
    uint8_t* new_mask = (uint8_t*)malloc(n);
    if (!mask || !new_mask) { C->oom = 1; free(mask); free(new_mask); memset(out, 0, n); return; }
    size_t seed_count = 0;
    for (size_t i = 0; i < n; ++i) { mask[i] = pz[i] < z_th; seed_count += mask[i]; }
    if (seed_count < 3) {
        size_t low[3];
        lowest3(pz, n, low);
        memset(mask, 0, n);
        mask[low[0]] = mask[low[1]] = mask[low[2]] = 1;
    }
    /* :185-217 */
    const float threshold = cfg->th_dist * (1.0f + 0.2f * rel_dist);
    int iters = 0;
    for (int iter = 0; iter < cfg->max_iter; ++iter) {
        size_t cnt = 0;
        for (size_t i = 0; i < n; ++i) cnt += mask[i];
        if (cnt < 3) break;
        const plane_t pl = fit_plane_pca(px, py, pz, mask, n, cnt);
        iters++;
        C->st.n_pca_iters++;
        C->st.n_point_iters += (int64_t)n;
        int same = 1;
        for (size_t i = 0; i < n; ++i) {
            const float dist = fabsf(dot3(px[i] - pl.c[0], py[i] - pl.c[1], pz[i] - pl.c[2], pl.nrm[0], pl.nrm[1], pl.nrm[2]));
            new_mask[i] = dist < threshold;
            same &= (new_mask[i] == mask[i]);
        }
        if (same) break;
        memcpy(mask, new_mask, n);
    }
    /* :220-228 */
    size_t n_in = 0;
    for (size_t i = 0; i < n; ++i) n_in += mask[i];
    const plane_t fin = fit_plane_pca(px, py, pz, mask, n, n_in);
    nd->iters = iters;
    nd->n_inliers = (int32_t)n_in;
    memcpy(nd->centroid, fin.c, sizeof(fin.c));
    memcpy(nd->normal, fin.nrm, sizeof(fin.nrm));
    nd->residual = fin.residual;
    /* :231-235 */
    const float split_threshold = cfg->th_dist * (1.0f + 1.5f * (float)depth);
    const size_t min_patch_size = (size_t)(50 + 10 * depth);
    if (fin.residual > split_threshold && depth < cfg->max_split_depth && n >= min_patch_size) {
        /* :238-250 population variances about computeCentroid of ALL node points */
        float cx = 0.f, cy = 0.f;
        for (size_t i = 0; i < n; ++i) { cx += px[i]; cy += py[i]; }
        cx = cx / (float)n; cy = cy / (float)n;
        float var_x = 0.f, var_y = 0.f;
        for (size_t i = 0; i < n; ++i) {
            const float dx = px[i] - cx, dy = py[i] - cy;
            var_x += dx * dx;
            var_y += dy * dy;
        }
        var_x = var_x / (float)n;
        var_y = var_y / (float)n;
        const int axis = (var_x > var_y) ? 0 : 1;
        const float* pv = axis == 0 ? px : py;
        /* :251-269 upper median of the sorted coordinate */
        float* sv = (float*)malloc(n * sizeof(float));
        float* buf = (float*)malloc(3 * n * sizeof(float));
        if (!sv || !buf) { C->oom = 1; free(sv); free(buf); free(mask); free(new_mask); memset(out, 0, n); return; }
        memcpy(sv, pv, n * sizeof(float));
        qsort(sv, n, sizeof(float), cmp_float);
        const float median = sv[n / 2];
        free(sv);
        /* :272-283 stable partition */
        size_t nl = 0;
        for (size_t i = 0; i < n; ++i) nl += (pv[i] <= median);
        float *lx = buf, *ly = buf + n, *lz = buf + 2 * n;
        size_t li = 0, ri = nl;
        for (size_t i = 0; i < n; ++i) {
            if (pv[i] <= median) { lx[li] = px[i]; ly[li] = py[i]; lz[li] = pz[i]; li++; }
            else { lx[ri] = px[i]; ly[ri] = py[i]; lz[ri] = pz[i]; ri++; }
        }
        nd->outcome = RPWO_NODE_SPLIT;
        nd->split_axis = axis;
        nd->median = median;
        C->st.n_splits++;
        if (fin.residual == FLT_MAX) C->st.n_splits_collapse++;
        free(mask); free(new_mask);
        /* :286-304 children inherit mean_dist (Q4); results are concatenated left-then-right and
         * handed back in THAT order (Q1: the reference never maps them back to parent order). */
        fit_plane_and_split(C, lx, ly, lz, nl, mean_dist, depth + 1, start, out);
        fit_plane_and_split(C, lx + nl, ly + nl, lz + nl, n - nl, mean_dist, depth + 1, start + nl, out + nl);
        free(buf);
        return;
    }
    memcpy(out, mask, n); /* :307 */
    nd->outcome = RPWO_NODE_FIT; C->st.n_leaves++;
    free(mask); free(new_mask);
}

/* ------------------------------------------------------------------------------------------
 * filterGroundPoints, RP/src/recursive_patchwork.cpp:310-426
 * ---------------------------------------------------------------------------------------- */
int rpwo_filter_ground(const rpwo_config* cfg, const float* xyz, size_t n, size_t stride,
                       uint8_t* labels_out, uint16_t* keys_out, float* dist_out, float* angle_out,
                       rpwo_node* nodes, size_t nodes_cap, size_t* n_nodes, rpwo_stats* stats) {
    if (!cfg || (!xyz && n) || stride < 3 || cfg->num_sectors < 1 || 8 * (int64_t)cfg->num_sectors > 0xFFF0) return -1;
    ctx_t C;
    memset(&C, 0, sizeof(C));
    C.cfg = cfg; C.nodes = nodes; C.nodes_cap = nodes_cap;
    C.st.n_points = (int64_t)n;
    const int S = cfg->num_sectors;
    const int P = NUM_RINGS * S;
    const float R = cfg->filtering_radius;

    uint8_t* labels = (uint8_t*)malloc(n ? n : 1);
    uint16_t* keys = (uint16_t*)malloc((n ? n : 1) * sizeof(uint16_t));
    float* dist = (float*)malloc((n ? n : 1) * sizeof(float));
    float* ang = (float*)malloc((n ? n : 1) * sizeof(float));
    uint32_t* count = (uint32_t*)calloc((size_t)P + 1, sizeof(uint32_t));
    if (!labels || !keys || !dist || !ang || !count) { free(labels); free(keys); free(dist); free(ang); free(count); return -1; }

    float ring_edges[NUM_RINGS + 1], sector_angle;
    rpwo_zone_model(cfg, ring_edges, &sector_angle);

    /* steps 1-5: clean (:19-35), range (:321, cuda_interface.cu:590), radius (:325, :607),
     * angle (:355, cuda_interface.cu:625-626), ring/sector membership (:360-378). */
    size_t n_zone = 0;
    for (size_t i = 0; i < n; ++i) {
        const float x = xyz[i * stride], y = xyz[i * stride + 1], z = xyz[i * stride + 2];
        dist[i] = 0.f; ang[i] = 0.f;
        if (!(isfinite(x) && isfinite(y) && isfinite(z))) { keys[i] = RPWO_KEY_DROPPED; labels[i] = RPWO_LABEL_DROPPED; continue; }
        C.st.n_clean++;
        const float d = sqrtf(x * x + y * y);
        dist[i] = d;
        if (!(d <= R)) { keys[i] = RPWO_KEY_BEYOND; labels[i] = RPWO_LABEL_BEYOND; continue; }
        n_zone++;
        labels[i] = RPWO_LABEL_NONGROUND;
        float a = atan2f(y, x);
        if (a < 0) a = (float)((double)a + (double)2.0f * M_PI);
        ang[i] = a;
        int ring = -1, sector = -1;
        for (int r = 0; r < NUM_RINGS; ++r)
            if (d >= ring_edges[r] && d < ring_edges[r + 1]) { ring = r; break; }
        for (int s = 0; s < S; ++s) {
            const float a0 = (float)s * sector_angle, a1 = (float)(s + 1) * sector_angle;
            if (a >= a0 && a < a1) { sector = s; break; }
        }
        if (ring < 0 || sector < 0) { keys[i] = RPWO_KEY_UNBINNED; continue; }
        keys[i] = (uint16_t)(ring * S + sector);
        count[keys[i]]++;
        C.st.n_binned++;
    }
    C.st.n_zone = (int64_t)n_zone;

    /* :339-341: fewer than 3 in-zone points -> everything cleaned is non-ground (labels already say so). */
    if (n_zone >= 3) {
        /* gather patches in input order (:372-378) */
        uint32_t* offs = (uint32_t*)malloc(((size_t)P + 1) * sizeof(uint32_t));
        size_t nb = (size_t)C.st.n_binned;
        float* sx = (float*)malloc((nb ? nb : 1) * 3 * sizeof(float));
        uint32_t* sidx = (uint32_t*)malloc((nb ? nb : 1) * sizeof(uint32_t));
        uint8_t* smask = (uint8_t*)malloc(nb ? nb : 1);
        if (!offs || !sx || !sidx || !smask) { free(offs); free(sx); free(sidx); free(smask); free(labels); free(keys); free(dist); free(ang); free(count); return -1; }
        float *sy = sx + nb, *sz = sx + 2 * nb;
        offs[0] = 0;
        for (int p = 0; p < P; ++p) offs[p + 1] = offs[p] + count[p];
        uint32_t* cur = (uint32_t*)malloc((size_t)P * sizeof(uint32_t));
        memcpy(cur, offs, (size_t)P * sizeof(uint32_t));
        for (size_t i = 0; i < n; ++i) {
            if (keys[i] >= RPWO_KEY_UNBINNED) continue;
            const uint32_t j = cur[keys[i]]++;
            sx[j] = xyz[i * stride]; sy[j] = xyz[i * stride + 1]; sz[j] = xyz[i * stride + 2];
            sidx[j] = (uint32_t)i;
        }
        free(cur);
        for (int p = 0; p < P; ++p) {
            const size_t np = count[p], o = offs[p];
            if (np == 0) continue; /* :380 */
            C.st.n_root_patches++;
            if ((int32_t)np > C.st.max_patch_points) C.st.max_patch_points = (int32_t)np;
            float mean_dist = 0.f; /* :383-387 */
            for (size_t j = 0; j < np; ++j) mean_dist += dist[sidx[o + j]];
            mean_dist = mean_dist / (float)np;
            C.root = p;
            fit_plane_and_split(&C, sx + o, sy + o, sz + o, np, mean_dist, 0, 0, smask + o);
            /* :393-397 result indexed by ORIGINAL patch order */
            for (size_t j = 0; j < np; ++j)
                if (smask[o + j]) labels[sidx[o + j]] = RPWO_LABEL_GROUND;
        }
        free(offs); free(sx); free(sidx); free(smask);
    }
    for (size_t i = 0; i < n; ++i) C.st.n_ground += (labels[i] == RPWO_LABEL_GROUND);

    if (labels_out) memcpy(labels_out, labels, n);
    if (keys_out) memcpy(keys_out, keys, n * sizeof(uint16_t));
    if (dist_out) memcpy(dist_out, dist, n * sizeof(float));
    if (angle_out) memcpy(angle_out, ang, n * sizeof(float));
    if (n_nodes) *n_nodes = C.n_nodes;
    if (stats) *stats = C.st;
    free(labels); free(keys); free(dist); free(ang); free(count);
    return C.oom ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Multi-LiDAR fusion front end: LidarFusion::fuseLidarPointClouds,
 * RP/src/lidar_fusion.cpp:42-86; processSingleLidar :88-108; applyRotation2D :110-126;
 * removeEgoVehicle :148-159; isPointInEgoRadius :184-187.
 * fused_xyz: 3 floats per surviving point, sensor after sensor; src[k]: index of fused point k in the
 * concatenation of all sensor inputs.  Returns the number of fused points.
 * ---------------------------------------------------------------------------------------- */
size_t rpwo_fuse(const rpwo_sensor* sensors, size_t n_sensors, size_t stride, float* fused_xyz, uint32_t* src) {
    size_t w = 0, base = 0;
    for (size_t s = 0; s < n_sensors; ++s) {
        const rpwo_sensor* S = &sensors[s];
        const int rotate = fabsf(S->rotation_deg) > 1e-6f;                         /* :99 */
        const float angle_rad = (float)((double)S->rotation_deg * M_PI / (double)180.0f); /* :111 */
        const float cos_a = cosf(angle_rad), sin_a = sinf(angle_rad);             /* :112-113 */
        for (size_t i = 0; i < S->n; ++i) {
            float x = S->xyz[i * stride], y = S->xyz[i * stride + 1];
            const float z = S->xyz[i * stride + 2];
            if (rotate) {
                const float rx = x * cos_a - y * sin_a;                           /* :120 */
                const float ry = x * sin_a + y * cos_a;                           /* :121 */
                x = rx; y = ry;
            }
            const float d = sqrtf(x * x + y * y);                                 /* :185 */
            if (d <= S->ego_radius) continue;                                     /* :186, :153 */
            if (fused_xyz) { fused_xyz[3 * w] = x; fused_xyz[3 * w + 1] = y; fused_xyz[3 * w + 2] = z; }
            if (src) src[w] = (uint32_t)(base + i);
            ++w;
        }
        base += S->n;
    }
    return w;
}

double rpwo_time_scan(const rpwo_config* cfg, const float* xyz, size_t n, size_t stride, int reps) {
    uint8_t* labels = (uint8_t*)malloc(n ? n : 1);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; ++r) rpwo_filter_ground(cfg, xyz, n, stride, labels, NULL, NULL, NULL, NULL, 0, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(labels);
    return ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (reps > 0 ? reps : 1);
}
