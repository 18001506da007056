/* oracle/rpw_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C interface of the CPU restatement of the reference's per-scan ground segmentation
 * (RP/src/recursive_patchwork.cpp:310-426 and everything it calls).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 */
#ifndef RPW_ORACLE_H
#define RPW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field-for-field mirror of PatchworkConfig, RP/include/recursive_patchwork.hpp:25-36
 * (bool stored as int32).  Same layout as rpwref_config (oracle/ref_driver.cpp) and
 * rpw_config (include/rpw_b200.h). */
typedef struct rpwo_config {
    float sensor_height;
    float max_range;
    int32_t num_sectors;
    int32_t max_iter;
    int32_t adaptive_seed_height;
    float th_seeds;
    float th_dist;
    float th_outlier;
    float filtering_radius;
    int32_t max_split_depth;
} rpwo_config;

/* Per-input-point key written by the binning stage. */
#define RPWO_KEY_DROPPED 0xFFFFu  /* non-finite, removed by cleanPoints          */
#define RPWO_KEY_BEYOND 0xFFFEu   /* d > filtering_radius                        */
#define RPWO_KEY_UNBINNED 0xFFFDu /* in zone but in no ring/sector   */
/* otherwise key = ring * num_sectors + sector */

/* Per-input-point label (the reference returns clouds only; SURVEY Q10 defines these). */
#define RPWO_LABEL_NONGROUND 0
#define RPWO_LABEL_GROUND 1
#define RPWO_LABEL_BEYOND 2
#define RPWO_LABEL_DROPPED 3

/* How a fitPlaneAndSplit call ended. */
#define RPWO_NODE_SMALL 1 /* n<3 or depth>max_split_depth: all false (:111-113)    */
#define RPWO_NODE_AREA 2  /* bbox area<25 at depth>0: all true (:126-129)          */
#define RPWO_NODE_FLAT 3  /* z range<0.05 and n>10: all true (:138-140)            */
#define RPWO_NODE_FIT 4   /* leaf: iterated plane mask returned (:307)             */
#define RPWO_NODE_SPLIT 5 /* internal node: split at the median (:234-305)         */

/* One fitPlaneAndSplit invocation.  (root, start, n) identifies the node: `start` is the
 * offset of its first point inside the root patch in the reference's concatenated return
 * order (left subtree first), so leaves tile [0, n_root) and their masks, concatenated, are
 * the vector the reference hands back. */
typedef struct rpwo_node {
    int32_t root;       /* ring * num_sectors + sector */
    int32_t depth;
    int32_t start;
    int32_t n;
    int32_t outcome;    /* RPWO_NODE_* */
    int32_t iters;      /* fitPlanePCA calls inside the loop (:199) */
    int32_t n_inliers;  /* inliers of the mask the final fit ran on (:220-228) */
    int32_t split_axis; /* 0 x, 1 y, -1 none */
    float centroid[3];  /* final plane (:228); zeros / (0,0,1) when n_inliers<3 */
    float normal[3];
    float residual;     /* FLT_MAX when n_inliers<3 */
    float median;       /* split value when outcome==SPLIT */
} rpwo_node;

typedef struct rpwo_stats {
    int64_t n_points, n_clean, n_zone, n_binned;
    int64_t n_ground;
    int64_t n_root_patches;   /* non-empty */
    int64_t n_nodes, n_leaves, n_splits, n_splits_collapse; /* collapse: residual==FLT_MAX route (Q3) */
    int64_t n_pca_iters;
    int64_t n_point_iters;    /* sum over loop iterations of node size */
    int32_t max_depth;
    int32_t max_patch_points;
} rpwo_stats;

void rpwo_default_config(rpwo_config* out);

/* Ring edges (9 floats) and sector angle exactly as RP/src/recursive_patchwork.cpp:344-352. */
void rpwo_zone_model(const rpwo_config* cfg, float ring_edges[9], float* sector_angle);

/* Full path.  xyz: n points `stride` floats apart.  Any output pointer may be NULL.
 * nodes: capacity nodes_cap; *n_nodes receives the number that WOULD be written.
 * Returns 0 on success, -1 on allocation failure / bad arguments. */
int rpwo_filter_ground(const rpwo_config* cfg, const float* xyz, size_t n, size_t stride,
                       uint8_t* labels_out, uint16_t* keys_out, float* dist_out, float* angle_out,
                       rpwo_node* nodes, size_t nodes_cap, size_t* n_nodes, rpwo_stats* stats);

/* 3x3 symmetric eigen-decomposition, Eigen 3.4.0 SelfAdjointEigenSolver<Matrix3f>::compute
 * restated (float).  a: row-major 3x3 (lower triangle read).  evals ascending, evecs[r*3+c]
 * column c = eigenvector c.  Returns 1 if converged. */
int rpwo_eig3_f32(const float a[9], float evals[3], float evecs[9]);

/* glibc-2.39 atan2f restated in IEEE float operations (see rpw_oracle.c); used to prove that
 * the device's sector angle can be made bit-identical to the host libm's. */
float rpwo_atan2f_restated(float y, float x);
uint64_t rpwo_atan2f_selfcheck(uint64_t n, uint64_t seed, float range);

/* Multi-LiDAR fusion front end (RP/src/lidar_fusion.cpp:42-126,148-159,184-187). */
typedef struct rpwo_sensor {
    const float* xyz;
    size_t n;
    float rotation_deg;
    float ego_radius;
} rpwo_sensor;
size_t rpwo_fuse(const rpwo_sensor* sensors, size_t n_sensors, size_t stride, float* fused_xyz, uint32_t* src);

/* seconds per call over `reps` calls (wall clock); labels only. */
double rpwo_time_scan(const rpwo_config* cfg, const float* xyz, size_t n, size_t stride, int reps);

#ifdef __cplusplus
}
#endif
#endif
