/* include/rpw_b200.h — C-ABI of the B200-native Recursive Patchwork ground-segmentation path.
 *
 * This is the drop-in boundary: plain C, opaque handle, plain pointers and sizes, integer status
 * codes, no exceptions, no CUDA / torch / Eigen types.  Everything below it is hand-written
 * sm_100a CUDA (librpw_b200.so); everything above it (the C++ class that mirrors the reference's
 * header, ROS2 node, bag loader, Python bindings) stays on the host side.
 *
 * Reference interface each entry point stands in for (RP = src/recursive_patchwork in the
 * reference repository):
 *   rpw_config                 struct PatchworkConfig            RP/include/recursive_patchwork.hpp:25-36
 *   rpw_create / rpw_destroy   RecursivePatchwork ctor / dtor    RP/include/recursive_patchwork.hpp:49-50
 *                              + cuda::CudaManager::initialize / cleanup  RP/include/cuda_interface.hpp:14-20
 *   rpw_set_config/get_config  setConfig / getConfig             RP/include/recursive_patchwork.hpp:66-67
 *   rpw_segment*               filterGroundPoints                RP/include/recursive_patchwork.hpp:53-54,
 *                                                                RP/src/recursive_patchwork.cpp:310-426
 *                              (which subsumes cuda::ops::computeDistances2D / filterPointsByRadius /
 *                               computeAngles / computePlaneDistances, RP/include/cuda_interface.hpp:73-88)
 *   rpw_zone_model             ring edges + sector angle         RP/src/recursive_patchwork.cpp:344-352
 *
 * Bit-exactness of the ring/sector keys is against the reference built on glibc <= 2.40 (this image: 2.39; ROS2
 * Humble / Jazzy targets: 2.35 / 2.39): the device restates that libm's atan2f (the fdlibm float algorithm) operation for
 * operation.  glibc 2.41 switched atan2f to the correctly rounded CORE-MATH routine; a reference built there computes
 * different angle bits itself (keys differ only for points within an ulp of a sector edge).  tests/test_oracle.py names
 * the libm it proved the sequence against.  Limits: num_sectors <= 128 (the reference has no limit; keys are 16-bit and
 * the scatter keeps per-warp counters per patch), < 2^32 - 65536 points per call, patches < 2^24 points.
 *
 * There is NO CPU fallback: without a usable CUDA device rpw_create fails with RPW_ERR_NO_DEVICE.
 *
 * Threading: a handle owns one device, one stream and its buffers; it is not re-entrant (one
 * call at a time per handle).  Distinct handles may be used concurrently from distinct threads;
 * that is how several GPUs (or several in-flight batches on one GPU) are driven.
 */
#ifndef RPW_B200_H
#define RPW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPW_ABI_VERSION 1

/* Status codes (the reference has no error convention on this path: CUDA errors are printed and
 * ignored, RP/cuda/cuda_wrapper.cu:119-122; the C++ shim turns non-zero into an exception). */
enum {
    RPW_OK = 0,
    RPW_ERR_BAD_ARG = 1,   /* NULL pointer, bad stride, bad config                           */
    RPW_ERR_NO_DEVICE = 2, /* no CUDA device / device index out of range / not sm_100         */
    RPW_ERR_CUDA = 3,      /* a CUDA runtime call or kernel failed; see rpw_last_error        */
    RPW_ERR_CAPACITY = 4,  /* more points / scans than the handle was created for            */
    RPW_ERR_ALLOC = 5      /* host or device allocation failed                                */
};

/* Field-for-field mirror of PatchworkConfig (bool widened to int32).  max_range and th_outlier are
 * carried for interface fidelity; like the reference, the path never reads them. */
typedef struct rpw_config {
    float sensor_height;          /* 1.2   */
    float max_range;              /* 150   */
    int32_t num_sectors;          /* 10    */
    int32_t max_iter;             /* 100   */
    int32_t adaptive_seed_height; /* 1     */
    float th_seeds;               /* 0.15  */
    float th_dist;                /* 0.2   */
    float th_outlier;             /* 0.08  */
    float filtering_radius;       /* 150   */
    int32_t max_split_depth;      /* 1000  */
} rpw_config;

/* Per-input-point label.  The reference returns two clouds and no labels; labels are the
 * north-star addition and determine the clouds: ground = points labelled 1 in input order;
 * non-ground = points labelled 0 in input order followed by points labelled 2 in input order
 * (RP/src/recursive_patchwork.cpp:402-419); points labelled 3 appear in neither. */
#define RPW_LABEL_NONGROUND 0u
#define RPW_LABEL_GROUND 1u
#define RPW_LABEL_BEYOND 2u  /* finite, sqrt(x^2+y^2) > filtering_radius */
#define RPW_LABEL_DROPPED 3u /* non-finite coordinate (cleanPoints)      */
#define RPW_LABEL_EGO 4u     /* rpw_segment_fused only: inside the sensor's ego radius, never part of the merged cloud */

/* Per-input-point binning key (debug / parity output). key = ring * num_sectors + sector. */
#define RPW_KEY_DROPPED 0xFFFFu
#define RPW_KEY_BEYOND 0xFFFEu
#define RPW_KEY_UNBINNED 0xFFFDu /* in radius but in no ring/sector (d < 1 m, d == R, angle == 2*pi) */
#define RPW_KEY_EGO 0xFFFCu      /* rpw_segment_fused only */

#define RPW_NUM_RINGS 8 /* RP/src/recursive_patchwork.cpp:345 */

/* How one fitPlaneAndSplit node ended (RP/src/recursive_patchwork.cpp:109-308). */
#define RPW_NODE_SMALL 1 /* n < 3 or depth > max_split_depth: all non-ground */
#define RPW_NODE_AREA 2  /* xy bounding box < 25 m^2 at depth > 0: all ground */
#define RPW_NODE_FLAT 3  /* z range < 0.05 m and n > 10: all ground           */
#define RPW_NODE_FIT 4   /* leaf: iterated plane mask                          */
#define RPW_NODE_SPLIT 5 /* split at the upper median of the wider axis        */

/* One node of the recursion, for parity checks (rpw_debug_nodes).  (scan, root, start, n)
 * identifies it: `start` is the node's offset inside its root patch, in the order in which the
 * reference concatenates child results (left subtree first). */
typedef struct rpw_node {
    int32_t scan;       /* index of the scan inside the batch */
    int32_t root;       /* ring * num_sectors + sector */
    int32_t depth;
    int32_t start;
    int32_t n;
    int32_t outcome;    /* RPW_NODE_* */
    int32_t iters;      /* plane fits inside the iteration loop */
    int32_t n_inliers;  /* inliers of the mask the final fit ran on */
    int32_t split_axis; /* 0 x, 1 y, -1 none */
    float centroid[3];  /* final plane; zeros / (0,0,1) / FLT_MAX when n_inliers < 3 */
    float normal[3];
    float residual;
    float median;       /* split value when outcome == RPW_NODE_SPLIT */
    float mean_dist;    /* root patch mean range inherited by the node */
} rpw_node;

typedef struct rpw_stats {
    uint64_t n_points;    /* input points over the whole call            */
    uint64_t n_ground;    /* label 1                                     */
    uint64_t n_nonground; /* label 0                                     */
    uint64_t n_beyond;    /* label 2                                     */
    uint64_t n_dropped;   /* label 3                                     */
    uint32_t n_levels;    /* recursion levels the device worklist ran    */
    uint32_t n_nodes;     /* fitPlaneAndSplit nodes processed            */
    uint64_t kernel_launches; /* kernels launched by this call           */
} rpw_stats;

typedef struct rpw_handle rpw_handle;

/* ---- lifecycle --------------------------------------------------------------------------- */
void rpw_default_config(rpw_config* out);

/* Ring edges (RPW_NUM_RINGS+1 floats) and sector angle, computed on the host with the same
 * libm calls as the reference (powf, double division) so that device keys are bit-exact. */
int rpw_zone_model(const rpw_config* cfg, float* ring_edges9, float* sector_angle);

/* max_total_points: capacity in points summed over one call's batch; max_batch: scans per call.
 * Device memory used is about 72 bytes per point of capacity. */
int rpw_create(const rpw_config* cfg, int device, size_t max_total_points, size_t max_batch, rpw_handle** out);
void rpw_destroy(rpw_handle* h);
int rpw_set_config(rpw_handle* h, const rpw_config* cfg);
int rpw_get_config(const rpw_handle* h, rpw_config* out);
/* Grows the handle's capacity to at least (max_total_points, max_batch) in place: the handle, its
 * streams, configuration, solver choice and debug switches stay, only the device / pinned buffers
 * whose size follows the capacity are replaced (waits for the handle's pending work first; results of
 * earlier calls -- rpw_last_clouds, rpw_debug_keys -- are gone afterwards).  A no-op when the handle
 * is already large enough; it never shrinks.  The reference's class holds no buffers at all
 * (RP/include/recursive_patchwork.hpp:70), so its callers never size anything: the C++ drop-in calls
 * this when a cloud is larger than any it has seen. */
int rpw_reserve(rpw_handle* h, size_t max_total_points, size_t max_batch);
int rpw_capacity(const rpw_handle* h, size_t* max_total_points, size_t* max_batch);

/* How the plane normal (smallest-eigenvalue eigenvector of the inlier covariance,
 * RP/src/recursive_patchwork.cpp:89-90) is computed on the device.
 *   RPW_SOLVER_EIGEN_QR: the operation sequence of Eigen 3.4.0's SelfAdjointEigenSolver<Matrix3f> in
 *     float — given the same covariance bits it returns the reference's bits.  The serial 3x3 QR is
 *     about half of every plane-fit iteration.
 *   RPW_SOLVER_CLOSED_FORM: closed-form FP64 eigenvector (Newton on the characteristic cubic + row cross
 *     products), ~1.25x faster end to end and more accurate than the reference's float QR, but not the
 *     reference's rounding: where the two smallest eigenvalues nearly coincide the float QR's answer is
 *     far from the true eigenvector and only the QR reproduces it (one stress scan in 64 fell to 99.5 %).
 *   RPW_SOLVER_HYBRID (default): the closed form wherever the eigenvector is well conditioned (the two
 *     smallest eigenvalues further apart than 2 % of the matrix scale: both solvers then agree to
 *     ~1e-6 rad), the QR sequence for the rest (a few percent of the solves).  Measured against
 *     RPW_SOLVER_EIGEN_QR: 1 label of 61.4 M differs over 512 ordinary scans, at most 31 of 262 k on the
 *     two-layer stress scans (tools/gpu_solver_agreement.py); every parity test runs with both.
 *   RPW_SOLVER_REFERENCE: RPW_SOLVER_EIGEN_QR plus every float reduction of fitPlaneAndSplit in the reference's
 *     ORDER -- computeCentroid / computeCovariance (RP/src/point_cloud_processor.cpp:58-86), the residual
 *     (RP/src/recursive_patchwork.cpp:98-104), the root patch's mean range (:383-387) and the split statistics
 *     (:240-249) are sequential float sums there, and the other solvers add in trees.  With it every decision
 *     is taken on the reference's bits: labels, node records (centroid, normal, residual, iteration counts) and
 *     clouds are IDENTICAL to the reference's strict-IEEE build, including the bistable and never-converging
 *     patches that amplify last-bit differences (tests/test_gpu_parity.py).  The sums are dependent chains
 *     (one lane per sum), so this mode is about eight times slower on batches (34 k against 278 k scans/s per B200);
 *     it is the verification mode.
 * Environment override at rpw_create: RPW_PLANE_SOLVER=0|1|2|3. */
#define RPW_SOLVER_EIGEN_QR 0
#define RPW_SOLVER_CLOSED_FORM 1
#define RPW_SOLVER_HYBRID 2
#define RPW_SOLVER_REFERENCE 3
int rpw_set_plane_solver(rpw_handle* h, int solver);
/* Selective form of the reference-order arithmetic, for any solver: a plane fit that needed more than
 * max_fast_iterations iterations is run again from its seeds with the reference's sequential sums and the QR solver
 * (long fits are the ones that amplify the moments' last bits); the root mean range and the split statistics are
 * always taken in the reference's order then.  0 = every fit (what RPW_SOLVER_REFERENCE selects), negative = off
 * (default).  Environment override at rpw_create: RPW_EXACT_REPLAY=K. */
int rpw_set_exact_replay(rpw_handle* h, int max_fast_iterations);

/* Use a caller-owned CUDA stream (cudaStream_t passed as void*; NULL = the handle's own stream).
 * All copies and kernels of later calls are enqueued on it. */
int rpw_set_stream(rpw_handle* h, void* cuda_stream);

/* Message for the last non-zero status on this handle (h == NULL: last rpw_create failure). */
const char* rpw_last_error(const rpw_handle* h);

/* ---- the path, host buffers in / host labels out ---------------------------------------- */
/* One scan.  xyz: n points, first three floats of every `stride_bytes` record (12 = the
 * reference's Point3D AoS, 16 = float4, any multiple of 4 up to 1024 = wider records with xyz in
 * front, which is how RP/src/rosbag_loader.cpp:226-254 reads a PointCloud2).  labels_out: n bytes.  stats may be NULL.
 * Synchronous: returns after labels_out is filled. */
int rpw_segment(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out, rpw_stats* stats);

/* A batch of scans (frames are independent; RP/include/recursive_patchwork.hpp:70 — the class
 * holds no per-scan state).  clouds[i] / labels_out[i]: host buffers of scan i. */
int rpw_segment_batch(rpw_handle* h, const float* const* clouds, const size_t* n, size_t batch, size_t stride_bytes,
                      uint8_t* const* labels_out, rpw_stats* stats);

/* Same, asynchronous: enqueues H2D copies, kernels and D2H copies on the handle's stream and
 * returns.  Host buffers must be pinned (rpw_host_alloc) and stay valid until rpw_wait. */
int rpw_segment_batch_async(rpw_handle* h, const float* const* clouds, const size_t* n, size_t batch, size_t stride_bytes,
                            uint8_t* const* labels_out);
int rpw_wait(rpw_handle* h, rpw_stats* stats);

/* One scan straight from a sensor_msgs/PointCloud2 data buffer (little-endian float32 x, y, z
 * fields at byte offsets off_x/off_y/off_z of every point_step-byte record): the device reads the
 * records in place, so the host-side pcl::fromROSMsg + copy loop of the node
 * (RP/src/recursive_patchwork_node.cpp:67-88) disappears.  The buffer is copied as it is (n_points * point_step
 * bytes) and the kernels read the fields in place.  (Moving only the 12 xyz bytes of every record with a strided copy
 * was measured and lost: 0.273 against 0.178 ms for a 120 k-point scan of 32-byte records; RPW_PC2_PACK=1 selects it.) */
int rpw_segment_pc2(rpw_handle* h, const void* data, size_t n_points, size_t point_step, size_t off_x, size_t off_y, size_t off_z,
                    uint8_t* labels_out, rpw_stats* stats);

/* One merged multi-LiDAR frame with the fusion front end folded into the binning kernel (SURVEY
 * section 8f row 1).  Replaces LidarFusion::fuseLidarPointClouds + filterGroundPoints
 * (RP/src/lidar_fusion.cpp:42-126, RP/src/main.cpp:245-268): every sensor's cloud is rotated about z
 * by rotation_deg when |rotation_deg| > 1e-6 (same float operations as applyRotation2D, :110-126),
 * points with sqrt(x^2+y^2) <= ego_radius are removed (:148-159, :184-187; label RPW_LABEL_EGO), the
 * rest is segmented as ONE cloud in sensor order — no host-side rotate / filter / concatenate passes.
 * labels_out[s]: sensors[s].n bytes.  Up to 8 sensors. */
typedef struct rpw_sensor_cloud {
    const float* xyz;   /* host, stride_bytes records */
    size_t n;
    float rotation_deg; /* LidarConfig::rotation_angle, RP/include/recursive_patchwork.hpp:39-44 */
    float ego_radius;   /* LidarConfig::ego_radius (2.5 default) */
} rpw_sensor_cloud;
int rpw_segment_fused(rpw_handle* h, const rpw_sensor_cloud* sensors, size_t n_sensors, size_t stride_bytes,
                      uint8_t* const* labels_out, rpw_stats* stats);

/* Result assembly (RP/src/recursive_patchwork.cpp:402-419) for the scans of this handle's LAST rpw_segment* call
 * (host or device path, single, batch, PointCloud2 or fused), done on the device by a stable compaction over
 * the labels: per scan the ground cloud (ground points in input order) and the non-ground cloud (non-ground
 * points in input order, then the beyond-radius points in input order), packed xyz (12 bytes per point);
 * fused frames in vehicle coordinates.  Dropped (non-finite) and ego points are in neither, as in the reference.
 * Both buffers hold 3 * (total points of the call) floats; the clouds of scan b start at record
 * (offset of scan b inside the call).  on_device != 0: device buffers (must not alias the input); else
 * host buffers, either may be NULL, only the used part of every scan's range is written.
 * counts: HOST array, 2 per scan: ground points, non-ground points.  The input of the last call must still
 * be alive (device path: the caller's d_points). */
int rpw_last_clouds(rpw_handle* h, float* ground_xyz, float* nonground_xyz, int on_device, uint64_t* counts);

/* RecursivePatchwork::sampleGroundAndObstacles (RP/src/recursive_patchwork.cpp:428-465) for the scan of this handle's
 * LAST single-scan call, computed on the device from the clouds of rpw_last_clouds: first a context sample of
 * min(sample_size, #ground) distinct ground points (the reference uses 2000 and an unseeded std::mt19937; seed = 0
 * seeds from std::random_device as it does, any other value makes the draw reproducible), then the obstacles: the
 * non-ground points with sqrtf(x*x+y*y) > ego_radius (reference: 2.5) and |z - target_height| <= base_tol (1.1 / 0.5),
 * in the non-ground cloud's order.  A scan without non-ground points returns its whole ground cloud (:435-437).
 * out_xyz: HOST buffer of 3 * out_cap_points floats (sample_size + #non-ground points always suffices). */
int rpw_sample_ground_and_obstacles(rpw_handle* h, float target_height, float base_tol, float ego_radius, size_t sample_size,
                                    uint64_t seed, float* out_xyz, size_t out_cap_points, size_t* n_ground_sample, size_t* n_obstacles);

/* Bird's-eye-view rasters of RP/src/visualization.cpp:18-113 for the scan of this handle's LAST single-scan call, drawn
 * on the device from the clouds of rpw_last_clouds.  bgr_out: HOST buffer of height * width * 3 bytes, row-major BGR
 * (the layout of the reference's cv::Mat CV_8UC3), black background; pixel = int((p - min) * size / (max - min)) with
 * the reference's float operations; where several points share a pixel the last one drawn wins, as in the reference.
 *   RPW_BEV_CLASSES: createGroundNonGroundImage(ground, non_ground) (:47-80): ground green, then non-ground red.
 *   RPW_BEV_HEIGHT_NONGROUND: createBEVImage(non_ground) (:18-45, main.cpp:297): (i, i, 255), i = clamp((z + 2) * 50).
 *   RPW_BEV_HEIGHT_ALL: createBEVImage over the ground cloud followed by the non-ground cloud. */
#define RPW_BEV_CLASSES 0
#define RPW_BEV_HEIGHT_NONGROUND 1
#define RPW_BEV_HEIGHT_ALL 2
int rpw_bev_image(rpw_handle* h, int mode, int width, int height, float x_min, float y_min, float x_max, float y_max, uint8_t* bgr_out);

/* One scan, with the two clouds the reference returns, in the reference's order (rpw_segment + rpw_last_clouds).
 * ground_xyz / nonground_xyz: caller buffers of 3*n floats each (either may be NULL). */
int rpw_segment_clouds(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out,
                       float* ground_xyz, size_t* n_ground, float* nonground_xyz, size_t* n_nonground);
/* Same without the last copy: *ground_xyz / *nonground_xyz point into the handle's pinned staging memory (packed xyz,
 * n_ground / n_nonground records) and stay valid until the next call on this handle.  One synchronisation per call:
 * input copy, pipeline, result assembly and the copies back are enqueued together.  labels_out may be NULL. */
int rpw_segment_clouds_view(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out,
                            const float** ground_xyz, size_t* n_ground, const float** nonground_xyz, size_t* n_nonground);

/* ---- the path, device-resident ----------------------------------------------------------- */
/* d_points: device pointer to packed float4 (x,y,z,ignored) records of all scans back to back;
 * scan_offsets: HOST array of batch+1 point offsets (scan i = [off[i], off[i+1])); d_labels:
 * device pointer, off[batch] bytes.  Asynchronous on the handle's stream. */
int rpw_segment_device(rpw_handle* h, const void* d_points, const uint64_t* scan_offsets, size_t batch, void* d_labels);

/* ---- parity / debug ---------------------------------------------------------------------- */
/* Binning keys of the last call, per input point, batch order (host buffer, n_total entries). */
int rpw_debug_keys(rpw_handle* h, uint16_t* keys_out, size_t n_total);
/* Record recursion nodes during later calls (costs a few stores per node). */
int rpw_debug_enable_nodes(rpw_handle* h, int enable);
/* Nodes of the last call; *count receives how many exist even if cap is smaller. */
int rpw_debug_nodes(rpw_handle* h, rpw_node* out, size_t cap, size_t* count);
/* Runs the device 3x3 symmetric eigensolver on `count` row-major matrices (host in / host out):
 * evals (3 per matrix, ascending), evecs (9 per matrix, column c = eigenvector c). */
int rpw_debug_eig3(rpw_handle* h, const float* mats, size_t count, float* evals, float* evecs);
/* Runs the plane-normal solver the fit kernel uses on `count` scatter matrices (6 floats each:
 * xx yx yy zx zy zz; one warp per matrix): mode 0 = closed-form FP64 smallest eigenvector, mode 1 = Eigen's QR sequence
 * (generic form, the one rpw_debug_eig3 runs), mode 2 = the same sequence restructured for latency on the
 * compiler's IEEE division / square root / reciprocal, mode 3 = mode 2 on branch-free copies of those operations'
 * fast paths with an IEEE fallback (what the fit kernel runs by default; modes 1, 2, 3 are bit-identical),
 * mode 4 = mode 3 without the fallback (normal (0, 0, 2) marks a solve whose operands left the fast range).
 * normals3: unit normals with z >= 0; cycles: SM cycles one call took. */
int rpw_debug_normal(rpw_handle* h, const float* scatter6, size_t count, int mode, float* normals3, uint32_t* cycles);
/* Runs the device restatement of libm atan2f on `count` (y, x) pairs. */
int rpw_debug_atan2(rpw_handle* h, const float* y, const float* x, size_t count, float* out);

/* Cycle accounting inside the fit kernel (thread 0 of every block; summed over blocks).  Reads and
 * clears the 16 counters, then enables/disables the accounting for later calls.  Slots: 0 load+bbox,
 * 1 seeds, 2 covariance pass, 3 eigensolve, 4 distance/mask pass, 5 final fit, 6 leaf label write,
 * 7 split, 8 fetch, 9 grid barrier, 10 nodes, 11 plane-fit iterations, 12 load loop alone (before its reductions),
 * 13 block reduction of the distance pass, 14 eigensolve alone (QR solver: cycles; hybrid solver: number of solves that
 * took the QR path -- 3-4 % on the C4 / C5 test scenes). */
int rpw_debug_fit_timing(rpw_handle* h, int enable, uint64_t* cycles16);

/* Per-node timeline of the fit kernels (scheduling analysis: which SM ran which node when).
 * out == NULL: (re)arms the trace with room for `cap` records (0 switches it off).  Otherwise copies
 * up to `cap` records of the calls since arming into `out`, stores how many nodes were seen in
 * *count (may exceed cap: the rest was dropped) and clears the trace.  Tracing slows the kernels a
 * little (one global timer read and one 32-byte store per node). */
typedef struct rpw_trace_rec {
    uint64_t t_start_ns, t_end_ns; /* %globaltimer */
    uint32_t sm;                   /* %smid */
    uint32_t n;                    /* points of the node */
    uint16_t depth, size_class;    /* size_class: level-0 class index, 0xFFFF for deeper levels */
    uint32_t iters;                /* plane-fit iterations */
} rpw_trace_rec;
int rpw_debug_fit_trace(rpw_handle* h, rpw_trace_rec* out, size_t cap, size_t* count);

/* ---- measurement -------------------------------------------------------------------------- */
/* Device time per kernel, measured with CUDA events recorded on the handle's stream around every
 * launch (bench.py's roofline numbers).  Kernel ids: 0 bin, 1 offsets, 2 scatter, 3 fit. */
#define RPW_PROF_KERNELS 4
typedef struct rpw_profile {
    double ms[RPW_PROF_KERNELS];          /* accumulated since rpw_profile_enable */
    uint64_t launches[RPW_PROF_KERNELS];
    uint32_t fit_grid_blocks;             /* persistent grid of the fit kernel */
    uint32_t fit_smem_points;             /* points a fit block keeps in shared memory */
} rpw_profile;
int rpw_profile_enable(rpw_handle* h, int enable); /* also resets the accumulators */
int rpw_profile_read(rpw_handle* h, rpw_profile* out);

/* Host<->device copy rate of this process on `device`, for bench.py's end-to-end ceiling: `reps` copies of `bytes`
 * between a pinned host buffer and a device buffer on a private stream, timed with CUDA events (seconds for all of
 * them in *seconds).  flags: bit 0 = device-to-host instead of host-to-device, bit 1 = write-combined host buffer
 * (cudaHostAllocWriteCombined), bit 2 = the path's own mix: every host-to-device copy of `bytes` runs together with a
 * device-to-host copy of bytes / 12 on a second stream (12 B of xyz in, 1 B of label out per point), the time is until both
 * are done.  Ranks that call it at the same time measure what they get from the host together. */
int rpw_copy_probe(int device, size_t bytes, int reps, int flags, double* seconds);

/* ---- utilities --------------------------------------------------------------------------- */
void* rpw_host_alloc(size_t bytes); /* pinned host memory (NULL on failure) */
void rpw_host_free(void* p);
/* Kernels launched through this handle since creation (bench.py's gpu_launches). */
uint64_t rpw_kernel_launches(const rpw_handle* h);
/* Single host-path scans (rpw_segment and what calls it) run as one captured CUDA graph per handle -- eleven kernels
 * with fixed parameters, re-captured when the configuration, solver, record layout, stream or capacity changes or a
 * scan outgrows the captured grids.  launches / captures so far; enable = 0 / 1 switches the graph path off / on,
 * negative leaves it as it is (environment: RPW_NO_GRAPH=1). */
int rpw_scan_graph(rpw_handle* h, int enable, uint64_t* launches, uint64_t* captures);
int rpw_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RPW_B200_H */
