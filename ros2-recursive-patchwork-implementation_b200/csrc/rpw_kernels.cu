// csrc/rpw_kernels.cu — the sm_100a kernels of the per-scan ground-segmentation path.
//
//   K1  rpw_bin_kernel      clean + range + radius + angle + ring/sector key, per-block patch
//                           histogram (RP/src/recursive_patchwork.cpp:315-378 steps 1-6a)
//   K1b rpw_offsets_kernel  per-scan exclusive scan of the block histograms -> stable offsets; per-size-class
//                           work lists of the non-empty root patches
//   K2  rpw_scatter_kernel  stable counting-sort scatter of (x, y, z, input index) into
//                           ring/sector patch segments (input order inside every patch, Q1)
//   K3  the plane fit (fitPlaneAndSplit) lives in rpw_fit.cuh and is instantiated by rpw_fit_fast.cu and
//       rpw_fit_replay.cu; this file only dispatches to them
//   K4  rpw_compact_count_kernel / rpw_compact_scatter_kernel   result assembly (:402-419): the ground and
//                           non-ground clouds in the reference's order, a stable compaction by label
//   dbg rpw_eig3_kernel / rpw_normal_kernel / rpw_atan2_kernel   unit-test entry points for the device math
//
// All of it is HBM/L2/shared-memory bound integer-and-float SIMT work; there is no dense
// contraction, so no tensor-core path.  Compiled with -fmad=false (see rpw_device.cuh).
#include "rpw_kernels.h"

namespace rpw {

// =============================================================================================
// K1: binning
// =============================================================================================
// Input records: VEC4 = packed float4 (x, y, z, ignored), one 16-byte load; otherwise any record of
// `stride` 4-byte words with x, y, z at word offsets ox, oy, oz — the reference's 12-byte Point3D
// (stride 3, offsets 0 1 2) and a PointCloud2 data buffer (point_step / 4, field offsets / 4;
// RP/src/recursive_patchwork_node.cpp:67-88, RP/src/rosbag_loader.cpp:226-254) are both this case.
template <bool VEC4>
__device__ __forceinline__ void load_xyz(const float* __restrict__ pts, uint64_t i, const PointLayout& L, float& x, float& y, float& z) {
    if (VEC4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pts) + i);
        x = v.x; y = v.y; z = v.z;
    } else {
        const float* p = pts + i * (uint64_t)L.stride;
        x = __ldg(p + L.ox); y = __ldg(p + L.oy); z = __ldg(p + L.oz);
    }
}

// THREADS: kBinThreads for batches (sixteen points per thread); 1024 for calls of one or two scans (four points per thread:
// a single scan is 30 blocks on 148 SMs, and what it waits for is the length of a thread's own loop, not throughput).
template <bool VEC4, int THREADS>
__global__ void __launch_bounds__(THREADS) rpw_bin_kernel(const float* __restrict__ pts, PointLayout lay, const uint64_t* __restrict__ scan_off,
                                                             const uint32_t* __restrict__ chunk_base, ZoneModel zm,
                                                             uint16_t* __restrict__ keys, uint8_t* __restrict__ labels,
                                                             uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ cls_count,
                                                             const FusionTable* __restrict__ fusion) {
    extern __shared__ uint32_t s_hist[];
    const int b = blockIdx.y, chunk = blockIdx.x;
    // the fit kernel's per-class work lists start empty; the offsets kernel, which runs after this
    // one, fills them
    if (b == 0 && chunk == 0 && threadIdx.x < kClsWords) cls_count[threadIdx.x] = 0;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    const int P = zm.num_patches;
    for (int p = threadIdx.x; p < P; p += THREADS) s_hist[p] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int k = 0; k < kBinChunk / THREADS; ++k) {
        const uint32_t i = base + k * THREADS + threadIdx.x;
        uint16_t key = kKeyDropped;
        const bool valid = i < n;
        if (valid) {
            float x, y, z;
            load_xyz<VEC4>(pts, off + i, lay, x, y, z);
            const bool ego = fusion != nullptr && fuse_point(*fusion, i, x, y);
            key = ego ? kKeyEgo : bin_key(x, y, z, zm);
            keys[off + i] = key;
            // points that never enter a patch get their final label here; patch points are
            // labelled by the fit kernel when their leaf finishes.
            if (key >= kKeySpecialMin) labels[off + i] = key == kKeyDropped ? 3 : (key == kKeyBeyond ? 2 : (key == kKeyEgo ? 4 : 0));
        }
        const uint32_t kk = (valid && key < kKeySpecialMin) ? key : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, kk);
        if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&s_hist[kk], __popc(peers));
    }
    __syncthreads();
    uint32_t* out = blk_hist + ((size_t)chunk_base[b] + chunk) * P;
    for (int p = threadIdx.x; p < P; p += THREADS) out[p] = s_hist[p];
}

// =============================================================================================
// K1b: per-scan offsets.  blk_hist[b][c][p] becomes the exclusive prefix over chunks c;
// patch_start[b][p] = scan base + exclusive prefix over patches; patch_start[b][P] = end.
// =============================================================================================
__global__ void __launch_bounds__(256) rpw_offsets_kernel(const uint64_t* __restrict__ scan_off, const uint32_t* __restrict__ chunk_base,
                                                         uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ patch_start,
                                                         uint32_t* __restrict__ cls_count, uint4* __restrict__ cls_list, uint32_t cls_cap,
                                                         ClassBounds cb, int P) {
    extern __shared__ uint32_t s_cnt[];  // [P] sizes, then [P] starts
    uint32_t* s_start = s_cnt + P;
    __shared__ uint32_t s_cls[kNumFitClasses], s_base[kNumFitClasses];
    const int b = blockIdx.x;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const int chunks = (int)((n + kBinChunk - 1) / kBinChunk);
    uint32_t* h = blk_hist + (size_t)chunk_base[b] * P;
    if (threadIdx.x < kNumFitClasses) s_cls[threadIdx.x] = 0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        uint32_t run = 0;
        int c = 0;
        for (; c + 4 <= chunks; c += 4) {
            const uint32_t v0 = h[(size_t)(c + 0) * P + p], v1 = h[(size_t)(c + 1) * P + p];
            const uint32_t v2 = h[(size_t)(c + 2) * P + p], v3 = h[(size_t)(c + 3) * P + p];
            h[(size_t)(c + 0) * P + p] = run; run += v0;
            h[(size_t)(c + 1) * P + p] = run; run += v1;
            h[(size_t)(c + 2) * P + p] = run; run += v2;
            h[(size_t)(c + 3) * P + p] = run; run += v3;
        }
        for (; c < chunks; ++c) {
            const uint32_t v = h[(size_t)c * P + p];
            h[(size_t)c * P + p] = run;
            run += v;
        }
        s_cnt[p] = run;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t carry = (uint32_t)off;  // sorted segment of scan b starts at its input offset
        for (int p0 = 0; p0 < P; p0 += 32) {
            const int p = p0 + threadIdx.x;
            const uint32_t v = p < P ? s_cnt[p] : 0;
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int)threadIdx.x >= d) inc += t;
            }
            if (p < P) {
                patch_start[(size_t)b * (P + 1) + p] = carry + inc - v;
                s_start[p] = carry + inc - v;
            }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) patch_start[(size_t)b * (P + 1) + P] = carry;
    }
    __syncthreads();
    // Work lists of the level-0 fit: every non-empty patch goes to the list of its size class (one
    // shared-memory counter per class and block, one global atomic per class and block).  The order
    // inside a list is arrival order; patches are independent, so it only affects scheduling.
    constexpr int kPer = (8 * kMaxSectors + 255) / 256;
    uint32_t rank[kPer];
    int cls[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int p = threadIdx.x + k * 256;
        cls[k] = -1;
        if (p < P && s_cnt[p] > 0) {
            int c = 0;
            while (c < kNumFitClasses - 1 && s_cnt[p] > cb.hi[c]) ++c;
            cls[k] = c;
            rank[k] = atomicAdd(&s_cls[c], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kNumFitClasses) s_base[threadIdx.x] = s_cls[threadIdx.x] ? atomicAdd(&cls_count[threadIdx.x], s_cls[threadIdx.x]) : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int p = threadIdx.x + k * 256;
        if (cls[k] >= 0) {
            const uint32_t slot = s_base[cls[k]] + rank[k];
            if (slot < cls_cap) cls_list[(size_t)cls[k] * cls_cap + slot] = make_uint4(s_start[p], s_cnt[p], (uint32_t)b * (uint32_t)P + (uint32_t)p, 0u);
        }
    }
}

// =============================================================================================
// K2: stable scatter.  Inside a block every warp owns a contiguous run of kBinChunk/8 points and
// walks it 32 at a time, so (block, warp, round, lane) order == input order; ranks inside a
// 32-group come from __match_any_sync.  No atomics claim slots, so the result is deterministic
// and stable regardless of scheduling (SURVEY Q1 needs that).
// =============================================================================================
// (64 registers, four blocks per SM.  Forcing five blocks changes nothing, six and more spill the key registers:
// 0.343 / 0.397 / 0.459 ms per 512 scans for 5 / 6 / 8.)
template <bool VEC4, int THREADS>
__global__ void __launch_bounds__(THREADS) rpw_scatter_kernel(const float* __restrict__ pts, PointLayout lay, const uint64_t* __restrict__ scan_off,
                                                                 const uint32_t* __restrict__ chunk_base,
                                                                 const uint16_t* __restrict__ keys, const uint32_t* __restrict__ blk_hist,
                                                                 const uint32_t* __restrict__ patch_start, float4* __restrict__ sorted, int P, const FusionTable* __restrict__ fusion) {
    extern __shared__ uint32_t s_off[];  // [warps][P]
    constexpr int kWarps = THREADS / 32;
    constexpr int kPerWarp = kBinChunk / kWarps;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kWarps * P; i += THREADS) s_off[i] = 0;
    __syncthreads();
    uint32_t* my = s_off + warp * P;
    const uint32_t wbase = base + warp * kPerWarp;
    // the warp's keys are read once, all loads in flight together, and kept for both phases
    constexpr int kRounds = kPerWarp / 32;
    uint32_t kreg[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        uint32_t kk = 0xFFFFFFFFu;
        if (i < n) { const uint16_t key = __ldg(keys + off + i); if (key < kKeySpecialMin) kk = key; }
        kreg[r] = kk;
    }
    // phase 1: per-warp counts
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint32_t kk = kreg[r];
        const unsigned peers = __match_any_sync(0xffffffffu, kk);
        if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) my[kk] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: exclusive prefix over warps + block offset + patch offset
    const uint32_t* bh = blk_hist + ((size_t)chunk_base[b] + chunk) * P;
    const uint32_t* ps = patch_start + (size_t)b * (P + 1);
    for (int p = threadIdx.x; p < P; p += THREADS) {
        uint32_t run = ps[p] + bh[p];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = s_off[w * P + p];
            s_off[w * P + p] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: ranks and scatter, four rounds per trip with their point loads issued together
    constexpr int kGroup = 4;
#pragma unroll
    for (int r0 = 0; r0 < kRounds; r0 += kGroup) {
        float x[kGroup], y[kGroup], z[kGroup];
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
            const uint32_t i = wbase + (r0 + u) * 32 + lane;
            x[u] = y[u] = z[u] = 0.f;
            if (kreg[r0 + u] != 0xFFFFFFFFu) load_xyz<VEC4>(pts, off + i, lay, x[u], y[u], z[u]);
        }
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
            const uint32_t i = wbase + (r0 + u) * 32 + lane;
            const uint32_t kk = kreg[r0 + u];
            const unsigned peers = __match_any_sync(0xffffffffu, kk);
            if (kk != 0xFFFFFFFFu) {
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                const uint32_t pos = my[kk] + rank;
                if (fusion != nullptr) fuse_point(*fusion, i, x[u], y[u]);  // the patches hold vehicle-frame coordinates
                sorted[pos] = make_float4(x[u], y[u], z[u], __uint_as_float((uint32_t)(off + i)));
            }
            __syncwarp();
            if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) my[kk] += __popc(peers);
            __syncwarp();
        }
    }
}

// =============================================================================================
// K4: result assembly on the device (RP/src/recursive_patchwork.cpp:402-419): the two clouds the
// reference returns, in the reference's order -- ground points in input order; non-ground points
// in input order followed by the beyond-radius points in input order.  Stable stream compaction
// by label: K4a counts labels per 4096-point chunk, K4b derives every chunk's bases from the
// counts of the chunks before it (no scan kernel, no atomics claiming slots: the order is
// deterministic) and writes packed 12-byte xyz records.  Fused frames are written in vehicle
// coordinates (the same fuse_point as K1); ego and non-finite points appear in neither cloud.
// Cloud slot i of scan b is record (scan_off[b] + i) of the output buffers.
// =============================================================================================
template <int THREADS>
__global__ void __launch_bounds__(THREADS) rpw_compact_count_kernel(const uint8_t* __restrict__ labels, const uint64_t* __restrict__ scan_off,
                                                                       const uint32_t* __restrict__ chunk_base, uint32_t* __restrict__ cnt) {
    __shared__ uint32_t s_c[3];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    if (threadIdx.x < 3) s_c[threadIdx.x] = 0;
    __syncthreads();
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 4
    for (int k = 0; k < kBinChunk / THREADS; ++k) {
        const uint32_t i = base + k * THREADS + threadIdx.x;
        const uint32_t l = i < n ? labels[off + i] : 255u;
        c0 += l == 0; c1 += l == 1; c2 += l == 2;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, d);
        c1 += __shfl_xor_sync(0xffffffffu, c1, d);
        c2 += __shfl_xor_sync(0xffffffffu, c2, d);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_c[0], c0); atomicAdd(&s_c[1], c1); atomicAdd(&s_c[2], c2); }
    __syncthreads();
    if (threadIdx.x < 3) cnt[((size_t)chunk_base[b] + chunk) * 4 + threadIdx.x] = s_c[threadIdx.x];
}

template <bool VEC4, int THREADS>
__global__ void __launch_bounds__(THREADS) rpw_compact_scatter_kernel(const float* __restrict__ pts, PointLayout lay, const uint8_t* __restrict__ labels,
                                                                         const uint64_t* __restrict__ scan_off, const uint32_t* __restrict__ chunk_base,
                                                                         const uint32_t* __restrict__ cnt, const FusionTable* __restrict__ fusion,
                                                                         float* __restrict__ ground, float* __restrict__ nonground,
                                                                         uint32_t* __restrict__ scan_counts, int packed) {
    constexpr int kWarps = THREADS / 32;
    constexpr int kPerWarp = kBinChunk / kWarps;
    __shared__ uint32_t s_base[5];          // ground, non-ground, beyond bases of this chunk; [3] label-0, [4] label-1 total of the scan
    __shared__ uint32_t s_warp[kWarps][3];  // per-warp counts, then exclusive offsets
    const int b = blockIdx.y, chunk = blockIdx.x;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 5) s_base[threadIdx.x] = 0;
    __syncthreads();
    // bases: label counts of the chunks before this one; total label-0 and beyond counts of the scan
    {
        const int chunks = (int)((n + kBinChunk - 1) / kBinChunk);
        const uint32_t* c = cnt + (size_t)chunk_base[b] * 4;
        uint32_t g = 0, ng = 0, by = 0, ng_all = 0, g_all = 0, by_all = 0;
        for (int q = threadIdx.x; q < chunks; q += THREADS) {
            const uint32_t a0 = c[q * 4 + 0], a1 = c[q * 4 + 1], a2 = c[q * 4 + 2];
            if (q < chunk) { ng += a0; g += a1; by += a2; }
            ng_all += a0; g_all += a1; by_all += a2;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, d); ng += __shfl_xor_sync(0xffffffffu, ng, d); by += __shfl_xor_sync(0xffffffffu, by, d);
            ng_all += __shfl_xor_sync(0xffffffffu, ng_all, d); g_all += __shfl_xor_sync(0xffffffffu, g_all, d);
            by_all += __shfl_xor_sync(0xffffffffu, by_all, d);
        }
        if (lane == 0) {
            atomicAdd(&s_base[0], g); atomicAdd(&s_base[1], ng); atomicAdd(&s_base[2], by); atomicAdd(&s_base[3], ng_all);
            atomicAdd(&s_base[4], g_all);
            if (chunk == 0) { atomicAdd(&scan_counts[2 * b], g_all); atomicAdd(&scan_counts[2 * b + 1], ng_all + by_all); }
        }
    }
    // per-warp counts over the warp's contiguous run of the chunk
    uint32_t wc0 = 0, wc1 = 0, wc2 = 0;
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = base + warp * kPerWarp + r * 32 + lane;
        const uint32_t l = i < n ? labels[off + i] : 255u;
        wc0 += __popc(__ballot_sync(0xffffffffu, l == 0));
        wc1 += __popc(__ballot_sync(0xffffffffu, l == 1));
        wc2 += __popc(__ballot_sync(0xffffffffu, l == 2));
    }
    if (lane == 0) { s_warp[warp][0] = wc0; s_warp[warp][1] = wc1; s_warp[warp][2] = wc2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t run = 0;
        for (int w = 0; w < kWarps; ++w) { const uint32_t v = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = run; run += v; }
    }
    __syncthreads();
    // Fewer than three points inside the radius: the reference returns ({}, cleaned points) before it ever separates
    // the beyond-radius points (RP/src/recursive_patchwork.cpp:339-341), i.e. ONE cloud in plain input order; label-2
    // points then take their place among the label-0 points instead of following them.
    const bool degen = s_base[3] + s_base[4] < 3u;
    // packed: both clouds in ONE buffer, the non-ground cloud right behind the scan's ground cloud (nonground == ground),
    // so that a single copy of (ground + non-ground) records brings both to the host
    const uint32_t ng0 = packed ? s_base[4] : 0u;
    uint32_t o_ng = ng0 + s_base[1] + s_warp[warp][0] + (degen ? s_base[2] + s_warp[warp][2] : 0u);
    uint32_t o_g = s_base[0] + s_warp[warp][1];
    uint32_t o_by = ng0 + s_base[3] + s_base[2] + s_warp[warp][2];  // beyond-radius points follow ALL label-0 points of the scan
    const unsigned lt = (1u << lane) - 1u;
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = base + warp * kPerWarp + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t l = valid ? labels[off + i] : 255u;
        const unsigned m1 = __ballot_sync(0xffffffffu, l == 1), m2 = __ballot_sync(0xffffffffu, l == 2);
        const unsigned m0 = __ballot_sync(0xffffffffu, l == 0) | (degen ? m2 : 0u);
        if (l <= 2u) {
            float x, y, z;
            load_xyz<VEC4>(pts, off + i, lay, x, y, z);
            if (fusion != nullptr) fuse_point(*fusion, i, x, y);
            float* dst;
            if (l == 1u) dst = ground + 3 * (off + o_g + __popc(m1 & lt));
            else if (l == 0u || degen) dst = nonground + 3 * (off + o_ng + __popc(m0 & lt));
            else dst = nonground + 3 * (off + o_by + __popc(m2 & lt));
            dst[0] = x; dst[1] = y; dst[2] = z;
        }
        o_ng += __popc(m0); o_g += __popc(m1); o_by += __popc(m2);
    }
}

// =============================================================================================
// K5: sampleGroundAndObstacles post-filter (RP/src/recursive_patchwork.cpp:428-465) on the two clouds K4
// left on the device: obstacles = non-ground points outside the ego radius (computeDistance2D > r, :64-75)
// whose height is within base_tol of target_height, order kept (the same stable compaction as K4 with a
// predicate instead of a label); ground context = the ground points at indices the host drew.
// =============================================================================================
__device__ __forceinline__ bool is_obstacle(const float* __restrict__ p, float target, float tol, float ego) {
    return range2d(p[0], p[1]) > ego && fabsf(p[2] - target) <= tol;
}

__global__ void __launch_bounds__(kBinThreads) rpw_obstacle_count_kernel(const float* __restrict__ xyz, uint32_t n, float target, float tol, float ego,
                                                                        uint32_t* __restrict__ cnt) {
    __shared__ uint32_t s_c;
    const uint32_t base = blockIdx.x * kBinChunk;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    uint32_t c = 0;
    for (int k = 0; k < kBinChunk / kBinThreads; ++k) {
        const uint32_t i = base + k * kBinThreads + threadIdx.x;
        c += i < n && is_obstacle(xyz + 3 * (size_t)i, target, tol, ego);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) cnt[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(kBinThreads) rpw_obstacle_scatter_kernel(const float* __restrict__ xyz, uint32_t n, float target, float tol, float ego,
                                                                          const uint32_t* __restrict__ cnt, float* __restrict__ out,
                                                                          uint32_t* __restrict__ total) {
    constexpr int kWarps = kBinThreads / 32;
    constexpr int kPerWarp = kBinChunk / kWarps;
    __shared__ uint32_t s_base, s_warp[kWarps];
    const uint32_t base = blockIdx.x * kBinChunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    {
        uint32_t before = 0, all = 0;
        for (uint32_t q = threadIdx.x; q < gridDim.x; q += kBinThreads) { const uint32_t v = cnt[q]; all += v; if (q < blockIdx.x) before += v; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, d); all += __shfl_xor_sync(0xffffffffu, all, d); }
        if (lane == 0) { atomicAdd(&s_base, before); if (blockIdx.x == 0) atomicAdd(total, all); }
    }
    uint32_t wc = 0;
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = base + warp * kPerWarp + r * 32 + lane;
        wc += __popc(__ballot_sync(0xffffffffu, i < n && is_obstacle(xyz + 3 * (size_t)i, target, tol, ego)));
    }
    if (lane == 0) s_warp[warp] = wc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < kWarps; ++w) { const uint32_t v = s_warp[w]; s_warp[w] = run; run += v; }
    }
    __syncthreads();
    uint32_t o = s_base + s_warp[warp];
    const unsigned lt = (1u << lane) - 1u;
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = base + warp * kPerWarp + r * 32 + lane;
        const bool keep = i < n && is_obstacle(xyz + 3 * (size_t)i, target, tol, ego);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            float* dst = out + 3 * (size_t)(o + __popc(m & lt));
            const float* src = xyz + 3 * (size_t)i;
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        }
        o += __popc(m);
    }
}

__global__ void rpw_gather_xyz_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ idx, uint32_t k, float* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const float* src = xyz + 3 * (size_t)idx[j];
    out[3 * (size_t)j] = src[0]; out[3 * (size_t)j + 1] = src[1]; out[3 * (size_t)j + 2] = src[2];
}

// =============================================================================================
// K6: bird's-eye-view rasters (RP/src/visualization.cpp:18-113) from K4's device-resident clouds.  The reference
// draws the points one after the other, so the LAST point that falls on a pixel colours it; here every
// point bids for its pixel with its draw-order index (atomicMax) and a second kernel colours each pixel
// from the winner.  Pixel coordinates with the reference's operations: int((p - min) * scale), truncation.
// =============================================================================================
__global__ void rpw_bev_bid_kernel(const float* __restrict__ xyz, uint32_t n, uint32_t order0, int width, int height, float x_min, float y_min,
                                   float x_scale, float y_scale, uint32_t* __restrict__ owner) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int x = (int)((xyz[3 * (size_t)j] - x_min) * x_scale);
    const int y = (int)((xyz[3 * (size_t)j + 1] - y_min) * y_scale);
    if (x >= 0 && x < width && y >= 0 && y < height) atomicMax(&owner[(size_t)y * width + x], order0 + j + 1u);
}

// mode 0: ground green / non-ground red (createGroundNonGroundImage, :47-80); mode 1: height colouring
// (createBEVImage, :18-45).  Draw order: the n_first points of cloud a, then cloud b.
__global__ void rpw_bev_paint_kernel(const uint32_t* __restrict__ owner, int n_pixels, int mode, const float* __restrict__ a, uint32_t n_first,
                                     const float* __restrict__ b, uint8_t* __restrict__ bgr) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    const uint32_t o = owner[p];
    uint8_t c0 = 0, c1 = 0, c2 = 0;  // cv::Scalar(0, 0, 0) background
    if (o) {
        const uint32_t j = o - 1u;
        if (mode == 0) {
            if (j < n_first) c1 = 255; else c2 = 255;  // Vec3b(0, 255, 0) / Vec3b(0, 0, 255)
        } else {
            const float z = j < n_first ? a[3 * (size_t)j + 2] : b[3 * (size_t)(j - n_first) + 2];
            const int intensity = (int)fminf(255.0f, fmaxf(0.0f, (z + 2.0f) * 50.0f));
            c0 = (uint8_t)intensity; c1 = (uint8_t)intensity; c2 = 255;
        }
    }
    bgr[3 * (size_t)p] = c0; bgr[3 * (size_t)p + 1] = c1; bgr[3 * (size_t)p + 2] = c2;
}

// =============================================================================================
// unit-test entry points
// =============================================================================================
__global__ void rpw_eig3_kernel(const float* __restrict__ mats, size_t count, float* __restrict__ evals, float* __restrict__ evecs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float* a = mats + i * 9;
    const Eig3 E = eig3_sym(a[0], a[3], a[4], a[6], a[7], a[8]);
    for (int k = 0; k < 3; ++k) evals[i * 3 + k] = E.val[k];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) evecs[i * 9 + r * 3 + c] = E.vec[r][c];
}

// One warp per scatter matrix (xx yx yy zx zy zz), the way the fit kernel calls the solvers; also
// reports the cycles one call took (mode 0: closed-form FP64, mode 1: Eigen's QR sequence).
__global__ void rpw_normal_kernel(const float* __restrict__ sc, size_t count, int mode, float* __restrict__ normals,
                                  uint32_t* __restrict__ cycles) {
    const size_t i = blockIdx.x;
    if (i >= count) return;
    const float* a = sc + i * 6;
    float s0 = a[0], s1 = a[1], s2 = a[2], s3 = a[3], s4 = a[4], s5 = a[5];
    float nx = 0.f, ny = 0.f, nz = 0.f;
    long long t0 = 0, t1 = 0;
    // two passes: the first warms the instruction cache, the second is the one timed and reported
    // (its input depends on the first result so that the compiler cannot merge them)
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {
        __syncwarp();
        t0 = clock64();
        if (mode == 1) {
            const Eig3 E = eig3_sym(s0, s1, s2, s3, s4, s5);
            nx = E.vec[0][0]; ny = E.vec[1][0]; nz = E.vec[2][0];
        } else if (mode == 2) {  // the latency-restructured QR on the compiler's IEEE operations (must equal mode 1 bit for bit)
            eig3_smallest_qr(s0, s1, s2, s3, s4, s5, nx, ny, nz);
        } else if (mode == 3) {  // what the fit kernel runs: branch-free arithmetic + IEEE fallback (must equal mode 1 too)
            ArithSpec ar;
            eig3_smallest_qr_t(ar, s0, s1, s2, s3, s4, s5, nx, ny, nz);
            if (!ar.ok()) eig3_smallest_qr_slow(s0, s1, s2, s3, s4, s5, nx, ny, nz);
        } else if (mode == 4) {  // mode 3 without the fallback: nz = 2 marks solves whose operands left the fast range
            ArithSpec ar;
            eig3_smallest_qr_t(ar, s0, s1, s2, s3, s4, s5, nx, ny, nz);
            if (!ar.ok()) { nx = 0.f; ny = 0.f; nz = 2.f; }
        } else {
            smallest_eigvec_psd(s0, s1, s2, s3, s4, s5, nx, ny, nz);
        }
        if (nz < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        t1 = clock64();
        if (nx != nx) s0 = nx;  // never true for finite input, but ties the second pass to the first
    }
    if (threadIdx.x == 0) {
        normals[i * 3] = nx; normals[i * 3 + 1] = ny; normals[i * 3 + 2] = nz;
        cycles[i] = (uint32_t)(t1 - t0);
    }
}

__global__ void rpw_atan2_kernel(const float* __restrict__ y, const float* __restrict__ x, size_t count, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = atan2f_libm(y[i], x[i]);
}

// =============================================================================================
// host-side launchers
// =============================================================================================
cudaError_t launch_bin(cudaStream_t st, const PointLayout& lay, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                       const ZoneModel& zm, uint16_t* keys, uint8_t* labels, uint32_t* blk_hist, uint32_t* cls_count,
                       const FusionTable* fusion, int max_chunks, int batch, int threads) {
    dim3 grid(max_chunks, batch);
    const size_t smem = (size_t)zm.num_patches * 4;
    if (threads == 1024) {
        if (lay.vec4) rpw_bin_kernel<true, 1024><<<grid, 1024, smem, st>>>(pts, lay, scan_off, chunk_base, zm, keys, labels, blk_hist, cls_count, fusion);
        else rpw_bin_kernel<false, 1024><<<grid, 1024, smem, st>>>(pts, lay, scan_off, chunk_base, zm, keys, labels, blk_hist, cls_count, fusion);
    } else {
        if (lay.vec4) rpw_bin_kernel<true, kBinThreads><<<grid, kBinThreads, smem, st>>>(pts, lay, scan_off, chunk_base, zm, keys, labels, blk_hist, cls_count, fusion);
        else rpw_bin_kernel<false, kBinThreads><<<grid, kBinThreads, smem, st>>>(pts, lay, scan_off, chunk_base, zm, keys, labels, blk_hist, cls_count, fusion);
    }
    return cudaGetLastError();
}

cudaError_t launch_offsets(cudaStream_t st, const uint64_t* scan_off, const uint32_t* chunk_base, uint32_t* blk_hist,
                           uint32_t* patch_start, uint32_t* cls_count, uint4* cls_list, uint32_t cls_cap, int P, int batch, int profile) {
    rpw_offsets_kernel<<<batch, 256, (size_t)(2 * P) * 4, st>>>(scan_off, chunk_base, blk_hist, patch_start, cls_count, cls_list, cls_cap,
                                                                fit_class_bounds(profile), P);
    return cudaGetLastError();
}

cudaError_t launch_scatter(cudaStream_t st, const PointLayout& lay, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                           const uint16_t* keys, const uint32_t* blk_hist, const uint32_t* patch_start, float4* sorted,
                           int P, const FusionTable* fusion, int max_chunks, int batch, int threads) {
    dim3 grid(max_chunks, batch);
    const size_t smem = (size_t)(threads / 32) * P * 4;
    if (threads == 1024) {
        if (lay.vec4) rpw_scatter_kernel<true, 1024><<<grid, 1024, smem, st>>>(pts, lay, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, P, fusion);
        else rpw_scatter_kernel<false, 1024><<<grid, 1024, smem, st>>>(pts, lay, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, P, fusion);
    } else {
        if (lay.vec4) rpw_scatter_kernel<true, kBinThreads><<<grid, kBinThreads, smem, st>>>(pts, lay, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, P, fusion);
        else rpw_scatter_kernel<false, kBinThreads><<<grid, kBinThreads, smem, st>>>(pts, lay, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, P, fusion);
    }
    return cudaGetLastError();
}

cudaError_t launch_compact(cudaStream_t st, const PointLayout& lay, const float* pts, const uint8_t* labels, const uint64_t* scan_off,
                           const uint32_t* chunk_base, uint32_t* cnt, const FusionTable* fusion, float* ground, float* nonground,
                           uint32_t* scan_counts, int max_chunks, int batch, int packed, int threads) {
    dim3 grid(max_chunks, batch);
    cudaError_t e = cudaMemsetAsync(scan_counts, 0, (size_t)batch * 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (threads == 1024) {  // calls of one or two scans: four points per thread instead of sixteen, like K1 / K2
        rpw_compact_count_kernel<1024><<<grid, 1024, 0, st>>>(labels, scan_off, chunk_base, cnt);
        if (lay.vec4) rpw_compact_scatter_kernel<true, 1024><<<grid, 1024, 0, st>>>(pts, lay, labels, scan_off, chunk_base, cnt, fusion, ground, nonground, scan_counts, packed);
        else rpw_compact_scatter_kernel<false, 1024><<<grid, 1024, 0, st>>>(pts, lay, labels, scan_off, chunk_base, cnt, fusion, ground, nonground, scan_counts, packed);
        return cudaGetLastError();
    }
    rpw_compact_count_kernel<kBinThreads><<<grid, kBinThreads, 0, st>>>(labels, scan_off, chunk_base, cnt);
    if (lay.vec4) rpw_compact_scatter_kernel<true, kBinThreads><<<grid, kBinThreads, 0, st>>>(pts, lay, labels, scan_off, chunk_base, cnt, fusion, ground, nonground, scan_counts, packed);
    else rpw_compact_scatter_kernel<false, kBinThreads><<<grid, kBinThreads, 0, st>>>(pts, lay, labels, scan_off, chunk_base, cnt, fusion, ground, nonground, scan_counts, packed);
    return cudaGetLastError();
}

cudaError_t launch_bev(cudaStream_t st, int mode, const float* a, uint32_t n_a, const float* b, uint32_t n_b, int width, int height,
                       float x_min, float y_min, float x_scale, float y_scale, uint32_t* owner, uint8_t* bgr) {
    const int n_pixels = width * height;
    cudaError_t e = cudaMemsetAsync(owner, 0, (size_t)n_pixels * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (n_a) rpw_bev_bid_kernel<<<(n_a + 255) / 256, 256, 0, st>>>(a, n_a, 0u, width, height, x_min, y_min, x_scale, y_scale, owner);
    if (n_b) rpw_bev_bid_kernel<<<(n_b + 255) / 256, 256, 0, st>>>(b, n_b, n_a, width, height, x_min, y_min, x_scale, y_scale, owner);
    rpw_bev_paint_kernel<<<(n_pixels + 255) / 256, 256, 0, st>>>(owner, n_pixels, mode, a, n_a, b, bgr);
    return cudaGetLastError();
}

cudaError_t launch_obstacles(cudaStream_t st, const float* nonground_xyz, uint32_t n, float target, float tol, float ego, uint32_t* cnt,
                             float* out, uint32_t* total) {
    cudaError_t e = cudaMemsetAsync(total, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    const unsigned grid = (n + kBinChunk - 1) / kBinChunk;
    rpw_obstacle_count_kernel<<<grid, kBinThreads, 0, st>>>(nonground_xyz, n, target, tol, ego, cnt);
    rpw_obstacle_scatter_kernel<<<grid, kBinThreads, 0, st>>>(nonground_xyz, n, target, tol, ego, cnt, out, total);
    return cudaGetLastError();
}

cudaError_t launch_gather_xyz(cudaStream_t st, const float* xyz, const uint32_t* idx, uint32_t k, float* out) {
    if (k == 0) return cudaSuccess;
    rpw_gather_xyz_kernel<<<(k + 255) / 256, 256, 0, st>>>(xyz, idx, k, out);
    return cudaGetLastError();
}

// The fit kernels live in two translation units (rpw_fit.cuh): the default ones and the ones that carry the
// reference-order arithmetic; the latter run whenever it is switched on.
cudaError_t fit_configure_fast(int smem_cap, int* blocks_per_sm);
cudaError_t fit_configure_replay(int smem_cap, int* blocks_per_sm);
cudaError_t launch_fit_roots_fast(cudaStream_t st, const FitArgs& args, int cls, unsigned grid);
cudaError_t launch_fit_roots_replay(cudaStream_t st, const FitArgs& args, int cls, unsigned grid);
cudaError_t launch_fit_levels_fast(cudaStream_t st, const FitArgs& args, int grid_blocks);
cudaError_t launch_fit_levels_replay(cudaStream_t st, const FitArgs& args, int grid_blocks);

cudaError_t fit_configure(int smem_cap, int* blocks_per_sm) {
    int a = 0, b = 0;
    cudaError_t e = fit_configure_fast(smem_cap, &a);
    if (e != cudaSuccess) return e;
    if ((e = fit_configure_replay(smem_cap, &b)) != cudaSuccess) return e;
    *blocks_per_sm = a < b ? a : b;
    return cudaSuccess;
}

cudaError_t launch_fit_roots(cudaStream_t st, const FitArgs& args, int cls, unsigned grid) {
    return args.fp.exact_replay >= 0 ? launch_fit_roots_replay(st, args, cls, grid) : launch_fit_roots_fast(st, args, cls, grid);
}

cudaError_t launch_fit_levels(cudaStream_t st, const FitArgs& args, int grid_blocks) {
    return args.fp.exact_replay >= 0 ? launch_fit_levels_replay(st, args, grid_blocks) : launch_fit_levels_fast(st, args, grid_blocks);
}

cudaError_t launch_eig3(cudaStream_t st, const float* mats, size_t count, float* evals, float* evecs) {
    rpw_eig3_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(mats, count, evals, evecs);
    return cudaGetLastError();
}

cudaError_t launch_normal(cudaStream_t st, const float* sc, size_t count, int mode, float* normals, uint32_t* cycles) {
    rpw_normal_kernel<<<(unsigned)count, 32, 0, st>>>(sc, count, mode, normals, cycles);
    return cudaGetLastError();
}

cudaError_t launch_atan2(cudaStream_t st, const float* y, const float* x, size_t count, float* out) {
    rpw_atan2_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(y, x, count, out);
    return cudaGetLastError();
}

}  // namespace rpw
