// synth/scangen.cpp — seeded synthetic point clouds of the shapes BASELINE.json names.
//
// Host-side input generation for tests and bench.py (no device code, not on the hot path).
// Every generator is deterministic in its seed and writes (x, y, z, 0) float4 records, the
// layout both the CUDA path and the CPU oracle consume.
//
//   C1  rpw_synth_testsuite   the reference test-suite cloud: same distributions and draw
//                             order as the generator in RP/test/test_recursive_patchwork.cpp:12-49
//                             (70 % ground x,y~U(-50,50) z~N(0,0.05); 30 % obstacles
//                             x,y~U(-30,30) z~U(0.5,3)), but seeded (the reference uses
//                             std::random_device, so it has no reproducible fixture).
//   C2  rpw_synth_spinning    64-beam x 1875-step spinning scan, ~120k returns, R = 80 m
//   C3                        = C2 with seeds base..base+4095 (bench.py / tests loop over seeds)
//   C4  rpw_synth_solidstate  3 x 120-degree solid-state sensors merged, banked track, ~300k
//   C5  rpw_synth_spinning    with beams=128, steps=2048, clutter=1: dense urban, two-layer
//                             ground that makes first fits collapse and patches split deep
//
// z is ground-referenced (ground near z = 0) as in the reference's own generator, because the
// reference's seed rule is an absolute-z test against sensor_height (recursive_patchwork.cpp:153).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846;

// Small counter-free PRNG helpers on top of mt19937 raw output (portable across libstdc++ builds;
// C1 alone uses the std:: distributions because it has to mirror the reference's generator).
struct Rng {
    std::mt19937 g;
    explicit Rng(uint32_t seed) : g(seed) {}
    double uni() { return (g() >> 5) * (1.0 / 134217728.0); }  // [0,1) 27 bits
    double uni(double a, double b) { return a + (b - a) * uni(); }
    int irange(int a, int b) { return a + int(uni() * (b - a + 1)); }  // inclusive
    bool has_spare = false;
    double spare = 0.0;
    double normal() {
        if (has_spare) { has_spare = false; return spare; }
        double u, v, s;
        do { u = 2.0 * uni() - 1.0; v = 2.0 * uni() - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
        const double m = std::sqrt(-2.0 * std::log(s) / s);
        spare = v * m; has_spare = true;
        return u * m;
    }
};

struct Box {  // yaw-rotated box standing on the (local) ground
    double cx, cy, z0, z1, hx, hy, c, s;  // centre, z range, half extents, cos/sin yaw
    double amin, amax;                    // azimuth interval seen from the sensor (may wrap)
    bool wraps;
};

void box_azimuth(Box& b) {
    double lo = 1e9, hi = -1e9;
    const double a0 = std::atan2(b.cy, b.cx);
    for (int k = 0; k < 4; ++k) {
        const double lx = (k & 1 ? b.hx : -b.hx), ly = (k & 2 ? b.hy : -b.hy);
        const double wx = b.cx + b.c * lx - b.s * ly, wy = b.cy + b.s * lx + b.c * ly;
        double da = std::atan2(wy, wx) - a0;
        while (da > kPi) da -= 2 * kPi;
        while (da < -kPi) da += 2 * kPi;
        lo = std::min(lo, da); hi = std::max(hi, da);
    }
    b.amin = a0 + lo - 1e-3; b.amax = a0 + hi + 1e-3;
    b.wraps = false;
}

inline bool az_in(const Box& b, double az) {
    double d = az - b.amin;
    d -= 2 * kPi * std::floor(d / (2 * kPi));
    return d <= (b.amax - b.amin);
}

// Ray (origin o, unit direction d) against a box; returns t of the first hit or +inf.
double ray_box(const Box& b, const double o[3], const double d[3]) {
    // into box frame
    const double ox = o[0] - b.cx, oy = o[1] - b.cy;
    const double lx = b.c * ox + b.s * oy, ly = -b.s * ox + b.c * oy;
    const double dx = b.c * d[0] + b.s * d[1], dy = -b.s * d[0] + b.c * d[1];
    double t0 = 0.0, t1 = 1e30;
    const double lo[3] = {-b.hx, -b.hy, b.z0}, hi[3] = {b.hx, b.hy, b.z1};
    const double oo[3] = {lx, ly, o[2]}, dd[3] = {dx, dy, d[2]};
    for (int k = 0; k < 3; ++k) {
        if (std::fabs(dd[k]) < 1e-12) {
            if (oo[k] < lo[k] || oo[k] > hi[k]) return 1e30;
        } else {
            double ta = (lo[k] - oo[k]) / dd[k], tb = (hi[k] - oo[k]) / dd[k];
            if (ta > tb) std::swap(ta, tb);
            t0 = std::max(t0, ta); t1 = std::min(t1, tb);
            if (t0 > t1) return 1e30;
        }
    }
    return t0 > 1e-6 ? t0 : 1e30;
}

}  // namespace

extern "C" {

// C1.  out: n float4 records.  Draw order and distributions as the reference generator.
void rpw_synth_testsuite(uint32_t seed, size_t n, float* out) {
    std::mt19937 gen(seed);
    std::normal_distribution<float> ground_z(0.0f, 0.05f);
    std::uniform_real_distribution<float> ground_xy(-50.0f, 50.0f);
    std::uniform_real_distribution<float> obstacle_xy(-30.0f, 30.0f);
    std::uniform_real_distribution<float> obstacle_z(0.5f, 3.0f);
    const size_t n_ground = static_cast<size_t>(n * 0.7);
    size_t i = 0;
    for (; i < n_ground; ++i) {
        out[4 * i + 0] = ground_xy(gen);
        out[4 * i + 1] = ground_xy(gen);
        out[4 * i + 2] = ground_z(gen);
        out[4 * i + 3] = 0.f;
    }
    for (; i < n; ++i) {
        out[4 * i + 0] = obstacle_xy(gen);
        out[4 * i + 1] = obstacle_xy(gen);
        out[4 * i + 2] = obstacle_z(gen);
        out[4 * i + 3] = 0.f;
    }
}

// C2 / C3 / C5.  Spinning multi-beam LiDAR, firing order azimuth-major (each step fires all
// beams).  beams x steps records are ALWAYS written: every ray returns something (ground,
// an obstacle box, or a building backdrop 60-110 m away, which lands beyond R = 80 m for part
// of the azimuth range and so exercises the beyond-radius path); `nan_per_million` of the
// returns are replaced by NaN records to exercise the cleaning path.
//   clutter = 0: planar ground (slightly tilted per seed) + 25-40 boxes (cars, poles, walls).
//   clutter = 1: gently undulating ground (40-90 m wavelengths) with a porous second layer 0.85-1.05 m
//                above it in six of the ten 36-degree sectors (kerbs / vegetation / vehicle bodies; half
//                of the beams there return from the layer), 15-25 boxes within 40 m, backdrop beyond R.  In a layered
//                sector the first plane fit lands between the two sheets, finds no inliers, collapses,
//                and the patch splits until its bbox drops under 25 m^2 (depth 4-6 in the outer rings).
// Returns the number of records written (= beams * steps).
size_t rpw_synth_spinning(uint32_t seed, int beams, int steps, int clutter, int nan_per_million, float* out) {
    Rng rng(seed * 2654435761u + 12345u);
    const double h = 1.73;  // sensor height above ground
    const double el_lo = (beams > 64 ? -25.0 : -24.8) * kPi / 180.0, el_hi = 2.0 * kPi / 180.0;
    const double pitch = rng.uni(-1.0, 1.0) * kPi / 180.0 * (clutter ? 0.3 : 1.0);
    const double roll = rng.uni(-1.0, 1.0) * kPi / 180.0 * (clutter ? 0.3 : 1.0);
    const double gx = std::tan(pitch), gy = std::tan(roll);  // ground z = gx*x + gy*y (+ undulation)

    // undulation: a few random sinusoids (clutter only)
    double ua[4], ukx[4], uky[4], uph[4];
    for (int k = 0; k < 4; ++k) {
        ua[k] = clutter ? rng.uni(0.02, 0.045) : 0.0;
        const double wl = rng.uni(40.0, 90.0), th = rng.uni(0, 2 * kPi);
        ukx[k] = 2 * kPi / wl * std::cos(th); uky[k] = 2 * kPi / wl * std::sin(th);
        uph[k] = rng.uni(0, 2 * kPi);
    }
    auto ground_z = [&](double x, double y) {
        double z = gx * x + gy * y;
        for (int k = 0; k < 4; ++k) z += ua[k] * std::sin(ukx[k] * x + uky[k] * y + uph[k]);
        return z;
    };
    // second-layer coverage (clutter only): six of the ten 36-degree sectors of the default zone model
    // (chosen per seed) carry a porous layer 0.85-1.05 m above the ground, all ranges, each with its
    // own height; inside such a sector every ground return comes from the layer with probability one
    // half.  Whole ring/sector patches are therefore two parallel sheets: the first plane fit lands
    // exactly between them, finds no inliers, collapses, and the patch splits until its
    // bounding box drops under 25 m^2 — depth 4-6 in the outer rings.
    const int n_blobs = clutter ? 10 : 0;
    std::vector<double> bh(n_blobs, 0.0);  // layer height per sector, 0 = no layer
    if (clutter) {
        int order[10];
        for (int k = 0; k < 10; ++k) order[k] = k;
        for (int k = 9; k > 0; --k) { const int j = rng.irange(0, k); std::swap(order[k], order[j]); }
        for (int k = 0; k < 6; ++k) bh[order[k]] = rng.uni(0.85, 1.05);
    }
    const int n_boxes = clutter ? rng.irange(15, 25) : rng.irange(25, 40);
    std::vector<Box> boxes(n_boxes);
    for (auto& b : boxes) {
        const double r = rng.uni(clutter ? 9.0 : 5.0, clutter ? 40.0 : 60.0), a = rng.uni(0, 2 * kPi), yaw = rng.uni(0, kPi);
        b.cx = r * std::cos(a); b.cy = r * std::sin(a);
        const double kind = rng.uni();
        double L, W, H;
        if (kind < 0.55) { L = 4.0; W = 1.8; H = 1.5; }                                   // car
        else if (kind < 0.75) { L = 0.3; W = 0.3; H = rng.uni(3.0, 6.0); }                 // pole
        else if (kind < 0.9) { L = rng.uni(6.0, 14.0); W = 0.3; H = rng.uni(1.5, 3.0); }   // wall
        else { L = rng.uni(5.0, 9.0); W = 2.5; H = rng.uni(2.5, 3.5); }                    // van / truck
        b.hx = L / 2; b.hy = W / 2; b.c = std::cos(yaw); b.s = std::sin(yaw);
        b.z0 = ground_z(b.cx, b.cy) - 0.05; b.z1 = b.z0 + H;
        box_azimuth(b);
    }
    // building backdrop: radius as a smooth function of azimuth, 60..110 m, height 25 m
    double ba[3], bp[3];
    for (int k = 0; k < 3; ++k) { ba[k] = rng.uni(4.0, 9.0); bp[k] = rng.uni(0, 2 * kPi); }
    auto backdrop_r = [&](double az) {
        // clutter scenes keep the backdrop beyond R = 80 m so that the outer rings stay two clean sheets
        return (clutter ? 110.0 : 85.0) + ba[0] * std::sin(az + bp[0]) + ba[1] * std::sin(2 * az + bp[1]) + ba[2] * std::sin(5 * az + bp[2]);
    };

    const double o[3] = {0.0, 0.0, h};
    std::vector<int> cand;
    size_t w = 0;
    const double az0 = rng.uni(0, 2 * kPi);
    for (int s = 0; s < steps; ++s) {
        const double az = az0 + 2 * kPi * s / steps;
        cand.clear();
        for (int k = 0; k < n_boxes; ++k) if (az_in(boxes[k], az)) cand.push_back(k);
        const double ca = std::cos(az), sa = std::sin(az);
        const double rb = backdrop_r(az);
        for (int b = 0; b < beams; ++b) {
            const double el = el_lo + (el_hi - el_lo) * b / (beams - 1);
            const double ce = std::cos(el), se = std::sin(el);
            const double d[3] = {ce * ca, ce * sa, se};
            // ground: f(t) = h + t*se - ground_z(t*dx, t*dy) = 0, a few quasi-Newton steps using
            // the tilted-plane slope (the undulation slopes are small)
            double tg = 1e30;
            {
                const double denom = se - (gx * d[0] + gy * d[1]);
                if (denom < -1e-6) {
                    double t = -h / denom;
                    for (int it = 0; it < 4; ++it) t -= (h + t * se - ground_z(t * d[0], t * d[1])) / denom;
                    if (t > 0) tg = t;
                }
            }
            double t = tg;
            int what = 0;  // 0 ground, 1 box, 2 backdrop
            for (int k : cand) {
                const double tb = ray_box(boxes[k], o, d);
                if (tb < t) { t = tb; what = 1; }
            }
            const double tbk = rb / ce;
            if (tbk < t) { t = tbk; what = 2; }
            double x = t * d[0], y = t * d[1], z = h + t * d[2];
            if (what == 0) {
                // second layer: inside a layered sector half of the beams return from the layer
                bool layered = false;
                if (n_blobs) {
                    double a = std::atan2(y, x);
                    if (a < 0) a += 2 * kPi;
                    int sec = (int)(a / (2 * kPi / 10.0));
                    sec = sec < 0 ? 0 : (sec > 9 ? 9 : sec);
                    if (bh[sec] > 0.0) {
                        if (rng.uni() < 0.5) z += bh[sec] + 0.03 * rng.normal();
                        else z += 0.02 * rng.normal();
                        layered = true;
                    }
                }
                if (!layered) z += 0.02 * rng.normal();
            } else {
                // range noise along the ray
                const double dn = 0.02 * rng.normal();
                x += dn * d[0]; y += dn * d[1]; z += dn * d[2];
            }
            float fx = (float)x, fy = (float)y, fz = (float)z;
            if (nan_per_million > 0 && (rng.g() % 1000000u) < (uint32_t)nan_per_million) {
                const uint32_t which = rng.g() % 3u;
                const float bad = (rng.g() & 1u) ? NAN : INFINITY;
                if (which == 0) fx = bad; else if (which == 1) fy = bad; else fz = bad;
            }
            out[4 * w + 0] = fx; out[4 * w + 1] = fy; out[4 * w + 2] = fz; out[4 * w + 3] = 0.f;
            ++w;
        }
    }
    return w;
}

// C4.  Three solid-state sensors (120-degree horizontal FOV each, yaw 0 / +120 / -120 as
// RP/src/lidar_fusion.cpp:20-36), cols x rows rays each, already rotated into the vehicle
// frame, merged sensor after sensor, ego returns (r <= 2.5 m, RP/src/lidar_fusion.cpp:184-187)
// removed.  Banked track: ground z = tan(bank) * y with bank 9-20 degrees, 1 m outer walls at
// lateral |y| = 8-12 m, catch fence / grandstand backdrop beyond.  Rays with no return inside
// max_range are dropped, so the count varies with the seed (~300k for 3 x 400 x 260).
// Returns the number of records written (<= 3 * cols * rows).
size_t rpw_synth_solidstate(uint32_t seed, int cols, int rows, float max_range, float* out) {
    Rng rng(seed * 2246822519u + 777u);
    const double h = 0.9;
    const double bank = rng.uni(9.0, 20.0) * kPi / 180.0 * (rng.uni() < 0.5 ? -1.0 : 1.0);
    const double tb = std::tan(bank);
    const double wall_l = rng.uni(8.0, 12.0), wall_r = -rng.uni(8.0, 12.0);
    const double stand_l = wall_l + rng.uni(6.0, 25.0), stand_r = wall_r - rng.uni(6.0, 25.0);
    const double heading = rng.uni(-0.05, 0.05);  // small yaw of the car w.r.t. the track axis
    const double ch = std::cos(heading), sh = std::sin(heading);
    const double yaws[3] = {0.0, 120.0 * kPi / 180.0, -120.0 * kPi / 180.0};
    const double el_lo = -22.0 * kPi / 180.0, el_hi = 4.0 * kPi / 180.0;
    // other cars on track
    const int n_boxes = rng.irange(3, 8);
    std::vector<Box> boxes(n_boxes);
    for (auto& b : boxes) {
        // track frame position
        const double tx = rng.uni(-80.0, 120.0), ty = rng.uni(wall_r + 1.5, wall_l - 1.5);
        b.cx = ch * tx - sh * ty; b.cy = sh * tx + ch * ty;
        b.hx = 2.4; b.hy = 0.95; b.c = ch; b.s = sh;
        b.z0 = tb * ty - 0.02; b.z1 = b.z0 + 1.1;
        box_azimuth(b);
    }
    size_t w = 0;
    const double o[3] = {0.0, 0.0, h};
    for (int sidx = 0; sidx < 3; ++sidx) {
        for (int r = 0; r < rows; ++r) {
            const double el = el_hi - (el_hi - el_lo) * r / (rows - 1);
            const double ce = std::cos(el), se = std::sin(el);
            for (int c = 0; c < cols; ++c) {
                const double az = yaws[sidx] + (-60.0 + 120.0 * (c + 0.5) / cols) * kPi / 180.0;
                const double d[3] = {ce * std::cos(az), ce * std::sin(az), se};
                // track-frame direction (rotate by -heading)
                const double dty = -sh * d[0] + ch * d[1];
                // banked plane z = tb * y_track, origin at y_track = 0
                double t = 1e30;
                int what = -1;
                const double denom = d[2] - tb * dty;
                if (denom < -1e-9) { t = -h / denom; what = 0; }
                // walls: planes y_track = wall_l / wall_r, 1 m above local ground
                for (int k = 0; k < 2; ++k) {
                    const double yw = k ? wall_r : wall_l;
                    if (std::fabs(dty) < 1e-9) continue;
                    const double tw = yw / dty;
                    if (tw > 1e-6 && tw < t) {
                        const double z = h + tw * d[2];
                        const double zg = tb * yw;
                        if (z >= zg - 0.05 && z <= zg + 1.0) { t = tw; what = 1; }
                        else if (z > zg + 1.0 && z <= zg + 6.0 && rng.uni() < 0.35) { t = tw + rng.uni(0.5, 3.0); what = 2; }  // catch fence, sparse
                        else if (z < zg - 0.05) { /* below ground on that side: ground hit comes first */ }
                    }
                }
                // grandstands / buildings behind the fences: vertical surfaces at |y_track| = stand_l / stand_r
                for (int k = 0; k < 2; ++k) {
                    const double ys = k ? stand_r : stand_l;
                    if (std::fabs(dty) < 1e-9) continue;
                    const double ts = ys / dty;
                    if (ts > 1e-6 && ts < t) {
                        const double z = h + ts * d[2];
                        if (z >= tb * ys - 0.5 && z <= tb * ys + 30.0) { t = ts; what = 4; }
                    }
                }
                for (const auto& b : boxes) {
                    const double tbx = ray_box(b, o, d);
                    if (tbx < t) { t = tbx; what = 3; }
                }
                if (what < 0 || t > max_range) continue;
                double x = t * d[0], y = t * d[1], z = h + t * d[2];
                const double dn = (what == 0 ? 0.0 : 0.02 * rng.normal());
                x += dn * d[0]; y += dn * d[1]; z += dn * d[2];
                if (what == 0) z += 0.015 * rng.normal();
                // ground-referenced: subtract nothing (ground passes through z=0 under the car)
                if (std::sqrt(x * x + y * y) <= 2.5) continue;  // ego removal
                out[4 * w + 0] = (float)x; out[4 * w + 1] = (float)y; out[4 * w + 2] = (float)z; out[4 * w + 3] = 0.f;
                ++w;
            }
        }
    }
    return w;
}

}  // extern "C"
