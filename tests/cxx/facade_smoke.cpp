// facade_smoke.cpp -- exercises the C++ facade (include/gem_b200/elevation_map.hpp) and the
// source-level shim (compat/gpu_process_shim.cpp, built against the stand-in Eigen) on a GPU and
// prints checksums that tests/test_cxx_facade.py compares with the Python path.
#include <cstdint>
#include <cstdio>
#include <vector>

#include <Eigen/Core>

#include "gem_b200/elevation_map.hpp"

// the nine reference entry points, re-exported by the shim
void Init_GPU_elevationmap(int, float, float, float);
void Move(float *, float, int, float *, int *, float *);
int Process_points(int *, float *, float *, float *, float *, float *, float *, float *, Eigen::Matrix4f, int, double, double,
                   float, float, float, Eigen::RowVector3f, Eigen::Matrix3f, Eigen::Matrix3f, Eigen::RowVector3f, Eigen::Matrix3f);
void Fuse(int, int, int *, int *, int *, int *, float *, float *, float *);
void Map_feature(int, float *, float *, int *, int *, int *, float *, float *, float *, float *);
void Raytracing(int);

static uint64_t fnv(const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main()
{
    const int L = 128, N = 50000;
    const float res = 0.1f;
    std::vector<gem_b200::PointXYZRGBICT> cloud(N);
    uint64_t s = 42;
    for (auto &p : cloud) {
        p.x = (float)((splitmix(s) % 100000) / 100000.0 * 12.0 - 6.0);
        p.y = (float)((splitmix(s) % 100000) / 100000.0 * 5.0 - 7.0); // behind the sensor: passes gpu_process.cu:393
        p.z = (float)((splitmix(s) % 100000) / 100000.0 * 1.0 - 0.5);
        p.pad = 1.0f;
        p.r = (unsigned char)(1 + splitmix(s) % 255); p.g = (unsigned char)(1 + splitmix(s) % 255); p.b = (unsigned char)(1 + splitmix(s) % 255); p.a = 255;
        p.covariance = 0; p.intensity = (float)(1 + splitmix(s) % 255); p.travers = 0;
    }
    double T[16] = {1, 0, 0, 0.3, 0, 1, 0, 3.0, 0, 0, 1, 0.2, 0, 0, 0, 1};
    gem_b200::LaserSensorProcessor laser;
    const gem_frame frame = gem_b200::makeFrame(T, laser, 0.0);
    const float pos[3] = {0.3f, 0.0f, 0.2f};

    // --- facade path -----------------------------------------------------------------------
    gem_b200::Layers a;
    std::vector<unsigned char> ortho, vis_rgb;
    std::vector<float> vis_xyz;
    int n_vis = 0;
    {
        gem_b200::ElevationMap map(L, res);
        map.move(pos);
        map.add(cloud.data(), cloud.size(), frame);
        map.update(0.0f);
        map.fuse(a);
        map.orthomosaic(ortho);
        n_vis = map.visualPoints(vis_xyz, vis_rgb);
        map.snapshot();
        map.clean();
        // drive 2 m along +x: the trailing rows of the snapshot leave the window and are harvested
        const float pos2[3] = {pos[0] + 2.0f, pos[1], pos[2]};
        float centre2[2], shift2[2];
        int start2[2];
        map.move(pos2, centre2, start2, shift2);
        std::vector<gem_b200::PointXYZRGBICT> left;
        const int n_left = map.harvest(centre2, shift2, left);
        int bad = 0;
        for (const auto &p : left) bad += !(p.x < centre2[0] - 0.5f * L * res) || !(p.travers >= 0.0f);
        std::printf("harvest points=%d bad=%d shift=(%g,%g)\n", n_left, bad, shift2[0], shift2[1]);
        if (bad || (shift2[0] > 0 && n_left == 0)) return 2;
        const gem_stats st = map.stats();
        std::printf("facade points_binned=%lld cells_touched=%lld\n", st.points_binned, st.cells_touched);
    }
    std::printf("facade elevation=%016llx variance=%016llx traver=%016llx color_r=%016llx\n",
                (unsigned long long)fnv(a.elevation.data(), a.elevation.size() * 4), (unsigned long long)fnv(a.variance.data(), a.variance.size() * 4),
                (unsigned long long)fnv(a.traver.data(), a.traver.size() * 4), (unsigned long long)fnv(a.color_r.data(), a.color_r.size() * 4));

    // --- shim path: the reference's nine functions, host arrays ---------------------------------
    Init_GPU_elevationmap(L, res, 2.5f, 0.7f);
    float p3[3] = {pos[0], pos[1], pos[2]}, centre[2], shift[2];
    int start[2];
    Move(p3, res, L, centre, start, shift);
    std::vector<float> x(N), y(N), z(N), var(N), xt(N), yt(N), zt(N), inten(N);
    std::vector<int> key(N), R(N), G(N), B(N);
    for (int i = 0; i < N; i++) { x[i] = cloud[i].x; y[i] = cloud[i].y; z[i] = cloud[i].z; R[i] = cloud[i].r; G[i] = cloud[i].g; B[i] = cloud[i].b; inten[i] = cloud[i].intensity; }
    Eigen::Matrix4f Tm;
    Eigen::Matrix3f Z3 = Eigen::Matrix3f::Zero(), I3 = Eigen::Matrix3f::Zero();
    Eigen::RowVector3f sj, pm;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Tm(i, j) = (float)T[4 * i + j];
    for (int j = 0; j < 3; j++) { sj(0, j) = (float)T[8 + j]; pm(0, j) = j == 2 ? 1.0f : 0.0f; I3(j, j) = 1.0f; }
    Process_points(key.data(), x.data(), y.data(), z.data(), var.data(), xt.data(), yt.data(), zt.data(), Tm, N, frame.rel_lower,
                   frame.rel_upper, laser.min_radius, laser.beam_angle, laser.beam_constant, sj, Z3, I3, pm, Z3);
    Fuse(L, N, key.data(), R.data(), G.data(), B.data(), inten.data(), zt.data(), var.data());
    const size_t C = (size_t)L * L;
    std::vector<float> e(C), v(C), ro(C), sl(C), tr(C), it(C);
    std::vector<int> cr(C), cg(C), cb(C);
    Map_feature(L, e.data(), v.data(), cr.data(), cg.data(), cb.data(), ro.data(), sl.data(), tr.data(), it.data());
    Raytracing(L);
    std::printf("shim key=%016llx elevation=%016llx variance=%016llx traver=%016llx\n", (unsigned long long)fnv(key.data(), N * 4),
                (unsigned long long)fnv(e.data(), C * 4), (unsigned long long)fnv(v.data(), C * 4), (unsigned long long)fnv(tr.data(), C * 4));
    // the shim's row-major elevation must equal the facade's column-major export where valid
    size_t mism = 0, valid = 0;
    for (int sx = 0; sx < L; sx++)
        for (int sy = 0; sy < L; sy++) {
            const float rm = e[(size_t)sx * L + sy], cm = a.elevation[(size_t)sy * L + sx];
            const bool shown = rm != -10.0f && tr[(size_t)sx * L + sy] != -10.0f && tr[(size_t)sx * L + sy] == tr[(size_t)sx * L + sy];
            if (shown) { valid++; if (rm != cm) mism++; } else if (cm == cm) mism++;
        }
    // show()'s cloud has one point per shown cell, and as many pixels of the orthomosaic can be non-black
    size_t lit = 0;
    for (size_t p = 0; p < ortho.size(); p += 3) lit += (ortho[p] | ortho[p + 1] | ortho[p + 2]) ? 1 : 0;
    if ((size_t)n_vis != valid || lit > valid || vis_xyz.size() != 3 * valid) mism++;
    std::printf("visual cloud points=%d orthomosaic lit pixels=%zu\n", n_vis, lit);
    std::printf("cross-check valid=%zu mismatches=%zu\n", valid, mism);
    return mism == 0 && valid > 100 ? 0 : 1;
}
