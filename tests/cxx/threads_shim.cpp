// threads_shim.cpp -- the reference node enters libgpu.so from three threads (ElevationMapping.cpp): the point-cloud
// thread calls Process_points OUTSIDE MapMutex_ and Fuse inside (:271-282), the map-update thread calls Mapvar_update
// (:286-300), and the spinner calls Move / Map_feature / Raytracing (:388-421).  This program does exactly that through
// the source-level shim (compat/gpu_process_shim.cpp), WITHOUT any lock of its own around Process_points, and checks
// that nothing fails and that the map stays sane.  With the handle mutex of libgem_b200 every call is atomic; the
// order of calls from different threads is whatever it is, like in the node.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

#include <Eigen/Core>

void Init_GPU_elevationmap(int, float, float, float);
void Move(float *, float, int, float *, int *, float *);
int Process_points(int *, float *, float *, float *, float *, float *, float *, float *, Eigen::Matrix4f, int, double, double,
                   float, float, float, Eigen::RowVector3f, Eigen::Matrix3f, Eigen::Matrix3f, Eigen::RowVector3f, Eigen::Matrix3f);
void Fuse(int, int, int *, int *, int *, int *, float *, float *, float *);
void Mapvar_update(int, float);
void Map_feature(int, float *, float *, int *, int *, int *, float *, float *, float *, float *);
void Raytracing(int);

static uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main()
{
    const int L = 256, N = 40000, ITER = 60;
    const float res = 0.1f;
    Init_GPU_elevationmap(L, res, 2.5f, 0.7f);
    std::mutex MapMutex_; // the node's own lock: Fuse, Mapvar_update, and the spinner's calls are inside, Process_points is not
    std::atomic<int> failures{0};
    std::atomic<bool> stop{false};

    std::thread pointcloud([&] { // ElevationMapping::processpoints
        uint64_t s = 7;
        std::vector<float> x(N), y(N), z(N), var(N), xt(N), yt(N), zt(N), inten(N);
        std::vector<int> key(N), R(N), G(N), B(N);
        for (int it = 0; it < ITER; it++) {
            for (int i = 0; i < N; i++) {
                x[i] = (float)((splitmix(s) % 100000) / 100000.0 * 20.0 - 10.0);
                y[i] = (float)((splitmix(s) % 100000) / 100000.0 * 9.0 - 11.0); // behind the sensor: passes gpu_process.cu:393
                z[i] = (float)((splitmix(s) % 100000) / 100000.0 * 1.0 - 0.5);
                R[i] = 1 + (int)(splitmix(s) % 255); G[i] = 1 + (int)(splitmix(s) % 255); B[i] = 1 + (int)(splitmix(s) % 255);
                inten[i] = (float)(1 + splitmix(s) % 255);
            }
            Eigen::Matrix4f T = Eigen::Matrix4f::Zero();
            T(0, 0) = T(1, 1) = T(2, 2) = T(3, 3) = 1.0f;
            T(0, 3) = 0.05f * it; T(1, 3) = 3.0f; T(2, 3) = 0.2f;
            Eigen::RowVector3f sJ; sJ(0, 0) = 0; sJ(0, 1) = 0; sJ(0, 2) = 1;
            Eigen::Matrix3f Z = Eigen::Matrix3f::Zero(), I = Eigen::Matrix3f::Zero();
            I(0, 0) = I(1, 1) = I(2, 2) = 1.0f;
            Eigen::RowVector3f P; P(0, 0) = 0; P(0, 1) = 0; P(0, 2) = 1;
            Process_points(key.data(), x.data(), y.data(), z.data(), var.data(), xt.data(), yt.data(), zt.data(), T, N, -5.0, 0.8, 0.018f,
                           0.0006f, 0.0015f, sJ, Z, I, P, Z); // NOT under MapMutex_ (ElevationMapping.cpp:271-276)
            int in_grid = 0;
            for (int i = 0; i < N; i++) {
                if (key[i] < -1 || key[i] >= L * L) failures++;
                in_grid += key[i] >= 0;
            }
            if (in_grid == 0) failures++;
            std::lock_guard<std::mutex> lk(MapMutex_);
            Fuse(L, N, key.data(), R.data(), G.data(), B.data(), inten.data(), zt.data(), var.data());
        }
        stop = true;
    });
    std::thread mapupdate([&] { // ElevationMapping::processmapcells
        while (!stop) {
            { std::lock_guard<std::mutex> lk(MapMutex_); Mapvar_update(L, 0.0f); }
            std::this_thread::yield();
        }
    });
    std::thread spinner([&] { // ElevationMapping::Callback
        std::vector<float> elev(L * L), var(L * L), rough(L * L), slope(L * L), traver(L * L), inten(L * L);
        std::vector<int> R(L * L), G(L * L), B(L * L);
        int it = 0;
        while (!stop) {
            float pos[3] = {0.05f * it, 0.0f, 0.2f}, centre[2], shift[2];
            int start[2];
            Move(pos, res, L, centre, start, shift);
            Map_feature(L, elev.data(), var.data(), R.data(), G.data(), B.data(), rough.data(), slope.data(), traver.data(), inten.data());
            for (int c = 0; c < L * L; c++)
                if (elev[c] != -10.0f && !(var[c] >= 1e-4f * 0.999f && std::isfinite(elev[c]))) failures++;
            Raytracing(L);
            it++;
        }
    });
    pointcloud.join(); mapupdate.join(); spinner.join();
    std::vector<float> elev(L * L), var(L * L), rough(L * L), slope(L * L), traver(L * L), inten(L * L);
    std::vector<int> R(L * L), G(L * L), B(L * L);
    Map_feature(L, elev.data(), var.data(), R.data(), G.data(), B.data(), rough.data(), slope.data(), traver.data(), inten.data());
    int valid = 0;
    for (int c = 0; c < L * L; c++) valid += elev[c] != -10.0f;
    std::printf("threads: failures=%d valid_cells=%d\n", failures.load(), valid);
    return (failures == 0 && valid > 1000) ? 0 : 1;
}
