// CPU driver of the lane LP solver (polytope_b200/csrc/lp_lane.cuh compiled as plain C++):
// test infrastructure for tests/test_cpu_lane_solver.py, which checks the solver's
// arithmetic against HiGHS without a GPU.  One LP per call, SingleLane policy.
//   lane_host <in.bin> <out.bin> [w]      (w: WaitingLane policy, every polish preceded by the full wait)
// in:  int32 B, m, n, then per LP: int32 rows, G[m][n], h[m], c[n] (doubles, row-major)
// out: per LP: int32 status, iters, polishes, pad; double fun; double x[NS]
#include <cstdio>
#include <cstdlib>
#include <vector>
// -DLANE_HOST_NS=12 / 16 builds the wide solver (lp_lane_wide.cuh, 9 <= n <= 16) instead.
#include "../polytope_b200/csrc/lp_lane_wide.cuh"

using namespace pb200::lane;
#ifndef LANE_HOST_NS
#define LANE_HOST_NS 8
#endif
constexpr int NS = LANE_HOST_NS;

struct HostData {
    const double* G; const double* hv; const double* cv;
    int m, n;
    std::vector<double> sv, zv, Lv;
    double& L(int e) { return Lv[e]; }
    int rows() const { return m; }
    void row(int i, double (&g)[NS]) const { for (int j = 0; j < NS; ++j) g[j] = j < n ? G[i * n + j] : 0.0; }
    double h(int i) const { return hv[i]; }
    double c(int j) const { return cv[j]; }
    double& s(int i) { return sv[i]; }
    double& z(int i) { return zv[i]; }
};

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[3];
    if (fread(hdr, sizeof(int), 3, f) != 3) return 4;
    const int B = hdr[0], m = hdr[1], n = hdr[2];
    std::vector<double> G((size_t)m * n), h(m), c(n);
    FILE* o = fopen(argv[2], "wb");
    for (int b = 0; b < B; ++b) {
        int rows;
        if (fread(&rows, sizeof(int), 1, f) != 1) return 5;
        if (fread(G.data(), sizeof(double), (size_t)m * n, f) != (size_t)m * n) return 5;
        if (fread(h.data(), sizeof(double), m, f) != (size_t)m) return 5;
        if (fread(c.data(), sizeof(double), n, f) != (size_t)n) return 5;
        HostData d{G.data(), h.data(), c.data(), rows, n, std::vector<double>(m), std::vector<double>(m),
                   std::vector<double>(NS * (NS + 1) / 2)};
        Result<NS> res;
        const bool waiting = argc > 3 && argv[3][0] == 'w';
        if (NS > 8) {
            if (waiting) lane_solve_wide<NS, HostData, WaitingLane>(d, true, n, res);
            else lane_solve_wide<NS, HostData, SingleLane>(d, true, n, res);
        } else {
            if (waiting) lane_solve<NS, HostData, WaitingLane>(d, true, n, res);
            else lane_solve<NS, HostData, SingleLane>(d, true, n, res);
        }
        int meta[4] = {res.status, res.iters, res.polishes, 0};
        fwrite(meta, sizeof(int), 4, o);
        fwrite(&res.fun, sizeof(double), 1, o);
        fwrite(res.x, sizeof(double), NS, o);
    }
    fclose(o);
    fclose(f);
    return 0;
}
