// TEST INFRASTRUCTURE ONLY -- host emulation of the binned K1 path (csrc/obs_binned.cuh, csrc/binplan.hpp).
// Built by tests/test_binned_host_emul.py with g++ and loaded through ctypes; nothing here is part of
// libvggp.so.  It runs, on the CPU and sequentially, exactly the __host__ __device__ lane functions the CUDA kernel
// k_obs_b1_binned runs (enter cell / per-observation update / flush), the host planner, and the index arithmetic of
// the gather kernel, so that the layout, the moment algebra and the padding rules are checked without a GPU.  What it
// cannot cover is the CUDA glue (TMA staging, work stealing, atomics, the device sort).
#define VGGP_HOST_EMUL
#include <vector>
#include <algorithm>
#include <numeric>
#include <string.h>
#include "../../variational-gridded-gaussian-processes_b200/csrc/obs_binned.cuh"

using namespace vggp;

namespace {

struct PlainAdder {
    template <typename T>
    void operator()(T* p, T v) const { *p += v; }
};

// c = clamp(searchsorted(mesh, x, right=False) - 1, 0, K-2); inside = mesh[0] <= x <= mesh[K-1]   (obs.cuh find_cell)
template <typename T>
int host_find_cell(const float* t, int K, T x, bool& inside) {
    inside = (x >= (T)t[0]) && (x <= (T)t[K - 1]);
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((T)t[mid] < x) lo = mid + 1; else hi = mid;
    }
    return std::min(std::max(lo - 1, 0), K - 2);
}

template <typename T, int D>
int run(const int* K, const float* knots, const void* const* xv, const void* yv, int64_t n, int run_cap,
        const int* stride, const int* band_off, const int* tab_off, const int* knot_off, const void* tabv,
        const void* alphav, void* galphav, void* gbandv, double* gs, int64_t* stats) {
    const T* y = reinterpret_cast<const T*>(yv);
    const T* tab = reinterpret_cast<const T*>(tabv);
    const T* alpha = reinterpret_cast<const T*>(alphav);
    T* galpha = reinterpret_cast<T*>(galphav);
    T* gband = reinterpret_cast<T*>(gbandv);
    const T* x[D];
    BinGeom<D> geo;
    int64_t ncells = 1;
    for (int d = 0; d < D; ++d) {
        x[d] = reinterpret_cast<const T*>(xv[d]);
        geo.K[d] = K[d]; geo.stride[d] = stride[d]; geo.band_off[d] = band_off[d]; geo.tab_off[d] = tab_off[d];
        geo.knot_off[d] = knot_off[d];
        ncells *= K[d] - 1;
    }
    // keys (k_cell_keys), stable sort (cub radix sort), histogram (k_bin_histogram)
    std::vector<uint32_t> key((size_t)n), perm((size_t)n), count((size_t)ncells + 1, 0u);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t k = 0;
        bool all_in = true;
        for (int d = 0; d < D; ++d) {
            bool inside;
            const int c = host_find_cell<T>(knots + knot_off[d], K[d], x[d][i], inside);
            k = k * (uint32_t)(K[d] - 1) + (uint32_t)c;
            all_in = all_in && inside;
        }
        key[(size_t)i] = all_in ? k : (uint32_t)ncells;
        ++count[key[(size_t)i]];
    }
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    BinLayout L;
    if (plan_bins(count.data(), ncells, run_cap, D, L) != 0) return -1;
    if (L.n != n) return -2;
    const BinOffsets o = bin_offsets(L.n_tasks, L.data_elems, (int)sizeof(T));
    std::vector<unsigned char> buf((size_t)o.bytes, 0xAB);      // poison: every byte the kernel reads must be written
    memset(buf.data(), 0, BIN_HEADER_BYTES);
    if (L.n_tasks > 0) {
        memcpy(buf.data() + o.task_off, L.task_off.data(), sizeof(int64_t) * L.n_tasks);
        memcpy(buf.data() + o.task_R, L.task_R.data(), sizeof(int32_t) * L.n_tasks);
        memcpy(buf.data() + o.run_cell, L.run_cell.data(), sizeof(uint32_t) * 32 * L.n_tasks);
        memcpy(buf.data() + o.run_n, L.run_n.data(), sizeof(int32_t) * 32 * L.n_tasks);
        memcpy(buf.data() + o.run_start, L.run_start.data(), sizeof(uint32_t) * 32 * L.n_tasks);
    }
    const int64_t* task_off = reinterpret_cast<const int64_t*>(buf.data() + o.task_off);
    const int* task_R = reinterpret_cast<const int*>(buf.data() + o.task_R);
    const uint32_t* run_cell = reinterpret_cast<const uint32_t*>(buf.data() + o.run_cell);
    const int* run_n = reinterpret_cast<const int*>(buf.data() + o.run_n);
    const uint32_t* run_start = reinterpret_cast<const uint32_t*>(buf.data() + o.run_start);
    T* data = reinterpret_cast<T*>(buf.data() + o.data);
    // gather: the body of k_bin_gather
    int64_t padded = 0;
    for (int64_t task = 0; task < L.n_tasks; ++task) {
        const int64_t elems = (int64_t)32 * task_R[task] * (D + 1);
        T* dst = data + task_off[task];
        for (int64_t e = 0; e < elems; ++e) {
            const BinSlot sl = bin_slot_of(e, D);
            const int64_t slot = task * 32 + sl.lane;
            const uint32_t cell = run_cell[slot];
            T v;
            if (sl.j < run_n[slot]) {
                const int64_t src = (int64_t)perm[(size_t)run_start[slot] + sl.j];
                v = (sl.arr < D) ? x[sl.arr < D ? sl.arr : 0][src] : y[src];
            } else if (sl.arr < D) {
                int c[D];
                bin_decode_cell<D>(cell != BIN_EMPTY ? cell : 0u, geo.K, c);
                v = (T)knots[knot_off[sl.arr] + c[sl.arr]];
                ++padded;
            } else {
                v = (T)0;
            }
            dst[e] = v;
        }
    }
    // sum y^2 outside (k_bin_sum_y2)
    double e_out = 0.0;
    for (int64_t i = L.n_inside; i < L.n; ++i) {
        const double v = (double)y[perm[(size_t)i]];
        e_out += v * v;
    }
    *reinterpret_cast<double*>(buf.data()) = e_out;
    // the kernel: tasks in order, lanes in order, groups of 4 observations
    const PlainAdder add;
    double etot = 0.0;
    for (int64_t task = 0; task < L.n_tasks; ++task) {
        for (int lane = 0; lane < 32; ++lane) {
            const int64_t slot = task * 32 + lane;
            const uint32_t cell = run_cell[slot];
            const int nrun = run_n[slot];
            const int groups = task_R[task] >> 2;
            const T* base = data + task_off[task] + lane * 4;
            const bool valid = cell != BIN_EMPTY;
            int c[D];
            bin_decode_cell<D>(valid ? cell : 0u, geo.K, c);
            BinLane<T, D> s;
            bin_lane_enter<T, D>(geo, s, c, tab, knots, alpha);
            for (int gi = 0; gi < groups; ++gi) {
                const T* g = base + (int64_t)gi * ((D + 1) * 128);
                const int left = nrun - 4 * gi;
                for (int j = 0; j < 4; ++j) {
                    T xx[D];
                    for (int d = 0; d < D; ++d) xx[d] = g[d * 128 + j];
                    bin_lane_obs<T, D>(s, xx, g[D * 128 + j], j < left);
                }
            }
            if (valid) etot += (double)bin_lane_flush<T, D>(geo, s, nrun, tab, galpha, gband, add);
        }
    }
    gs[0] += etot + *reinterpret_cast<const double*>(buf.data());
    gs[1] = (double)n;
    stats[0] = L.n_inside; stats[1] = L.n_runs; stats[2] = L.n_tasks; stats[3] = L.data_elems; stats[4] = o.bytes;
    stats[5] = padded;
    return 0;
}

}  // namespace

extern "C" {

// dtype: 0 float32, 1 float64.  `knots` is the concatenated float32 knot block addressed by knot_off; `tab` the
// per-cell tables [pe0 pe1 pe2 qe0 qe1 qe2 h rh] x K_d per dimension addressed by tab_off (obs dtype).
// galpha (M), gband (sum 4 K_d) and gs[2] must be zero on entry.  stats[6]: n_inside, n_runs, n_tasks, data_elems,
// bytes, padded x-slots.
int emul_binned_run(int dtype, int D, const int* K, const float* knots, const void* const* x, const void* y, int64_t n,
                    int run_cap, const int* stride, const int* band_off, const int* tab_off, const int* knot_off,
                    const void* tab, const void* alpha, void* galpha, void* gband, double* gs, int64_t* stats) {
#define EMUL_CASE(T, DD) return run<T, DD>(K, knots, x, y, n, run_cap, stride, band_off, tab_off, knot_off, tab, alpha, galpha, gband, gs, stats)
    if (dtype == 0) {
        if (D == 1) EMUL_CASE(float, 1);
        if (D == 2) EMUL_CASE(float, 2);
        if (D == 3) EMUL_CASE(float, 3);
    } else {
        if (D == 1) EMUL_CASE(double, 1);
        if (D == 2) EMUL_CASE(double, 2);
        if (D == 3) EMUL_CASE(double, 3);
    }
    return -3;
}

// planner only: layout summary for a given per-cell histogram.  out[8]: n, n_inside, n_runs, n_tasks, data_elems,
// max task R, min task R, sum of run lengths; run_n_out (32 * n_tasks, optional) receives the per-slot run lengths.
int emul_plan_bins(const uint32_t* count, int64_t ncells, int run_cap, int D, int64_t* out, int32_t* run_n_out,
                   int64_t run_n_cap) {
    BinLayout L;
    const int rc = plan_bins(count, ncells, run_cap, D, L);
    if (rc) return rc;
    out[0] = L.n; out[1] = L.n_inside; out[2] = L.n_runs; out[3] = L.n_tasks; out[4] = L.data_elems;
    int mx = 0, mn = 1 << 30;
    for (int r : L.task_R) { mx = std::max(mx, r); mn = std::min(mn, r); }
    out[5] = mx; out[6] = L.n_tasks ? mn : 0;
    int64_t tot = 0;
    for (int v : L.run_n) tot += v;
    out[7] = tot;
    if (run_n_out)
        for (int64_t i = 0; i < (int64_t)L.run_n.size() && i < run_n_cap; ++i) run_n_out[i] = L.run_n[(size_t)i];
    return 0;
}

}
