// TEST INFRASTRUCTURE ONLY -- runs the library's CUDA kernels (csrc/obs.cuh, csrc/obs_binned.cuh) on the CPU under the
// SIMT emulator of fake_cuda/cuda_runtime.h.  Built by tests/test_device_emul.py with g++; never part of libvggp.so.
// The host glue below repeats what csrc/vggp.cu does around these kernels (argument structs, launch geometry); the
// device radix sort is replaced by std::stable_sort.
#include "fake_cuda/cuda_runtime.h"      // must come first: the kernels include <cuda_runtime.h> (same include guard)
#include <numeric>

namespace vggp {
// dynamic shared memory of the kernels: `extern __shared__ ... smraw[]` inside a kernel of namespace vggp names this.
// Defined BEFORE the kernels are included, so that the block-scope extern declarations bind to a known definition.
thread_local __attribute__((aligned(128))) unsigned char smraw[228 * 1024];
}
#include "../../variational-gridded-gaussian-processes_b200/csrc/obs_binned.cuh"
#include "../../variational-gridded-gaussian-processes_b200/csrc/obs.cuh"

namespace vggp {
thread_local char g_err[512] = {0};
unsigned long long g_launches = 0;
}

using namespace vggp;

namespace {

struct Tables {
    std::vector<unsigned char> bytes;
    int table_bytes, knots_byte_off;
};

template <typename T>
Tables make_tables(int D, const int* K, const float* knots, const T* tab) {
    int ktot = 0;
    for (int d = 0; d < D; ++d) ktot += K[d];
    Tables t;
    const int band_bytes = (int)(((size_t)8 * ktot * sizeof(T) + 15) / 16 * 16);
    const int knot_bytes = (int)(((size_t)ktot * 4 + 15) / 16 * 16);
    t.knots_byte_off = band_bytes;
    t.table_bytes = band_bytes + knot_bytes;
    t.bytes.assign((size_t)t.table_bytes, 0);
    memcpy(t.bytes.data(), tab, (size_t)8 * ktot * sizeof(T));
    memcpy(t.bytes.data() + band_bytes, knots, (size_t)ktot * 4);
    return t;
}

MeshView make_mesh(const float* t, int K) {
    MeshView mv;
    mv.t = t; mv.K = K; mv.t0 = t[0];
    mv.inv_h = (float)((double)(K - 1) / ((double)t[K - 1] - (double)t[0]));
    mv.nearly_uniform = 1;
    mv.tfirst = t[0];
    mv.tlast = t[K - 1];
    for (int k = 0; k < K; ++k) {
        const float gf = (t[k] - mv.t0) * mv.inv_h;
        if (fabsf(gf - (float)k) > 1.25f) mv.nearly_uniform = 0;
    }
    return mv;
}

template <typename T, int D>
int run(int layout, const int* K, const float* knots, const void* const* xv, const void* yv, int64_t n, int run_len,
        int blocks_cap, const void* tabv, const void* alphav, void* gbufv, int64_t scalar_off, int64_t* stats) {
    const T* y = reinterpret_cast<const T*>(yv);
    int64_t M = 1, ncells = 1;
    int knot_off[D], tab_off[D], band_off[D], stride[D], ktot = 0;
    for (int d = 0; d < D; ++d) {
        knot_off[d] = ktot; tab_off[d] = 8 * ktot; band_off[d] = 4 * ktot;
        ktot += K[d];
        M *= K[d];
        ncells *= K[d] - 1;
    }
    for (int d = 0; d < D; ++d) {
        int s = 1;
        for (int f = d + 1; f < D; ++f) s *= K[f];
        stride[d] = s;
    }
    // per-dimension table offsets are in elements of T inside the table block: [8 K_0 | 8 K_1 | ...]
    Tables tb = make_tables<T>(D, K, knots, reinterpret_cast<const T*>(tabv));
    MeshView mesh[D];
    const T* x[D];
    for (int d = 0; d < D; ++d) {
        mesh[d] = make_mesh(knots + knot_off[d], K[d]);
        x[d] = reinterpret_cast<const T*>(xv[d]);
    }
    T* gb = reinterpret_cast<T*>(gbufv);
    double* gs = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(gbufv) + scalar_off);
    unsigned int counter = 0;

    // ---- cell keys (k_cell_keys) + stable sort (stands in for cub::DeviceRadixSort::SortPairs) ----
    std::vector<uint32_t> keys((size_t)std::max<int64_t>(n, 1)), idx((size_t)std::max<int64_t>(n, 1)), perm((size_t)std::max<int64_t>(n, 1));
    if (n > 0) {
        KeyArgs<T, D> ka;
        ka.n = n;
        for (int d = 0; d < D; ++d) { ka.x[d] = x[d]; ka.mesh[d] = mesh[d]; }
        const int blocks = (int)std::min<int64_t>((n + 255) / 256, 6);
        cuda_emul::launch(dim3(blocks), dim3(256), [&] { k_cell_keys<T, D>(ka, (uint32_t)ncells, keys.data(), idx.data()); });
        std::iota(perm.begin(), perm.end(), 0u);
        std::stable_sort(perm.begin(), perm.begin() + n, [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    }

    if (layout == 0 || layout == 1) {
        // ---- packed layout, k_obs_b1 (the kernel verified on B200): layout 0 = cell-sorted, 1 = input order ----
        PackGeom geo;
        geo.R = run_len;
        const int64_t lanes = (n + run_len - 1) / run_len;
        geo.nwarps = (lanes + 31) / 32;
        geo.n_packed = geo.nwarps * 32 * run_len;
        std::vector<T> xp[D], yp((size_t)std::max<int64_t>(geo.n_packed, 1));
        GatherArgs<T, D> ga;
        ga.geo = geo; ga.n = n; ga.perm = layout == 0 ? perm.data() : nullptr;
        for (int d = 0; d < D; ++d) {
            xp[d].resize((size_t)std::max<int64_t>(geo.n_packed, 1));
            ga.x[d] = x[d]; ga.xp[d] = xp[d].data();
        }
        ga.y = y; ga.yp = yp.data();
        if (geo.n_packed > 0)
            cuda_emul::launch(dim3((unsigned)std::min<int64_t>((geo.n_packed + 255) / 256, 6)), dim3(256), [&] { k_pack_gather<T, D>(ga); });
        PackedArgs<T, D> a;
        a.geo = geo;
        for (int d = 0; d < D; ++d) {
            a.xp[d] = xp[d].data(); a.mesh[d] = mesh[d]; a.stride[d] = stride[d]; a.band_off[d] = band_off[d];
            a.tab_off[d] = tab_off[d]; a.knot_off[d] = knot_off[d];
        }
        a.yp = yp.data();
        a.table_bytes = tb.table_bytes; a.knots_byte_off = tb.knots_byte_off; a.tables = tb.bytes.data();
        a.alpha = reinterpret_cast<const T*>(alphav);
        a.galpha = gb; a.gband = gb + M; a.gs = gs; a.n_real = (double)n; a.counter = &counter;
        int64_t blocks = (geo.nwarps + (OBS_THREADS / 32) - 1) / (OBS_THREADS / 32);
        blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, blocks_cap));
        if (n > 0) cuda_emul::launch(dim3((unsigned)blocks), dim3(OBS_THREADS), [&] { k_obs_b1<T, D>(a); });
        stats[0] = geo.nwarps; stats[1] = geo.n_packed;
        return 0;
    }

    // ---- binned layout: histogram, host plan, gather, sum y^2 outside, k_obs_b1_binned[_tma] ----
    std::vector<uint32_t> count((size_t)ncells + 1, 0u);
    if (n > 0)
        cuda_emul::launch(dim3((unsigned)std::min<int64_t>((n + 255) / 256, 6)), dim3(256), [&] { k_bin_histogram(keys.data(), n, count.data()); });
    BinLayout L;
    if (plan_bins(count.data(), ncells, run_len, D, L) != 0) return -1;
    if (L.n != n) return -2;
    const BinOffsets o = bin_offsets(L.n_tasks, L.data_elems, (int)sizeof(T));
    std::vector<unsigned char> bufv((size_t)o.bytes + 256, 0xAB);
    unsigned char* buf = bufv.data() + (256 - ((uintptr_t)bufv.data() & 255)) % 256;
    memset(buf, 0, BIN_HEADER_BYTES);
    if (L.n_tasks > 0) {
        memcpy(buf + o.task_off, L.task_off.data(), sizeof(int64_t) * L.n_tasks);
        memcpy(buf + o.task_R, L.task_R.data(), sizeof(int32_t) * L.n_tasks);
        memcpy(buf + o.run_cell, L.run_cell.data(), sizeof(uint32_t) * 32 * L.n_tasks);
        memcpy(buf + o.run_n, L.run_n.data(), sizeof(int32_t) * 32 * L.n_tasks);
        memcpy(buf + o.run_start, L.run_start.data(), sizeof(uint32_t) * 32 * L.n_tasks);
        BinGatherArgs<T, D> ga;
        for (int d = 0; d < D; ++d) { ga.x[d] = x[d]; ga.knots[d] = knots + knot_off[d]; ga.K[d] = K[d]; }
        ga.y = y; ga.perm = perm.data(); ga.buf = buf;
        ga.off_task_off = o.task_off; ga.off_task_R = o.task_R; ga.off_run_cell = o.run_cell; ga.off_run_n = o.run_n;
        ga.off_run_start = o.run_start; ga.off_data = o.data; ga.n_tasks = (int)L.n_tasks;
        cuda_emul::launch(dim3((unsigned)std::min<int64_t>(L.n_tasks, 5)), dim3(256), [&] { k_bin_gather<T, D>(ga); });
    }
    if (L.n > L.n_inside) {
        const int64_t nout = L.n - L.n_inside;
        cuda_emul::launch(dim3((unsigned)std::min<int64_t>((nout + 255) / 256, 3)), dim3(256),
                          [&] { k_bin_sum_y2<T>(y, perm.data(), L.n_inside, L.n, reinterpret_cast<double*>(buf)); });
    }
    BinnedArgs<T, D> a;
    for (int d = 0; d < D; ++d) {
        a.geo.K[d] = K[d]; a.geo.stride[d] = stride[d]; a.geo.band_off[d] = band_off[d]; a.geo.tab_off[d] = tab_off[d];
        a.geo.knot_off[d] = knot_off[d];
    }
    a.buf = buf;
    a.off_task_off = o.task_off; a.off_task_R = o.task_R; a.off_run_cell = o.run_cell; a.off_run_n = o.run_n; a.off_data = o.data;
    a.n_tasks = (int)L.n_tasks;
    a.knots_byte_off = tb.knots_byte_off; a.tables = tb.bytes.data();
    a.alpha = reinterpret_cast<const T*>(alphav);
    // band sums go through replicas (3 here, so that CTAs share and do not share one) and k_band_reduce, as in the library
    const int band_total = band_off[D - 1] + 4 * K[D - 1];
    std::vector<T> rep((size_t)3 * band_total, (T)0);
    a.galpha = gb; a.gband = rep.data(); a.n_rep = 3; a.band_rep_stride = band_total;
    a.gs = gs; a.n_real = (double)n; a.counter = &counter;
    int64_t blocks = (L.n_tasks + BIN_WARPS - 1) / BIN_WARPS;
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, blocks_cap));
    if (n > 0) {
        if (layout == 3) cuda_emul::launch(dim3((unsigned)blocks), dim3(BIN_THREADS), [&] { k_obs_b1_binned_tma<T, D>(a); });
        else cuda_emul::launch(dim3((unsigned)blocks), dim3(BIN_THREADS), [&] { k_obs_b1_binned<T, D>(a); });
        cuda_emul::launch(dim3((unsigned)((band_total + 15) / 16)), dim3(256),
                          [&] { k_band_reduce<T>(rep.data(), 3, (int64_t)band_total, band_total, gb + M, &counter); });
        for (T v : rep) if (v != (T)0) return -7;      // the reduce kernel leaves the replicas cleared
    }
    stats[0] = L.n_tasks; stats[1] = L.data_elems; stats[2] = L.n_inside; stats[3] = L.n_runs;
    return 0;
}

}  // namespace

extern "C" {

// layout: 0 packed cell-sorted (k_obs_b1), 1 packed input order (k_obs_b1), 2 binned + LDG stream, 3 binned + TMA ring.
// run_len: run length per lane (packed, multiple of 4) or run_cap (binned).  blocks_cap: resident CTAs to emulate.
// gbuf: zeroed by the caller, layout [d alpha (M) | bands (4 sum K_d) | pad8 | 8 float64 scalars at scalar_off].
int emul_device_run(int dtype, int D, int layout, const int* K, const float* knots, const void* const* x, const void* y,
                    int64_t n, int run_len, int blocks_cap, const void* tab, const void* alpha, void* gbuf,
                    int64_t scalar_off, int64_t* stats) {
#define EMUL_CASE(T, DD) return run<T, DD>(layout, K, knots, x, y, n, run_len, blocks_cap, tab, alpha, gbuf, scalar_off, stats)
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

unsigned long long emul_yields() { return cuda_emul::yields(); }

}
