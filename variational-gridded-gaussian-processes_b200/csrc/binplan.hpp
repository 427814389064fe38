// Host-side planner of the "binned" observation layout consumed by k_obs_b1_binned (obs_binned.cuh).
// Pure C++ (no CUDA): it only looks at the per-cell observation counts.
//
// The cell-sorted observation stream is cut into RUNS: one run = all observations of one grid cell, or an equal
// share of them when the cell holds more than `run_cap`.  Runs are ordered by length (longest first) and grouped 32
// at a time into warp TASKS; lane l of the warp that takes task t walks run 32 t + l.  Because the 32 runs of a task
// have (almost) the same length, every lane enters and leaves its cell at the same instruction: the per-cell work of
// the fused kernel (load the cell's constants, flush its gradient sums) runs once per task with all lanes active,
// and the per-observation loop needs no cell-membership test at all.  Tasks are handed out longest-first (LPT) to
// persistent warps.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace vggp {

constexpr uint32_t BIN_EMPTY = 0xffffffffu;     // run_cell of a lane slot without a run (last task only)
constexpr int BIN_HEADER_BYTES = 256;           // device header: [0] double sum of y^2 over observations outside the mesh

struct BinLayout {
    int64_t n = 0, n_inside = 0, n_runs = 0, n_tasks = 0, data_elems = 0;
    std::vector<uint32_t> run_cell;    // [32 n_tasks] flat cell id (row-major over cells) or BIN_EMPTY
    std::vector<int32_t> run_n;        // [32 n_tasks] observations in the run (0 for an empty slot)
    std::vector<uint32_t> run_start;   // [32 n_tasks] position of the run's first observation in the cell-sorted stream
    std::vector<int64_t> task_off;     // [n_tasks] element offset of the task's data
    std::vector<int32_t> task_R;       // [n_tasks] padded run length of the task (multiple of 4)
};

// cell_count: [ncells + 1] observations per flat cell id; entry ncells counts the observations outside the mesh.
// D + 1 arrays (x_1..x_D, y) are streamed per observation.  Returns 0, or -1 for a bad argument.
inline int plan_bins(const uint32_t* cell_count, int64_t ncells, int run_cap, int D, BinLayout& out) {
    if (!cell_count || ncells < 0 || run_cap < 4 || D < 1) return -1;
    run_cap = run_cap / 4 * 4;
    out = BinLayout();
    struct Run { uint32_t cell, start; int32_t len; };
    std::vector<Run> runs;
    int64_t pos = 0;
    for (int64_t c = 0; c < ncells; ++c) {
        const int64_t cnt = cell_count[c];
        if (cnt > 0) {
            const int64_t k = (cnt + run_cap - 1) / run_cap;       // equal shares, sizes differ by at most 1
            const int64_t base = cnt / k, extra = cnt % k;
            int64_t s = pos;
            for (int64_t i = 0; i < k; ++i) {
                const int64_t len = base + (i < extra ? 1 : 0);
                runs.push_back(Run{(uint32_t)c, (uint32_t)s, (int32_t)len});
                s += len;
            }
        }
        pos += cnt;
    }
    out.n_inside = pos;
    out.n = pos + cell_count[ncells];
    out.n_runs = (int64_t)runs.size();
    // longest first; ties keep the cell order (stable), which keeps neighbouring cells in the same task
    std::stable_sort(runs.begin(), runs.end(), [](const Run& a, const Run& b) { return a.len > b.len; });
    out.n_tasks = (out.n_runs + 31) / 32;
    const int64_t slots = out.n_tasks * 32;
    out.run_cell.assign((size_t)slots, BIN_EMPTY);
    out.run_n.assign((size_t)slots, 0);
    out.run_start.assign((size_t)slots, 0u);
    out.task_off.resize((size_t)out.n_tasks);
    out.task_R.resize((size_t)out.n_tasks);
    int64_t off = 0;
    for (int64_t t = 0; t < out.n_tasks; ++t) {
        const int R = (runs[(size_t)(t * 32)].len + 3) / 4 * 4;
        out.task_off[(size_t)t] = off;
        out.task_R[(size_t)t] = R;
        off += (int64_t)32 * R * (D + 1);
        for (int l = 0; l < 32 && t * 32 + l < out.n_runs; ++l) {
            const Run& r = runs[(size_t)(t * 32 + l)];
            out.run_cell[(size_t)(t * 32 + l)] = r.cell;
            out.run_n[(size_t)(t * 32 + l)] = r.len;
            out.run_start[(size_t)(t * 32 + l)] = r.start;
        }
    }
    out.data_elems = off;
    return 0;
}

inline int64_t bin_align(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Byte offsets of the sections of the binned device buffer:
//   [header | task_off (i64) | task_R (i32) | run_cell (u32) | run_n (i32) | run_start (u32) | data (obs dtype)]
struct BinOffsets { int64_t task_off, task_R, run_cell, run_n, run_start, data, bytes; };

inline BinOffsets bin_offsets(int64_t n_tasks, int64_t data_elems, int elem_size) {
    BinOffsets o;
    int64_t b = BIN_HEADER_BYTES;
    o.task_off = b;  b = bin_align(b + 8 * n_tasks, 256);
    o.task_R = b;    b = bin_align(b + 4 * n_tasks, 256);
    o.run_cell = b;  b = bin_align(b + 4 * 32 * n_tasks, 256);
    o.run_n = b;     b = bin_align(b + 4 * 32 * n_tasks, 256);
    o.run_start = b; b = bin_align(b + 4 * 32 * n_tasks, 256);
    o.data = b;      b = bin_align(b + (int64_t)elem_size * data_elems, 256);
    o.bytes = b;
    return o;
}

}  // namespace vggp
