// libpnb200: B200-native nonlocal operator assembly (C ABI in include/pnb200.h).
//
// Dense assembly design:
//   * the N x N output is cut into PNB_TD x PNB_TD DoF tiles; a CTA owns the
//     tiles of one (row group, column group) pair and is the ONLY writer of
//     those entries (row-tile ownership, no floating point atomics);
//   * per tile the CTA walks over the cell pairs (cells touching the row DoFs)
//     x (cells touching the column DoFs) in 16x16 sub-batches: classify the
//     pair (shared vertices / quadrature order), evaluate low-order regular
//     pairs one per thread, queue singular and high-order pairs in a
//     deterministic list and evaluate them one per warp, stage the 3x3 cross
//     blocks in shared memory and fold them into the tile accumulator with an
//     entry-centric gather in fixed order;
//   * the cell-diagonal blocks (both dofs on the same cell; they receive a
//     contribution from EVERY other cell) are reduced per CTA, staged per
//     (group, cell) and reduced/scattered by row-owning threads afterwards;
//   * only tiles with row tile <= column tile are computed; the mirror image
//     is written by the same CTA.
#include "../../include/pnb200.h"
#include "pnb_pair.cuh"

#include <algorithm>
#include <time.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <map>

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static double wall_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static thread_local std::string g_err;
struct pnb_problem;
extern "C" const char *pnb_last_error(void) { return g_err.c_str(); }
extern "C" int pnb_version(void) { return 100; }
extern "C" int pnb_far_max_order(void) { return PNB_FAR_MAX_ORDER; }

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PNB_ERR_NO_DEVICE : PNB_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    } while (0)

// device attributes are queried once per device and attribute (some of the queries take a millisecond)
static int device_attr(cudaDeviceAttr attr, int device)
{
    static std::mutex mu;
    static std::map<std::pair<int, int>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({(int)attr, device});
    if (it != cache.end()) return it->second;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, attr, device) != cudaSuccess) { cudaGetLastError(); v = 0; }
    cache[{(int)attr, device}] = v;
    return v;
}

// Entry points run on the problem's device and hand the calling thread its previous current device back on every
// exit path (the caller -- torch -- keeps allocating on whatever device is current).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        err = cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(dev)          \
    DeviceGuard device_guard_(dev); \
    CK(device_guard_.err)

extern "C" int pnb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}


// ---------------------------------------------------------------------------
// caching device allocator: problems are created and destroyed per assembly in the end-to-end path, and
// cudaMalloc / cudaFree of the multi-GB staging buffers cost ~100 ms each.  Freed blocks are kept (per
// device) and handed out again; pnb_release_cached_memory() returns them to the driver.
// ---------------------------------------------------------------------------
#include <mutex>
#include <map>
#include <thread>
struct PoolBlock { void *p; size_t bytes; int dev; };
static std::mutex g_pool_mu;
static std::vector<PoolBlock> g_pool_free;
static std::map<void *, PoolBlock> g_pool_used;
static size_t g_pool_cached = 0;

static void pool_trim_locked(size_t limit, int dev_only)
{
    for (size_t i = 0; i < g_pool_free.size() && g_pool_cached > limit;) {
        if (dev_only >= 0 && g_pool_free[i].dev != dev_only) { i++; continue; }
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(g_pool_free[i].dev);
        cudaFree(g_pool_free[i].p);
        cudaSetDevice(cur);
        g_pool_cached -= g_pool_free[i].bytes;
        g_pool_free.erase(g_pool_free.begin() + i);
    }
}

static cudaError_t pool_malloc(void **out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = bytes <= (1u << 20) ? ((bytes + 255) & ~(size_t)255) : ((bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1));
    std::lock_guard<std::mutex> lock(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool_free.size(); i++) {
        const PoolBlock &b = g_pool_free[i];
        if (b.dev != dev || b.bytes < want || b.bytes > want + want / 4 + (1u << 20)) continue;
        if (best < 0 || b.bytes < g_pool_free[best].bytes) best = i;
    }
    if (best >= 0) {
        PoolBlock b = g_pool_free[best];
        g_pool_free.erase(g_pool_free.begin() + best);
        g_pool_cached -= b.bytes;
        g_pool_used[b.p] = b;
        *out = b.p;
        return cudaSuccess;
    }
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim_locked(0, dev);      // give the cached blocks of this device back and retry
        e = cudaMalloc(&d, want);
        if (e != cudaSuccess) return e;
    }
    g_pool_used[d] = PoolBlock{d, want, dev};
    *out = d;
    return cudaSuccess;
}

static void pool_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = g_pool_used.find(p);
    if (it == g_pool_used.end()) { cudaFree(p); return; }
    g_pool_free.push_back(it->second);
    g_pool_cached += it->second.bytes;
    g_pool_used.erase(it);
    static const size_t limit = (size_t)(getenv("PNB_POOL_LIMIT_GB") ? atof(getenv("PNB_POOL_LIMIT_GB")) : 48.) << 30;
    pool_trim_locked(limit, -1);
}

extern "C" int pnb_release_cached_memory(void)
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    pool_trim_locked(0, -1);
    return 0;
}


// Dynamic shared memory opt-in: function attributes are process-wide state, so concurrent host threads must not
// set problem-dependent values; every kernel is opted in once per device to the device maximum.
template <class K> static void smem_optin(K kernel, int device)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, bool> done;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair((const void *)kernel, device);
    if (done.count(key)) return;
    int smem_blk = 0;
    smem_blk = device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaFuncAttributes attr;
    if (smem_blk > 0 && cudaFuncGetAttributes(&attr, kernel) == cudaSuccess)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_blk - (int)attr.sharedSizeBytes);
    // all of the unified L1 as shared memory: two CTAs of ~100 KB must be resident per SM (with the default carve-out
    // the driver kept a single CTA of the near evaluator and of the order-2 kernel on an SM: 8 warps instead of 16)
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    done[key] = true;
}

// ---------------------------------------------------------------------------
// mesh.hVector / mesh.h / mesh.hmin as the reference computes them (hdeltaCy, fem/PyNucleus_fem/meshCy.pyx:1654-1732):
// per cell the longest edge, over the mesh the longest and the SHORTEST edge, every edge length being
// sqrt(mydot(e, e)) with mydot = BLAS ddot (base/PyNucleus_base/opt_true_blas.pxi:125-141).  The ddot of the BLAS the
// reference links (OpenBLAS through scipy) accumulates with fused multiply-adds, x1*x1 + fl(x0*x0) in ONE rounding;
// getQuadOrder takes logarithms of these numbers and rounds up, so the last bit decides the quadrature order of
// some pairs on large meshes.  Host arithmetic only (no device needed).
// ---------------------------------------------------------------------------
extern "C" int pnb_mesh_edge_lengths(int32_t dim, int32_t num_cells, const double *vertices, const int32_t *cells,
                                     double *h, double *h_max, double *h_min)
{
    if (!vertices || !cells || !h || !h_max || !h_min) return fail(PNB_ERR_ARG, "null argument");
    if (dim != 1 && dim != 2) return fail(PNB_ERR_UNSUPPORTED, "pnb_mesh_edge_lengths: dim must be 1 or 2");
    double hmax = 0., hmin = 100.;       // the reference starts from 100 (meshCy.pyx:1661)
    for (int32_t c = 0; c < num_cells; c++) {
        double hl = 0.;
        if (dim == 1) {
            hl = fabs(vertices[cells[2 * (size_t)c + 1]] - vertices[cells[2 * (size_t)c]]);
            hmin = std::min(hmin, hl);
        } else {
            const double *v0 = vertices + 2 * (size_t)cells[3 * (size_t)c], *v1 = vertices + 2 * (size_t)cells[3 * (size_t)c + 1],
                         *v2 = vertices + 2 * (size_t)cells[3 * (size_t)c + 2];
            const double e[3][2] = {{v2[0] - v1[0], v2[1] - v1[1]}, {v2[0] - v0[0], v2[1] - v0[1]}, {v1[0] - v0[0], v1[1] - v0[1]}};
            for (int j = 0; j < 3; j++) {
                const double hS = sqrt(fma(e[j][1], e[j][1], e[j][0] * e[j][0]));
                hmin = std::min(hmin, hS);
                hl = std::max(hl, hS);
            }
        }
        hmax = std::max(hmax, hl);
        h[c] = hl;
    }
    *h_max = hmax;
    *h_min = hmin;
    return 0;
}

// ---------------------------------------------------------------------------
// problem object
// ---------------------------------------------------------------------------
struct TileSched {
    int ntiles, G, ngroups;
    const int *tile_ptr;    // ntiles+1
    const int *tile_cells;  // cell ids per tile, ascending
    const int *tile_loc;    // packed tile-local dof index of each vertex (8 bits each, 0xFF = not in tile)
    const double *tile_box; // ntiles x 2 x dim: bounding box (lo, hi per coordinate) of the vertices of the tile's cells
    // batched independent blocks (pnb_mesh_t.num_blocks > 0; nullptr otherwise): cells of different blocks never interact,
    // the output holds one dense operator per block
    const int *cell_block;  // nc: block of a cell
    const int *tile_block;  // ntiles: block of a tile
    const int *blk_tile0;   // nblocks: first tile
    const int *blk_group0;  // nblocks: first tile group
    const int *blk_n;       // nblocks: number of dofs (leading dimension of the block's operator)
    const long long *blk_out;   // nblocks: offset of the block's operator in the output
    const int *blk_fptr;    // nblocks+1: boundary facets of the block
    int dgroups;            // first dimension of DXp / DYp (groups per block; ngroups without blocks)
    const int *home;        // nc: home tile of a cell
    const int *units;       // 2 x nunits (row group, col group)
    int nunits;
    int nunits_all;         // capacity of the unit buffers
    double *DXp, *DYp;      // ngroups x nc x ND staging of cell-diagonal blocks
    double *Dbnd;           // nc x ND boundary contributions
    double *D;              // nc x ND reduced
    const int *dof_ptr;     // N+1: dof -> incident cells
    const int *dof_cells;   // packed (cell*4 + local index)
    int *err;               // [0]: max order requested beyond tables
    unsigned long long *counters;  // [0] evaluated pairs (far pass), [1] evaluated pairs (near pass)
    int own_t0, own_t1;     // row tiles [own_t0, own_t1) are owned (written) by this problem instance
    const unsigned char *cell_mask;  // several GPUs (pnb_dist_plan): cells whose surface terms this instance computes; nullptr: by home tile
    int maxcells;           // largest cell list of a tile
    int *tileflag;          // ntiles x tf_width: tile pair holds pairs for the near pass (batched blocks: column tile counted
                            // from the first tile of the block, tf_width = most tiles of a block)
    int tf_width;
    int *unitflag;          // nunits: unit holds flagged tiles
    int *nearunits;         // compacted list of flagged units, [nunits] = count
};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct GroupSched;
struct GroupHost;
struct pnb_problem {
    int device = 0;
    GroupSched *G = nullptr;          // 2D cell-group path (pnb_group.cuh)
    GroupHost *gh = nullptr;
    std::vector<int> h_cells, h_dofs, h_home; // host copies for the lazily built schedules
    std::vector<double> h_vertices;           // mesh vertices (tile bounding boxes of the finite-horizon path)
    std::vector<unsigned char> h_labels;      // cell labels of piecewise variable kernels (empty: constant kernel)
    bool tiles_ready = false;
    bool finite = false;        // finite horizon: DoF-tile path only
    int nblocks = 0;            // batched independent blocks (pnb_mesh_t.num_blocks)
    std::vector<int> h_blk_cell, h_blk_dof, h_blk_facet;    // block ranges (host)
    std::vector<int> h_block_units;                          // unit list of the batched problem
    long long blocks_out_doubles = 0;                        // size of the batched output
    cudaStream_t copy_stream = nullptr;             // device -> host copies of finished row panels
    cudaEvent_t pev[PNB_ROW_PANELS] = {};
    cudaEvent_t bev[2] = {};                        // surface-term kernel launched ahead of the schedule
    bool early_boundary = false;
    bool host_panels = false;                       // the last assembly copied its rows panel by panel
    bool ordered_classes = false;   // piecewise kernels with an unsymmetric class table or reversed singular pairs: DoF-tile path only
    int path = 0;               // 0: default, 1: DoF-tile path for whole 2D operators too (pnb_problem_set_path)
    int pow_eoff = 240;      // PowTab::eoff of this problem
    int row_part = 0, row_nparts = 1;   // row-owner kernels: this instance assembles the rows of one part (pnb_problem_set_row_part)
    std::vector<double> h_centers, h_h;
    std::vector<int4> h_grid;         // lane grids of the near evaluator per order
    std::vector<int> h_reg_n;         // node count of the regular cell rule per order
    int part = 0, nparts = 1;         // share of the units this problem instance evaluates (multi-GPU)
    bool dist = false;                // several GPUs: parts own groups (pnb_dist_plan)
    DProblem P{};
    TileSched S{};
    std::vector<void *> allocs;       // everything to free
    std::vector<void *> rule_allocs;  // regular tables (replaced by set_rules)
    FarRule far_rules[PNB_FAR_MAX_ORDER + 1];
    int far_mask = 0;   // bit o set: order o is handled by the thread-per-pair evaluator
    int64_t stats[8] = {0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *stage = nullptr;      // device staging of the host-output entry point
    size_t stage_bytes = 0;
    double timings[4] = {0};
    cudaEvent_t kev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // group path: f2 | near list + evaluator | mix | symmetrize
    double ktimings[4] = {0};
    int64_t distinct_pairs = 0;
    // host copies needed later
    int dim = 0, nc = 0, N = 0, nb = 0;
    bool has_singular = false;
};

template <class T> static int upload(pnb_problem *p, const T *host, size_t count, const T **dev, bool rule = false)
{
    void *d = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(pool_malloc(&d, bytes));
    (rule ? p->rule_allocs : p->allocs).push_back(d);
    if (count) CK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)d;
    return 0;
}

template <class T> static int dalloc(pnb_problem *p, size_t count, T **dev, bool zero = true)
{
    void *d = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(pool_malloc(&d, bytes));
    p->allocs.push_back(d);
    if (zero) CK(cudaMemset(d, 0, bytes));
    *dev = (T *)d;
    return 0;
}

static int upload_rule(pnb_problem *p, const pnb_rule_t &r, DRule *out, bool rule_alloc)
{
    out->n = r.n;
    out->rows = r.rows;
    out->bary = nullptr;
    out->w = nullptr;
    if (r.n <= 0) return 0;
    if (upload(p, r.bary, (size_t)r.rows * r.n, &out->bary, rule_alloc)) return PNB_ERR_CUDA;
    if (upload(p, r.w, (size_t)r.n, &out->w, rule_alloc)) return PNB_ERR_CUDA;
    return 0;
}


// host side of PowTab (see pnb_device.cuh); long double keeps the table entries correctly rounded
static void build_powtab(PowTab *t, double scal, double expo, int eoff)
{
    t->scal = scal;
    t->expo = expo;
    t->eoff = eoff;
    t->pad = 0;
    t->coef[0] = 1.;
    for (int k = 1; k < 8; k++) t->coef[k] = t->coef[k - 1] * (expo - k + 1) / k;
    for (int i = 0; i < 128; i++) {
        t->IT[i].x = 1.0 / (1.0 + (i + 0.5) / 128.0);
        const long double m0 = 1.0L / (long double)t->IT[i].x;
        t->IT[i].y = (double)powl(m0, (long double)expo);
    }
    for (int k = 0; k < 256; k++) t->T1[k] = (double)((long double)scal * exp2l((long double)expo * (long double)(k - eoff)));
}

extern "C" int pnb_problem_set_rules(pnb_problem *p, const pnb_rules_t *rules)
{
    if (!p || !rules) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    for (void *d : p->rule_allocs) pool_free(d);
    p->rule_allocs.clear();
    const int nvc = p->dim + 1;
    int rc = 0;
    p->has_singular = rules->vertex.n > 0;
    rc |= upload_rule(p, rules->identical, &p->P.q_id, true);
    rc |= upload_rule(p, rules->edge, &p->P.q_edge, true);
    rc |= upload_rule(p, rules->vertex, &p->P.q_vertex, true);
    rc |= upload_rule(p, rules->bedge, &p->P.bq_edge, true);
    rc |= upload_rule(p, rules->bvertex, &p->P.bq_vertex, true);
    if (rc) return PNB_ERR_CUDA;
    const int mo = rules->max_order;
    std::vector<DRule> cell(mo + 1), facet(mo + 1);
    memset(cell.data(), 0, sizeof(DRule) * (mo + 1));
    memset(facet.data(), 0, sizeof(DRule) * (mo + 1));
    for (int o = 1; o <= mo; o++) {
        if (rules->cell[o].rows != nvc) return fail(PNB_ERR_ARG, "cell rule must have dim+1 barycentric rows");
        if (upload_rule(p, rules->cell[o], &cell[o], true)) return PNB_ERR_CUDA;
        if (rules->facet && rules->facet[o].n > 0)
            if (upload_rule(p, rules->facet[o], &facet[o], true)) return PNB_ERR_CUDA;
    }
    if (upload(p, cell.data(), (size_t)mo + 1, &p->P.reg_cell, true)) return PNB_ERR_CUDA;
    if (upload(p, facet.data(), (size_t)mo + 1, &p->P.reg_facet, true)) return PNB_ERR_CUDA;
    p->P.max_order = mo;
    p->P.reg_derived = nullptr; p->P.reg_doff = nullptr; p->P.reg_nmax = 0; p->P.reg_grid = nullptr;
    if (p->dim == 2) {
        // per node PNB_DER2 double2: (w, w b0) (w b1, w b2) (q0,q1) (q2,q3) (q4,q5) with q = w b_a b_b (a<=b), (b0,b1) (b2,0)
        std::vector<int> doff(mo + 2, 0);
        std::vector<double> der;
        for (int o = 1; o <= mo; o++) {
            const pnb_rule_t &r = rules->cell[o];
            doff[o] = (int)(der.size() / (2 * PNB_DER2));
            p->P.reg_nmax = std::max(p->P.reg_nmax, r.n);
            for (int i = 0; i < r.n; i++) {
                der.push_back(r.w[i]);
                for (int k = 0; k < 3; k++) der.push_back(r.w[i] * r.bary[k * r.n + i]);
                for (int a = 0; a < 3; a++)
                    for (int b = a; b < 3; b++) der.push_back(r.w[i] * r.bary[a * r.n + i] * r.bary[b * r.n + i]);
                for (int k = 0; k < 3; k++) der.push_back(r.bary[k * r.n + i]);
                der.push_back(0.);
            }
        }
        doff[mo + 1] = (int)(der.size() / (2 * PNB_DER2));
        // lane groups of the near evaluator per order: W lanes per item (they split the row tiles of PNB_NEAR_R rows),
        // K = 32 / W items per warp; the pair (W, K) with the best lane utilisation whose column nodes fit the
        // per-warp buffer of PNB_NEAR_WARP_POINTS points; ties: fewer lanes per item (longer loops, shorter reduction)
        std::vector<int4> grid(mo + 1, make_int4(32, 1, 1, 0));
        for (int o = 1; o <= mo; o++) {
            const int n = rules->cell[o].n;
            if (n <= 0) continue;
            const int nsl = std::max(1, (n * n + 2047) / 2048), ncols = (n + nsl - 1) / nsl;
            const int ntiles = (n + PNB_NEAR_R - 1) / PNB_NEAR_R;
            double best = -1.;
            for (int W = 1; W <= 32; W++) {
                const int K = 32 / W;
                if (K * (3 + ncols) > PNB_NEAR_WARP_POINTS) continue;
                const int tp = (ntiles + W - 1) / W;
                if ((int64_t)tp * PNB_NEAR_R * ncols > 4096) continue;        // bound on the work of one lane
                const double u = (double)n / ((double)tp * W * PNB_NEAR_R) * ((double)K * W / 32.);
                if (u > best + 1e-9) { best = u; grid[o] = make_int4(W, K, tp, 0); }
            }
        }
        if (upload(p, der.data(), der.size(), &p->P.reg_derived, true)) return PNB_ERR_CUDA;
        if (upload(p, doff.data(), doff.size(), &p->P.reg_doff, true)) return PNB_ERR_CUDA;
        if (upload(p, grid.data(), grid.size(), &p->P.reg_grid, true)) return PNB_ERR_CUDA;
        p->h_grid = grid;
        p->h_reg_n.assign(mo + 1, 0);
        for (int o = 1; o <= mo; o++) p->h_reg_n[o] = rules->cell[o].n;
    }
    // low-order 2D rules for the thread-per-pair evaluator (orders 2..5, node counts fixed at compile time)
    memset(p->far_rules, 0, sizeof(p->far_rules));
    p->far_mask = 0;
    if (p->dim == 2)
        for (int o = 2; o <= std::min(mo, PNB_FAR_MAX_ORDER); o++) {
            const pnb_rule_t &r = rules->cell[o];
            if (r.n != far_expected_nodes(o)) continue;
            FarRule &F = p->far_rules[o];
            F.n = r.n;
            for (int i = 0; i < r.n; i++) {
                F.w[i] = r.w[i];
                for (int k = 0; k < 3; k++) {
                    F.bary[k][i] = r.bary[k * r.n + i];
                    F.wb[k][i] = r.w[i] * r.bary[k * r.n + i];
                }
                int e = 0;
                for (int a = 0; a < 3; a++)
                    for (int b = a; b < 3; b++) F.qq[e++][i] = r.w[i] * r.bary[a * r.n + i] * r.bary[b * r.n + i];
            }
            p->far_mask |= 1 << o;
        }
    if (upload(p, p->far_rules, (size_t)PNB_FAR_MAX_ORDER + 1, &p->P.far_rules, true)) return PNB_ERR_CUDA;
    return 0;
}

static void destroy_group_host(pnb_problem *p);
extern "C" void pnb_problem_destroy(pnb_problem *p)
{
    if (p && p->copy_stream) {
        cudaStreamSynchronize(p->copy_stream);
        cudaStreamDestroy(p->copy_stream);
        p->copy_stream = nullptr;
    }
    if (p)
        for (auto &e : p->pev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (p)
        for (auto &e : p->bev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (!p) return;
    DeviceGuard guard(p->device);
    for (void *d : p->allocs) pool_free(d);
    for (void *d : p->rule_allocs) pool_free(d);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    for (auto &e : p->kev) if (e) cudaEventDestroy(e);
    if (p->stage) pool_free(p->stage);
    destroy_group_host(p);
    delete p;
}

// Cell lists of the DoF tiles, cut into batches of PNB_SB cells that share NO vertex.  The cross blocks of the pairs of
// (row batch) x (column batch) then hit pairwise distinct tile entries, so that they can be added to the
// shared-memory tile straight from registers without conflicts (and in an order that is fixed by the schedule).
static int build_tile_schedule(pnb_problem *p)
{
    if (p->tiles_ready) return 0;
    TileSched &S = p->S;
    const int nc = p->nc, N = p->N, nvc = p->dim + 1, TD = PNB_TD;
    const int *cells = p->h_cells.data(), *dofs = p->h_dofs.data();
    std::vector<std::vector<int>> tcells(S.ntiles);
    for (int c = 0; c < nc; c++) {
        int tl[3], nt = 0;
        for (int m = 0; m < nvc; m++) {
            const int d = dofs[(size_t)c * nvc + m];
            if (d >= 0) {
                const int t = d / TD;
                bool seen = false;
                for (int k = 0; k < nt; k++) seen |= tl[k] == t;
                if (!seen) tl[nt++] = t;
            }
        }
        if (nt == 0) tl[nt++] = p->h_home[c];
        for (int k = 0; k < nt; k++) tcells[tl[k]].push_back(c);
    }
    std::vector<int> tptr(S.ntiles + 1, 0), tlist, tloc;
    {
        std::vector<std::vector<int>> bcells;      // batches of the current tile
        std::vector<std::vector<int>> bverts;
        for (int t = 0; t < S.ntiles; t++) {
            bcells.clear();
            bverts.clear();
            size_t first_open = 0;
            for (int c : tcells[t]) {
                const int *v = cells + (size_t)c * nvc;
                size_t b = first_open;
                for (; b < bcells.size(); b++) {
                    if ((int)bcells[b].size() >= PNB_SB) continue;
                    bool clash = false;
                    for (int x : bverts[b])
                        for (int m = 0; m < nvc; m++) clash |= x == v[m];
                    if (!clash) break;
                }
                if (b == bcells.size()) { bcells.emplace_back(); bverts.emplace_back(); }
                bcells[b].push_back(c);
                for (int m = 0; m < nvc; m++) bverts[b].push_back(v[m]);
                while (first_open < bcells.size() && (int)bcells[first_open].size() >= PNB_SB) first_open++;
            }
            for (auto &bc : bcells) {
                for (int k = 0; k < PNB_SB; k++) {
                    const int c = k < (int)bc.size() ? bc[k] : -1;
                    tlist.push_back(c);
                    int packed = 0x00FFFFFF;
                    if (c >= 0) {
                        packed = 0;
                        for (int m = 0; m < 3; m++) {
                            int l = 0xFF;
                            if (m < nvc) {
                                const int d = dofs[(size_t)c * nvc + m];
                                if (d >= 0 && d / TD == t) l = d - t * TD;
                            }
                            packed |= l << (8 * m);
                        }
                    }
                    tloc.push_back(packed);
                }
            }
            tptr[t + 1] = (int)tlist.size();
        }
    }
    {
        const int dim = p->dim;
        std::vector<double> box((size_t)S.ntiles * 2 * dim);
        for (int t = 0; t < S.ntiles; t++) {
            for (int l = 0; l < dim; l++) { box[((size_t)t * 2) * dim + l] = 1e300; box[((size_t)t * 2 + 1) * dim + l] = -1e300; }
            for (int c : tcells[t])
                for (int m = 0; m < nvc; m++)
                    for (int l = 0; l < dim; l++) {
                        const double v = p->h_vertices[(size_t)cells[(size_t)c * nvc + m] * dim + l];
                        box[((size_t)t * 2) * dim + l] = std::min(box[((size_t)t * 2) * dim + l], v);
                        box[((size_t)t * 2 + 1) * dim + l] = std::max(box[((size_t)t * 2 + 1) * dim + l], v);
                    }
        }
        if (upload(p, box.data(), box.size(), &S.tile_box)) return PNB_ERR_CUDA;
    }
    std::vector<int> units;
    if (p->nblocks > 0) {
        // batched blocks: the group pairs inside every block
        const int align = S.G * TD;
        for (int k = 0; k < p->nblocks; k++) {
            const int g0 = p->h_blk_dof[k] / align, g1 = (p->h_blk_dof[k + 1] + align - 1) / align;
            for (int gr = g0; gr < g1; gr++)
                for (int gc = gr; gc < g1; gc++) { units.push_back(gr); units.push_back(gc); }
        }
        p->h_block_units = units;
    } else
    // heavy (near-diagonal) units first
    for (int dg = 0; dg < S.ngroups; dg++)
        for (int gr = 0; gr + dg < S.ngroups; gr++) { units.push_back(gr); units.push_back(gr + dg); }
    S.nunits = (int)units.size() / 2;
    S.nunits_all = S.nunits;
    const int ND = nvc * (nvc + 1) / 2;
    int rc = 0;
    rc |= upload(p, tptr.data(), tptr.size(), &S.tile_ptr);
    rc |= upload(p, tlist.data(), tlist.size(), &S.tile_cells);
    rc |= upload(p, tloc.data(), tloc.size(), &S.tile_loc);
    rc |= upload(p, units.data(), units.size(), &S.units);
    rc |= dalloc(p, (size_t)S.dgroups * nc * ND, &S.DXp);
    rc |= dalloc(p, (size_t)S.dgroups * nc * ND, &S.DYp);
    S.maxcells = PNB_SB;
    for (int t = 0; t < S.ntiles; t++) S.maxcells = std::max(S.maxcells, tptr[t + 1] - tptr[t]);
    rc |= dalloc(p, (size_t)S.ntiles * S.tf_width, &S.tileflag);
    rc |= dalloc(p, (size_t)S.nunits, &S.unitflag);
    rc |= dalloc(p, (size_t)S.nunits + 1, &S.nearunits);
    if (rc) return PNB_ERR_CUDA;
    p->tiles_ready = true;
    (void)N;
    return 0;
}

extern "C" int pnb_problem_create(const pnb_mesh_t *mesh, const pnb_dofmap_t *dm, const pnb_kernel_t *kernel,
                                  const pnb_rules_t *rules, int device, pnb_problem **out)
{
    if (!mesh || !dm || !kernel || !rules || !out) return fail(PNB_ERR_ARG, "null argument");
    if (mesh->dim != 1 && mesh->dim != 2) return fail(PNB_ERR_UNSUPPORTED, "only 1D and 2D meshes are supported");
    if (dm->dofs_per_element != mesh->dim + 1) return fail(PNB_ERR_UNSUPPORTED, "only P1 DoFMaps are supported");
    // kernels of the form C |x-y|^singularity chi(|x-y| <= horizon): fractional (singularity = -d-2s), constant /
    // indicator (0), inverse distance / peridynamic (-1)  (kernelsCy.pyx:75-183, 273-386)
    if (kernel->kernel_type != PNB_KERNEL_FRACTIONAL && kernel->kernel_type != PNB_KERNEL_INDICATOR &&
        kernel->kernel_type != PNB_KERNEL_PERIDYNAMIC)
        return fail(PNB_ERR_UNSUPPORTED, "kernel type not supported");
    if (kernel->kernel_type != PNB_KERNEL_FRACTIONAL && !std::isfinite(kernel->horizon2))
        return fail(PNB_ERR_UNSUPPORTED, "integrable kernels need a finite horizon");
    if (kernel->dim != mesh->dim) return fail(PNB_ERR_ARG, "Kernel dimension must match dm.mesh dimension");
    if (kernel->horizon2 <= 0. || kernel->horizon2 != kernel->horizon2) return fail(PNB_ERR_ARG, "horizon must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PNB_ERR_NO_DEVICE, "no CUDA device: libpnb200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(PNB_ERR_ARG, "invalid device index");
    ON_DEVICE(device);

    const double tc0 = wall_ms();
    pnb_problem *p = new pnb_problem();
    p->device = device;
    const int dim = mesh->dim, nvc = dim + 1, nc = mesh->num_cells, N = dm->num_dofs, nb = mesh->num_bfacets;
    p->dim = dim; p->nc = nc; p->N = N; p->nb = nb;
    DProblem &P = p->P;
    P.dim = dim; P.nc = nc; P.nv = mesh->num_vertices; P.N = N; P.nb = nb;

    // precomputeSimplices (nonlocalOperator_{SCALAR}.pxi:111-126): same summation order
    std::vector<double> simplices((size_t)nc * nvc * dim), centers((size_t)nc * dim, 0.), hcell(nc);
    const double fac = 1. / nvc;
    for (int c = 0; c < nc; c++) {
        for (int m = 0; m < nvc; m++) {
            const int k = mesh->cells[(size_t)c * nvc + m];
            if (k < 0 || k >= mesh->num_vertices) { delete p; return fail(PNB_ERR_ARG, "cell refers to a vertex out of range"); }
            for (int l = 0; l < dim; l++) {
                const double v = mesh->vertices[(size_t)k * dim + l];
                simplices[((size_t)c * nvc + m) * dim + l] = v;
                centers[(size_t)c * dim + l] += v;
            }
        }
        for (int l = 0; l < dim; l++) centers[(size_t)c * dim + l] *= fac;
        // get_h_simplex (nonlocalOperator.pyx:114-118, 152-160)
        const double *sx = &simplices[(size_t)c * nvc * dim];
        if (dim == 1) hcell[c] = fabs(sx[1] - sx[0]);
        else {
            double hmax = 0.;
            for (int i = 0; i < 2; i++)
                for (int j = i + 1; j < 3; j++) {
                    const double h2 = (sx[2 * j] - sx[2 * i]) * (sx[2 * j] - sx[2 * i]) + (sx[2 * j + 1] - sx[2 * i + 1]) * (sx[2 * j + 1] - sx[2 * i + 1]);
                    hmax = std::max(hmax, h2);
                }
            hcell[c] = sqrt(hmax);
        }
    }
    std::vector<double> bsimplices((size_t)nb * dim * dim), bcenters((size_t)nb * dim, 0.), bvol(nb), bh(nb);
    const double bfac = 1. / dim;
    for (int f = 0; f < nb; f++) {
        for (int m = 0; m < dim; m++) {
            const int k = mesh->bfacets[(size_t)f * dim + m];
            for (int l = 0; l < dim; l++) {
                const double v = mesh->vertices[(size_t)k * dim + l];
                bsimplices[((size_t)f * dim + m) * dim + l] = v;
                bcenters[(size_t)f * dim + l] += v;
            }
        }
        for (int l = 0; l < dim; l++) bcenters[(size_t)f * dim + l] *= bfac;
        if (dim == 1) { bvol[f] = 1.; bh[f] = 1.; }
        else {
            const double *sx = &bsimplices[(size_t)f * 4];
            double h2 = 0.;
            for (int k = 0; k < 2; k++) h2 += (sx[2 + k] - sx[k]) * (sx[2 + k] - sx[k]);
            bvol[f] = bh[f] = sqrt(h2);
        }
    }
    p->h_cells.assign(mesh->cells, mesh->cells + (size_t)nc * nvc);
    p->h_vertices.assign(mesh->vertices, mesh->vertices + (size_t)mesh->num_vertices * dim);
    p->h_dofs.assign(dm->dofs, dm->dofs + (size_t)nc * nvc);
    if (dim == 2) {
        p->h_centers = centers;
        p->h_h.assign(mesh->h, mesh->h + nc);
    }
    int rc = 0;
    rc |= upload(p, simplices.data(), simplices.size(), &P.simplices);
    rc |= upload(p, centers.data(), centers.size(), &P.centers);
    rc |= upload(p, mesh->cells, (size_t)nc * nvc, &P.cells);
    rc |= upload(p, dm->dofs, (size_t)nc * nvc, &P.dofs);
    rc |= upload(p, mesh->vol, (size_t)nc, &P.vol);
    rc |= upload(p, mesh->h, (size_t)nc, &P.h);
    rc |= upload(p, hcell.data(), (size_t)nc, &P.hcell);
    rc |= upload(p, mesh->bfacets, (size_t)nb * dim, &P.bfacets);
    rc |= upload(p, bsimplices.data(), bsimplices.size(), &P.bsimplices);
    rc |= upload(p, bcenters.data(), bcenters.size(), &P.bcenters);
    rc |= upload(p, bvol.data(), (size_t)nb, &P.bvol);
    rc |= upload(p, bh.data(), (size_t)nb, &P.bh);
    if (rc) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }

    // table-driven power for both kernels, per-cell logs for the fast order selection
    {
        // third table: boundary kernel divided by |x-y| (the unit vector of the surface form folded into the power)
        PowTab tabs[3];
        const double scal[3] = {kernel->scaling, kernel->bscaling, kernel->bscaling};
        const double expo[3] = {0.5 * kernel->singularity, 0.5 * kernel->bsingularity, 0.5 * kernel->bsingularity - 0.5};
        // exponent window of the tables: top above 4 diam^2 (no two points of the mesh are further apart than the
        // diagonal of its bounding box), 256 binary exponents down from there
        int ex = 0;
        frexp(4. * mesh->diam * mesh->diam, &ex);
        p->pow_eoff = 254 - ex;
        for (int t = 0; t < 3; t++) {
            build_powtab(&tabs[t], scal[t], expo[t], p->pow_eoff);
            tabs[t].horizon2 = std::isfinite(kernel->horizon2) ? kernel->horizon2 : INFINITY;
        }
        const PowTab *dt = nullptr;
        rc |= upload(p, tabs, 3, &dt);
        P.pow_int = dt;
        P.pow_bnd = dt + 1;
        P.pow_bnd_unit = dt + 2;
        std::vector<float> lhf(nc), ahf(nc);
        const double H0 = mesh->diam / sqrt(8.);
        for (int c = 0; c < nc; c++) {
            lhf[c] = (float)log(mesh->h[c]);
            ahf[c] = (float)fabs(log(mesh->h[c] / H0));
        }
        rc |= upload(p, lhf.data(), (size_t)nc, &P.lhf);
        rc |= upload(p, ahf.data(), (size_t)nc, &P.ahf);
        std::vector<float> lhc(nc), ahc(nc), lhb(nb), ahb(nb);
        for (int c = 0; c < nc; c++) { lhc[c] = (float)log(hcell[c]); ahc[c] = (float)fabs(log(hcell[c] / H0)); }
        for (int f = 0; f < nb; f++) { lhb[f] = (float)log(bh[f]); ahb[f] = (float)fabs(log(bh[f] / H0)); }
        rc |= upload(p, lhc.data(), (size_t)nc, &P.lhcf);
        rc |= upload(p, ahc.data(), (size_t)nc, &P.ahcf);
        rc |= upload(p, lhb.data(), (size_t)nb, &P.lhbf);
        rc |= upload(p, ahb.data(), (size_t)nb, &P.ahbf);
        if (rc) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }
    }
    P.labels = nullptr; P.blabels = nullptr; P.active_class = 0; P.pair_orientation = 0; P.pair_filter = 0;
    for (int l = 0; l < 4; l++) P.pair_class[l] = P.bpair_class[l] = 0;
    if (kernel->cell_labels) {
        if (nb > 0 && !kernel->bfacet_labels) { pnb_problem_destroy(p); return fail(PNB_ERR_ARG, "cell labels without boundary facet labels"); }
        rc |= upload(p, kernel->cell_labels, (size_t)nc, &P.labels);
        rc |= upload(p, kernel->bfacet_labels, (size_t)nb, &P.blabels);
        if (rc) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }
        P.active_class = kernel->active_class;
        for (int l1 = 0; l1 < 4; l1++)
            for (int l2 = 0; l2 < 4; l2++) {
                if (kernel->pair_class[l1 * 4 + l2] != kernel->pair_class[l2 * 4 + l1]) p->path = 1;
                P.pair_class[l1] |= (unsigned)kernel->pair_class[l1 * 4 + l2] << (8 * l2);
                P.bpair_class[l1] |= (unsigned)kernel->bpair_class[l1 * 4 + l2] << (8 * l2);
            }
        p->h_labels.assign(kernel->cell_labels, kernel->cell_labels + nc);
        P.pair_orientation = kernel->pair_orientation ? 1 : 0;
        P.pair_filter = kernel->pair_filter == 1 ? 1 : 0;
        if (P.pair_orientation || P.pair_filter) p->path = 1;
        p->ordered_classes = p->path == 1;
    }
    P.s = kernel->s; P.C = kernel->scaling; P.Cb = kernel->bscaling;
    P.sing = kernel->singularity; P.bsing = kernel->bsingularity;
    P.expo = 0.5 * kernel->singularity;         // kernelsCy.pyx:159-183: -d/2 - s for the fractional kernel
    P.bexpo = 0.5 * kernel->bsingularity;       // kernelsCy.pyx:216-240
    P.horizon2 = std::isfinite(kernel->horizon2) ? kernel->horizon2 : INFINITY;
    p->finite = std::isfinite(kernel->horizon2);
    P.H0 = mesh->diam / sqrt(8.);               // nonlocalOperator_{SCALAR}.pxi:435
    const double Nord = kernel->order_num_dofs > 0 ? kernel->order_num_dofs : N;
    if (dim == 2) {
        P.c_int = (0.5 * kernel->target_order + 0.5) * log(Nord * (P.H0 * P.H0));
        P.c_bnd = (0.5 * kernel->btarget_order + 0.25) * log(Nord * (P.H0 * P.H0));
    } else {
        P.c_int = (kernel->target_order + 2.) * log(Nord * P.H0);
        P.c_bnd = (kernel->btarget_order + 1.) * log(Nord * P.H0);
    }

    const double tc1 = wall_ms();
    // ---- DoF tiles: ownership of cells and of the cell-diagonal blocks (both assembly paths) ----------------
    TileSched &S = p->S;
    const int TD = PNB_TD;
    S.ntiles = std::max(1, (N + TD - 1) / TD);
    S.G = 2;
    S.ngroups = (S.ntiles + S.G - 1) / S.G;
    S.dgroups = S.ngroups;
    S.tf_width = S.ntiles;
    S.cell_block = S.tile_block = S.blk_tile0 = S.blk_group0 = S.blk_n = S.blk_fptr = nullptr;
    S.blk_out = nullptr;
    std::vector<int> cell_block;
    if (mesh->num_blocks > 0) {
        // batched independent blocks: contiguous ranges of cells / dofs / boundary facets per block
        const int nbk = mesh->num_blocks, align = S.G * TD;
        if (!mesh->block_cell_ptr || !mesh->block_dof_ptr || !mesh->block_facet_ptr) { pnb_problem_destroy(p); return fail(PNB_ERR_ARG, "block ranges missing"); }
        p->nblocks = nbk;
        p->h_blk_cell.assign(mesh->block_cell_ptr, mesh->block_cell_ptr + nbk + 1);
        p->h_blk_dof.assign(mesh->block_dof_ptr, mesh->block_dof_ptr + nbk + 1);
        p->h_blk_facet.assign(mesh->block_facet_ptr, mesh->block_facet_ptr + nbk + 1);
        if (p->h_blk_cell[0] != 0 || p->h_blk_cell[nbk] != nc || p->h_blk_dof[0] != 0 || p->h_blk_dof[nbk] != N || p->h_blk_facet[0] != 0 ||
            p->h_blk_facet[nbk] != nb) { pnb_problem_destroy(p); return fail(PNB_ERR_ARG, "block ranges do not cover the mesh"); }
        std::vector<int> tile_block(S.ntiles, 0), blk_tile0(nbk), blk_group0(nbk), blk_n(nbk);
        std::vector<long long> blk_out(nbk);
        cell_block.assign(nc, 0);
        long long off = 0;
        int dg = 1;
        for (int k = 0; k < nbk; k++) {
            const int d0 = p->h_blk_dof[k], d1 = p->h_blk_dof[k + 1];
            if (d0 % align != 0 || d1 <= d0 || (d1 % align != 0 && k + 1 < nbk) || p->h_blk_cell[k + 1] < p->h_blk_cell[k] ||
                p->h_blk_facet[k + 1] < p->h_blk_facet[k]) { pnb_problem_destroy(p); return fail(PNB_ERR_ARG, "block dof ranges must be non-empty and start at multiples of pnb_block_alignment()"); }
            blk_tile0[k] = d0 / TD;
            blk_group0[k] = d0 / align;
            blk_n[k] = d1 - d0;
            blk_out[k] = off;
            off += (long long)(d1 - d0) * (d1 - d0);
            dg = std::max(dg, (d1 - d0 + align - 1) / align);
            for (int t = d0 / TD; t < (d1 + TD - 1) / TD; t++) tile_block[t] = k;
            for (int c = p->h_blk_cell[k]; c < p->h_blk_cell[k + 1]; c++) cell_block[c] = k;
        }
        p->blocks_out_doubles = off;
        S.dgroups = dg;
        S.tf_width = dg * S.G;
        rc |= upload(p, cell_block.data(), cell_block.size(), &S.cell_block);
        rc |= upload(p, tile_block.data(), tile_block.size(), &S.tile_block);
        rc |= upload(p, blk_tile0.data(), blk_tile0.size(), &S.blk_tile0);
        rc |= upload(p, blk_group0.data(), blk_group0.size(), &S.blk_group0);
        rc |= upload(p, blk_n.data(), blk_n.size(), &S.blk_n);
        rc |= upload(p, blk_out.data(), blk_out.size(), &S.blk_out);
        rc |= upload(p, p->h_blk_facet.data(), p->h_blk_facet.size(), &S.blk_fptr);
        if (rc) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }
        p->path = 1;        // the DoF-tile kernels
    }
    std::vector<int> home(nc);
    int64_t live = 0;
    for (int c = 0; c < nc; c++) {
        int mind = -1;
        for (int m = 0; m < nvc; m++) {
            const int d = dm->dofs[(size_t)c * nvc + m];
            if (d >= N) { pnb_problem_destroy(p); return fail(PNB_ERR_ARG, "dof index out of range"); }
            if (d >= 0 && (mind < 0 || d < mind)) mind = d;
        }
        if (mind >= 0) {
            home[c] = mind / TD; live++;
            if (!cell_block.empty() && (mind < p->h_blk_dof[cell_block[c]] || mind >= p->h_blk_dof[cell_block[c] + 1])) {
                pnb_problem_destroy(p);
                return fail(PNB_ERR_ARG, "a cell refers to a dof outside its block");
            }
        }
        // no dofs: spread such cells evenly (blocks: first tile of the block); they only feed cell-diagonal blocks of other cells
        else home[c] = cell_block.empty() ? (int)(((int64_t)c * S.ntiles) / std::max(nc, 1)) : p->h_blk_dof[cell_block[c]] / TD;
    }
    // pairs c1<=c2 that the reference does not skip (at least one non-negative dof)
    {
        const int64_t dead = nc - live;
        p->distinct_pairs = (int64_t)nc * (nc + 1) / 2 - dead * (dead + 1) / 2;
    }
    S.own_t0 = 0;
    S.own_t1 = S.ntiles;
    // dof -> cells
    std::vector<int> dptr(N + 1, 0), dcells;
    for (int c = 0; c < nc; c++)
        for (int m = 0; m < nvc; m++) { const int d = dm->dofs[(size_t)c * nvc + m]; if (d >= 0) dptr[d + 1]++; }
    for (int i = 0; i < N; i++) dptr[i + 1] += dptr[i];
    dcells.resize(dptr[N]);
    {
        std::vector<int> pos(dptr.begin(), dptr.end() - 1);
        for (int c = 0; c < nc; c++)
            for (int m = 0; m < nvc; m++) { const int d = dm->dofs[(size_t)c * nvc + m]; if (d >= 0) dcells[pos[d]++] = c * 4 + m; }
    }
    const int ND = nvc * (nvc + 1) / 2;
    rc |= upload(p, home.data(), home.size(), &S.home);
    rc |= upload(p, dptr.data(), dptr.size(), &S.dof_ptr);
    rc |= upload(p, dcells.data(), dcells.size(), &S.dof_cells);
    rc |= dalloc(p, (size_t)nc * ND, &S.Dbnd);
    rc |= dalloc(p, (size_t)nc * ND, &S.D);
    rc |= dalloc(p, 4, &S.err);
    rc |= dalloc(p, 8, &S.counters);
    if (rc) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }
    p->h_home = home;
    // the tile cell lists are only needed by the DoF-tile kernels (1D, row ranges): built on first use in 2D
    if (dim == 1 && build_tile_schedule(p)) { pnb_problem_destroy(p); return PNB_ERR_CUDA; }
    const double tc2 = wall_ms();
    rc = pnb_problem_set_rules(p, rules);
    if (rc) { pnb_problem_destroy(p); return rc; }
    if (getenv("PNB_BENCH_VERBOSE"))
        fprintf(stderr, "pnb_problem_create: mesh + upload %.1f ms, tile schedule %.1f ms, tables %.1f ms\n", tc1 - tc0, tc2 - tc1, wall_ms() - tc2);
    *out = p;
    return 0;
}

// ---------------------------------------------------------------------------
// classification / order kernels
// ---------------------------------------------------------------------------
__global__ void classify_kernel(DProblem P, int boundary, int64_t np, const int *pairs, int *panel, int *perm1, int *perm2)
{
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= np) return;
    int p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
    const int a = pairs[2 * n], b = pairs[2 * n + 1];
    const int pan = boundary ? panel_boundary(P, a, b, p1, p2) : panel_interior(P, a, b, p1, p2);
    panel[n] = pan;
    const int nvc = P.dim + 1;
    if (perm1)
        for (int k = 0; k < nvc; k++) perm1[n * nvc + k] = p1[k];
    if (perm2)
        for (int k = 0; k < nvc; k++) perm2[n * nvc + k] = (boundary && k >= P.dim) ? 0 : p2[k];
}

// one thread per c1, loops over c2 >= c1 (and the facets)
__global__ void max_order_kernel(DProblem P, int zero_exterior, int *out, unsigned long long *hist)
{
    const int c1 = blockIdx.x * blockDim.x + threadIdx.x;
    int best = 0;
    if (c1 < P.nc) {
        int p1[3], p2[3];
        for (int c2 = c1; c2 < P.nc; c2++) {
            const int pan = panel_interior(P, c1, c2, p1, p2);
            if (pan == PNB_IGNORED_PANEL) continue;
            best = max(best, pan);
            if (hist && pan >= -3 && pan < 256) atomicAdd(&hist[3 + pan], 1ull);
        }
        if (zero_exterior)
            for (int f = 0; f < P.nb; f++) best = max(best, panel_boundary(P, c1, f, p1, p2));
    }
    atomicMax(out, best);
}

extern "C" int pnb_block_alignment(void) { return 2 * PNB_TD; }

extern "C" int pnb_problem_set_path(pnb_problem *p, int path)
{
    if (!p || path < 0 || path > 1) return fail(PNB_ERR_ARG, "invalid argument");
    if (p->nblocks > 0 && path != 1) return fail(PNB_ERR_ARG, "batched blocks use the DoF-tile path");
    if (p->ordered_classes && path != 1) return fail(PNB_ERR_ARG, "ordered pair classes / pair orientation use the DoF-tile path");
    p->path = path;
    return 0;
}

extern "C" int pnb_max_order(pnb_problem *p, int zero_exterior, int32_t *max_order_out)
{
    if (!p || !max_order_out) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    CK(cudaMemset(p->S.err, 0, sizeof(int) * 4));
    max_order_kernel<<<(p->nc + 127) / 128, 128>>>(p->P, zero_exterior, p->S.err + 1, nullptr);
    CK(cudaGetLastError());
    int v = 0;
    CK(cudaMemcpy(&v, p->S.err + 1, sizeof(int), cudaMemcpyDeviceToHost));
    *max_order_out = v;
    return 0;
}

// sparsity pattern of getDense(trySparsification=True) (nonlocalAssembly_{SCALAR}.pxi:1293-1332): every pair of dofs of
// every cell pair that is not ignored.  One thread per first cell; the stores of a one are idempotent.
__global__ void sparsity_mask_kernel(DProblem P, unsigned char *mask, int64_t ld)
{
    const int c1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c1 >= P.nc) return;
    const int nvc = P.dim + 1;
    int d1[3], p1[3], p2[3];
    bool any1 = false;
    for (int m = 0; m < nvc; m++) { d1[m] = P.dofs[(size_t)c1 * nvc + m]; any1 |= d1[m] >= 0; }
    for (int c2 = c1; c2 < P.nc; c2++) {
        int d2[3];
        bool any2 = false;
        for (int m = 0; m < nvc; m++) { d2[m] = P.dofs[(size_t)c2 * nvc + m]; any2 |= d2[m] >= 0; }
        if (!any1 && !any2) continue;
        if (panel_interior(P, c1, c2, p1, p2) == PNB_IGNORED_PANEL) continue;
        for (int a = 0; a < 2 * nvc; a++) {
            const int I = a < nvc ? d1[a] : d2[a - nvc];
            if (I < 0) continue;
            for (int b = 0; b < 2 * nvc; b++) {
                const int J = b < nvc ? d1[b] : d2[b - nvc];
                if (J >= 0) mask[(size_t)I * ld + J] = 1;
            }
        }
    }
}

extern "C" int pnb_sparsity_mask(pnb_problem *p, unsigned char *dmask, int64_t ld)
{
    if (!p || !dmask) return fail(PNB_ERR_ARG, "null argument");
    if (ld < p->N) return fail(PNB_ERR_ARG, "leading dimension smaller than num_dofs");
    ON_DEVICE(p->device);
    CK(cudaMemset2D(dmask, (size_t)ld, 0, (size_t)p->N, (size_t)p->N));
    sparsity_mask_kernel<<<(p->nc + 127) / 128, 128>>>(p->P, dmask, ld);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int pnb_panel_histogram(pnb_problem *p, int64_t *hist)
{
    if (!p || !hist) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 259 * sizeof(unsigned long long)));
    CK(cudaMemset(d, 0, 259 * sizeof(unsigned long long)));
    CK(cudaMemset(p->S.err, 0, sizeof(int) * 4));
    max_order_kernel<<<(p->nc + 127) / 128, 128>>>(p->P, 0, p->S.err + 1, d);
    cudaError_t e = cudaMemcpy(hist, d, 259 * sizeof(int64_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CK(e);
    return 0;
}

extern "C" int pnb_classify_pairs(pnb_problem *p, int boundary, int64_t npairs, const int32_t *pairs, int32_t *panel,
                                  int32_t *perm1, int32_t *perm2)
{
    if (!p || !pairs || !panel) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    const int nvc = p->dim + 1;
    int *dpairs = nullptr, *dpanel = nullptr, *dp1 = nullptr, *dp2 = nullptr;
    const size_t n = (size_t)std::max<int64_t>(npairs, 1);
    CK(cudaMalloc(&dpairs, n * 2 * sizeof(int)));
    CK(cudaMalloc(&dpanel, n * sizeof(int)));
    CK(cudaMalloc(&dp1, n * nvc * sizeof(int)));
    CK(cudaMalloc(&dp2, n * nvc * sizeof(int)));
    int rc = 0;
    if (npairs > 0) {
        cudaMemcpy(dpairs, pairs, (size_t)npairs * 2 * sizeof(int), cudaMemcpyHostToDevice);
        classify_kernel<<<(unsigned)((npairs + 255) / 256), 256>>>(p->P, boundary, npairs, dpairs, dpanel, dp1, dp2);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(panel, dpanel, (size_t)npairs * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && perm1) e = cudaMemcpy(perm1, dp1, (size_t)npairs * nvc * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && perm2) e = cudaMemcpy(perm2, dp2, (size_t)npairs * nvc * sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(PNB_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(dpairs); cudaFree(dpanel); cudaFree(dp1); cudaFree(dp2);
    return rc;
}

// ---------------------------------------------------------------------------
// local matrices of a list of pairs (parity entry point)
// ---------------------------------------------------------------------------
// writes the reduced lane sums of a SINGULAR pair into the reference's local index space
template <int DIM>
__device__ __forceinline__ void singular_to_local(const double *acc, int lane, int common, const int *perm1, const int *perm2,
                                                  double scale, double *out /* NL, zero-initialised */)
{
    constexpr int NV = PairDims<DIM>::NV, NR = 2 * NV - 1, NA = NR * (NR + 1) / 2;
    // lane l owns compact entry l = (I,J), I<=J over NR rows
    int I = 0, J = 0, k = 0;
    bool mine = false;
    double v = 0.;
#pragma unroll
    for (int a = 0; a < NR; a++)
#pragma unroll
        for (int b = a; b < NR; b++) {
            if (k == lane) { I = a; J = b; mine = true; v = acc[k]; }
            k++;
        }
    (void)NA;
    const int rows = 2 * NV - common;
    if (!mine || J >= rows) return;
    // perm: dofs on the reordered simplices -> usual local numbering (nonlocalOperator_{SCALAR}.pxi:351-376)
    int i = I < NV ? perm1[I] : NV + perm2[I - NV + common];
    int j = J < NV ? perm1[J] : NV + perm2[J - NV + common];
    const int kk = j < i ? tri_idx(2 * NV, j, i) : tri_idx(2 * NV, i, j);
    out[kk] = v * scale;
}

template <int DIM>
__global__ void local_matrices_kernel(DProblem P, int boundary, int path, int far_mask, int64_t np, const int *pairs,
                                      int *panel_out, double *contrib, int *err)
{
    constexpr int NV = PairDims<DIM>::NV, NL = PairDims<DIM>::NL, ND = PairDims<DIM>::ND;
    const int lane = threadIdx.x & 31;
    const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (n >= np) return;
    const int a = pairs[2 * n], b = pairs[2 * n + 1];
    int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
    if (boundary) {
        const int pan = panel_boundary(P, a, b, p1, p2);
        if (lane == 0) panel_out[n] = pan;
        double *out = contrib + n * ND;
        if (pan > P.max_order) { if (lane == 0) atomicMax(err, pan); return; }
        double acc[ND];
        lanes_boundary<DIM>(P, a, b, pan, p1, p2, lane, 32, acc);
        warp_allreduce<ND>(acc);
        // vol factors: regular vol1*vol2 (nonlocalOperator_{SCALAR}.pxi:1071), singular 2D -2 vol1 vol2
        // (fractionalLaplacian2D.pyx:1375), singular 1D vol1 (fractionalLaplacian1D.pyx:724)
        double scale;
        if (pan >= 1) scale = P.vol[a] * P.bvol[b];
        else scale = DIM == 2 ? -2.0 * P.vol[a] * P.bvol[b] : P.vol[a];
        if (lane < ND) out[lane] = 0.;
        __syncwarp();
        int k = 0;
#pragma unroll
        for (int I = 0; I < NV; I++)
#pragma unroll
            for (int J = I; J < NV; J++) {
                if (k == lane) {
                    const int i = pan >= 1 ? I : p1[I], j = pan >= 1 ? J : p1[J];
                    const int kk = j < i ? tri_idx(NV, j, i) : tri_idx(NV, i, j);
                    out[kk] = acc[k] * scale;
                }
                k++;
            }
        return;
    }
    const int pan = panel_interior(P, a, b, p1, p2);
    if (lane == 0) panel_out[n] = pan;
    double *out = contrib + n * NL;
    if (lane < NL) out[lane] = 0.;
    __syncwarp();
    if (pan == PNB_IGNORED_PANEL) return;
    if (pan > P.max_order) { if (lane == 0) atomicMax(err, pan); return; }
    if (pan >= 1) {
        const double vol = P.vol[a] * P.vol[b];
        if (path == 1 && DIM == 2 && pan >= 2 && pan <= PNB_FAR_MAX_ORDER && ((far_mask >> pan) & 1) &&
            !(P.horizon2 < INFINITY && pair_relative_position(P, a, b) == 2)) {
            if (lane == 0) {
                double s1[3][2], s2[3][2], xy[9], xx[6], yy[6];
                load_simplex<2>(P.simplices, a, 3, s1);
                load_simplex<2>(P.simplices, b, 3, s2);
                const PowCtx kv(P.pow_int);
                far_eval_2d(P.far_rules[pan], s1, s2, kv, true, xy, xx, yy);
                for (int I = 0; I < 3; I++)
                    for (int J = I; J < 3; J++) {
                        out[tri_idx(6, I, J)] = xx[tri_idx(3, I, J)] * vol;
                        out[tri_idx(6, 3 + I, 3 + J)] = yy[tri_idx(3, I, J)] * vol;
                    }
                for (int I = 0; I < 3; I++)
                    for (int J = 0; J < 3; J++) out[tri_idx(6, I, 3 + J)] = xy[3 * I + J] * vol;
            }
            return;
        }
        double acc[NL];
        if (P.horizon2 < INFINITY && pair_relative_position(P, a, b) == 2) lanes_cut_interior<DIM>(P, a, b, pan, lane, 32, acc);
        else lanes_regular_interior<DIM>(P, a, b, pan, lane, 32, acc);
        warp_allreduce<NL>(acc);
#pragma unroll
        for (int k = 0; k < NL; k++)
            if (k == lane) out[k] = acc[k] * vol;
    } else {
        constexpr int NR = 2 * NV - 1, NA = NR * (NR + 1) / 2;
        double acc[NA];
        lanes_singular_interior<DIM>(P, a, b, pan, p1, p2, lane, 32, acc);
        warp_allreduce<NA>(acc);
        // vol = 4 vol1 vol2 in 2D (fractionalLaplacian2D.pyx:851), vol1 vol2 in 1D (fractionalLaplacian1D.pyx:378)
        const double scale = (DIM == 2 ? 4.0 : 1.0) * P.vol[a] * P.vol[b];
        singular_to_local<DIM>(acc, lane, -pan, p1, p2, scale, out);
    }
}

extern "C" int pnb_local_matrices(pnb_problem *p, int boundary, int path, int64_t npairs, const int32_t *pairs,
                                  int32_t *panel, double *contrib)
{
    if (!p || !pairs || !panel || !contrib) return fail(PNB_ERR_ARG, "null argument");
    if (!p->has_singular) return fail(PNB_ERR_ARG, "problem was created without quadrature tables");
    ON_DEVICE(p->device);
    const int nvc = p->dim + 1;
    const int nloc = boundary ? nvc * (nvc + 1) / 2 : (2 * nvc) * (2 * nvc + 1) / 2;
    int *dpairs = nullptr, *dpanel = nullptr;
    double *dc = nullptr;
    const size_t n = (size_t)std::max<int64_t>(npairs, 1);
    CK(cudaMalloc(&dpairs, n * 2 * sizeof(int)));
    CK(cudaMalloc(&dpanel, n * sizeof(int)));
    CK(cudaMalloc(&dc, n * nloc * sizeof(double)));
    int rc = 0;
    if (npairs > 0) {
        cudaMemcpy(dpairs, pairs, (size_t)npairs * 2 * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemset(p->S.err, 0, 4 * sizeof(int));
        const unsigned blocks = (unsigned)((npairs * 32 + 255) / 256);
        if (p->dim == 2) local_matrices_kernel<2><<<blocks, 256>>>(p->P, boundary, path, p->far_mask, npairs, dpairs, dpanel, dc, p->S.err);
        else local_matrices_kernel<1><<<blocks, 256>>>(p->P, boundary, path, 0, npairs, dpairs, dpanel, dc, p->S.err);
        cudaError_t e = cudaGetLastError();
        int herr[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) e = cudaMemcpy(herr, p->S.err, 4 * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(panel, dpanel, (size_t)npairs * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(contrib, dc, (size_t)npairs * nloc * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(PNB_ERR_CUDA, cudaGetErrorString(e));
        else if (herr[0] > 0) rc = fail(PNB_ERR_ORDER, "regular quadrature order " + std::to_string(herr[0]) + " exceeds the supplied tables (max_order " + std::to_string(p->P.max_order) + ")");
    }
    cudaFree(dpairs); cudaFree(dpanel); cudaFree(dc);
    return rc;
}

// ---------------------------------------------------------------------------
// dense assembly: tile kernel
// ---------------------------------------------------------------------------
// cell data of one batch of PNB_SB cells, structure-of-arrays (bank-conflict free for lane = cell)
template <int DIM> struct CellBatch {
    static constexpr int NV = DIM + 1;
    double sx[NV * DIM][PNB_SB];  // simplex coordinates [vertex*DIM + axis][cell]
    double cx[DIM][PNB_SB];       // centers
    double vol[PNB_SB];
    int v[NV][PNB_SB];            // global vertex ids
    float lh[PNB_SB], ah[PNB_SB]; // (float) log h, |log(h/H0)|
    int cell[PNB_SB], loc[PNB_SB], home[PNB_SB], any[PNB_SB];
};


template <int DIM, bool NEAR> struct TileSmem {
    static constexpr int NV = PairDims<DIM>::NV, NX = PairDims<DIM>::NX, ND = PairDims<DIM>::ND;
    PowTab pw;
    FarRule far[NEAR ? 1 : PNB_FAR_MAX_ORDER + 1];
    double acc[PNB_TD][PNB_TD + 1];
    double dxy[PNB_SB * PNB_SB][2 * ND];   // cell-diagonal blocks of the pairs of the sub-batch (slot = k1*SB+k2)
    double nv[NEAR ? PNB_SB * PNB_SB : 1][NX];   // near pass: cross blocks kept for the mirrored update of diagonal tiles
    unsigned char slotD[PNB_SB * PNB_SB];  // dxy[slot] holds data of this sub-batch
    double partial[NEAR ? 64 : 1][PairDims<DIM>::NL];   // near pass: slice sums of split pairs
    CellBatch<DIM> rb, cb;
    int list[PNB_SB * PNB_SB];
    int listpanel[PNB_SB * PNB_SB];
    int warpcnt[PNB_THREADS / 32];
    int clscnt[(PNB_FAR_MAX_ORDER - 1) * (PNB_THREADS / 32)];
    int nlist;
    int anynear;
    int anyD;
    int cursor[2];      // near pass: next item of the sub-batch (items are taken by the warps as they become free)
    // followed by DXs[maxcells][ND], DYs[maxcells][ND] (dynamic)
};

// getQuadOrder (fractionalLaplacian2D.pyx:622-642) in single precision.  Returns the order when the
// result is certain (the value handed to ceil() is farther than the error bound from an integer), else -1
// and the caller repeats the selection with the reference's exact double arithmetic.
__device__ __forceinline__ int fast_order_2d(double d2, float lh1, float lh2, float ah1, float ah2, float cf, float sf)
{
    // error of the FP32 evaluation: ~1e-6 absolute in the logarithms and 6e-8 relative in the divisions, numerators up to
    // ~20 over denominators >= 0.4 -> below 1e-4; values closer than MARGIN to an integer are re-decided in FP64
    const float MARGIN = 4e-4f;
    const float Ld = 0.5f * __logf((float)d2);
    const float l1 = Ld - lh1, l2 = Ld - lh2;
    const float m = fmaxf(ah1, ah2);
    const float num1 = cf + (sf - 1.f) * ah2 + m - sf * l2;
    const float num2 = cf + (sf - 1.f) * ah1 + m - sf * l1;
    const float f1 = __fdividef(num1, fmaxf(l1, 0.f) + 0.4f);
    const float f2 = __fdividef(num2, fmaxf(l2, 0.f) + 0.4f);
    const float g = fmaxf(f1, f2);
    if (g <= 2.f - MARGIN) return 2;
    const float k = ceilf(g);
    if (k - g > MARGIN && g - (k - 1.f) > MARGIN && g < 250.f) return (int)k;
    return -1;
}

template <int DIM>
__device__ __forceinline__ void load_batch(const DProblem &P, const TileSched &S, CellBatch<DIM> &b, int beg, int off, int tid)
{
    constexpr int NV = DIM + 1;
    if (tid < PNB_SB) {
        const int c = S.tile_cells[beg + off + tid];
        const bool ok = c >= 0;
        b.cell[tid] = c;
        b.loc[tid] = S.tile_loc[beg + off + tid];
        b.home[tid] = ok ? S.home[c] : -1;
        int any = 0;
        if (ok) {
#pragma unroll
            for (int m = 0; m < NV; m++) {
                any |= P.dofs[(size_t)c * NV + m] >= 0;
                b.v[m][tid] = P.cells[(size_t)c * NV + m];
            }
            b.vol[tid] = P.vol[c];
            b.lh[tid] = P.lhf[c];
            b.ah[tid] = P.ahf[c];
#pragma unroll
            for (int j = 0; j < DIM; j++) b.cx[j][tid] = P.centers[(size_t)c * DIM + j];
        }
        b.any[tid] = any;
    } else if (tid >= 32 && tid < 32 + PNB_SB * NV * DIM) {
        const int e = tid - 32, comp = e / PNB_SB, k = e - comp * PNB_SB;
        const int c = S.tile_cells[beg + off + k];
        if (c >= 0) b.sx[comp][k] = P.simplices[(size_t)c * (NV * DIM) + comp];
    }
}

// address of entry (row, col) of the operator: one matrix with leading dimension ld whose first row is row tile own_t0,
// or (batched blocks) the dense operator of the block that holds tile t
__device__ __forceinline__ double *entry_ptr(const TileSched &S, double *A, int64_t ld, int t, int row, int col)
{
    if (S.tile_block) {
        const int k = S.tile_block[t];
        const int o = S.blk_tile0[k] * PNB_TD;
        return A + S.blk_out[k] + (size_t)(row - o) * S.blk_n[k] + (col - o);
    }
    return A + (size_t)(row - S.own_t0 * PNB_TD) * ld + col;
}

// adds the NV x NV cross block X of the pair (row cell, column cell) to the tile; conflict free within a sub-batch
template <int DIM>
__device__ __forceinline__ void scatter_block(double (*acc)[PNB_TD + 1], int rloc, int cloc, const double *X, bool transposed)
{
    constexpr int NV = DIM + 1;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const int a = (rloc >> (8 * i)) & 0xFF;
        if (a == 0xFF) continue;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int b = (cloc >> (8 * j)) & 0xFF;
            if (b == 0xFF) continue;
            if (!transposed) acc[a][b] += X[i * NV + j];
            else acc[b][a] += X[i * NV + j];
        }
    }
}

// NEAR = false: far pass.  Evaluates the regular pairs of order 2..5 one per thread (binned by order), writes
//               the tiles (A = ...) and flags the tiles that hold other pairs.
// NEAR = true:  near pass over the flagged units.  Evaluates singular pairs and the remaining regular pairs one
//               per warp and adds to the tiles (A += ...).  Same CTA <-> tile ownership in both passes: no
//               atomics on floating point data anywhere.
// FIN: finite horizon (REMOTE / CUT classes, re-triangulated pairs).  The infinite-horizon instantiation of the near pass
// does not carry the registers and the stack of the re-triangulation (236 registers, 720 bytes of stack).  (Forcing two
// CTAs per SM on it -- 128 registers, spills -- doubled its time: measured.)
template <int DIM, bool NEAR, bool FIN>
__global__ void __launch_bounds__(PNB_THREADS, NEAR ? 1 : 2)
tile_kernel(DProblem P, TileSched S, double *A, int64_t ld, int far_mask)
{
    constexpr int NV = PairDims<DIM>::NV, NX = PairDims<DIM>::NX, ND = PairDims<DIM>::ND, NL = PairDims<DIM>::NL;
    constexpr int TD = PNB_TD, SB = PNB_SB, NW = PNB_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<DIM, NEAR> &sm = *reinterpret_cast<TileSmem<DIM, NEAR> *>(smem_raw);
    double *DXs = reinterpret_cast<double *>(smem_raw + sizeof(TileSmem<DIM, NEAR>));
    double *DYs = DXs + (size_t)S.maxcells * ND;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int unit = NEAR ? S.nearunits[blockIdx.x] : blockIdx.x;
    const int gr = S.units[2 * unit], gc = S.units[2 * unit + 1];
    // batched blocks: the cell-diagonal staging is indexed by the group inside the block
    const int goff = S.tile_block ? S.blk_group0[S.tile_block[gr * S.G]] : 0;
    unsigned long long my_pairs = 0;
    const float cf = (float)P.c_int, sf = (float)fmax(-0.5 * (P.sing + 2), 0.);
    constexpr bool finite = FIN;
    {   // stage the power table
        const double *src = reinterpret_cast<const double *>(P.pow_int);
        double *dst = reinterpret_cast<double *>(&sm.pw);
        for (int e = tid; e < (int)(sizeof(PowTab) / sizeof(double)); e += PNB_THREADS) dst[e] = src[e];
        if (!NEAR && DIM == 2) {
            const double *fs = reinterpret_cast<const double *>(P.far_rules);
            double *fd = reinterpret_cast<double *>(&sm.far[0]);
            for (int e = tid; e < (int)((PNB_FAR_MAX_ORDER + 1) * sizeof(FarRule) / sizeof(double)); e += PNB_THREADS) fd[e] = fs[e];
        }
    }
    __syncthreads();
    const PowCtx kv(&sm.pw);
    bool unit_near = false;

    for (int rt = gr * S.G; rt < min((gr + 1) * S.G, S.ntiles); rt++)
        for (int ct = max(gc * S.G, rt); ct < min((gc + 1) * S.G, S.ntiles); ct++) {
            if (S.tile_block && S.tile_block[rt] != S.tile_block[ct]) continue;
            const size_t tfi = (size_t)rt * S.tf_width + (S.tile_block ? ct - S.blk_tile0[S.tile_block[rt]] : ct);
            if (NEAR && !S.tileflag[tfi]) continue;
            if ((finite || P.pair_filter == 1) && rt != ct) {
                // finite horizon: tiles whose bounding boxes are further apart than the horizon hold only REMOTE pairs;
                // touching pairs only (pair_filter 1): cells that share a vertex sit in tiles whose boxes share that point
                double g2 = 0.;
                for (int l = 0; l < DIM; l++) {
                    const double lo1 = S.tile_box[((size_t)rt * 2) * DIM + l], hi1 = S.tile_box[((size_t)rt * 2 + 1) * DIM + l];
                    const double lo2 = S.tile_box[((size_t)ct * 2) * DIM + l], hi2 = S.tile_box[((size_t)ct * 2 + 1) * DIM + l];
                    const double g = fmax(0., fmax(lo2 - hi1, lo1 - hi2));
                    g2 += g * g;
                }
                // strict margin: a pair is REMOTE when its smallest vertex distance is >= the horizon
                if (P.pair_filter == 1 ? g2 > 0. : g2 > P.horizon2 * (1. + 1e-12)) {
                    // the far pass defines every entry of the matrix: a skipped tile (and its mirror image) is zero
                    if (!NEAR) {
                        const bool own_r = rt >= S.own_t0 && rt < S.own_t1, own_c = ct >= S.own_t0 && ct < S.own_t1;
                        const int r0 = rt * TD, c0 = ct * TD;
                        for (int e = tid; e < TD * TD; e += PNB_THREADS) {
                            const int a = e / TD, b = e - a * TD;
                            if (r0 + a < P.N && c0 + b < P.N) {
                                if (own_r) *entry_ptr(S, A, ld, rt, r0 + a, c0 + b) = 0.;
                                if (own_c) *entry_ptr(S, A, ld, rt, c0 + b, r0 + a) = 0.;
                            }
                        }
                    }
                    continue;
                }
            }
            const bool diag = rt == ct;
            // row-block ownership: the direct image belongs to the owner of row tile rt, the mirror image to the
            // owner of row tile ct; tiles that touch no owned row are not computed by this GPU
            const bool own_r = rt >= S.own_t0 && rt < S.own_t1, own_c = ct >= S.own_t0 && ct < S.own_t1;
            if (!own_r && !own_c) continue;
            const int rbeg = S.tile_ptr[rt], nR = S.tile_ptr[rt + 1] - rbeg;
            const int cbeg = S.tile_ptr[ct], nC = S.tile_ptr[ct + 1] - cbeg;
            __syncthreads();
            for (int e = tid; e < TD * (TD + 1); e += PNB_THREADS) (&sm.acc[0][0])[e] = 0.;
            for (int e = tid; e < nR * ND; e += PNB_THREADS) DXs[e] = 0.;
            for (int e = tid; e < nC * ND; e += PNB_THREADS) DYs[e] = 0.;
            if (tid == 0) sm.anynear = 0;
            for (int rb = 0; rb < nR; rb += SB) {
                __syncthreads();
                load_batch<DIM>(P, S, sm.rb, rbeg, rb, tid);
                for (int cb = diag ? rb : 0; cb < nC; cb += SB) {
                    __syncthreads();    // S0: previous sub-batch done
                    load_batch<DIM>(P, S, sm.cb, cbeg, cb, tid);
                    if (tid == 0) sm.anyD = 0;
                    __syncthreads();    // S1: batches visible
                    // ---- phase 1: classify every pair of the sub-batch ----
                    const int k1 = tid / SB, k2 = tid % SB;
                    const int K1 = sm.rb.cell[k1], K2 = sm.cb.cell[k2];
                    int todo = 0;  // 0 nothing, >0 regular order (queued), <0 singular panel (queued)
                    int relpos = 0;
                    int cls = 0;   // far pass: order 2..5 of a pair for the thread-per-pair evaluator, else 0
                    bool countD = false;
                    sm.slotD[tid] = 0;
                    // diagonal tiles: every unordered pair once (batches rb <= cb; inside a batch k1 <= k2)
                    if (K1 >= 0 && K2 >= 0 && K1 != K2 ? (!diag || rb < cb || k1 < k2) : (K1 >= 0 && K1 == K2 && diag)) {
                        countD = sm.rb.home[k1] == rt && sm.cb.home[k2] == ct;
                        const bool rin = (sm.rb.loc[k1] & 0x00FFFFFF) != 0x00FFFFFF, cin = (sm.cb.loc[k2] & 0x00FFFFFF) != 0x00FFFFFF;
                        if ((sm.rb.any[k1] || sm.cb.any[k2]) && (countD || (rin && cin)) &&
                            (!P.labels || pnb_class_active(P, P.labels[min(K1, K2)], P.labels[max(K1, K2)]))) {
                            int panel;
                            if (K1 == K2) panel = -NV;
                            else {
                                int v1[NV], v2[NV];
#pragma unroll
                                for (int m = 0; m < NV; m++) { v1[m] = sm.rb.v[m][k1]; v2[m] = sm.cb.v[m][k2]; }
                                panel = -shared_vertices(v1, NV, v2, NV);
                                // pair filter 1: touching pairs only (see pnb_kernel_t.pair_filter)
                                if (panel == 0 && P.pair_filter == 1) panel = PNB_IGNORED_PANEL;
                                if (panel == 0 && finite) {
                                    relpos = pair_relative_position(P, min(K1, K2), max(K1, K2));
                                    if (relpos == 1) panel = PNB_IGNORED_PANEL;      // REMOTE
                                }
                                if (panel == 0) {
                                    if (DIM == 2) {
                                        const double a = sm.rb.cx[0][k1] - sm.cb.cx[0][k2], b = sm.rb.cx[DIM - 1][k1] - sm.cb.cx[DIM - 1][k2];
                                        panel = fast_order_2d(a * a + b * b, sm.rb.lh[k1], sm.cb.lh[k2], sm.rb.ah[k1], sm.cb.ah[k2], cf, sf);
                                    } else panel = -1;
                                    if (panel < 0) {
                                        // getPanelType evaluates (c1 <= c2): keep the operand order of the reference
                                        const int c1 = min(K1, K2), c2 = max(K1, K2);
                                        const double d = center_distance(P.centers + (size_t)c1 * DIM, P.centers + (size_t)c2 * DIM, DIM);
                                        panel = quad_order_interior(P, P.h[c1], P.h[c2], d);
                                    }
                                }
                            }
                            const bool is_far = DIM == 2 && panel >= 2 && panel <= PNB_FAR_MAX_ORDER && ((far_mask >> panel) & 1) && relpos != 2;
                            if (panel == PNB_IGNORED_PANEL) {}
                            else if (panel > P.max_order) atomicMax(S.err, panel);
                            else if (is_far) cls = NEAR ? 0 : panel;
                            else todo = relpos == 2 ? panel + PNB_CUT_FLAG : panel;      // cut pairs: regular order + flag
                        }
                    }
                    // ---- ordered binning: far pass by order, near pass in slot order ----
                    unsigned mybal = 0, mybalx = 0;
                    bool heavy = false;
                    if (!NEAR) {
                        if (todo != 0) sm.anynear = 1;
#pragma unroll
                        for (int c = 2; c <= PNB_FAR_MAX_ORDER; c++) {
                            const unsigned bc = __ballot_sync(0xffffffffu, cls == c);
                            if (lane == 0) sm.clscnt[(c - 2) * NW + warp] = __popc(bc);
                            if (cls == c) mybal = bc;
                        }
                    } else {
                        // expensive items (pairs cut by the horizon, regular orders above 8) go to the front of the list: the
                        // warps take the items in list order, so the long ones start first
                        heavy = todo >= PNB_CUT_FLAG || (todo > 8 && todo < PNB_CUT_FLAG);
                        mybal = __ballot_sync(0xffffffffu, todo != 0);
                        mybalx = __ballot_sync(0xffffffffu, heavy);
                        if (lane == 0) { sm.warpcnt[warp] = __popc(mybal); sm.clscnt[warp] = __popc(mybalx); }
                    }
                    __syncthreads();    // S2
                    if (!NEAR) {
                        const int me = (cls - 2) * NW + warp;
                        int pos = 0, tot = 0;
#pragma unroll 4
                        for (int q = 0; q < (PNB_FAR_MAX_ORDER - 1) * NW; q++) {
                            const int c = sm.clscnt[q];
                            if (q < me) pos += c;
                            tot += c;
                        }
                        if (cls != 0) sm.list[pos + __popc(mybal & ((1u << lane) - 1))] = tid | (countD ? 0x100 : 0) | (cls << 12);
                        if (tid == 0) sm.nlist = tot;
                    } else {
                        int posx = 0, posl = 0, tot = 0, totx = 0;
                        for (int w = 0; w < NW; w++) {
                            if (w < warp) { posx += sm.clscnt[w]; posl += sm.warpcnt[w] - sm.clscnt[w]; }
                            tot += sm.warpcnt[w];
                            totx += sm.clscnt[w];
                        }
                        if (todo != 0) {
                            const unsigned lt = (1u << lane) - 1;
                            const int pos = heavy ? posx + __popc(mybalx & lt) : totx + posl + __popc(mybal & ~mybalx & lt);
                            sm.list[pos] = tid | (countD ? 0x100 : 0);
                            sm.listpanel[pos] = todo;
                        }
                        if (tid == 0) { sm.nlist = tot; sm.cursor[0] = sm.cursor[1] = 0; }
                    }
                    __syncthreads();    // S3
                    const int nlist = sm.nlist;
                    if (nlist == 0) continue;     // uniform across the CTA
                    // ---- phase 2: evaluate ----
                    double X[NX];
                    int rl = 0x00FFFFFF, cl = 0x00FFFFFF;   // tile-local dofs of the pair this thread scatters
                    bool have = false;
                    if (!NEAR) {
                        if (DIM == 2 && tid < nlist) {
                            const int item = sm.list[tid];
                            const int slot = item & 0xFF, order = item >> 12;
                            const bool cD = (item & 0x100) != 0;
                            const int a1 = slot / SB, a2 = slot % SB;
                            my_pairs++;
                            double s1[3][2], s2[3][2], xx[6], yy[6];
#pragma unroll
                            for (int m = 0; m < 3; m++) {
                                s1[m][0] = sm.rb.sx[(m * DIM) % (NV * DIM)][a1];
                                s1[m][1] = sm.rb.sx[(m * DIM + 1) % (NV * DIM)][a1];
                                s2[m][0] = sm.cb.sx[(m * DIM) % (NV * DIM)][a2];
                                s2[m][1] = sm.cb.sx[(m * DIM + 1) % (NV * DIM)][a2];
                            }
                            const double sc = 2.0 * sm.rb.vol[a1] * sm.cb.vol[a2];
                            double xy[9];
                            far_eval_2d(sm.far[NEAR ? 0 : order], s1, s2, kv, cD, xy, xx, yy);
                            if (cD) {
#pragma unroll
                                for (int k = 0; k < 6; k++) {
                                    sm.dxy[slot][k % (2 * ND)] = xx[k] * sc;
                                    sm.dxy[slot][(ND + k) % (2 * ND)] = yy[k] * sc;
                                }
                                sm.slotD[slot] = 1;
                                sm.anyD = 1;
                            }
#pragma unroll
                            for (int k = 0; k < NX; k++) X[k] = xy[k % 9] * sc;
                            rl = sm.rb.loc[a1];
                            cl = sm.cb.loc[a2];
                            have = true;
                            scatter_block<DIM>(sm.acc, rl, cl, X, false);
                        }
                    } else {
                        // One warp per (pair, slice): sub-batches with few queued pairs split every pair into S slices of
                        // its quadrature nodes so that all warps stay busy; slice sums are combined in fixed order.
                        const int Sl = nlist >= 32 ? 1 : (nlist >= 16 ? 2 : (nlist >= 8 ? 4 : 8));
                        constexpr int NRr = 2 * NV - 1, NA = NRr * (NRr + 1) / 2;
                        for (int pass = 0; pass < (Sl > 1 ? 2 : 1); pass++) {
                            if (pass == 1) __syncthreads();
                            const int nitems = pass == 0 ? nlist * Sl : nlist;
                            // The items of a sub-batch differ in cost by orders of magnitude (cut pairs, high orders);
                            // dealt round robin, the warps waited at the closing barrier for the unlucky one (ncu: barrier
                            // 3.4 of the stall cycles per issue).  They are taken from a counter instead: every pair
                            // of a sub-batch updates entries of its own (the cells of a batch share no vertex), so the
                            // result does not depend on which warp takes which item.
                            for (;;) {
                                int it = 0;
                                if (lane == 0) it = atomicAdd(&sm.cursor[pass], 1);
                                it = __shfl_sync(0xffffffffu, it, 0);
                                if (it >= nitems) break;
                                const int q = pass == 0 ? it / Sl : it, sl = pass == 0 ? it - q * Sl : 0;
                                const int slot = sm.list[q] & 0xFF;
                                const bool cD = (sm.list[q] & 0x100) != 0;
                                const bool cut = FIN && sm.listpanel[q] >= PNB_CUT_FLAG;
                                const int panel = cut ? sm.listpanel[q] - PNB_CUT_FLAG : sm.listpanel[q];
                                const int Ka = sm.rb.cell[slot / SB], Kb = sm.cb.cell[slot % SB];
                                // reference orientation of singular pairs: smaller cell index first
                                const bool swapped = panel < 0 && (P.pair_orientation ? Ka < Kb : Ka > Kb);
                                const int c1 = swapped ? Kb : Ka, c2 = swapped ? Ka : Kb;
                                int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
                                int pan = panel;
                                if (panel < 0) pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.cells + (size_t)c2 * NV, NV, c1 == c2, p1, p2);
                                double acc[NL];
                                if (pass == 0) {
                                    if (lane == 0 && sl == 0) my_pairs++;
                                    if (panel >= 1) {
                                        // cut pairs are evaluated in the reference's orientation (smaller cell first): the
                                        // re-triangulation is not symmetric in its two arguments
                                        if (FIN && cut) lanes_cut_interior<DIM>(P, min(Ka, Kb), max(Ka, Kb), panel, sl * 32 + lane, 32 * Sl, acc);
                                        else lanes_regular_interior<DIM>(P, Ka, Kb, panel, sl * 32 + lane, 32 * Sl, acc);
                                        warp_allreduce<NL>(acc);
                                    } else {
                                        lanes_singular_interior<DIM>(P, c1, c2, pan, p1, p2, sl * 32 + lane, 32 * Sl, acc);
                                        warp_allreduce<NA>(acc);
                                    }
                                    if (Sl > 1) {
#pragma unroll
                                        for (int k = 0; k < NL; k++)
                                            if (k == lane) sm.partial[it][k] = acc[k];
                                        continue;
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < NL; k++) {
                                        double v = 0.;
                                        for (int ss = 0; ss < Sl; ss++) v += sm.partial[q * Sl + ss][k];
                                        acc[k] = v;
                                    }
                                }
                                if (cut && Ka > Kb) {
                                    // evaluated as (Kb, Ka): back to the local numbering of (Ka, Kb)
                                    double t[NL];
#pragma unroll
                                    for (int k = 0; k < NL; k++) t[k] = acc[k];
#pragma unroll
                                    for (int I = 0; I < 2 * NV; I++)
#pragma unroll
                                        for (int J = I; J < 2 * NV; J++) {
                                            const int i2 = (I + NV) % (2 * NV), j2 = (J + NV) % (2 * NV);
                                            acc[tri_idx(2 * NV, I, J)] = t[i2 <= j2 ? tri_idx(2 * NV, i2, j2) : tri_idx(2 * NV, j2, i2)];
                                        }
                                }
                                double myv = 0.;         // lane k < NX: entry k of the cross block
                                double myd = 0.;         // lane k < 2*ND: entry k of (dx, dy)
                                if (panel >= 1) {
                                    const double sc = 2.0 * P.vol[Ka] * P.vol[Kb];
                                    int k = 0;
#pragma unroll
                                    for (int I = 0; I < 2 * NV; I++)
#pragma unroll
                                        for (int J = I; J < 2 * NV; J++) {
                                            const double v = acc[k] * sc;
                                            if (I < NV && J >= NV) { if (lane == I * NV + (J - NV)) myv = v; }
                                            else if (J < NV) { if (lane == tri_idx(NV, I, J)) myd = v; }
                                            else { if (lane == ND + tri_idx(NV, I - NV, J - NV)) myd = v; }
                                            k++;
                                        }
                                } else {
                                    const double sc = (c1 == c2 ? 1.0 : 2.0) * (DIM == 2 ? 4.0 : 1.0) * P.vol[c1] * P.vol[c2];
                                    const int common = -pan, rows = 2 * NV - common;
                                    int k = 0;
#pragma unroll
                                    for (int I = 0; I < NRr; I++)
#pragma unroll
                                        for (int J = I; J < NRr; J++) {
                                            if (J < rows) {
                                                const double v = acc[k] * sc;
                                                int i = I < NV ? p1[I] : NV + p2[I - NV + common];
                                                int j = J < NV ? p1[J] : NV + p2[J - NV + common];
                                                if (j < i) { const int t = i; i = j; j = t; }
                                                // (i,j) in the reference's 2NV x 2NV local numbering of (c1,c2)
                                                if (i < NV && j >= NV) {
                                                    const int e = !swapped ? i * NV + (j - NV) : (j - NV) * NV + i;
                                                    if (lane == e) myv = v;
                                                } else {
                                                    const bool first = j < NV;   // block of c1
                                                    const int a = first ? i : i - NV, b = first ? j : j - NV;
                                                    const bool to_dx = first != swapped;
                                                    if (lane == (to_dx ? 0 : ND) + tri_idx(NV, a, b)) myd = v;
                                                }
                                            }
                                            k++;
                                        }
                                }
                                // lanes 0..NX-1 add the cross block (distinct entries), lanes 0..2ND-1 store the diagonal blocks
                                if (lane < NX) {
                                    const int i = lane / NV, j = lane - i * NV;
                                    const int a = (sm.rb.loc[slot / SB] >> (8 * i)) & 0xFF, b = (sm.cb.loc[slot % SB] >> (8 * j)) & 0xFF;
                                    if (a != 0xFF && b != 0xFF) sm.acc[a][b] += myv;
                                    sm.nv[q][lane] = myv;
                                }
                                if (cD && lane < 2 * ND) sm.dxy[slot][lane] = myd;
                                if (cD && lane == 0) { sm.slotD[slot] = 1; sm.anyD = 1; }
                            }
                        }
                    }
                    __syncthreads();    // S4: direct updates done
                    if (diag) {
                        // mirror image inside a diagonal tile: second conflict-free round
                        if (!NEAR) {
                            if (have) scatter_block<DIM>(sm.acc, rl, cl, X, true);
                        } else {
                            for (int q = tid / NX; q < nlist; q += PNB_THREADS / NX) {
                                const int e = tid % NX;
                                if (tid / NX >= PNB_THREADS / NX) break;
                                const int slot = sm.list[q] & 0xFF;
                                const int i = e / NV, j = e - i * NV;
                                const int a = (sm.rb.loc[slot / SB] >> (8 * i)) & 0xFF, b = (sm.cb.loc[slot % SB] >> (8 * j)) & 0xFF;
                                if (a != 0xFF && b != 0xFF) sm.acc[b][a] += sm.nv[q][e];
                            }
                        }
                    }
                    // ---- cell-diagonal blocks: reduce over the sub-batch into the per-tile accumulators ----
                    if (sm.anyD) {
                        if (tid < SB * ND) {
                            const int kk1 = tid / ND, comp = tid - kk1 * ND;
                            if (sm.rb.cell[kk1] >= 0 && sm.rb.home[kk1] == rt) {
                                double sacc = 0.;
                                for (int kk2 = 0; kk2 < SB; kk2++)
                                    if (sm.slotD[kk1 * SB + kk2]) sacc += sm.dxy[kk1 * SB + kk2][comp];
                                DXs[(rb + kk1) * ND + comp] += sacc;
                            }
                        } else if (tid < 2 * SB * ND) {
                            const int t2 = tid - SB * ND;
                            const int kk2 = t2 / ND, comp = t2 - kk2 * ND;
                            if (sm.cb.cell[kk2] >= 0 && sm.cb.home[kk2] == ct) {
                                double sacc = 0.;
                                for (int kk1 = 0; kk1 < SB; kk1++)
                                    if (sm.slotD[kk1 * SB + kk2]) sacc += sm.dxy[kk1 * SB + kk2][ND + comp];
                                DYs[(cb + kk2) * ND + comp] += sacc;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- write the tile (and its mirror image); flush the cell-diagonal partial sums ----
            const int r0 = rt * TD, c0 = ct * TD;
            for (int e = tid; e < TD * TD; e += PNB_THREADS) {
                const int a = e / TD, b = e - a * TD;
                if (own_r && r0 + a < P.N && c0 + b < P.N) {
                    double *dst = entry_ptr(S, A, ld, rt, r0 + a, c0 + b);
                    // diagonal tiles: acc[a][b] and acc[b][a] hold the same terms summed in different orders;
                    // their mean is bitwise symmetric (and deterministic)
                    const double v = diag ? 0.5 * (sm.acc[a][b] + sm.acc[b][a]) : sm.acc[a][b];
                    *dst = NEAR ? *dst + v : v;
                }
            }
            if (!diag && own_c)
                for (int e = tid; e < TD * TD; e += PNB_THREADS) {
                    const int b = e / TD, a = e - b * TD;
                    if (r0 + a < P.N && c0 + b < P.N) {
                        double *dst = entry_ptr(S, A, ld, rt, c0 + b, r0 + a);
                        *dst = NEAR ? *dst + sm.acc[a][b] : sm.acc[a][b];
                    }
                }
            for (int e = tid; e < nR * ND; e += PNB_THREADS) {
                const int K = S.tile_cells[rbeg + e / ND];
                if (K >= 0 && DXs[e] != 0.) S.DXp[((size_t)(gc - goff) * P.nc + K) * ND + (e % ND)] += DXs[e];
            }
            for (int e = tid; e < nC * ND; e += PNB_THREADS) {
                const int K = S.tile_cells[cbeg + e / ND];
                if (K >= 0 && DYs[e] != 0.) S.DYp[((size_t)(gr - goff) * P.nc + K) * ND + (e % ND)] += DYs[e];
            }
            if (!NEAR && sm.anynear) {
                if (tid == 0) S.tileflag[tfi] = 1;
                unit_near = true;
            }
        }
    if (!NEAR && unit_near && tid == 0) S.unitflag[unit] = 1;
    // pair counter (statistics only; integer atomics)
    for (int off = 16; off > 0; off >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, off);
    if (lane == 0 && my_pairs) atomicAdd(S.counters + (NEAR ? 1 : 0), my_pairs);
}

// ordered compaction of the flagged units (single block; nunits is small)
__global__ void compact_units_kernel(TileSched S)
{
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int u0 = 0; u0 < S.nunits; u0 += blockDim.x) {
        const int u = u0 + threadIdx.x;
        const bool f = u < S.nunits && S.unitflag[u];
        // block-wide ordered scan via ballots
        __shared__ int wc[32];
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int pos = base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1));
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) pos += wc[w];
        if (f) S.nearunits[pos] = u;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += wc[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) S.nearunits[S.nunits] = base;
}

// ---------------------------------------------------------------------------
// zero-exterior facet loop (nonlocalAssembly_{SCALAR}.pxi:1430-1448): one warp per
// cell, lanes over the boundary facets, each lane integrates whole pairs and
// keeps a private sum; fixed butterfly at the end.
// ---------------------------------------------------------------------------
// getQuadOrder of the boundary class (fractionalLaplacian2D.pyx:1226-1253) in single precision; -1 when the value
// handed to ceil() is within the error bound of an integer (the caller then repeats it in double precision)
__device__ __forceinline__ int fast_order_boundary_2d(double d2, float lh1, float lh2, float ah1, float ah2, float cb, float sf)
{
    // error of the FP32 evaluation: ~1e-6 absolute in the logarithms and 6e-8 relative in the divisions, numerators up to
    // ~20 over denominators >= 0.4 -> below 1e-4; values closer than MARGIN to an integer are re-decided in FP64
    const float MARGIN = 4e-4f;
    const float Ld = 0.5f * __logf((float)d2);
    const float l1 = fmaxf(Ld - lh1, 0.f), l2 = fmaxf(Ld - lh2, 0.f);
    const float m = fmaxf(ah1, ah2);
    const float num1 = cb + m + (sf - 1.f) * ah2 - sf * l2;
    const float num2 = cb + m + (sf - 1.f) * ah1 - sf * l1;
    const float g = fmaxf(__fdividef(num1, l1 + 0.35f), __fdividef(num2, l2 + 0.35f));
    if (g <= 2.f - MARGIN) return 2;
    const float k = ceilf(g);
    if (k - g > MARGIN && g - (k - 1.f) > MARGIN && g < 250.f) return (int)k;
    return -1;
}

#define PNB_BND_LANE_ORDER 5

template <int DIM>
__global__ void boundary_kernel(DProblem P, TileSched S)
{
    constexpr int NV = PairDims<DIM>::NV, ND = PairDims<DIM>::ND;
    const int lane = threadIdx.x & 31;
    const int c1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c1 >= P.nc) return;
    bool any = false;
    for (int m = 0; m < NV; m++) any |= P.dofs[(size_t)c1 * NV + m] >= 0;
    if (!any) return;
    if (S.cell_mask ? !S.cell_mask[c1] : (S.home[c1] < S.own_t0 || S.home[c1] >= S.own_t1)) return;
    double tot[ND];
#pragma unroll
    for (int k = 0; k < ND; k++) tot[k] = 0.;
    // batched blocks: the surface of the cell's own block
    const int fbeg = S.cell_block ? S.blk_fptr[S.cell_block[c1]] : 0, fend = S.cell_block ? S.blk_fptr[S.cell_block[c1] + 1] : P.nb;
    // regular facets: lane-per-facet
    for (int f0 = fbeg; f0 < fend; f0 += 32) {
        const int f = f0 + lane;
        int pan = 0;
        int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
        const bool mine = f < fend && (!P.labels || pnb_bclass_active(P, P.labels[c1], P.blabels[f]));
        if (mine) {
            if (DIM == 2) {
                pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.bfacets + (size_t)f * 2, 2, false, p1, p2);
                if (pan == 0) {
                    const double a = P.centers[(size_t)c1 * 2] - P.bcenters[(size_t)f * 2], b = P.centers[(size_t)c1 * 2 + 1] - P.bcenters[(size_t)f * 2 + 1];
                    pan = fast_order_boundary_2d(a * a + b * b, P.lhcf[c1], P.lhbf[f], P.ahcf[c1], P.ahbf[f], (float)P.c_bnd,
                                                 (float)fmax(0.5 * (-P.bsing - 1.), 0.));
                    if (pan < 0) pan = panel_boundary(P, c1, f, p1, p2);
                }
            } else pan = panel_boundary(P, c1, f, p1, p2);
        }
        // regular facets of low order: one per lane; the few facets close to the cell (orders above PNB_BND_LANE_ORDER,
        // up to ~1500 node pairs) would keep one lane busy while the others idle: they go to the whole warp below
        if (mine && pan > P.max_order) { atomicMax(S.err, pan); pan = 0; }
        if (mine && pan >= 1 && pan <= PNB_BND_LANE_ORDER) {
            double acc[ND];
            lanes_boundary<DIM>(P, c1, f, pan, p1, p2, 0, 1, acc);
            const double sc = P.vol[c1] * P.bvol[f];
#pragma unroll
            for (int k = 0; k < ND; k++) tot[k] += acc[k] * sc;
        }
        // singular facets and regular facets of high order of this chunk: whole warp per facet, in facet order
        unsigned sing = __ballot_sync(0xffffffffu, mine && (pan < 0 || pan > PNB_BND_LANE_ORDER));
        while (sing) {
            const int src = __ffs(sing) - 1;
            sing &= sing - 1;
            const int fs = f0 + src;
            const int pans = __shfl_sync(0xffffffffu, pan, src);
            int q1[3], q2[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                q1[k] = __shfl_sync(0xffffffffu, p1[k], src);
                q2[k] = __shfl_sync(0xffffffffu, p2[k], src);
            }
            double acc[ND];
            lanes_boundary<DIM>(P, c1, fs, pans, q1, q2, lane, 32, acc);
            const double sc = pans >= 1 ? P.vol[c1] * P.bvol[fs] : (DIM == 2 ? -2.0 * P.vol[c1] * P.bvol[fs] : P.vol[c1]);
            // entry (I,J) over permuted dofs -> local (perm1[I], perm1[J]) (regular facets: identity)
            int k = 0;
#pragma unroll
            for (int I = 0; I < NV; I++)
#pragma unroll
                for (int J = I; J < NV; J++) {
                    const int i = q1[I], j = q1[J];
                    const int kk = j < i ? tri_idx(NV, j, i) : tri_idx(NV, i, j);
#pragma unroll
                    for (int t = 0; t < ND; t++)
                        if (t == kk) tot[t] += acc[k] * sc;
                    k++;
                }
        }
    }
    warp_allreduce<ND>(tot);
#pragma unroll
    for (int k = 0; k < ND; k++)
        if (k == lane) S.Dbnd[(size_t)c1 * ND + k] = tot[k];
}

// D[K] = sum_g DXp[g][K] + sum_g DYp[g][K] + Dbnd[K], fixed order
__global__ void reduce_D_kernel(TileSched S, int nc, int ND, int use_bnd)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)nc * ND) return;
    // only cells whose home tile is owned have seen all their partners; the others are completed by the
    // owner of their home tile (exchanged by the caller between pnb_dense_rows_begin and _end)
    const int hm = S.home[e / ND];
    if (hm < S.own_t0 || hm >= S.own_t1) { S.D[e] = 0.; return; }
    double s = 0.;
    for (int g = 0; g < S.dgroups; g++) s += S.DXp[(size_t)g * nc * ND + e];
    for (int g = 0; g < S.dgroups; g++) s += S.DYp[(size_t)g * nc * ND + e];
    if (use_bnd) s += S.Dbnd[e];
    S.D[e] = s;
}

// row-owned scatter of the cell-diagonal blocks (addToMatrixElemSym semantics,
// nonlocalAssembly_{SCALAR}.pxi:152-168, 204-221): thread I adds to row I only
__global__ void scatter_D_kernel(DProblem P, TileSched S, double *A, int64_t ld)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x + S.own_t0 * PNB_TD;
    if (I >= P.N || I >= S.own_t1 * PNB_TD) return;
    double *Ablk = A;
    A -= (size_t)S.own_t0 * PNB_TD * ld;
    const int NV = P.dim + 1, ND = NV * (NV + 1) / 2;
    for (int t = S.dof_ptr[I]; t < S.dof_ptr[I + 1]; t++) {
        const int K = S.dof_cells[t] >> 2, p = S.dof_cells[t] & 3;
        for (int q = 0; q < NV; q++) {
            const int J = P.dofs[(size_t)K * NV + q];
            if (J < 0) continue;
            const int kk = p <= q ? tri_idx(NV, p, q) : tri_idx(NV, q, p);
            if (S.tile_block) *entry_ptr(S, Ablk, ld, I / PNB_TD, I, J) += S.D[(size_t)K * ND + kk];
            else A[(size_t)I * ld + J] += S.D[(size_t)K * ND + kk];
        }
    }
}


#include "pnb_group.cuh"

// ---------------------------------------------------------------------------
// cell-group schedule (host)
// ---------------------------------------------------------------------------
struct GroupHost {
    bool ready = false;
    int GC = 0;
    std::vector<GUnit> f2_units, mix_units, near_units;   // f2 / mix sorted by phase
    const GUnit *d_f2 = nullptr, *d_mix = nullptr, *d_near = nullptr;
    size_t smem_f2 = 0, smem_mix = 0, smem_near = 0;
    int nphase = 0;
};

static uint64_t hilbert_index(uint32_t x, uint32_t y)
{
    uint64_t d = 0;
    for (uint32_t s = 32768; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        d += (uint64_t)s * s * ((3 * rx) ^ ry);
        if (!ry) {
            if (rx) { x = 65535 - x; y = 65535 - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}

// upper bound of the real-valued quadrature order (the argument of ceil in getQuadOrder,
// fractionalLaplacian2D.pyx:622-642) over all pairs with d >= dmin, h1 in group 1, h2 in group 2
static double order_upper_bound(const DProblem &P, double dmin, double h1max, double h2max, double a1min, double a1max,
                                double a2min, double a2max)
{
    if (!(dmin > 0.)) return 1e30;
    const double s = fmax(-0.5 * (P.sing + 2), 0.);
    const double amax = fmax(a1max, a2max);
    auto f = [&](double hthis_max, double ath_min, double ath_max, double hother_max) {
        // num = c + (s-1) |log(h_this/H0)| + max(|log h1/H0|, |log h2/H0|) - s log(d/h_this);  den = max(log(d/h_other),0) + 0.4
        const double num = P.c_int + (s - 1. < 0. ? (s - 1.) * ath_min : (s - 1.) * ath_max) + amax - s * log(dmin / hthis_max);
        const double den = fmax(log(dmin / hother_max), 0.) + 0.4;
        return num <= 0. ? 0. : num / den;
    };
    return fmax(f(h2max, a2min, a2max, h1max), f(h1max, a1min, a1max, h2max));
}


// ---------------------------------------------------------------------------
// cell-group path: schedule construction and launch sequence
// ---------------------------------------------------------------------------
struct GroupGeom {
    std::vector<int> gptr, gcells, gloc, gdptr, gdofs, color;
    std::vector<std::vector<int>> adj;     // groups sharing a vertex (sorted, includes self)
    std::vector<double> box;               // per group: xmin, xmax, ymin, ymax, hmax, amin, amax
    std::vector<int> labelrange;           // per group: smallest and largest cell label (piecewise variable kernels)
    int ngroups = 0, cap = 0, maxld = 0, ncolors = 0;
};

static void build_group_geometry(const pnb_problem *p, int GC, const std::vector<int> &order, GroupGeom &gg)
{
    const int nc = p->nc;
    const int *cells = p->h_cells.data(), *dofs = p->h_dofs.data();
    gg = GroupGeom();
    gg.ngroups = std::max(1, (nc + GC - 1) / GC);
    gg.gptr.assign(1, 0);
    gg.gdptr.assign(1, 0);
    std::vector<int> grp(nc, 0);
    struct PerGroup { std::vector<int> gcells, gloc, ld; };
    std::vector<PerGroup> pg(gg.ngroups);
    auto do_group = [&](int g) {
        std::vector<std::vector<int>> bcells;
        PerGroup &out = pg[g];
        const int c0 = g * GC, c1 = std::min(nc, c0 + GC);
        // local dofs
        std::vector<int> ld;
        for (int k = c0; k < c1; k++)
            for (int m = 0; m < 3; m++) { const int d = dofs[(size_t)order[k] * 3 + m]; if (d >= 0) ld.push_back(d); }
        std::sort(ld.begin(), ld.end());
        ld.erase(std::unique(ld.begin(), ld.end()), ld.end());
        // batches of <= PNB_SB cells sharing no vertex.  Most-constrained-first (DSATUR-like) assignment to the
        // emptiest admissible batch: reaches the minimum number of batches (full batches, no padding) on
        // triangulations, where first-fit leaves ~15-25 % of the slots empty.
        const int B0 = std::max(1, (c1 - c0 + PNB_SB - 1) / PNB_SB);
        bcells.assign(B0, std::vector<int>());
        {
            const int n = c1 - c0;
            std::vector<int> lv((size_t)n * 3);           // group-local vertex ids
            std::vector<int> vids;
            for (int k = 0; k < n; k++)
                for (int m = 0; m < 3; m++) vids.push_back(cells[(size_t)order[c0 + k] * 3 + m]);
            std::vector<int> uv(vids);
            std::sort(uv.begin(), uv.end());
            uv.erase(std::unique(uv.begin(), uv.end()), uv.end());
            for (size_t e = 0; e < vids.size(); e++) lv[e] = (int)(std::lower_bound(uv.begin(), uv.end(), vids[e]) - uv.begin());
            // the greedy choice occasionally needs one batch more than the minimum; a few restarts with a rotated scan
            // order (deterministic) almost always find a minimal colouring, which keeps the slot count of the groups
            // (shared memory per warp of the unit kernels) at its minimum
            std::vector<std::vector<int>> bestb;
            // cells around every group-local vertex: the batches forbidden to a cell are kept per cell and updated when a
            // neighbour is assigned (the selection scans one word per cell)
            std::vector<int> vptr(uv.size() + 1, 0), vlist((size_t)n * 3);
            for (int e = 0; e < n * 3; e++) vptr[lv[e] + 1]++;
            for (size_t v = 0; v < uv.size(); v++) vptr[v + 1] += vptr[v];
            {
                std::vector<int> pos(vptr.begin(), vptr.end() - 1);
                for (int e = 0; e < n * 3; e++) vlist[pos[lv[e]]++] = e / 3;
            }
            std::vector<unsigned long long> forb(n);
            for (int attempt = 0; attempt < 12; attempt++) {
            bcells.assign(B0, std::vector<int>());
            const int rot = (int)(((long long)attempt * 37) % std::max(n, 1));
            std::fill(forb.begin(), forb.end(), 0ull);
            std::vector<int> cnt(B0, 0);
            std::vector<char> done(n, 0);
            for (int it = 0; it < n; it++) {
                unsigned long long full = 0;
                for (int b = 0; b < (int)cnt.size(); b++) if (cnt[b] >= PNB_SB) full |= 1ull << b;
                const unsigned long long all = cnt.size() >= 64 ? ~0ull : ((1ull << cnt.size()) - 1);
                int best = -1, bestf = 1 << 30;
                unsigned long long bestmask = 0;
                for (int k0 = 0; k0 < n; k0++) {
                    const int k = k0 + rot < n ? k0 + rot : k0 + rot - n;
                    if (done[k]) continue;
                    const unsigned long long feas = ~(forb[k] | full) & all;
                    const int nf = __builtin_popcountll(feas);
                    if (nf < bestf) { bestf = nf; best = k; bestmask = feas; if (nf == 0) break; }
                }
                int b = -1;
                if (bestmask == 0) {
                    if (cnt.size() >= 63) b = -2;
                    else { cnt.push_back(0); bcells.emplace_back(); b = (int)cnt.size() - 1; }
                } else {
                    for (int q = 0; q < (int)cnt.size(); q++)
                        if (((bestmask >> q) & 1) && (b < 0 || cnt[q] < cnt[b])) b = q;
                }
                if (b == -2) {      // pathological mesh: one cell per further batch
                    bcells.emplace_back();
                    bcells.back().push_back(order[c0 + best]);
                    done[best] = 1;
                    continue;
                }
                cnt[b]++;
                bcells[b].push_back(order[c0 + best]);
                for (int m = 0; m < 3; m++) {
                    const int v = lv[best * 3 + m];
                    for (int e = vptr[v]; e < vptr[v + 1]; e++) forb[vlist[e]] |= 1ull << b;
                }
                done[best] = 1;
            }
            if (bestb.empty() || bcells.size() < bestb.size()) bestb = bcells;
            if ((int)bestb.size() <= B0) break;
            }
            bcells = bestb;
        }
        for (int k = c0; k < c1; k++) grp[order[k]] = g;
        // fullest batches first
        std::stable_sort(bcells.begin(), bcells.end(), [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
        for (auto &bc : bcells) {
            if (bc.empty()) continue;
            for (int k = 0; k < PNB_SB; k++) {
                const int c = k < (int)bc.size() ? bc[k] : -1;
                out.gcells.push_back(c);
                int packed = 0x00FFFFFF;
                if (c >= 0) {
                    packed = 0;
                    for (int m = 0; m < 3; m++) {
                        const int d = dofs[(size_t)c * 3 + m];
                        int l = 0xFF;
                        if (d >= 0) l = (int)(std::lower_bound(ld.begin(), ld.end(), d) - ld.begin());
                        packed |= l << (8 * m);
                    }
                }
                out.gloc.push_back(packed);
            }
        }
        out.ld.swap(ld);
    };
    {
        // groups are independent: a few host threads (the colouring is quadratic in the group size)
        const int nt = std::max(1, std::min(16, std::min((int)std::thread::hardware_concurrency(), gg.ngroups / 8)));
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t]() { for (int g = t; g < gg.ngroups; g += nt) do_group(g); });
        for (auto &x : th) x.join();
    }
    for (int g = 0; g < gg.ngroups; g++) {
        const PerGroup &o = pg[g];
        gg.maxld = std::max(gg.maxld, (int)o.ld.size());
        gg.gcells.insert(gg.gcells.end(), o.gcells.begin(), o.gcells.end());
        gg.gloc.insert(gg.gloc.end(), o.gloc.begin(), o.gloc.end());
        gg.gptr.push_back((int)gg.gcells.size());
        gg.cap = std::max(gg.cap, gg.gptr[g + 1] - gg.gptr[g]);
        gg.gdofs.insert(gg.gdofs.end(), o.ld.begin(), o.ld.end());
        gg.gdptr.push_back((int)gg.gdofs.size());
    }
    // adjacency through shared vertices
    gg.adj.assign(gg.ngroups, std::vector<int>());
    {
        // groups around every vertex (CSR), then all pairs of groups that meet at a vertex
        const int nv = p->P.nv;
        std::vector<int> vptr(nv + 1, 0), vgrp((size_t)nc * 3);
        for (int c = 0; c < nc; c++)
            for (int m = 0; m < 3; m++) vptr[cells[(size_t)c * 3 + m] + 1]++;
        for (int v = 0; v < nv; v++) vptr[v + 1] += vptr[v];
        {
            std::vector<int> pos(vptr.begin(), vptr.end() - 1);
            for (int c = 0; c < nc; c++)
                for (int m = 0; m < 3; m++) vgrp[pos[cells[(size_t)c * 3 + m]]++] = grp[c];
        }
        for (int v = 0; v < nv; v++) {
            int *b = &vgrp[vptr[v]], *e = &vgrp[vptr[v + 1]];
            if (e - b < 2) continue;
            std::sort(b, e);
            e = std::unique(b, e);
            if (e - b < 2) continue;
            for (int *a = b; a < e; a++)
                for (int *q = b; q < e; q++) gg.adj[*a].push_back(*q);
        }
        for (int g = 0; g < gg.ngroups; g++) {
            auto &l = gg.adj[g];
            l.push_back(g);
            std::sort(l.begin(), l.end());
            l.erase(std::unique(l.begin(), l.end()), l.end());
        }
    }
    // greedy colouring
    gg.color.assign(gg.ngroups, -1);
    for (int g = 0; g < gg.ngroups; g++) {
        unsigned long long used = 0;
        for (int a : gg.adj[g]) if (a != g && gg.color[a] >= 0) used |= 1ull << gg.color[a];
        int c = 0;
        while ((used >> c) & 1) c++;
        gg.color[g] = c;
        gg.ncolors = std::max(gg.ncolors, c + 1);
    }
    gg.labelrange.assign((size_t)gg.ngroups * 2, 0);
    if (!p->h_labels.empty())
        for (int g = 0; g < gg.ngroups; g++) {
            int lo = 255, hi = 0;
            for (int k = g * GC; k < std::min(nc, (g + 1) * GC); k++) {
                lo = std::min(lo, (int)p->h_labels[order[k]]);
                hi = std::max(hi, (int)p->h_labels[order[k]]);
            }
            gg.labelrange[(size_t)g * 2] = lo;
            gg.labelrange[(size_t)g * 2 + 1] = hi;
        }
    // boxes of the cell centers, extreme mesh sizes
    gg.box.assign((size_t)gg.ngroups * 7, 0.);
    for (int g = 0; g < gg.ngroups; g++) {
        double *b = &gg.box[(size_t)g * 7];
        b[0] = b[2] = 1e300; b[1] = b[3] = -1e300; b[4] = 0.; b[5] = 1e300; b[6] = 0.;
        for (int k = g * GC; k < std::min(nc, (g + 1) * GC); k++) {
            const int c = order[k];
            const double x = p->h_centers[(size_t)c * 2], y = p->h_centers[(size_t)c * 2 + 1], h = p->h_h[c];
            const double a = fabs(log(h / p->P.H0));
            b[0] = std::min(b[0], x); b[1] = std::max(b[1], x);
            b[2] = std::min(b[2], y); b[3] = std::max(b[3], y);
            b[4] = std::max(b[4], h); b[5] = std::min(b[5], a); b[6] = std::max(b[6], a);
        }
    }
}

struct GroupHostFull : GroupHost {
    GroupGeom gg;
    // several GPUs (pnb_dist_plan): part of every group, owner / position tables of the group-local dofs, rows of this part
    std::vector<int> gr_part, rows, gcnt;
    std::vector<unsigned char> gown, gpos;
    std::vector<signed char> kinds;         // ngroups x ngroups (I <= J): unit kind, -1 = no pair of this problem instance
    std::vector<int> ticket;                // ngroups x ngroups: list position | list << 30 of the units of this instance
    bool dist = false;
    std::vector<long long> uoff_all;        // ngroups x ngroups: start of the unit's fragments in THIS part's staging, -1 = none
    long long stage_doubles = 0;
    bool dist_ready = false;
    std::vector<void *> dist_allocs;
    DistApply apply{};
    const unsigned char *d_cell_mask = nullptr;
    int far_mask = -1, max_order = -1, part = -1, nparts = -1;
    std::vector<void *> unit_allocs, near_allocs;
    const int2 *d_items = nullptr;
    const int *d_perm = nullptr;
    const int4 *d_chunks = nullptr;
    double *d_R = nullptr, *d_F = nullptr;
    int nitems = 0, npairs = 0, nchunks = 0;
    int near_nmax = 1;                      // largest node count of a regular rule with near items
    int mix_warps = PNB_MW_MAX;             // warps per CTA of gmix_kernel (what fits into shared memory)
    // row panels: the unit lists are ordered by the first panel of rows (by dof index) that a unit touches, so that the rows
    // of panel k are final once the units of the panels <= k are done (their copy to the host overlaps the later panels)
    int npanels = 1, panel_tiles = 0;       // panel = panel_tiles row tiles of 32 rows
    std::vector<int> f2_pbeg, mix_pbeg;     // npanels + 1: list positions
    void *mix_scratch = nullptr;            // unit blocks of the resident CTAs of gmix_kernel
    size_t mix_scratch_bytes = 0;
    bool near_ready = false;
};

static int build_group_schedule(pnb_problem *p)
{
    const int nc = p->nc;
    const double tw0 = wall_ms();
    if (!p->gh) p->gh = new GroupHostFull();
    if (!p->G) p->G = new GroupSched();
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    GroupSched &G = *p->G;
    int smem_sm = 0;
    smem_sm = device_attr(cudaDevAttrMaxSharedMemoryPerMultiprocessor, p->device);
    if (smem_sm <= 0) smem_sm = 228 * 1024;
    int smem_blk = 0;
    smem_blk = device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device);
    const size_t budget = smem_blk > 0 ? (size_t)smem_blk : (size_t)smem_sm - 1024;     // one 512-thread CTA per SM
    if (!gh->ready) {
        // Hilbert order of the cell centers
        double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
        for (int c = 0; c < nc; c++)
            for (int l = 0; l < 2; l++) { lo[l] = std::min(lo[l], p->h_centers[(size_t)c * 2 + l]); hi[l] = std::max(hi[l], p->h_centers[(size_t)c * 2 + l]); }
        std::vector<uint64_t> key(nc);
        for (int c = 0; c < nc; c++) {
            uint32_t q[2];
            for (int l = 0; l < 2; l++) {
                const double t = hi[l] > lo[l] ? (p->h_centers[(size_t)c * 2 + l] - lo[l]) / (hi[l] - lo[l]) : 0.;
                q[l] = (uint32_t)std::min(65535., std::max(0., t * 65535.));
            }
            key[c] = hilbert_index(q[0], q[1]);
        }
        // stable order by key = order of the packed (key, cell) words (keys are below 2^32)
        std::vector<int> order(nc);
        {
            std::vector<uint64_t> packed(nc);
            for (int c = 0; c < nc; c++) packed[c] = (key[c] << 32) | (uint32_t)c;
            std::sort(packed.begin(), packed.end());
            for (int c = 0; c < nc; c++) order[c] = (int)(packed[c] & 0xFFFFFFFFu);
        }
        const double tg0 = wall_ms();
        // PNB_GROUP_CELLS: cells per group (experiments; a multiple of 32)
        const int forced = getenv("PNB_GROUP_CELLS") ? atoi(getenv("PNB_GROUP_CELLS")) : 0;
        const int cand[] = {128, 96, 64, 32};   // even numbers of full batches (two column batches per step)
        for (int GC : cand) {
            if (forced > 0) GC = forced;
            build_group_geometry(p, GC, order, gh->gg);
            gh->GC = GC;
            const int ldS = gh->gg.maxld + 1;
            if (forced > 0 || (gf2_smem_bytes(gh->gg.cap, gh->gg.maxld, ldS) <= budget && gmix_warps(gh->gg.cap, gh->gg.maxld, budget) >= 4 &&
                               gh->gg.maxld < 255))
                break;
        }
        const GroupGeom &gg = gh->gg;
        const double tg1 = wall_ms();
        if (gg.maxld >= 255) return fail(PNB_ERR_UNSUPPORTED, "cell group with more than 254 local dofs");
        G.ngroups = gg.ngroups; G.cap = gg.cap; G.maxld = gg.maxld; G.ldS = gg.maxld + 1; G.ncolors = gg.ncolors;
        G.nbmax = gg.cap / PNB_SB;
        int rc = 0;
        rc |= upload(p, gg.gptr.data(), gg.gptr.size(), &G.gptr);
        rc |= upload(p, gg.gcells.data(), gg.gcells.size(), &G.gcells);
        rc |= upload(p, gg.gloc.data(), gg.gloc.size(), &G.gloc);
        rc |= upload(p, gg.gdptr.data(), gg.gdptr.size(), &G.gdptr);
        rc |= upload(p, gg.gdofs.data(), gg.gdofs.size(), &G.gdofs);
        rc |= dalloc(p, (size_t)gg.ngroups * nc * 6, &G.Dp);
        {
            std::vector<int> aptr(1, 0), alist;
            for (int g = 0; g < gg.ngroups; g++) {
                alist.insert(alist.end(), gg.adj[g].begin(), gg.adj[g].end());
                aptr.push_back((int)alist.size());
            }
            rc |= upload(p, aptr.data(), aptr.size(), &G.adjptr);
            rc |= upload(p, alist.data(), alist.size(), &G.adj);
            rc |= dalloc(p, 2 * PNB_ROW_PANELS + 2, &G.counters_i);
        }
        if (rc) return PNB_ERR_CUDA;
        const double tg2 = wall_ms();
        G.err = p->S.err;
        G.counters = p->S.counters;
        gh->smem_f2 = gf2_smem_bytes(G.cap, G.maxld, G.ldS);
        gh->mix_warps = std::max(1, gmix_warps(G.cap, G.maxld, budget));
        gh->smem_mix = gmix_smem_bytes(G.cap, G.maxld, gh->mix_warps);
        {
            // incidence lists of the group-local dofs: (cell slot, local vertex) in ascending order (counting sort per group)
            std::vector<int> iptr(1, 0);
            std::vector<unsigned short> ilist;
            ilist.reserve(gg.gcells.size() * 3);
            std::vector<int> cnt;
            for (int g = 0; g < gg.ngroups; g++) {
                const int nld = gg.gdptr[g + 1] - gg.gdptr[g], ns = gg.gptr[g + 1] - gg.gptr[g];
                cnt.assign(nld + 1, 0);
                for (int sl = 0; sl < ns; sl++) {
                    if (gg.gcells[gg.gptr[g] + sl] < 0) continue;
                    const int packed = gg.gloc[gg.gptr[g] + sl];
                    for (int m = 0; m < 3; m++) {
                        const int l = (packed >> (8 * m)) & 0xFF;
                        if (l != 0xFF) cnt[l + 1]++;
                    }
                }
                for (int l = 0; l < nld; l++) cnt[l + 1] += cnt[l];
                const size_t base = ilist.size();
                ilist.resize(base + cnt[nld]);
                for (int l = 0; l < nld; l++) iptr.push_back((int)(base + cnt[l + 1]));
                for (int sl = 0; sl < ns; sl++) {
                    if (gg.gcells[gg.gptr[g] + sl] < 0) continue;
                    const int packed = gg.gloc[gg.gptr[g] + sl];
                    for (int m = 0; m < 3; m++) {
                        const int l = (packed >> (8 * m)) & 0xFF;
                        if (l != 0xFF) ilist[base + cnt[l]++] = (unsigned short)(sl * 4 + m);
                    }
                }
            }
            if (ilist.empty()) ilist.push_back(0);
            if (upload(p, iptr.data(), iptr.size(), &G.gincptr) || upload(p, ilist.data(), ilist.size(), &G.ginc)) return PNB_ERR_CUDA;
        }
        gh->smem_near = gnear_list_smem_bytes(G.cap);
        gh->ready = true;
        if (getenv("PNB_BENCH_VERBOSE"))
            fprintf(stderr, "group path: GC %d, %d groups, cap %d, maxld %d, %d colours, smem f2/mix/near %zu/%zu/%zu; host ms: hilbert %.1f, "
                    "geometry %.1f, uploads %.1f, incidences %.1f\n", gh->GC,
                    gg.ngroups, gg.cap, gg.maxld, gg.ncolors, gh->smem_f2, gh->smem_mix, gh->smem_near, tg0 - tw0, tg1 - tg0, tg2 - tg1,
                    wall_ms() - tg2);
    }
    if (gh->far_mask == p->far_mask && gh->max_order == p->P.max_order && gh->part == p->part && gh->nparts == p->nparts &&
        gh->dist == p->dist)
        return 0;
    gh->dist_ready = false;
    const double tw1 = wall_ms();
    // ---- unit kinds (depend on the tables) ----
    const GroupGeom &gg = gh->gg;
    // far_top: all orders 2..far_top are taken by the thread-per-pair evaluator
    int far_top = 1;
    while (far_top < PNB_FAR_MAX_ORDER && far_top < p->P.max_order && ((p->far_mask >> (far_top + 1)) & 1)) far_top++;
    const bool far_full = far_top >= 3;
    const bool far_2 = far_top >= 2;
    const int ncol = gg.ncolors;
    gh->nphase = ncol * ncol;
    // first row panel touched by the dofs of every group
    const int NP = p->dist ? 1 : PNB_ROW_PANELS;
    const int ntile32 = std::max(1, (p->N + 31) / 32);
    gh->panel_tiles = (ntile32 + NP - 1) / NP;
    gh->npanels = (ntile32 + gh->panel_tiles - 1) / gh->panel_tiles;
    std::vector<int> gpanel(gg.ngroups, 0);
    for (int g = 0; g < gg.ngroups; g++) {
        int mn = p->N;
        for (int k = gg.gdptr[g]; k < gg.gdptr[g + 1]; k++) mn = std::min(mn, gg.gdofs[k]);
        gpanel[g] = std::min(gh->npanels - 1, (mn / 32) / gh->panel_tiles);
    }
    std::vector<std::vector<GUnit>> f2((size_t)gh->nphase * gh->npanels), mix((size_t)gh->nphase * gh->npanels);
    gh->near_units.clear();
    int nslots = 0;
    // kind of every unit (depends on the tables)
    gh->kinds.assign((size_t)gg.ngroups * gg.ngroups, -1);
    for (int I = 0; I < gg.ngroups; I++) {
        const double *b1 = &gg.box[(size_t)I * 7];
        for (int J = I; J < gg.ngroups; J++) {
            const double *b2 = &gg.box[(size_t)J * 7];
            int kind = 2;
            if (!std::binary_search(gg.adj[I].begin(), gg.adj[I].end(), J)) {
                const double dx = std::max(0., std::max(b2[0] - b1[1], b1[0] - b2[1]));
                const double dy = std::max(0., std::max(b2[2] - b1[3], b1[2] - b2[3]));
                const double ub = order_upper_bound(p->P, sqrt(dx * dx + dy * dy), b1[4], b2[4], b1[5], b1[6], b2[5], b2[6]);
                if (far_2 && ub <= 2. - 1e-6) kind = 0;
                else if (far_full && ub <= far_top - 1e-6) kind = 1;
            }
            if (!p->h_labels.empty()) {
                const int *lI = &gg.labelrange[(size_t)I * 2], *lJ = &gg.labelrange[(size_t)J * 2];
                if (lI[0] == lI[1] && lJ[0] == lJ[1]) {
                    if (!pnb_class_active(p->P, lI[0], lJ[0])) continue;      // no pair of this unit belongs to the class
                } else if (kind == 0) kind = 1;                                // mixed labels: classified pair by pair
            }
            gh->kinds[(size_t)I * gg.ngroups + J] = (signed char)kind;
        }
    }
    if (gh->nparts != p->nparts) gh->gr_part.clear();
    if (p->dist && gh->gr_part.empty()) {
        // several GPUs: contiguous ranges of groups (Hilbert order) per part, balanced by the estimated work of the
        // units in the group's row and column (a part evaluates about half of the units that touch its groups)
        static const double kind_cost[3] = {1., 3.5, 9.};
        std::vector<double> gcost(gg.ngroups, 0.);
        double total = 0.;
        for (int I = 0; I < gg.ngroups; I++)
            for (int J = I; J < gg.ngroups; J++) {
                const int k = gh->kinds[(size_t)I * gg.ngroups + J];
                if (k < 0) continue;
                const double c = kind_cost[k] * (I == J ? 0.5 : 1.);
                gcost[I] += 0.5 * c; gcost[J] += 0.5 * c;
                total += c;
            }
        gh->gr_part.assign(gg.ngroups, 0);
        double acc = 0.;
        for (int g = 0; g < gg.ngroups; g++) {
            // the group goes to the part in whose share of the total its midpoint falls
            const double mid = acc + 0.5 * gcost[g];
            gh->gr_part[g] = total > 0. ? std::min(p->nparts - 1, (int)(mid / total * p->nparts)) : 0;
            acc += gcost[g];
        }
    }
    for (int I = 0; I < gg.ngroups; I++) {
        for (int J = I; J < gg.ngroups; J++) {
            const int kind = gh->kinds[(size_t)I * gg.ngroups + J];
            if (kind < 0) continue;
            GUnit u{I, J, kind, -1};
            const int ph = std::min(gpanel[I], gpanel[J]) * gh->nphase + gg.color[I] * ncol + gg.color[J];
            if (p->dist) {
                // the unit is evaluated by the part of its row group or of its column group, alternating
                const int rI = gh->gr_part[I], rJ = gh->gr_part[J];
                if ((((I + J) & 1) ? rJ : rI) != p->part) continue;
            }
            if (kind == 2) { u.slot = nslots++; gh->near_units.push_back(u); }
            (kind == 0 ? f2 : mix)[ph].push_back(u);
        }
    }
    gh->f2_units.clear(); gh->mix_units.clear();
    gh->f2_pbeg.assign(1, 0); gh->mix_pbeg.assign(1, 0);
    for (int ph = 0; ph < gh->nphase * gh->npanels; ph++) {
        // near units first: they are the longest
        std::stable_sort(mix[ph].begin(), mix[ph].end(), [](const GUnit &a, const GUnit &b) { return a.kind > b.kind; });
        gh->f2_units.insert(gh->f2_units.end(), f2[ph].begin(), f2[ph].end());
        gh->mix_units.insert(gh->mix_units.end(), mix[ph].begin(), mix[ph].end());
        if ((ph + 1) % gh->nphase == 0) {
            gh->f2_pbeg.push_back((int)gh->f2_units.size());
            gh->mix_pbeg.push_back((int)gh->mix_units.size());
        }
    }
    for (void *d : gh->unit_allocs) pool_free(d);
    gh->unit_allocs.clear();
    auto up = [&](const std::vector<GUnit> &v, const GUnit **dev) -> int {
        void *d = nullptr;
        CK(pool_malloc(&d, std::max<size_t>(v.size(), 1) * sizeof(GUnit)));
        gh->unit_allocs.push_back(d);
        if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(GUnit), cudaMemcpyHostToDevice));
        *dev = (const GUnit *)d;
        return 0;
    };
    if (up(gh->f2_units, &gh->d_f2) || up(gh->mix_units, &gh->d_mix) || up(gh->near_units, &gh->d_near)) return PNB_ERR_CUDA;
    {
        // list positions of the units (phase-major order: neighbours in the list never touch the same entries of U)
        std::vector<int> &ticket = gh->ticket;
        ticket.assign((size_t)gg.ngroups * gg.ngroups, -1);
        for (size_t k = 0; k < gh->f2_units.size(); k++) ticket[(size_t)gh->f2_units[k].I * gg.ngroups + gh->f2_units[k].J] = (int)k;
        for (size_t k = 0; k < gh->mix_units.size(); k++)
            ticket[(size_t)gh->mix_units[k].I * gg.ngroups + gh->mix_units[k].J] = (int)k | (1 << 30);
        void *d = nullptr;
        CK(pool_malloc(&d, ticket.size() * sizeof(int)));
        gh->unit_allocs.push_back(d);
        CK(cudaMemcpy(d, ticket.data(), ticket.size() * sizeof(int), cudaMemcpyHostToDevice));
        G.ticket = (const int *)d;
        CK(pool_malloc(&d, std::max<size_t>(gh->f2_units.size() + gh->mix_units.size(), 1) * sizeof(int)));
        gh->unit_allocs.push_back(d);
        G.done = (int *)d;
        G.nlist0 = (int)gh->f2_units.size();
    }
    gh->near_ready = false;
    gh->far_mask = p->far_mask;
    gh->max_order = p->P.max_order;
    gh->part = p->part;
    gh->nparts = p->nparts;
    gh->dist = p->dist;
    if (getenv("PNB_BENCH_VERBOSE"))
        fprintf(stderr, "group path: units f2 %zu, mix %zu (near %zu), phases %d; host ms: groups %.1f, units %.1f\n",
                gh->f2_units.size(), gh->mix_units.size(), gh->near_units.size(), gh->nphase, tw1 - tw0, wall_ms() - tw1);
    return 0;
}

// second half of the schedule: the near pair list.  Built after the f2 kernel has been launched, so that the host work
// and the two list kernels overlap with it (the list depends on mesh and tables only and is reused by later assemblies)
static int build_near_list(pnb_problem *p)
{
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    GroupSched &G = *p->G;
    if (gh->near_ready) return 0;
    const int nslots = (int)gh->near_units.size();
    const double tw2 = wall_ms();
    // ---- near pair list: count, allocate, fill (depends on mesh and tables only; reused by every assembly) ----
    for (void *d : gh->near_allocs) pool_free(d);
    gh->near_allocs.clear();
    gh->nitems = 0;
    if (!gh->near_units.empty()) {
        int *cursor = nullptr;
        CK(pool_malloc((void **)&cursor, 4 * sizeof(int)));
        gh->near_allocs.push_back(cursor);
        CK(cudaMemset(cursor, 0, 4 * sizeof(int)));
        smem_optin(gnear_list_kernel, p->device);
        int *bins = nullptr, *binbase = nullptr, *perm = nullptr;
        CK(pool_malloc((void **)&bins, 128 * sizeof(int)));
        gh->near_allocs.push_back(bins);
        CK(pool_malloc((void **)&binbase, 64 * sizeof(int)));
        gh->near_allocs.push_back(binbase);
        CK(cudaMemset(bins, 0, 128 * sizeof(int)));
        gnear_list_kernel<<<(unsigned)gh->near_units.size(), PNB_THREADS, gh->smem_near>>>(p->P, G, gh->d_near, p->far_mask, 0, cursor, nullptr, nullptr, nullptr, nullptr, bins, nullptr, nullptr);
        int tot[2] = {0, 0};
        CK(cudaMemcpy(tot, cursor, sizeof(tot), cudaMemcpyDeviceToHost));
        int hb[64], hbase[64];
        CK(cudaMemcpy(hb, bins, sizeof(hb), cudaMemcpyDeviceToHost));
        // processing order: highest orders first (longest items), singular pairs last; chunks of 8 warps
        std::vector<int4> chunks;
        {
            int off = 0;
            gh->near_nmax = 1;
            for (int t = 63; t >= 0; t--) {
                const int key = t;
                hbase[key] = off;
                if (key >= 1 && hb[key] > 0 && key <= p->P.max_order) gh->near_nmax = std::max(gh->near_nmax, p->h_reg_n[key]);
                const int K = key >= 1 ? (key < (int)p->h_grid.size() ? p->h_grid[key].y : 1) : 1;
                const int per = (PNB_NEAR_THREADS / 32) * K;
                for (int q = 0; q < hb[key]; q += per) chunks.push_back(make_int4(key, off + q, std::min(per, hb[key] - q), 0));
                off += hb[key];
            }
        }
        CK(cudaMemcpy(binbase, hbase, sizeof(hbase), cudaMemcpyHostToDevice));
        int4 *pairs = nullptr;
        int2 *items = nullptr;
        int *nearbase = nullptr;
        double *R = nullptr;
        int4 *dchunks = nullptr;
        CK(pool_malloc((void **)&pairs, std::max<size_t>(tot[0], 1) * sizeof(int4)));
        gh->near_allocs.push_back(pairs);
        CK(pool_malloc((void **)&items, std::max<size_t>(tot[1], 1) * sizeof(int2)));
        gh->near_allocs.push_back(items);
        CK(pool_malloc((void **)&perm, std::max<size_t>(tot[1], 1) * sizeof(int)));
        gh->near_allocs.push_back(perm);
        CK(pool_malloc((void **)&nearbase, (size_t)nslots * G.nbmax * G.nbmax * sizeof(int)));
        gh->near_allocs.push_back(nearbase);
        unsigned char *nearrow = nullptr;
        CK(pool_malloc((void **)&nearrow, (size_t)nslots * G.nbmax * G.nbmax * PNB_SB));
        gh->near_allocs.push_back(nearrow);
        if (tot[0] >= (1 << 28)) return fail(PNB_ERR_UNSUPPORTED, "more than 2^28 near pairs");
        CK(pool_malloc((void **)&R, std::max<size_t>(tot[1], 1) * PairDims<2>::NL * sizeof(double)));
        gh->near_allocs.push_back(R);
        double *F = nullptr;
        CK(pool_malloc((void **)&F, std::max<size_t>(tot[0], 1) * 21 * sizeof(double)));
        gh->near_allocs.push_back(F);
        G.F = F;
        gh->d_F = F;
        CK(pool_malloc((void **)&dchunks, std::max<size_t>(chunks.size(), 1) * sizeof(int4)));
        gh->near_allocs.push_back(dchunks);
        if (!chunks.empty()) CK(cudaMemcpy(dchunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice));
        CK(cudaMemset(cursor, 0, 4 * sizeof(int)));
        gnear_list_kernel<<<(unsigned)gh->near_units.size(), PNB_THREADS, gh->smem_near>>>(p->P, G, gh->d_near, p->far_mask, 1, cursor, pairs, items, nearbase, nearrow, bins, binbase, perm);
        CK(cudaDeviceSynchronize());
        gh->d_perm = perm;
        gh->d_chunks = dchunks;
        gh->nchunks = (int)chunks.size();
        G.npairs = pairs;
        G.nearbase = nearbase;
        G.nearrow = nearrow;
        G.R = R;
        gh->d_items = items;
        gh->d_R = R;
        gh->nitems = tot[1];
        gh->npairs = tot[0];
    }
    gh->near_ready = true;
    if (getenv("PNB_BENCH_VERBOSE"))
        fprintf(stderr, "group path: near pairs %d in %d items; host ms: near list %.1f\n", gh->npairs, gh->nitems, wall_ms() - tw2);
    return 0;
}

// part / nparts: share of the units evaluated by this instance (nparts > 1: dA is a full N x N scratch that holds
// this share of U + U^T afterwards; the caller sums the shares of all instances).  own tiles: cells whose home
// tile is owned get their boundary terms here.
static int build_dist_tables(pnb_problem *p);
// host_out != nullptr (pinned host memory, leading dimension host_ld): the rows of a panel are copied to the host as soon
// as the units that touch them are done, on a second stream, while the later panels are assembled
static int run_group_path(pnb_problem *p, int zero_exterior, double *dA, int64_t ld, int part = 0, int nparts = 1, int own_t0 = 0, int own_t1 = -1,
                          bool dist = false, double *host_out = nullptr, int64_t host_ld = 0)
{
    p->part = part;
    p->nparts = nparts;
    p->dist = dist;
    // one GPU: the surface terms do not depend on the unit schedule; their kernel runs while the host builds it
    const bool early_boundary = !dist && nparts == 1;
    p->early_boundary = early_boundary;
    if (early_boundary) {
        TileSched &S0 = p->S;
        S0.own_t0 = own_t0;
        S0.own_t1 = own_t1 < 0 ? S0.ntiles : own_t1;
        S0.cell_mask = nullptr;
        for (auto &e : p->bev) if (!e) cudaEventCreate(&e);
        cudaMemsetAsync(S0.err, 0, 4 * sizeof(int));
        cudaMemsetAsync(S0.counters, 0, 8 * sizeof(unsigned long long));
        cudaMemsetAsync(S0.Dbnd, 0, (size_t)p->nc * 6 * sizeof(double));
        cudaEventRecord(p->bev[0]);
        if (zero_exterior && p->nb > 0) boundary_kernel<2><<<(unsigned)(((size_t)p->nc * 32 + 255) / 256), 256>>>(p->P, S0);
        cudaEventRecord(p->bev[1]);
    }
    if (build_group_schedule(p)) return PNB_ERR_CUDA;
    if (dist && build_dist_tables(p)) return PNB_ERR_CUDA;
    if (!dist && p->G) p->G->dist.nparts = 0;
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    GroupSched &G = *p->G;
    TileSched &S = p->S;
    const int nc = p->nc, N = p->N;
    S.own_t0 = own_t0;
    S.own_t1 = own_t1 < 0 ? S.ntiles : own_t1;
    S.cell_mask = dist ? gh->d_cell_mask : nullptr;
    // unit slots of the other instances stay untouched: start from zero
    if (nparts > 1) cudaMemsetAsync(G.Dp, 0, (size_t)G.ngroups * nc * 6 * sizeof(double));
    for (auto &e : p->ev) if (!e) cudaEventCreate(&e);
    smem_optin(gf2_kernel, p->device);
    smem_optin(gmix_kernel, p->device);
    F2Rule R;
    {
        const FarRule &F = p->far_rules[2];
        for (int q = 0; q < 3; q++) {
            R.w[q] = F.w[q];
            for (int k = 0; k < 3; k++) { R.bary[k][q] = F.bary[k][q]; R.wphi[q][k] = F.wb[k][q]; }
            for (int e = 0; e < 6; e++) R.qq[e][q] = F.qq[e][q];
        }
        R.c[0] = 1.;
        for (int k = 1; k < 8; k++) R.c[k] = R.c[k - 1] * (p->P.expo - k + 1) / k;     // = PowTab::coef
        R.eoff = p->pow_eoff - 1023;
        R.pad = 0;
    }
    if (!early_boundary) {
        cudaMemsetAsync(S.err, 0, 4 * sizeof(int));
        cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned long long));
        cudaMemsetAsync(S.Dbnd, 0, (size_t)nc * 6 * sizeof(double));
    }
    cudaEventRecord(p->ev[0]);
    if (!dist) cudaMemset2DAsync(dA, (size_t)ld * sizeof(double), 0, (size_t)N * sizeof(double), N);
    int launches = 0;
    {
        // one persistent launch per unit list; the units take tickets in list order and order their updates of U
        // among themselves (g_wait_predecessors)
        int nsm = 148;
        nsm = device_attr(cudaDevAttrMultiProcessorCount, p->device);
        const int nf = (int)gh->f2_units.size(), nm = (int)gh->mix_units.size();
        cudaMemsetAsync(G.counters_i, 0, (2 * PNB_ROW_PANELS + 2) * sizeof(int));
        cudaMemsetAsync(G.done, 0, std::max<size_t>((size_t)nf + nm, 1) * sizeof(int));
        for (auto &e : p->kev) if (!e) cudaEventCreate(&e);
        cudaEventRecord(p->kev[0]);
        // device output: one launch per list.  Host output: the order-2 units of the first panel, the near evaluator, then
        // panel by panel (order-2 units, other units, symmetrisation of the panel's rows, copy on the second stream)
        const bool panels = host_out != nullptr && !dist && gh->npanels > 1;
        p->host_panels = panels;
        const int f2_first_end = panels ? gh->f2_pbeg[1] : nf;
        if (f2_first_end > 0) {
            gf2_kernel<<<std::min(f2_first_end, 2 * nsm), PNB_F2T, gh->smem_f2>>>(p->P, G, gh->d_f2, 0, f2_first_end, G.counters_i, dA, ld, R);
            launches++;
        }
        cudaEventRecord(p->kev[1]);
        // the near pair list is built (first assembly only) while the f2 kernel runs; its results feed the mix kernel
        if (build_near_list(p)) return PNB_ERR_CUDA;
        if (gh->nitems > 0) {
            const int wpb = PNB_NEAR_THREADS / 32;
            // rule table of the largest order that has items; per warp PNB_NEAR_WARP_POINTS points
            // two CTAs per SM: the table holds at most the nodes that fit beside the power table and the warp buffers
            int smem_sm = 0;
            smem_sm = device_attr(cudaDevAttrMaxSharedMemoryPerMultiprocessor, p->device);
            if (smem_sm <= 0) smem_sm = 228 * 1024;
            const size_t fixed = sizeof(PowTabS) + (size_t)wpb * PNB_NEAR_WARP_POINTS * sizeof(double2);
            const size_t per_cta = (size_t)smem_sm / 2 - 1024 - 256;
            const int fit = per_cta > fixed ? (int)((per_cta - fixed) / (PNB_DER2 * sizeof(double2))) : 0;
            const int der_nodes = std::max(1, std::min(gh->near_nmax, fit));
            const size_t smem_eval = fixed + (size_t)PNB_DER2 * der_nodes * sizeof(double2);
            smem_optin(gnear_eval_kernel, p->device);
            int nsm = 148;
            nsm = device_attr(cudaDevAttrMultiProcessorCount, p->device);
            const int grid = std::min(gh->nchunks, 2 * nsm);
            gnear_eval_kernel<<<grid, PNB_NEAR_THREADS, smem_eval>>>(p->P, G.npairs, gh->d_items, gh->d_perm, gh->d_chunks, gh->nchunks, gh->d_R,
                                                                der_nodes, PNB_NEAR_WARP_POINTS);
            gnear_finalize_kernel<<<(gh->npairs + 255) / 256, 256>>>(p->P, G.npairs, gh->npairs, gh->d_R, gh->d_F);
            launches += 2;
        }
        cudaEventRecord(p->kev[2]);
        if (nm > 0) {
            // unit blocks of the resident CTAs (one per SM): global scratch, L2 resident
            const int ncta = std::min(nm, nsm);
            const size_t need = (size_t)ncta * gmix_scratch_doubles(G.cap, G.maxld, G.ldS) * sizeof(double);
            if (gh->mix_scratch_bytes < need) {
                if (gh->mix_scratch) pool_free(gh->mix_scratch);
                gh->mix_scratch = nullptr;
                gh->mix_scratch_bytes = 0;
                CK(pool_malloc(&gh->mix_scratch, need));
                gh->mix_scratch_bytes = need;
            }
            if (!panels) {
                gmix_kernel<<<ncta, gh->mix_warps * 32, gh->smem_mix>>>(p->P, G, gh->d_mix, 0, nm, G.counters_i + 1, dA, ld, p->far_mask,
                                                                        (double *)gh->mix_scratch);
                launches++;
            }
        }
        if (panels) {
            if (!p->copy_stream) CK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
            for (auto &e : p->pev) if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            const unsigned nt = (unsigned)((N + 31) / 32);
            for (int k = 0; k < gh->npanels; k++) {
                const int f0 = gh->f2_pbeg[k], f1 = gh->f2_pbeg[k + 1], m0 = gh->mix_pbeg[k], m1 = gh->mix_pbeg[k + 1];
                if (k > 0 && f1 > f0) {
                    gf2_kernel<<<std::min(f1 - f0, 2 * nsm), PNB_F2T, gh->smem_f2>>>(p->P, G, gh->d_f2, f0, f1, G.counters_i + 2 * k, dA, ld, R);
                    launches++;
                }
                if (m1 > m0) {
                    gmix_kernel<<<std::min(m1 - m0, nsm), gh->mix_warps * 32, gh->smem_mix>>>(p->P, G, gh->d_mix, m0, m1, G.counters_i + 2 * k + 1, dA,
                                                                                              ld, p->far_mask, (double *)gh->mix_scratch);
                    launches++;
                }
                // rows of the panel: U + U^T for the tiles right of the diagonal tile (both images), final from here on
                const int t0 = k * gh->panel_tiles, t1 = std::min((int)nt, t0 + gh->panel_tiles);
                symmetrize_kernel<<<dim3(nt, (unsigned)(t1 - t0)), 256>>>(dA, ld, N, t0);
                launches++;
                cudaEventRecord(p->pev[k]);
                cudaStreamWaitEvent(p->copy_stream, p->pev[k], 0);
                const int r0 = t0 * 32, r1 = std::min(N, t1 * 32);
                CK(cudaMemcpy2DAsync(host_out + (size_t)r0 * host_ld, (size_t)host_ld * sizeof(double), dA + (size_t)r0 * ld, (size_t)ld * sizeof(double),
                                     (size_t)N * sizeof(double), (size_t)(r1 - r0), cudaMemcpyDeviceToHost, p->copy_stream));
            }
        }
    }
    cudaEventRecord(p->kev[3]);
    const bool panels_done = p->host_panels;
    if (!dist && !panels_done) {
        const unsigned nt = (unsigned)((N + 31) / 32);
        symmetrize_kernel<<<dim3(nt, nt), 256>>>(dA, ld, N, 0);
        launches++;
    }
    cudaEventRecord(p->kev[4]);
    cudaEventRecord(p->ev[1]);
    if (zero_exterior && p->nb > 0) {
        if (!early_boundary) boundary_kernel<2><<<(unsigned)(((size_t)nc * 32 + 255) / 256), 256>>>(p->P, S);
        launches++;
    }
    cudaEventRecord(p->ev[2]);
    greduce_D_kernel<<<(unsigned)(((size_t)nc * 6 + 255) / 256), 256>>>(G, S.D, S.Dbnd, nc, zero_exterior && p->nb > 0);
    launches++;
    cudaEventRecord(p->ev[3]);
    p->stats[2] = launches;
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// several GPUs: parts own groups and the rows of their dofs (no N x N scratch, no collective over matrix entries)
// ---------------------------------------------------------------------------
// Tables of the distributed assembly, built after the unit lists: which part owns every dof (the part of the first
// group, in Hilbert order, that holds it), where the fragments of every unit start in the staging buffer of every
// part, the rows of this part and the lookup tables of dist_apply_kernel.
static int build_dist_tables(pnb_problem *p)
{
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    GroupSched &G = *p->G;
    if (gh->dist_ready) return 0;
    const GroupGeom &gg = gh->gg;
    const int ng = gg.ngroups, W = p->nparts, me = p->part, N = p->N, nc = p->nc;
    if (W > PNB_MAX_PARTS) return fail(PNB_ERR_ARG, "too many parts");
    const double tw0 = wall_ms();
    for (void *d : gh->dist_allocs) pool_free(d);
    gh->dist_allocs.clear();
    // owner of every dof
    std::vector<int> owner(N, -1);
    std::vector<int> d2g_cnt(N + 1, 0);
    for (int g = 0; g < ng; g++)
        for (int k = gg.gdptr[g]; k < gg.gdptr[g + 1]; k++) {
            const int d = gg.gdofs[k];
            if (owner[d] < 0) owner[d] = gh->gr_part[g];
            d2g_cnt[d + 1]++;
        }
    for (int d = 0; d < N; d++) d2g_cnt[d + 1] += d2g_cnt[d];
    std::vector<int2> d2g(gg.gdofs.size());
    {
        std::vector<int> fill(d2g_cnt.begin(), d2g_cnt.end() - 1);
        for (int g = 0; g < ng; g++)
            for (int k = gg.gdptr[g]; k < gg.gdptr[g + 1]; k++) d2g[fill[gg.gdofs[k]]++] = make_int2(g, k - gg.gdptr[g]);
    }
    gh->gown.assign(gg.gdofs.size(), 0);
    gh->gpos.assign(gg.gdofs.size(), 0);
    gh->gcnt.assign((size_t)ng * W, 0);
    for (int g = 0; g < ng; g++)
        for (int k = gg.gdptr[g]; k < gg.gdptr[g + 1]; k++) {
            const int o = owner[gg.gdofs[k]];
            gh->gown[k] = (unsigned char)o;
            gh->gpos[k] = (unsigned char)gh->gcnt[(size_t)g * W + o]++;
        }
    gh->rows.clear();
    for (int d = 0; d < N; d++)
        if (owner[d] == me) gh->rows.push_back(d);
    // fragments of every unit in the staging buffer of part o: rows of I owned by o (nldJ values each), then rows of J
    // owned by o (nldI values each); units in lexicographic order
    const size_t nf2 = gh->f2_units.size(), nmix = gh->mix_units.size();
    std::vector<long long> uoff_f2(std::max<size_t>(nf2, 1) * W, 0), uoff_mix(std::max<size_t>(nmix, 1) * W, 0);
    gh->uoff_all.assign((size_t)ng * ng, -1);
    std::vector<long long> run(W, 0);
    for (int I = 0; I < ng; I++) {
        const int nldI = gg.gdptr[I + 1] - gg.gdptr[I];
        for (int J = I; J < ng; J++) {
            if (gh->kinds[(size_t)I * ng + J] < 0) continue;
            const int nldJ = gg.gdptr[J + 1] - gg.gdptr[J];
            const int tk = gh->ticket[(size_t)I * ng + J];
            long long *mine = tk < 0 ? nullptr : ((tk >> 30) ? &uoff_mix[(size_t)(tk & 0x3FFFFFFF) * W] : &uoff_f2[(size_t)tk * W]);
            for (int o = 0; o < W; o++) {
                const long long sz = (long long)gh->gcnt[(size_t)I * W + o] * nldJ + (long long)gh->gcnt[(size_t)J * W + o] * nldI;
                if (mine) mine[o] = run[o];
                if (o == me && sz > 0) gh->uoff_all[(size_t)I * ng + J] = run[o];
                run[o] += sz;
            }
        }
    }
    gh->stage_doubles = run[me];
    // groups by colour
    std::vector<int> colptr(gg.ncolors + 1, 0), collist(ng);
    for (int g = 0; g < ng; g++) colptr[gg.color[g] + 1]++;
    for (int c = 0; c < gg.ncolors; c++) colptr[c + 1] += colptr[c];
    {
        std::vector<int> fill(colptr.begin(), colptr.end() - 1);
        for (int g = 0; g < ng; g++) collist[fill[gg.color[g]]++] = g;
    }
    // cells whose surface terms this part computes: the cells of its groups
    std::vector<unsigned char> mask(nc, 0);
    for (int g = 0; g < ng; g++)
        if (gh->gr_part[g] == me)
            for (int k = gg.gptr[g]; k < gg.gptr[g + 1]; k++)
                if (gg.gcells[k] >= 0) mask[gg.gcells[k]] = 1;
    auto up = [&](const void *host, size_t bytes, const void **dev) -> int {
        void *d = nullptr;
        CK(pool_malloc(&d, std::max<size_t>(bytes, 16)));
        gh->dist_allocs.push_back(d);
        if (bytes) CK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
        *dev = d;
        return 0;
    };
    DistSched &D = G.dist;
    DistApply &X = gh->apply;
    int rc = 0;
    rc |= up(gh->gown.data(), gh->gown.size(), (const void **)&D.gown);
    rc |= up(gh->gpos.data(), gh->gpos.size(), (const void **)&D.gpos);
    rc |= up(gh->gcnt.data(), gh->gcnt.size() * sizeof(int), (const void **)&D.gcnt);
    rc |= up(uoff_f2.data(), uoff_f2.size() * sizeof(long long), (const void **)&D.uoff_f2);
    rc |= up(uoff_mix.data(), uoff_mix.size() * sizeof(long long), (const void **)&D.uoff_mix);
    rc |= up(gh->rows.data(), gh->rows.size() * sizeof(int), (const void **)&X.rows);
    rc |= up(d2g_cnt.data(), d2g_cnt.size() * sizeof(int), (const void **)&X.d2g_ptr);
    rc |= up(d2g.data(), d2g.size() * sizeof(int2), (const void **)&X.d2g);
    rc |= up(colptr.data(), colptr.size() * sizeof(int), (const void **)&X.colptr);
    rc |= up(collist.data(), collist.size() * sizeof(int), (const void **)&X.collist);
    rc |= up(gh->uoff_all.data(), gh->uoff_all.size() * sizeof(long long), (const void **)&X.uoff);
    rc |= up(mask.data(), mask.size(), (const void **)&gh->d_cell_mask);
    if (rc) return PNB_ERR_CUDA;
    X.nrows = (int)gh->rows.size();
    D.nparts = W;
    D.part = me;
    gh->dist_ready = true;
    if (getenv("PNB_BENCH_VERBOSE"))
        fprintf(stderr, "dist plan: part %d of %d, %d rows, staging %.2f GB; host ms %.1f\n", me, W, X.nrows, gh->stage_doubles * 8e-9,
                wall_ms() - tw0);
    return 0;
}

// Several GPUs, 2D.  Part `part` of `nparts` (one problem instance per GPU) owns a contiguous range of cell groups
// (cells ordered along a Hilbert curve, ranges balanced by estimated work) and the rows of the dofs that first appear in
// its groups.  Every cell pair is evaluated once over all parts: a unit of two groups by the part of its row group or
// of its column group, alternating.  Replaces the reference's cell-range split + Allreduce of the N x N matrix
// (nonlocalAssembly_{SCALAR}.pxi:1280-1285, 1449-1450): per GPU only the owned rows and a staging buffer of about
// twice their size exist, and no collective moves matrix entries.
extern "C" int pnb_dist_plan(pnb_problem *p, int32_t nparts, int32_t part, int32_t *num_rows, int64_t *staging_doubles)
{
    if (!p || !num_rows || !staging_doubles) return fail(PNB_ERR_ARG, "null argument");
    if (p->dim != 2) return fail(PNB_ERR_UNSUPPORTED, "pnb_dist_plan: 2D only (1D problems use row blocks)");
    if (p->finite) return fail(PNB_ERR_UNSUPPORTED, "pnb_dist_plan: finite horizon operators use row blocks (pnb_dense_rows_begin/_end)");
    if (!p->has_singular) return fail(PNB_ERR_ARG, "problem was created without quadrature tables");
    if (nparts < 1 || nparts > PNB_MAX_PARTS || part < 0 || part >= nparts) return fail(PNB_ERR_ARG, "invalid part");
    ON_DEVICE(p->device);
    p->part = part;
    p->nparts = nparts;
    p->dist = true;
    if (build_group_schedule(p)) return PNB_ERR_CUDA;
    if (build_dist_tables(p)) return PNB_ERR_CUDA;
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    *num_rows = (int32_t)gh->rows.size();
    *staging_doubles = gh->stage_doubles;
    return 0;
}

// global dof of every owned row, ascending (host, num_rows entries)
extern "C" int pnb_dist_rows(pnb_problem *p, int32_t *rows)
{
    if (!p || !rows) return fail(PNB_ERR_ARG, "null argument");
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    if (!gh || !gh->dist_ready) return fail(PNB_ERR_ARG, "pnb_dist_rows: call pnb_dist_plan first");
    memcpy(rows, gh->rows.data(), gh->rows.size() * sizeof(int));
    return 0;
}

// Evaluates the units of this part and stores their blocks, row by row, into the staging buffers of the row owners.
// stage_ptrs: host array of nparts device pointers, the staging buffer (staging_doubles of pnb_dist_plan) of every
// part as seen from this device (own allocation, or peer memory opened with pnb_ipc_import).  Asynchronous; the
// cell-diagonal blocks of this part's share are in the pnb_dense_cell_blocks buffer afterwards.
extern "C" int pnb_dist_eval(pnb_problem *p, int zero_exterior, double *const *stage_ptrs)
{
    if (!p || !stage_ptrs) return fail(PNB_ERR_ARG, "null argument");
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    if (!gh || !p->dist) return fail(PNB_ERR_ARG, "pnb_dist_eval: call pnb_dist_plan first");
    ON_DEVICE(p->device);
    for (int o = 0; o < p->nparts; o++) p->G->dist.stage[o] = stage_ptrs[o];
    return run_group_path(p, zero_exterior, nullptr, 0, p->part, p->nparts, 0, -1, true);
}

// largest regular quadrature order requested beyond the supplied tables by the last pnb_dist_eval (0: none).
// Synchronises the device.  The caller takes the maximum over all parts and, if it is positive, raises the tables
// on ALL parts (pnb_problem_set_rules) and repeats the evaluation: the decision must be collective.
extern "C" int pnb_dist_status(pnb_problem *p, int32_t *order_needed)
{
    if (!p || !order_needed) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    CK(cudaDeviceSynchronize());
    int herr[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(herr, p->S.err, sizeof(herr), cudaMemcpyDeviceToHost));
    if (herr[1] > 0) return fail(PNB_ERR_CUDA, "internal error: a unit classified as far holds a near pair");
    *order_needed = herr[0];
    return 0;
}

// Owned rows of the operator from the staged fragments (all parts must have finished pnb_dist_eval: the caller
// synchronises them) plus, if use_cell_blocks, the cell-diagonal blocks of the pnb_dense_cell_blocks buffer (summed
// over the parts by the caller).  A_rows: device, num_rows x num_dofs, row k = global row rows[k].
extern "C" int pnb_dist_apply(pnb_problem *p, int use_cell_blocks, double *A_rows, int64_t ld)
{
    if (!p || !A_rows) return fail(PNB_ERR_ARG, "null argument");
    GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
    if (!gh || !gh->dist_ready) return fail(PNB_ERR_ARG, "pnb_dist_apply: call pnb_dist_plan first");
    if (ld < p->N) return fail(PNB_ERR_ARG, "leading dimension smaller than num_dofs");
    ON_DEVICE(p->device);
    DistApply X = gh->apply;
    X.stage = p->G->dist.stage[p->part];
    if (X.nrows > 0) dist_apply_kernel<<<X.nrows, 256>>>(p->P, *p->G, X, p->S.dof_ptr, p->S.dof_cells, p->S.D, use_cell_blocks, A_rows, ld);
    p->stats[2] += 1;
    cudaEventRecord(p->ev[4]);
    cudaError_t e = cudaEventSynchronize(p->ev[4]);
    if (e == cudaSuccess) e = cudaGetLastError();
    unsigned long long hcnt[8] = {0};
    if (e == cudaSuccess) e = cudaMemcpy(hcnt, p->S.counters, sizeof(hcnt), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(PNB_ERR_CUDA, std::string("distributed assembly: ") + cudaGetErrorString(e));
    float ms;
    for (int k = 0; k < 2; k++) { cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]); p->timings[k] = ms; }
    cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]);
    float ms2 = 0.f;
    cudaEventElapsedTime(&ms2, p->ev[3], p->ev[4]);
    p->timings[2] = ms + ms2;      // reduction of the cell blocks + (exchange by the caller) + apply
    cudaEventElapsedTime(&ms, p->ev[0], p->ev[4]);
    p->timings[3] = ms;
    p->stats[0] = (int64_t)(hcnt[0] + hcnt[1] + hcnt[2]);
    p->stats[3] = (int64_t)hcnt[1];
    p->stats[4] = (int64_t)hcnt[2];
    p->stats[1] = p->distinct_pairs;
    if (p->kev[0])
        for (int k = 0; k < 4; k++) {
            float kms = 0.f;
            if (cudaEventElapsedTime(&kms, p->kev[k], p->kev[k + 1]) == cudaSuccess) p->ktimings[k] = kms;
            else { cudaGetLastError(); p->ktimings[k] = 0.; }
        }
    return 0;
}

// plain device allocations that can be shared with the other processes of the node (peer memory over NVLink)
extern "C" int pnb_device_alloc(int device, int64_t bytes, void **dptr)
{
    if (!dptr || bytes < 0) return fail(PNB_ERR_ARG, "invalid argument");
    ON_DEVICE(device);
    CK(cudaMalloc(dptr, (size_t)std::max<int64_t>(bytes, 256)));
    return 0;
}

extern "C" int pnb_device_free(int device, void *dptr)
{
    ON_DEVICE(device);
    if (dptr) CK(cudaFree(dptr));
    return 0;
}

// handle: 64 bytes (cudaIpcMemHandle_t) that another process of the node turns into a device pointer
extern "C" int pnb_ipc_export(int device, void *dptr, unsigned char *handle)
{
    if (!dptr || !handle) return fail(PNB_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    ON_DEVICE(device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return 0;
}

extern "C" int pnb_ipc_import(int device, const unsigned char *handle, void **dptr)
{
    if (!dptr || !handle) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int pnb_ipc_close(int device, void *dptr)
{
    ON_DEVICE(device);
    if (dptr) CK(cudaIpcCloseMemHandle(dptr));
    return 0;
}

static void destroy_group_host(pnb_problem *p)
{
    if (p->gh) {
        GroupHostFull *gh = static_cast<GroupHostFull *>(p->gh);
        for (void *d : gh->unit_allocs) pool_free(d);
        for (void *d : gh->near_allocs) pool_free(d);
        for (void *d : gh->dist_allocs) pool_free(d);
        if (gh->mix_scratch) pool_free(gh->mix_scratch);
        delete gh;
        p->gh = nullptr;
    }
    delete p->G;
    p->G = nullptr;
}

static int check_rows(pnb_problem *p, int32_t row_begin, int32_t row_end)
{
    if (row_begin < 0 || row_end > p->N || row_begin >= row_end) return fail(PNB_ERR_ARG, "invalid row range");
    if (row_begin % PNB_TD != 0 || (row_end % PNB_TD != 0 && row_end != p->N))
        return fail(PNB_ERR_ARG, "row blocks must start and end at multiples of " + std::to_string(PNB_TD) + " (pnb_row_granularity)");
    return 0;
}

extern "C" int pnb_row_granularity(void) { return PNB_TD; }

// tile passes + boundary kernel + reduction of the cell-diagonal blocks of the owned cells
static int dense_rows_begin_impl(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end, double *dA, int64_t ld,
                                 double *host_out, int64_t host_ld);
extern "C" int pnb_dense_rows_begin(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end, double *dA, int64_t ld)
{
    return dense_rows_begin_impl(p, zero_exterior, row_begin, row_end, dA, ld, nullptr, 0);
}

static int dense_rows_begin_impl(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end, double *dA, int64_t ld,
                                 double *host_out, int64_t host_ld)
{
    if (!p || !dA) return fail(PNB_ERR_ARG, "null argument");
    p->host_panels = false;
    if (!p->has_singular) return fail(PNB_ERR_ARG, "problem was created without quadrature tables");
    if (check_rows(p, row_begin, row_end)) return PNB_ERR_ARG;
    if (ld < p->N) return fail(PNB_ERR_ARG, "leading dimension smaller than num_dofs");
    ON_DEVICE(p->device);
    const int nc = p->nc, nvc = p->dim + 1, ND = nvc * (nvc + 1) / 2;
    TileSched &S = p->S;
    // 2D, whole operator, infinite horizon: cell-group path (pnb_problem_set_path(p, 1) selects the DoF-tile path)
    if (p->nblocks > 0 && (row_begin != 0 || row_end != p->N)) return fail(PNB_ERR_ARG, "batched blocks are assembled as a whole");
    if (p->dim == 2 && !p->finite && row_begin == 0 && row_end == p->N && p->path == 0)
        return run_group_path(p, zero_exterior, dA, ld, 0, 1, 0, -1, false, host_out, host_ld);
    if (build_tile_schedule(p)) return PNB_ERR_CUDA;
    S.cell_mask = nullptr;
    S.own_t0 = row_begin / PNB_TD;
    S.own_t1 = (row_end + PNB_TD - 1) / PNB_TD;
    // units: group pairs (gr <= gc) that hold a tile touching an owned row tile, near-diagonal first
    {
        const int g0 = S.own_t0 / S.G, g1 = (S.own_t1 - 1) / S.G;
        std::vector<int> units;
        if (p->nblocks > 0) units = p->h_block_units;
        else
        for (int dg = 0; dg < S.ngroups; dg++)
            for (int gr = 0; gr + dg < S.ngroups; gr++) {
                const int gc = gr + dg;
                if ((gr >= g0 && gr <= g1) || (gc >= g0 && gc <= g1)) { units.push_back(gr); units.push_back(gc); }
            }
        S.nunits = (int)units.size() / 2;
        CK(cudaMemcpy(const_cast<int *>(S.units), units.data(), units.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    for (auto &e : p->ev) if (!e) cudaEventCreate(&e);
    cudaMemsetAsync(S.DXp, 0, (size_t)S.dgroups * nc * ND * sizeof(double));
    cudaMemsetAsync(S.DYp, 0, (size_t)S.dgroups * nc * ND * sizeof(double));
    cudaMemsetAsync(S.Dbnd, 0, (size_t)nc * ND * sizeof(double));
    cudaMemsetAsync(S.err, 0, 4 * sizeof(int));
    cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned long long));
    cudaMemsetAsync(S.tileflag, 0, (size_t)S.ntiles * S.tf_width * sizeof(int));
    cudaMemsetAsync(S.unitflag, 0, (size_t)S.nunits_all * sizeof(int));
    cudaEventRecord(p->ev[0]);
    int launches = 0;
    int nnear_units = 0;
    if (p->dim == 2) {
        const size_t smem = sizeof(TileSmem<2, true>) + 2 * (size_t)S.maxcells * ND * sizeof(double);
        const size_t smem_far = sizeof(TileSmem<2, false>) + 2 * (size_t)S.maxcells * ND * sizeof(double);
        if (p->finite) {
            smem_optin(tile_kernel<2, false, true>, p->device);
            smem_optin(tile_kernel<2, true, true>, p->device);
            tile_kernel<2, false, true><<<S.nunits, PNB_THREADS, smem_far>>>(p->P, S, dA, ld, p->far_mask);
        } else {
            smem_optin(tile_kernel<2, false, false>, p->device);
            smem_optin(tile_kernel<2, true, false>, p->device);
            tile_kernel<2, false, false><<<S.nunits, PNB_THREADS, smem_far>>>(p->P, S, dA, ld, p->far_mask);
        }
        compact_units_kernel<<<1, 1024>>>(S);
        cudaMemcpy(&nnear_units, S.nearunits + S.nunits, sizeof(int), cudaMemcpyDeviceToHost);
        if (nnear_units > 0) {
            if (p->finite) tile_kernel<2, true, true><<<nnear_units, PNB_THREADS, smem>>>(p->P, S, dA, ld, p->far_mask);
            else tile_kernel<2, true, false><<<nnear_units, PNB_THREADS, smem>>>(p->P, S, dA, ld, p->far_mask);
        }
    } else {
        const size_t smem = sizeof(TileSmem<1, true>) + 2 * (size_t)S.maxcells * ND * sizeof(double);
        const size_t smem_far = sizeof(TileSmem<1, false>) + 2 * (size_t)S.maxcells * ND * sizeof(double);
        if (p->finite) {
            smem_optin(tile_kernel<1, false, true>, p->device);
            smem_optin(tile_kernel<1, true, true>, p->device);
            tile_kernel<1, false, true><<<S.nunits, PNB_THREADS, smem_far>>>(p->P, S, dA, ld, 0);
        } else {
            smem_optin(tile_kernel<1, false, false>, p->device);
            smem_optin(tile_kernel<1, true, false>, p->device);
            tile_kernel<1, false, false><<<S.nunits, PNB_THREADS, smem_far>>>(p->P, S, dA, ld, 0);
        }
        compact_units_kernel<<<1, 1024>>>(S);
        cudaMemcpy(&nnear_units, S.nearunits + S.nunits, sizeof(int), cudaMemcpyDeviceToHost);
        if (nnear_units > 0) {
            if (p->finite) tile_kernel<1, true, true><<<nnear_units, PNB_THREADS, smem>>>(p->P, S, dA, ld, 0);
            else tile_kernel<1, true, false><<<nnear_units, PNB_THREADS, smem>>>(p->P, S, dA, ld, 0);
        }
    }
    launches += nnear_units > 0 ? 3 : 2;
    cudaEventRecord(p->ev[1]);
    if (zero_exterior && p->nb > 0) {
        const unsigned blocks = (unsigned)(((size_t)nc * 32 + 255) / 256);
        if (p->dim == 2) boundary_kernel<2><<<blocks, 256>>>(p->P, S);
        else boundary_kernel<1><<<blocks, 256>>>(p->P, S);
        launches++;
    }
    cudaEventRecord(p->ev[2]);
    reduce_D_kernel<<<(unsigned)(((size_t)nc * ND + 255) / 256), 256>>>(S, nc, ND, zero_exterior && p->nb > 0);
    launches++;
    cudaEventRecord(p->ev[3]);
    p->stats[2] = launches;
    CK(cudaGetLastError());
    return 0;
}

// device buffer of the cell-diagonal blocks: num_cells x (dim+1)(dim+2)/2 doubles.  After _begin it holds the
// complete blocks of the cells whose home row tile is owned and zeros elsewhere; with several row blocks the
// caller sums the buffers of all owners (disjoint supports: the sum is exact) before calling _end.
extern "C" int pnb_dense_cell_blocks(pnb_problem *p, double **dptr, int64_t *count)
{
    if (!p || !dptr || !count) return fail(PNB_ERR_ARG, "null argument");
    const int nvc = p->dim + 1;
    *dptr = p->S.D;
    *count = (int64_t)p->nc * (nvc * (nvc + 1) / 2);
    return 0;
}

// copies between the cell-block buffer and a caller device buffer (direction != 0: caller -> problem)
extern "C" int pnb_dense_cell_blocks_copy(pnb_problem *p, double *device_buf, int to_problem)
{
    if (!p || !device_buf) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    const int nvc = p->dim + 1;
    const size_t bytes = (size_t)p->nc * (nvc * (nvc + 1) / 2) * sizeof(double);
    if (to_problem) CK(cudaMemcpyAsync(p->S.D, device_buf, bytes, cudaMemcpyDeviceToDevice));
    else CK(cudaMemcpyAsync(device_buf, p->S.D, bytes, cudaMemcpyDeviceToDevice));
    return 0;
}

// Zero-exterior surface terms alone: per cell the symmetric (dim+1) x (dim+1) block (upper triangle, row-major,
// num_cells x (dim+1)(dim+2)/2 doubles, host buffer) of sum over the boundary facets.  The H2 near field of the
// regional operator subtracts these (assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1889-1912).
extern "C" int pnb_boundary_cell_blocks(pnb_problem *p, double *host_out)
{
    if (!p || !host_out) return fail(PNB_ERR_ARG, "null argument");
    if (!p->has_singular) return fail(PNB_ERR_ARG, "problem was created without quadrature tables");
    ON_DEVICE(p->device);
    const int nc = p->nc, nvc = p->dim + 1, ND = nvc * (nvc + 1) / 2;
    TileSched &S = p->S;
    S.own_t0 = 0;
    S.own_t1 = S.ntiles;
    S.cell_mask = nullptr;
    CK(cudaMemsetAsync(S.Dbnd, 0, (size_t)nc * ND * sizeof(double)));
    CK(cudaMemsetAsync(S.err, 0, 4 * sizeof(int)));
    if (p->nb > 0) {
        const unsigned blocks = (unsigned)(((size_t)nc * 32 + 255) / 256);
        if (p->dim == 2) boundary_kernel<2><<<blocks, 256>>>(p->P, S);
        else boundary_kernel<1><<<blocks, 256>>>(p->P, S);
        CK(cudaGetLastError());
    }
    int herr[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(herr, S.err, sizeof(herr), cudaMemcpyDeviceToHost));
    if (herr[0] > 0) return fail(PNB_ERR_ORDER, "regular quadrature order " + std::to_string(herr[0]) + " exceeds the supplied tables (max_order " + std::to_string(p->P.max_order) + ")");
    CK(cudaMemcpy(host_out, S.Dbnd, (size_t)nc * ND * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// scatter of the cell-diagonal blocks into the owned rows; collects errors, counters and timings
extern "C" int pnb_dense_rows_end(pnb_problem *p, int32_t row_begin, int32_t row_end, double *dA, int64_t ld)
{
    if (!p || !dA) return fail(PNB_ERR_ARG, "null argument");
    if (check_rows(p, row_begin, row_end)) return PNB_ERR_ARG;
    ON_DEVICE(p->device);
    TileSched &S = p->S;
    if (S.own_t0 != row_begin / PNB_TD) return fail(PNB_ERR_ARG, "pnb_dense_rows_end does not match pnb_dense_rows_begin");
    const int nrows = row_end - row_begin;
    scatter_D_kernel<<<(nrows + 127) / 128, 128>>>(p->P, S, dA, ld);
    p->stats[2] += 1;
    cudaEventRecord(p->ev[4]);
    cudaError_t e = cudaEventSynchronize(p->ev[4]);
    if (e == cudaSuccess) e = cudaGetLastError();
    int herr[4] = {0, 0, 0, 0};
    unsigned long long hcnt[8] = {0};
    if (e == cudaSuccess) e = cudaMemcpy(herr, S.err, sizeof(herr), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(hcnt, S.counters, sizeof(hcnt), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(PNB_ERR_CUDA, std::string("dense assembly: ") + cudaGetErrorString(e));
    if (herr[0] > 0) return fail(PNB_ERR_ORDER, "regular quadrature order " + std::to_string(herr[0]) + " exceeds the supplied tables (max_order " + std::to_string(p->P.max_order) + ")");
    if (herr[1] > 0) return fail(PNB_ERR_CUDA, "internal error: a unit classified as far holds a near pair");
    float ms;
    for (int k = 0; k < 2; k++) { cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]); p->timings[k] = ms; }
    if (p->early_boundary && p->bev[0] && cudaEventElapsedTime(&ms, p->bev[0], p->bev[1]) == cudaSuccess) p->timings[1] = ms;
    p->early_boundary = false;
    cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]);
    float ms2;
    cudaEventElapsedTime(&ms2, p->ev[3], p->ev[4]);
    p->timings[2] = ms + ms2;      // reduce + (exchange by the caller) + scatter
    cudaEventElapsedTime(&ms, p->ev[0], p->ev[4]);
    p->timings[3] = ms;
    p->stats[0] = (int64_t)(hcnt[0] + hcnt[1] + hcnt[2]);
    p->stats[3] = (int64_t)hcnt[1];      // pairs of the near evaluator / near pass
    p->stats[4] = (int64_t)hcnt[2];      // pairs of the uniform order-2 units
    if (p->kev[0]) {
        for (int k = 0; k < 4; k++) {
            float kms = 0.f;
            if (cudaEventElapsedTime(&kms, p->kev[k], p->kev[k + 1]) == cudaSuccess) p->ktimings[k] = kms;
            else { cudaGetLastError(); p->ktimings[k] = 0.; }
        }
    }
    p->stats[1] = p->distinct_pairs;
    return 0;
}

// entries of the matrix that receive cell-diagonal blocks (pairs of dofs of one cell): gathered after the scatter
__global__ void gather_entries_kernel(const double *__restrict__ A, int64_t ld, const int2 *__restrict__ ij, int n, double *__restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = A[(size_t)ij[e].x * ld + ij[e].y];
}

// Host output assembled panel by panel (run_group_path): the rows went to the host before the cell-diagonal blocks were
// complete (they sum over all partner cells, i.e. are final only after the last unit).  The entries they touch -- the
// pairs of dofs of one cell, ~7 per row -- are read back after the scatter and written over the copied values.
static int fix_cell_block_entries(pnb_problem *p, const double *dA, int64_t ld, double *A, int64_t host_ld)
{
    const int nvc = p->dim + 1, nc = p->nc;
    std::vector<int2> ij;
    ij.reserve((size_t)nc * nvc * nvc);
    for (int c = 0; c < nc; c++)
        for (int a = 0; a < nvc; a++) {
            const int I = p->h_dofs[(size_t)c * nvc + a];
            if (I < 0) continue;
            for (int b = 0; b < nvc; b++) {
                const int J = p->h_dofs[(size_t)c * nvc + b];
                if (J >= 0) ij.push_back(make_int2(I, J));
            }
        }
    const int n = (int)ij.size();
    if (n == 0) return 0;
    int2 *dij = nullptr;
    double *dv = nullptr;
    CK(pool_malloc((void **)&dij, (size_t)n * sizeof(int2)));
    CK(pool_malloc((void **)&dv, (size_t)n * sizeof(double)));
    std::vector<double> v(n);
    cudaError_t e = cudaMemcpy(dij, ij.data(), (size_t)n * sizeof(int2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        gather_entries_kernel<<<(n + 255) / 256, 256>>>(dA, ld, dij, n, dv);
        e = cudaMemcpy(v.data(), dv, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    pool_free(dij);
    pool_free(dv);
    if (e != cudaSuccess) return fail(PNB_ERR_CUDA, std::string("cell block entries: ") + cudaGetErrorString(e));
    // scattered writes into a large array: a few host threads (duplicates carry the same value)
    const int nt = std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            const int e0 = (int)((int64_t)n * t / nt), e1 = (int)((int64_t)n * (t + 1) / nt);
            for (int q = e0; q < e1; q++) A[(size_t)ij[q].x * host_ld + ij[q].y] = v[q];
        });
    for (auto &x : th) x.join();
    return 0;
}

extern "C" int pnb_dense_assemble(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end, double *A,
                                  int64_t ld, int a_on_device)
{
    if (!p || !A) return fail(PNB_ERR_ARG, "null argument");
    if (row_begin != 0 || row_end != p->N)
        return fail(PNB_ERR_ARG, "pnb_dense_assemble builds the whole operator; row blocks go through pnb_dense_rows_begin/_end");
    if (ld < p->N) return fail(PNB_ERR_ARG, "leading dimension smaller than num_dofs");
    if (p->nblocks > 0 && !a_on_device) return fail(PNB_ERR_UNSUPPORTED, "batched blocks: device output only");
    ON_DEVICE(p->device);
    const int N = p->N;
    double *dA = A;
    if (!a_on_device) {
        // device staging buffer, kept for the life of the problem
        const size_t need = (size_t)N * N * sizeof(double);
        if (p->stage_bytes < need) {
            if (p->stage) pool_free(p->stage);
            p->stage = nullptr;
            p->stage_bytes = 0;
            CK(pool_malloc(&p->stage, need));
            p->stage_bytes = need;
        }
        dA = (double *)p->stage;
    }
    const int64_t dld = a_on_device ? ld : N;
    const bool verbose = getenv("PNB_BENCH_VERBOSE") != nullptr;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t0 = now();
    // pinned host buffer: finished row panels are copied while the assembly continues (cell-group path)
    bool pinned = false;
    if (!a_on_device) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, A) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    int rc = dense_rows_begin_impl(p, zero_exterior, 0, N, dA, dld, pinned ? A : nullptr, ld);
    const double t1 = now();
    if (!rc) rc = pnb_dense_rows_end(p, 0, N, dA, dld);
    const double t2 = now();
    if (rc && p->host_panels && p->copy_stream) cudaStreamSynchronize(p->copy_stream);     // no copy may outlive a failed call
    if (!rc && !a_on_device && p->host_panels) {
        cudaError_t e = cudaStreamSynchronize(p->copy_stream);
        if (e != cudaSuccess) rc = fail(PNB_ERR_CUDA, std::string("copy back: ") + cudaGetErrorString(e));
        else rc = fix_cell_block_entries(p, dA, dld, A, ld);
    } else if (!rc && !a_on_device) {
        cudaError_t e;
        if (ld == N) e = cudaMemcpy(A, dA, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToHost);
        else e = cudaMemcpy2D(A, (size_t)ld * sizeof(double), dA, (size_t)N * sizeof(double), (size_t)N * sizeof(double), N, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(PNB_ERR_CUDA, std::string("copy back: ") + cudaGetErrorString(e));
    }
    if (verbose) fprintf(stderr, "pnb_dense_assemble: begin %.1f ms, end %.1f ms, copy %.1f ms (device timers: tiles %.1f)\n", t1 - t0, t2 - t1, now() - t2, p->timings[0]);
    return rc;
}

extern "C" int pnb_dense_stats(pnb_problem *p, int64_t *stats)
{
    if (!p || !stats) return fail(PNB_ERR_ARG, "null argument");
    memcpy(stats, p->stats, sizeof(p->stats));
    return 0;
}

// 2D group path: device time of the last assembly per kernel: [0] uniform order-2 units, [1] near pair list (first
// assembly) + near evaluator, [2] all other units, [3] symmetrisation
extern "C" int pnb_dense_kernel_timings(pnb_problem *p, double *ms)
{
    if (!p || !ms) return fail(PNB_ERR_ARG, "null argument");
    memcpy(ms, p->ktimings, sizeof(p->ktimings));
    return 0;
}

extern "C" int pnb_dense_timings(pnb_problem *p, double *ms)
{
    if (!p || !ms) return fail(PNB_ERR_ARG, "null argument");
    memcpy(ms, p->timings, sizeof(p->timings));
    return 0;
}

// ---------------------------------------------------------------------------
// H2 far-field kernel blocks (assembleFarFieldInteractions, clusterMethodCy.pyx:2153-2238):
// block b holds -2 gamma(xi_i, xi_j) at the tensor Chebyshev nodes of the two cluster boxes,
// xi = (hi-lo)*0.5*(eta_p+1) + lo with the 1D nodes eta supplied by the host (np.cos as in the
// reference), multi-index with the last dimension fastest (productIterator, :2120-2151).
// One CTA per block, threads over the m1^d x m2^d entries.
// ---------------------------------------------------------------------------
__global__ void farfield_kernel(const PowTab *__restrict__ tab, int dim, const double *__restrict__ boxes1,
                                const double *__restrict__ boxes2, const int *__restrict__ m1s, const int *__restrict__ m2s,
                                const double *__restrict__ eta, const int *__restrict__ eta_ptr,
                                const long long *__restrict__ offsets, double *__restrict__ out)
{
    const int b = blockIdx.x;
    const int m1 = m1s[b], m2 = m2s[b];
    const int n1 = dim == 1 ? m1 : m1 * m1, n2 = dim == 1 ? m2 : m2 * m2;
    const double *e1 = eta + eta_ptr[m1], *e2 = eta + eta_ptr[m2];
    const double *bx = boxes1 + (size_t)b * dim * 2, *by = boxes2 + (size_t)b * dim * 2;
    const PowCtx kv(tab);
    double *o = out + offsets[b];
    for (int e = threadIdx.x; e < n1 * n2; e += blockDim.x) {
        const int i = e / n2, j = e - i * n2;
        double d2 = 0.;
        for (int a = 0; a < dim; a++) {
            const int pi = dim == 1 ? i : (a == 0 ? i / m1 : i % m1);
            const int pj = dim == 1 ? j : (a == 0 ? j / m2 : j % m2);
            const double x = PNB_ADD(PNB_MUL(PNB_MUL(PNB_SUB(bx[2 * a + 1], bx[2 * a]), 0.5), PNB_ADD(e1[pi], 1.0)), bx[2 * a]);
            const double y = PNB_ADD(PNB_MUL(PNB_MUL(PNB_SUB(by[2 * a + 1], by[2 * a]), 0.5), PNB_ADD(e2[pj], 1.0)), by[2 * a]);
            const double t = PNB_SUB(x, y);
            d2 = a == 0 ? PNB_MUL(t, t) : PNB_ADD(d2, PNB_MUL(t, t));
        }
        o[e] = kv(d2) * -2.0;
    }
}

extern "C" int pnb_farfield_blocks(pnb_problem *p, int64_t nblk, const double *boxes1, const double *boxes2,
                                   const int32_t *m1, const int32_t *m2, int32_t max_m, const double *eta,
                                   const int32_t *eta_ptr, const int64_t *offsets, double *out)
{
    if (!p || !boxes1 || !boxes2 || !m1 || !m2 || !eta || !eta_ptr || !offsets || !out) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(p->device);
    if (nblk == 0) return 0;
    const int dim = p->dim;
    const int64_t total = offsets[nblk];
    double *db1 = nullptr, *db2 = nullptr, *deta = nullptr, *dout = nullptr;
    int *dm1 = nullptr, *dm2 = nullptr, *dptr = nullptr;
    long long *doff = nullptr;
    const size_t nb = (size_t)nblk, neta = (size_t)eta_ptr[max_m + 1];
    int rc = 0;
    cudaError_t e = cudaSuccess;
#define FF(call) if (e == cudaSuccess) e = (call)
    FF(cudaMalloc(&db1, nb * dim * 2 * sizeof(double)));
    FF(cudaMalloc(&db2, nb * dim * 2 * sizeof(double)));
    FF(cudaMalloc(&dm1, nb * sizeof(int)));
    FF(cudaMalloc(&dm2, nb * sizeof(int)));
    FF(cudaMalloc(&deta, neta * sizeof(double)));
    FF(cudaMalloc(&dptr, ((size_t)max_m + 2) * sizeof(int)));
    FF(cudaMalloc(&doff, (nb + 1) * sizeof(long long)));
    FF(cudaMalloc(&dout, (size_t)total * sizeof(double)));
    FF(cudaMemcpy(db1, boxes1, nb * dim * 2 * sizeof(double), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(db2, boxes2, nb * dim * 2 * sizeof(double), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(dm1, m1, nb * sizeof(int), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(dm2, m2, nb * sizeof(int), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(deta, eta, neta * sizeof(double), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(dptr, eta_ptr, ((size_t)max_m + 2) * sizeof(int), cudaMemcpyHostToDevice));
    FF(cudaMemcpy(doff, offsets, (nb + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    if (e == cudaSuccess) {
        farfield_kernel<<<(unsigned)nblk, 256>>>(p->P.pow_int, dim, db1, db2, dm1, dm2, deta, dptr, doff, dout);
        e = cudaGetLastError();
    }
    FF(cudaMemcpy(out, dout, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost));
#undef FF
    if (e != cudaSuccess) rc = fail(PNB_ERR_CUDA, std::string("far-field blocks: ") + cudaGetErrorString(e));
    cudaFree(db1); cudaFree(db2); cudaFree(dm1); cudaFree(dm2); cudaFree(deta); cudaFree(dptr); cudaFree(doff); cudaFree(dout);
    return rc;
}

// ---------------------------------------------------------------------------
// dense matvec  y = A x  (row block), HBM-bound: one warp per row, 16-byte loads
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matvec_kernel(const double *__restrict__ A, int64_t nrows, int64_t ncols, int64_t ld,
                                                      const double *__restrict__ x, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t row = warp; row < nrows; row += nwarps) {
        const double *a = A + row * ld;
        double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
        if (vec_ok) {
            const double2 *a2 = reinterpret_cast<const double2 *>(a);
            const double2 *x2 = reinterpret_cast<const double2 *>(x);
            const int64_t n2 = ncols >> 1;
            int64_t j = lane;
            for (; j + 96 < n2; j += 128) {
                const double2 v0 = __ldcs(a2 + j), v1 = __ldcs(a2 + j + 32), v2 = __ldcs(a2 + j + 64), v3 = __ldcs(a2 + j + 96);
                const double2 w0 = x2[j], w1 = x2[j + 32], w2 = x2[j + 64], w3 = x2[j + 96];
                s0 += v0.x * w0.x + v0.y * w0.y;
                s1 += v1.x * w1.x + v1.y * w1.y;
                s2 += v2.x * w2.x + v2.y * w2.y;
                s3 += v3.x * w3.x + v3.y * w3.y;
            }
            for (; j < n2; j += 32) {
                const double2 v0 = __ldcs(a2 + j);
                const double2 w0 = x2[j];
                s0 += v0.x * w0.x + v0.y * w0.y;
            }
            if ((ncols & 1) && lane == 0) s1 += a[ncols - 1] * x[ncols - 1];
        } else {
            for (int64_t j = lane; j < ncols; j += 32) s0 += a[j] * x[j];
        }
        double s = (s0 + s1) + (s2 + s3);
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) y[row] = s;
    }
}

extern "C" int pnb_dense_matvec(int device, const double *A, int64_t num_rows, int64_t num_cols, int64_t ld,
                                const double *x, double *y, void *stream)
{
    if (!A || !x || !y) return fail(PNB_ERR_ARG, "null argument");
    if (pnb_device_count() == 0) return fail(PNB_ERR_NO_DEVICE, "no CUDA device: libpnb200 has no CPU fallback");
    ON_DEVICE(device);
    int sms = 148;
    sms = device_attr(cudaDevAttrMultiProcessorCount, device);
    const int64_t want = (num_rows * 32 + 255) / 256;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    matvec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A, num_rows, num_cols, ld, x, y);
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// FP64 FMA peak microbenchmark
// ---------------------------------------------------------------------------
__global__ void fp64_peak_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int pnb_fp64_peak(int device, double *tflops)
{
    if (!tflops) return fail(PNB_ERR_ARG, "null argument");
    if (pnb_device_count() == 0) return fail(PNB_ERR_NO_DEVICE, "no CUDA device");
    ON_DEVICE(device);
    int sms = 148;
    sms = device_attr(cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double *d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads>>>(d, 1000);
    double best = 0.;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    CK(cudaGetLastError());
    *tflops = best;
    return 0;
}

#include "pnb_h2.cuh"
#include "pnb_element.cuh"
#include "pnb_varorder.cuh"
#include "pnb_krylov.cuh"
