// H2 operator on the device: y = Anear x + sum over admissible cluster pairs V1 K12 V2^T x
// (H2Matrix.matvec, nl/PyNucleus_nl/clusterMethodCy.pyx:2269-2295; upwardPass / downwardPass :1093-1176), and the leaf
// moments V[dof, alpha] = int phi_dof(x) L_alpha(x) dx (tree_node.enterLeafValues, :1205-1325).
//
// The coefficient vectors of all tree nodes are stacked into two device vectors (up, down); node n owns the slots
// [coef_ptr[n], coef_ptr[n+1]).  One matvec = 2 L + 5 launches (L = tree depth), every value has exactly one writer and
// every sum a fixed order: no atomics, bitwise reproducible.
//   csr_matvec_kernel        y = Anear x                                   warp per row, 32-bit column indices
//   h2_leaf_up_kernel        up[leaf] = V^T x[dofs(leaf)]                  CTA per leaf, thread per coefficient
//   h2_transfer_up_kernel    up[parent] = sum_children T_c up[child]       CTA per parent of one level
//   h2_far_kernel            tmp[pair] = K12 up[n2]                        CTA per admissible pair (balanced: a node has 0..30 pairs)
//   h2_far_sum_kernel        down[n1] = sum of its pairs, list order       CTA per node
//   h2_transfer_down_kernel  down[child] += T_c^T down[parent]             CTA per child of one level
//   h2_leaf_down_kernel      y[dofs(leaf)] += V down[leaf]                 CTA per leaf, warp per dof
// The blocks T, K (m^d x m^d) are streamed through shared memory by all threads of the CTA (one round trip per chunk) and
// multiplied from there.  All of it is HBM / L2 bound streaming of the blocks (V, T, K, Anear), each read once per product;
// at N = 12 097 (9 levels) the product is bound by the 23 dependent launches (0.23 ms), at N = 48 769 it takes 0.58 ms
// against 3.33 ms of the dense product (profiles/r2_h2_matvec.txt).
#pragma once

struct pnb_h2 {
    int device = 0;
    int num_dofs = 0, num_nodes = 0, num_leaves = 0, num_levels = 0, ncoef = 0;
    std::vector<void *> allocs;
    // tree
    const int *coef_ptr = nullptr;        // num_nodes + 1
    const int *child_ptr = nullptr, *child_list = nullptr;          // children of a node
    const int *level_ptr = nullptr, *level_nodes = nullptr;         // nodes by level (root = level 0)
    const int *leaf_node = nullptr, *leaf_dof_ptr = nullptr, *leaf_dofs = nullptr;
    const long long *leaf_val_ptr = nullptr;   // num_leaves: start of V[ndofs][mm] (row-major)
    double *leaf_values = nullptr;
    const long long *transfer_ptr = nullptr;   // num_nodes: start of T[mm_parent][mm_node], -1 for the root
    const double *transfer = nullptr;
    const int *far_ptr = nullptr, *far_row = nullptr, *far_col = nullptr;   // CSR over target nodes n1 -> source nodes n2
    const long long *far_tmp = nullptr;                             // start of the pair's product in `tmp`
    double *tmp = nullptr;
    int num_far = 0;
    const long long *far_blk = nullptr;                             // start of K[mm1][mm2] per pair
    const double *far_blocks = nullptr;
    // near field
    long long near_nnz = 0;
    const long long *near_indptr = nullptr;
    const int *near_indices = nullptr;
    const double *near_data = nullptr;
    double *up = nullptr, *down = nullptr;
    const int *parent = nullptr;
    std::vector<int> h_level_ptr;
    int max_mm = 0, max_leaf_dofs = 0;
    // the launch sequence of a product as a CUDA graph over internal input / output vectors (one per flavour: with /
    // without the near field); built on first use
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    cudaStream_t cap_stream = nullptr;
    double *xin = nullptr, *yout = nullptr;
    bool graph_failed = false;
};

template <class T> static int h2_upload(pnb_h2 *h, const T *host, size_t count, const T **dev)
{
    void *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T)));
    h->allocs.push_back(d);
    if (count) CK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)d;
    return 0;
}

__global__ void __launch_bounds__(256) csr_matvec_kernel(int n, const long long *__restrict__ indptr, const int *__restrict__ indices,
                                                         const double *__restrict__ data, const double *__restrict__ x,
                                                         double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
        const long long b = indptr[r], e = indptr[r + 1];
        double s0 = 0., s1 = 0.;
        long long k = b + lane;
        for (; k + 32 < e; k += 64) {
            const double a0 = __ldcs(data + k), a1 = __ldcs(data + k + 32);
            s0 = fma(a0, __ldg(x + indices[k]), s0);
            s1 = fma(a1, __ldg(x + indices[k + 32]), s1);
        }
        if (k < e) s0 = fma(__ldcs(data + k), __ldg(x + indices[k]), s0);
        double s = s0 + s1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) y[r] = s;
    }
}

__global__ void __launch_bounds__(128) h2_leaf_up_kernel(const int *__restrict__ leaf_node, const int *__restrict__ coef_ptr,
                                                         const int *__restrict__ leaf_dof_ptr, const int *__restrict__ leaf_dofs,
                                                         const long long *__restrict__ leaf_val_ptr, const double *__restrict__ V,
                                                         const double *__restrict__ x, double *__restrict__ up)
{
    extern __shared__ double xs[];
    const int l = blockIdx.x, n = leaf_node[l];
    const int c0 = coef_ptr[n], mm = coef_ptr[n + 1] - c0;
    const int d0 = leaf_dof_ptr[l], nd = leaf_dof_ptr[l + 1] - d0;
    const double *Vl = V + leaf_val_ptr[l];
    for (int d = threadIdx.x; d < nd; d += blockDim.x) xs[d] = x[leaf_dofs[d0 + d]];
    __syncthreads();
    for (int a = threadIdx.x; a < mm; a += blockDim.x) {
        double s = 0.;
        for (int d = 0; d < nd; d++) s = fma(Vl[(size_t)d * mm + a], xs[d], s);
        up[c0 + a] = s;
    }
}

// The blocks (T, K) are small (m^d x m^d doubles): a CTA streams a block through shared memory in row chunks with all its
// threads (one round trip to L2 / HBM per chunk instead of one per row), then multiplies from shared memory.
#define PNB_H2_STAGE 4096      // doubles of shared memory per CTA for a chunk of a block
#define PNB_H2_MAXV 1024       // largest vector staged next to it (m^d <= 1024)

// out[i] (+)= sum_j M[i][j] v[j], M rows x cols row-major in global memory, v in shared memory; warp per row
__device__ __forceinline__ void h2_stage_rows(const double *__restrict__ M, int rows, int cols, const double *v, double *out, bool add,
                                              double *stage, int tid, int nt)
{
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int R = max(1, PNB_H2_STAGE / cols);
    for (int r0 = 0; r0 < rows; r0 += R) {
        const int nr = min(R, rows - r0);
        const double *src = M + (size_t)r0 * cols;
        const int cnt = nr * cols;
        if (cols <= PNB_H2_STAGE) {
            for (int e = tid; e < cnt; e += nt) stage[e] = __ldcs(src + e);
            __syncthreads();
        }
        for (int i = warp; i < nr; i += nw) {
            double t = 0.;
            if (cols <= PNB_H2_STAGE) for (int j = lane; j < cols; j += 32) t = fma(stage[i * cols + j], v[j], t);
            else for (int j = lane; j < cols; j += 32) t = fma(src[(size_t)i * cols + j], v[j], t);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) out[r0 + i] = add ? out[r0 + i] + t : t;
        }
        __syncthreads();
    }
}

// parents of one level, CTA per node: up[parent] = sum over the children (list order) of T_c up[child]
__global__ void __launch_bounds__(256) h2_transfer_up_kernel(const int *__restrict__ nodes, const int *__restrict__ coef_ptr,
                                                             const int *__restrict__ child_ptr, const int *__restrict__ child_list,
                                                             const long long *__restrict__ transfer_ptr, const double *__restrict__ T,
                                                             double *up)
{
    __shared__ double stage[PNB_H2_STAGE], v[PNB_H2_MAXV], acc[PNB_H2_MAXV];
    const int n = nodes[blockIdx.x];
    const int cb = child_ptr[n], ce = child_ptr[n + 1];
    if (cb == ce) return;       // leaf: filled by h2_leaf_up_kernel
    const int c0 = coef_ptr[n], mm = coef_ptr[n + 1] - c0;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int q = cb; q < ce; q++) {
        const int c = child_list[q];
        const int cc0 = coef_ptr[c], mc = coef_ptr[c + 1] - cc0;
        for (int j = tid; j < mc; j += nt) v[j] = up[cc0 + j];
        __syncthreads();
        h2_stage_rows(T + transfer_ptr[c], mm, mc, v, acc, q > cb, stage, tid, nt);
    }
    for (int i = tid; i < mm; i += nt) up[c0 + i] = acc[i];
}

// CTA per admissible pair (n1, n2), in the order of the target nodes: tmp[pair] = K12 up[n2]
__global__ void __launch_bounds__(256) h2_far_kernel(const int *__restrict__ coef_ptr, const int *__restrict__ far_row,
                                                     const int *__restrict__ far_col, const long long *__restrict__ far_blk,
                                                     const long long *__restrict__ far_tmp, const double *__restrict__ K,
                                                     const double *__restrict__ up, double *__restrict__ tmp)
{
    __shared__ double stage[PNB_H2_STAGE], v[PNB_H2_MAXV];
    const int q = blockIdx.x;
    const int n1 = far_row[q], n2 = far_col[q];
    const int m1 = coef_ptr[n1 + 1] - coef_ptr[n1], cc0 = coef_ptr[n2], m2 = coef_ptr[n2 + 1] - cc0;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < m2; j += nt) v[j] = up[cc0 + j];
    __syncthreads();
    h2_stage_rows(K + far_blk[q], m1, m2, v, tmp + far_tmp[q], false, stage, tid, nt);
}

// down[n1] = sum of its pairs' vectors in list order; nodes without a pair get zeros
__global__ void __launch_bounds__(128) h2_far_sum_kernel(int num_nodes, const int *__restrict__ coef_ptr, const int *__restrict__ far_ptr,
                                                         const long long *__restrict__ far_tmp, const double *__restrict__ tmp,
                                                         double *__restrict__ down)
{
    const int n = blockIdx.x;
    const int c0 = coef_ptr[n], mm = coef_ptr[n + 1] - c0;
    const int pb = far_ptr[n], pe = far_ptr[n + 1];
    for (int i = threadIdx.x; i < mm; i += blockDim.x) {
        double s = 0.;
        for (int q = pb; q < pe; q++) s += tmp[far_tmp[q] + i];
        down[c0 + i] = s;
    }
}

// nodes of one level (not the root), CTA per node: down[node] += T^T down[parent]; the block is staged in row chunks,
// thread per column
__global__ void __launch_bounds__(256) h2_transfer_down_kernel(const int *__restrict__ nodes, const int *__restrict__ parent,
                                                               const int *__restrict__ coef_ptr,
                                                               const long long *__restrict__ transfer_ptr, const double *__restrict__ T,
                                                               double *down)
{
    __shared__ double stage[PNB_H2_STAGE], v[PNB_H2_MAXV];
    const int n = nodes[blockIdx.x], p = parent[n];
    const int c0 = coef_ptr[n], mm = coef_ptr[n + 1] - c0;
    const int p0 = coef_ptr[p], mp = coef_ptr[p + 1] - p0;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < mp; i += nt) v[i] = down[p0 + i];
    const double *Tn = T + transfer_ptr[n];
    const int R = max(1, PNB_H2_STAGE / mm);
    for (int r0 = 0; r0 < mp; r0 += R) {
        const int nr = min(R, mp - r0);
        __syncthreads();
        if (mm <= PNB_H2_STAGE)
            for (int e = tid; e < nr * mm; e += nt) stage[e] = Tn[(size_t)r0 * mm + e];
        __syncthreads();
        for (int j = tid; j < mm; j += nt) {
            double s = 0.;
            if (mm <= PNB_H2_STAGE) for (int i = 0; i < nr; i++) s = fma(stage[i * mm + j], v[r0 + i], s);
            else for (int i = 0; i < nr; i++) s = fma(Tn[(size_t)(r0 + i) * mm + j], v[r0 + i], s);
            down[c0 + j] += s;
        }
    }
}

// y[dofs(leaf)] += V down[leaf]; warp per dof (every dof belongs to exactly one leaf)
__global__ void __launch_bounds__(256) h2_leaf_down_kernel(const int *__restrict__ leaf_node, const int *__restrict__ coef_ptr,
                                                           const int *__restrict__ leaf_dof_ptr, const int *__restrict__ leaf_dofs,
                                                           const long long *__restrict__ leaf_val_ptr, const double *__restrict__ V,
                                                           const double *__restrict__ down, double *__restrict__ y, int accumulate)
{
    const int l = blockIdx.x, n = leaf_node[l];
    const int c0 = coef_ptr[n], mm = coef_ptr[n + 1] - c0;
    const int d0 = leaf_dof_ptr[l], nd = leaf_dof_ptr[l + 1] - d0;
    const double *Vl = V + leaf_val_ptr[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int d = warp; d < nd; d += nw) {
        double t = 0.;
        for (int a = lane; a < mm; a += 32) t = fma(Vl[(size_t)d * mm + a], down[c0 + a], t);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) {
            const int r = leaf_dofs[d0 + d];
            y[r] = accumulate ? y[r] + t : t;
        }
    }
}

// ---- leaf moments -----------------------------------------------------------------------------------------------
// enterLeafValues (clusterMethodCy.pyx:1205-1325): for the cells around the dofs of a leaf, V[dof, alpha] += vol(K)
// sum_q w_q lambda_k(q) L_alpha(x_q) with the tensor Lagrange polynomials L_alpha on the Chebyshev nodes of the leaf's
// box; the 1D factors prod_{l' != l} (x - xi_l') / beta_l are formed from prefix / suffix products, with the
// reference's special case |x - xi_l| <= 1e-9 -> 1.  CTA per leaf; the cells are visited one after the other
// (ascending), thread alpha owns column alpha of V: one writer per value, fixed order.
struct H2LeafJob {
    const int *leaf_node, *coef_ptr, *leaf_dof_ptr;
    const long long *leaf_val_ptr;
    const int *cell_ptr;        // num_leaves+1: cells around the dofs of the leaf
    const int *cells;           // cell index
    const int *cell_pos;        // (dim+1) per listed cell: local position of its dofs in the leaf, -1 = not in the leaf
    const double *vertices;     // nv x dim
    const int *mesh_cells;      // nc x (dim+1)
    const double *vol;          // nc
    const double *boxes;        // num_leaves x dim x 2
    const int *orders;          // num_leaves: interpolation order m
    const int *rule_n;          // per interpolation order m: rule of order m+2
    const long long *rule_bary_ptr, *rule_w_ptr;
    const double *bary, *w;     // rules back to back: (dim+1) x nq, nq
    const double *eta;          // Chebyshev nodes of all orders back to back (as numpy computed them), eta_ptr[m] = start
    const int *eta_ptr;
    int dim;
};

template <int DIM>
__global__ void __launch_bounds__(128) h2_leaf_values_kernel(H2LeafJob J, double *__restrict__ V)
{
    constexpr int NV = DIM + 1, MAXM = 32;
    extern __shared__ double sh[];
    const int l = blockIdx.x, n = J.leaf_node[l];
    const int m = J.orders[l];
    const int nq = J.rule_n[m];
    const double *__restrict__ bary = J.bary + J.rule_bary_ptr[m], *__restrict__ wq = J.w + J.rule_w_ptr[m];
    const int mm = J.coef_ptr[n + 1] - J.coef_ptr[n];
    const int nd = J.leaf_dof_ptr[l + 1] - J.leaf_dof_ptr[l];
    double *Vl = V + J.leaf_val_ptr[l];
    double *xi = sh;                          // [DIM][m]
    double *beta = xi + DIM * MAXM;           // [DIM][m]
    double *fac = beta + DIM * MAXM;          // [nq][DIM][m]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < nd * mm; e += nt) Vl[e] = 0.;
    for (int e = tid; e < DIM * m; e += nt) {
        const int k = e / m, p = e - k * m;
        const double lo = J.boxes[((size_t)l * DIM + k) * 2], hi = J.boxes[((size_t)l * DIM + k) * 2 + 1];
        // (box[:, 1]-box[:, 0]) * 0.5 * (eta+1) + box[:, 0]   (clusterMethodCy.pyx:1259-1262)
        xi[k * MAXM + p] = (hi - lo) * 0.5 * (J.eta[J.eta_ptr[m] + p] + 1.0) + lo;
    }
    __syncthreads();
    for (int e = tid; e < DIM * m; e += nt) {
        const int k = e / m, p = e - k * m;
        double b = 1.;
        for (int q = 0; q < m; q++)
            if (q != p) b *= xi[k * MAXM + p] - xi[k * MAXM + q];
        beta[k * MAXM + p] = b;
    }
    __syncthreads();
    for (int ci = J.cell_ptr[l]; ci < J.cell_ptr[l + 1]; ci++) {
        const int c = J.cells[ci];
        // 1D Lagrange factors at the quadrature nodes of this cell
        for (int e = tid; e < nq * DIM; e += nt) {
            const int q = e / DIM, k = e - q * DIM;
            double x = 0.;
#pragma unroll
            for (int v = 0; v < NV; v++) x += bary[v * nq + q] * J.vertices[(size_t)J.mesh_cells[(size_t)c * NV + v] * DIM + k];
            double *f = fac + ((size_t)q * DIM + k) * MAXM;
            // prefix products, then suffix products folded in
            double pre = 1.;
            for (int p = 0; p < m; p++) { f[p] = pre; pre *= x - xi[k * MAXM + p]; }
            double suf = 1.;
            for (int p = m - 1; p >= 0; p--) {
                const double d = x - xi[k * MAXM + p];
                const double om = fabs(d) <= 1e-9 ? beta[k * MAXM + p] : f[p] * suf;
                f[p] = om / beta[k * MAXM + p];
                suf *= d;
            }
        }
        __syncthreads();
        const double vol = J.vol[c];
        for (int a = tid; a < mm; a += nt) {
            int idx[DIM];
            if (DIM == 1) idx[0] = a;
            else { idx[0] = a / m; idx[DIM - 1] = a - idx[0] * m; }     // last dimension fastest
            double acc[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] = 0.;
            for (int q = 0; q < nq; q++) {
                double L = 1.;
#pragma unroll
                for (int k = 0; k < DIM; k++) L *= fac[((size_t)q * DIM + k) * MAXM + idx[k]];
                const double wl = wq[q] * L;
#pragma unroll
                for (int v = 0; v < NV; v++) acc[v] = fma(bary[v * nq + q], wl, acc[v]);
            }
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int pos = J.cell_pos[(size_t)ci * NV + v];
                if (pos >= 0) Vl[(size_t)pos * mm + a] += vol * acc[v];
            }
        }
        __syncthreads();
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
extern "C" int pnb_h2_destroy(pnb_h2 *h)
{
    if (!h) return 0;
    {
        DeviceGuard g(h->device);
        for (auto &e : h->gexec) if (e) cudaGraphExecDestroy(e);
        if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
        for (void *d : h->allocs) cudaFree(d);
    }
    delete h;
    return 0;
}

__global__ void h2_narrow_indices_kernel(long long n, const long long *__restrict__ in, int *__restrict__ out)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

static int h2_create_impl(int device, const pnb_h2_desc_t *D, pnb_h2 *h)
{
    const int nn = D->num_nodes, dim = D->dim;
    h->device = device;
    h->num_dofs = D->num_dofs;
    h->num_nodes = nn;
    h->num_leaves = D->num_leaves;
    h->ncoef = D->coef_ptr[nn];
    for (int l = 0; l < D->num_leaves; l++) h->max_leaf_dofs = std::max(h->max_leaf_dofs, D->leaf_dof_ptr[l + 1] - D->leaf_dof_ptr[l]);
    for (int n = 0; n < nn; n++) h->max_mm = std::max(h->max_mm, D->coef_ptr[n + 1] - D->coef_ptr[n]);
    // children lists (node order) and nodes by level
    std::vector<int> child_ptr(nn + 1, 0), child_list, level_ptr, level_nodes(nn);
    int nlev = 0;
    for (int n = 0; n < nn; n++) {
        if (D->parent[n] >= nn || D->level[n] < 0) return fail(PNB_ERR_ARG, "invalid tree");
        if (D->parent[n] >= 0) child_ptr[D->parent[n] + 1]++;
        nlev = std::max(nlev, D->level[n] + 1);
    }
    for (int n = 0; n < nn; n++) child_ptr[n + 1] += child_ptr[n];
    child_list.resize(child_ptr[nn]);
    {
        std::vector<int> fill(child_ptr.begin(), child_ptr.end() - 1);
        for (int n = 0; n < nn; n++)
            if (D->parent[n] >= 0) child_list[fill[D->parent[n]]++] = n;
    }
    level_ptr.assign(nlev + 1, 0);
    for (int n = 0; n < nn; n++) level_ptr[D->level[n] + 1]++;
    for (int l = 0; l < nlev; l++) level_ptr[l + 1] += level_ptr[l];
    {
        std::vector<int> fill(level_ptr.begin(), level_ptr.end() - 1);
        for (int n = 0; n < nn; n++) level_nodes[fill[D->level[n]]++] = n;
    }
    h->num_levels = nlev;
    h->h_level_ptr = level_ptr;
    // far pairs by target node, list order kept inside a node
    std::vector<int> far_ptr(nn + 1, 0), far_col(D->num_far), far_row(D->num_far);
    std::vector<long long> far_blk(D->num_far), far_tmp(D->num_far + 1, 0);
    h->num_far = D->num_far;
    for (int k = 0; k < D->num_far; k++) far_ptr[D->far_n1[k] + 1]++;
    for (int n = 0; n < nn; n++) far_ptr[n + 1] += far_ptr[n];
    {
        std::vector<int> fill(far_ptr.begin(), far_ptr.end() - 1);
        for (int k = 0; k < D->num_far; k++) {
            const int q = fill[D->far_n1[k]]++;
            far_col[q] = D->far_n2[k];
            far_row[q] = D->far_n1[k];
            far_blk[q] = D->far_ptr[k];
        }
        for (int q = 0; q < D->num_far; q++) far_tmp[q + 1] = far_tmp[q] + (D->coef_ptr[far_row[q] + 1] - D->coef_ptr[far_row[q]]);
    }
    if (h->max_mm > PNB_H2_MAXV) return fail(PNB_ERR_UNSUPPORTED, "more than 1024 coefficients per cluster");
    std::vector<long long> leaf_val_ptr(D->num_leaves + 1, 0);
    for (int l = 0; l < D->num_leaves; l++) {
        const int n = D->leaf_node[l];
        leaf_val_ptr[l + 1] = leaf_val_ptr[l] + (long long)(D->leaf_dof_ptr[l + 1] - D->leaf_dof_ptr[l]) * (D->coef_ptr[n + 1] - D->coef_ptr[n]);
    }
    std::vector<long long> tptr(D->transfer_ptr, D->transfer_ptr + nn);
    const int *dparent = nullptr;
    if (h2_upload(h, D->coef_ptr, (size_t)nn + 1, &h->coef_ptr) || h2_upload(h, child_ptr.data(), child_ptr.size(), &h->child_ptr) ||
        h2_upload(h, child_list.data(), child_list.size(), &h->child_list) || h2_upload(h, level_ptr.data(), level_ptr.size(), &h->level_ptr) ||
        h2_upload(h, level_nodes.data(), level_nodes.size(), &h->level_nodes) || h2_upload(h, D->parent, (size_t)nn, &dparent) ||
        h2_upload(h, D->leaf_node, (size_t)D->num_leaves, &h->leaf_node) ||
        h2_upload(h, D->leaf_dof_ptr, (size_t)D->num_leaves + 1, &h->leaf_dof_ptr) ||
        h2_upload(h, D->leaf_dofs, (size_t)D->leaf_dof_ptr[D->num_leaves], &h->leaf_dofs) ||
        h2_upload(h, leaf_val_ptr.data(), leaf_val_ptr.size(), &h->leaf_val_ptr) ||
        h2_upload(h, tptr.data(), tptr.size(), &h->transfer_ptr) || h2_upload(h, D->transfer, (size_t)D->transfer_size, &h->transfer) ||
        h2_upload(h, far_ptr.data(), far_ptr.size(), &h->far_ptr) || h2_upload(h, far_col.data(), far_col.size(), &h->far_col) ||
        h2_upload(h, far_blk.data(), far_blk.size(), &h->far_blk) || h2_upload(h, D->far_blocks, (size_t)D->far_size, &h->far_blocks) ||
        h2_upload(h, far_row.data(), far_row.size(), &h->far_row) || h2_upload(h, far_tmp.data(), far_tmp.size(), &h->far_tmp))
        return PNB_ERR_CUDA;
    h->parent = dparent;
    const size_t nV = (size_t)leaf_val_ptr[D->num_leaves];
    {
        void *d = nullptr;
        CK(cudaMalloc(&d, std::max<size_t>(nV, 1) * sizeof(double)));
        h->allocs.push_back(d);
        h->leaf_values = (double *)d;
        CK(cudaMalloc(&d, std::max<size_t>(h->ncoef, 1) * sizeof(double) * 2));
        h->allocs.push_back(d);
        h->up = (double *)d;
        h->down = h->up + h->ncoef;
        CK(cudaMalloc(&d, std::max<size_t>((size_t)far_tmp[D->num_far], 1) * sizeof(double)));
        h->allocs.push_back(d);
        h->tmp = (double *)d;
    }
    if (D->leaf_values) CK(cudaMemcpy(h->leaf_values, D->leaf_values, nV * sizeof(double), cudaMemcpyHostToDevice));
    else {
        if (!D->leaf_cell_ptr || !D->leaf_cells || !D->leaf_cell_pos || !D->leaf_boxes || !D->leaf_orders || !D->vertices || !D->cells ||
            !D->vol || !D->rule_bary || !D->rule_w || !D->rule_n || !D->rule_bary_ptr || !D->rule_w_ptr || !D->eta || !D->eta_ptr)
            return fail(PNB_ERR_ARG, "leaf moments: mesh, rule and cell lists needed");
        H2LeafJob J;
        const int max_m = D->max_m;
        int max_nq = 0;
        for (int l = 0; l < D->num_leaves; l++) {
            const int m = D->leaf_orders[l];
            if (m < 1 || m > max_m || D->rule_n[m] <= 0) return fail(PNB_ERR_ARG, "leaf moments: no rule for an interpolation order");
            max_nq = std::max(max_nq, D->rule_n[m]);
        }
        if (max_m > 32) return fail(PNB_ERR_UNSUPPORTED, "interpolation order > 32");
        const size_t ncl = (size_t)D->leaf_cell_ptr[D->num_leaves];
        J.leaf_node = h->leaf_node; J.coef_ptr = h->coef_ptr; J.leaf_dof_ptr = h->leaf_dof_ptr; J.leaf_val_ptr = h->leaf_val_ptr;
        J.dim = dim;
        // scratch copies, freed with the handle (small against the blocks)
        if (h2_upload(h, D->leaf_cell_ptr, (size_t)D->num_leaves + 1, &J.cell_ptr) || h2_upload(h, D->leaf_cells, ncl, &J.cells) ||
            h2_upload(h, D->leaf_cell_pos, ncl * (dim + 1), &J.cell_pos) || h2_upload(h, D->vertices, (size_t)D->num_vertices * dim, &J.vertices) ||
            h2_upload(h, D->cells, (size_t)D->num_cells * (dim + 1), &J.mesh_cells) || h2_upload(h, D->vol, (size_t)D->num_cells, &J.vol) ||
            h2_upload(h, D->leaf_boxes, (size_t)D->num_leaves * dim * 2, &J.boxes) || h2_upload(h, D->leaf_orders, (size_t)D->num_leaves, &J.orders) ||
            h2_upload(h, D->rule_bary, (size_t)D->rule_bary_size, &J.bary) || h2_upload(h, D->rule_w, (size_t)D->rule_w_size, &J.w) ||
            h2_upload(h, D->rule_n, (size_t)max_m + 1, &J.rule_n) ||
            h2_upload(h, (const long long *)D->rule_bary_ptr, (size_t)max_m + 1, &J.rule_bary_ptr) ||
            h2_upload(h, (const long long *)D->rule_w_ptr, (size_t)max_m + 1, &J.rule_w_ptr) ||
            h2_upload(h, D->eta, (size_t)D->eta_ptr[max_m + 1], &J.eta) || h2_upload(h, D->eta_ptr, (size_t)max_m + 2, &J.eta_ptr))
            return PNB_ERR_CUDA;
        const size_t smem = (size_t)(2 * dim * 32 + (size_t)max_nq * dim * 32) * sizeof(double);
        if (smem > 200 * 1024) return fail(PNB_ERR_UNSUPPORTED, "leaf moment rule too large");
        if (D->num_leaves > 0) {
            if (dim == 2) {
                CK(cudaFuncSetAttribute(h2_leaf_values_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                h2_leaf_values_kernel<2><<<D->num_leaves, 128, smem>>>(J, h->leaf_values);
            } else {
                CK(cudaFuncSetAttribute(h2_leaf_values_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                h2_leaf_values_kernel<1><<<D->num_leaves, 128, smem>>>(J, h->leaf_values);
            }
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
        }
    }
    if (D->near_indptr) {
        long long nnz = 0;
        CK(cudaMemcpy(&nnz, D->near_indptr + D->num_dofs, sizeof(long long), cudaMemcpyDeviceToHost));
        h->near_nnz = nnz;
        void *d = nullptr;
        CK(cudaMalloc(&d, ((size_t)D->num_dofs + 1) * sizeof(long long)));
        h->allocs.push_back(d);
        CK(cudaMemcpy(d, D->near_indptr, ((size_t)D->num_dofs + 1) * sizeof(long long), cudaMemcpyDeviceToDevice));
        h->near_indptr = (const long long *)d;
        CK(cudaMalloc(&d, std::max<size_t>(nnz, 1) * sizeof(int)));
        h->allocs.push_back(d);
        if (nnz) h2_narrow_indices_kernel<<<(unsigned)((nnz + 255) / 256), 256>>>(nnz, (const long long *)D->near_indices, (int *)d);
        CK(cudaGetLastError());
        h->near_indices = (const int *)d;
        CK(cudaMalloc(&d, std::max<size_t>(nnz, 1) * sizeof(double)));
        h->allocs.push_back(d);
        CK(cudaMemcpy(d, D->near_data, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice));
        h->near_data = (const double *)d;
        CK(cudaDeviceSynchronize());
    }
    return 0;
}

extern "C" int pnb_h2_create(int device, const pnb_h2_desc_t *desc, pnb_h2 **out)
{
    if (!desc || !out) return fail(PNB_ERR_ARG, "null argument");
    if (desc->dim < 1 || desc->dim > 2) return fail(PNB_ERR_UNSUPPORTED, "dim must be 1 or 2");
    if (pnb_device_count() == 0) return fail(PNB_ERR_NO_DEVICE, "no CUDA device: libpnb200 has no CPU fallback");
    ON_DEVICE(device);
    pnb_h2 *h = new pnb_h2;
    const int rc = h2_create_impl(device, desc, h);
    if (rc) { pnb_h2_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int pnb_h2_leaf_values(pnb_h2 *h, double *out)
{
    if (!h || !out) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(h->device);
    long long n = 0;
    CK(cudaMemcpy(&n, h->leaf_val_ptr + h->num_leaves, sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out, h->leaf_values, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// the kernels of one product, in order, on stream st
static void h2_enqueue(pnb_h2 *h, const double *x, double *y, bool near, cudaStream_t st)
{
    if (near) {
        int sms = device_attr(cudaDevAttrMultiProcessorCount, h->device);
        if (sms <= 0) sms = 148;
        const int blocks = std::max(1, std::min((h->num_dofs * 32 + 255) / 256, sms * 8));
        csr_matvec_kernel<<<blocks, 256, 0, st>>>(h->num_dofs, h->near_indptr, h->near_indices, h->near_data, x, y);
    }
    if (h->num_leaves > 0) {
        h2_leaf_up_kernel<<<h->num_leaves, 128, h->max_leaf_dofs * sizeof(double), st>>>(h->leaf_node, h->coef_ptr, h->leaf_dof_ptr, h->leaf_dofs,
                                                                                        h->leaf_val_ptr, h->leaf_values, x, h->up);
        for (int l = h->num_levels - 2; l >= 0; l--) {        // parents, deepest level first
            const int n0 = h->h_level_ptr[l], cnt = h->h_level_ptr[l + 1] - n0;
            if (cnt > 0)
                h2_transfer_up_kernel<<<cnt, 256, 0, st>>>(h->level_nodes + n0, h->coef_ptr, h->child_ptr, h->child_list, h->transfer_ptr,
                                                           h->transfer, h->up);
        }
        if (h->num_far > 0)
            h2_far_kernel<<<h->num_far, 256, 0, st>>>(h->coef_ptr, h->far_row, h->far_col, h->far_blk, h->far_tmp, h->far_blocks, h->up, h->tmp);
        h2_far_sum_kernel<<<h->num_nodes, 128, 0, st>>>(h->num_nodes, h->coef_ptr, h->far_ptr, h->far_tmp, h->tmp, h->down);
        for (int l = 1; l < h->num_levels; l++) {             // root's children first
            const int n0 = h->h_level_ptr[l], cnt = h->h_level_ptr[l + 1] - n0;
            if (cnt > 0)
                h2_transfer_down_kernel<<<cnt, 256, 0, st>>>(h->level_nodes + n0, h->parent, h->coef_ptr, h->transfer_ptr, h->transfer,
                                                             h->down);
        }
        h2_leaf_down_kernel<<<h->num_leaves, 256, 0, st>>>(h->leaf_node, h->coef_ptr, h->leaf_dof_ptr, h->leaf_dofs, h->leaf_val_ptr,
                                                          h->leaf_values, h->down, y, near ? 1 : 0);
    }
}

// captures the launch sequence over the internal vectors xin / yout (2 L + 5 dependent launches: the launch gaps are a third
// of the product at 12k DoFs)
static void h2_build_graph(pnb_h2 *h, bool near)
{
    const int gi = near ? 0 : 1;
    const char *env = getenv("PNB_H2_GRAPH");
    if (env && atoi(env) == 0) { h->graph_failed = true; return; }
    if (!h->xin) {
        void *d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(h->num_dofs, 1) * sizeof(double) * 2) != cudaSuccess) { cudaGetLastError(); h->graph_failed = true; return; }
        h->allocs.push_back(d);
        h->xin = (double *)d;
        h->yout = h->xin + h->num_dofs;
    }
    if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError(); h->graph_failed = true; return;
    }
    device_attr(cudaDevAttrMultiProcessorCount, h->device);      // cached before the capture
    cudaGraph_t g = nullptr;
    bool ok = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        h2_enqueue(h, h->xin, h->yout, near, h->cap_stream);
        ok = cudaStreamEndCapture(h->cap_stream, &g) == cudaSuccess && g != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&h->gexec[gi], g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (!ok) { cudaGetLastError(); h->gexec[gi] = nullptr; h->graph_failed = true; }
}

extern "C" int pnb_h2_matvec(pnb_h2 *h, const double *x, double *y, int far_only, void *stream)
{
    if (!h || !x || !y) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool near = !far_only && h->near_indptr;
    const int gi = near ? 0 : 1;
    if (!h->gexec[gi] && !h->graph_failed) h2_build_graph(h, near);
    if (h->gexec[gi]) {
        const size_t bytes = (size_t)h->num_dofs * sizeof(double);
        CK(cudaMemcpyAsync(h->xin, x, bytes, cudaMemcpyDeviceToDevice, st));
        CK(cudaGraphLaunch(h->gexec[gi], st));
        CK(cudaMemcpyAsync(y, h->yout, bytes, cudaMemcpyDeviceToDevice, st));
    } else h2_enqueue(h, x, y, near, st);
    CK(cudaGetLastError());
    return 0;
}
