// Dense assembly for finite elements other than P1 (P0 and P2 in 1D and 2D; P1 for cross-checking): row-owner kernel.
//
// The reference treats every element through the same local matrices, built from the shape functions of the DoFMap
// (getLocalShapeFunction, fractionalLaplacian2D.pyx:644-813 / fractionalLaplacian1D.pyx:255-339 for the singular PSI
// tables, nonlocalOperator_{SCALAR}.pxi:549-600 for the regular ones, :988-1020 / fractionalLaplacian2D.pyx:1255-1314
// for the surface terms), with (2 dpe)(2 dpe + 1)/2 = 78 local entries for P2 triangles.  Here ONE WARP OWNS ONE ROW I of
// the operator: for every cell c1 around dof I and every cell c2 it evaluates, with the lanes over the quadrature nodes,
// only row I of the local matrix of the pair,
//      a(I, slot) = sum_q w_q gamma(x_q, y_q) psi_I(q) * { phi_slot(x_q) for the slots of the first cell,
//                                                          -phi_slot(y_q) for the slots of the second cell },
// psi_I = phi_I|first(x) - phi_I|second(y), and adds it to the row -- first the slots of the first cell, then those of the
// second (the dofs of one cell are distinct), pair after pair: every entry has one writer and a fixed summation order
// (no atomics, bitwise reproducible).  Dofs shared by the two cells need no special PSI rows: their two parts land on the
// same entry.  Pairs are classified lane-per-partner (32 at a time, partner cells that share no vertex and hence no dof)
// with the same panel / order functions as the P1 path and evaluated in the reference's orientation (smaller cell index
// first) with its permutations of the shared vertices; regular pairs of low order are evaluated one per lane.
// The kernel value is re-evaluated for every row dof of a pair (2 dpe times): this path trades speed for generality and
// is meant for the problem sizes P2 is used at; the P1 production path is pnb_group.cuh.
#pragma once

#define PNB_ELEM_THREAD_ORDER 5    // regular pairs up to this order are evaluated one per lane (<= 49 node pairs)

template <int DIM, int PORD> struct ElemDims {
    static constexpr int NV = DIM + 1;
    static constexpr int DPE = PORD == 0 ? 1 : (PORD == 1 ? NV : (PORD == 2 ? (DIM == 1 ? 3 : 6) : 4));     // P3: intervals only
};

// shape functions in the cell's own vertex order (DoFMaps.pyx:1854-1880 P1, :1932-2005 P2: vertices, then the edges
// (0,1), (1,2), (0,2); 1D: the two vertices, then the cell)
template <int DIM, int PORD> __device__ __forceinline__ void elem_shape(const double *lam, double *phi)
{
    if (PORD == 0) phi[0] = 1.;      // piecewise constants (DoFMaps.pyx:1776-1786)
    else if (PORD == 1) {
#pragma unroll
        for (int k = 0; k <= DIM; k++) phi[k] = lam[k];
    } else if (PORD == 2) {
#pragma unroll
        for (int k = 0; k <= DIM; k++) phi[k] = lam[k] * (2. * lam[k] - 1.);
        phi[DIM + 1] = 4. * lam[0] * lam[1];
        if (DIM == 2) {
            phi[4] = 4. * lam[1] * lam[2];
            phi[5] = 4. * lam[0] * lam[2];
        }
    } else {
        // cubic elements on an interval (DoFMaps.pyx:2034-2078, 2113-2122): the two vertices, then the cell dofs at 1/3, 2/3
        phi[0] = 4.5 * lam[0] * (lam[0] - 1. / 3.) * (lam[0] - 2. / 3.);
        phi[1] = 4.5 * lam[1] * (lam[1] - 1. / 3.) * (lam[1] - 2. / 3.);
        phi[2] = 13.5 * lam[0] * lam[1] * (lam[0] - 1. / 3.);
        phi[3] = 13.5 * lam[1] * lam[0] * (lam[1] - 1. / 3.);
    }
}

// Smooth factor on top of the power law C |x-y|^e of the problem's tables, as a function of |x-y|^2:
//   1  exp(-a |x-y|)        tempered fractional kernels (kernelsCy.pyx:186-213), exponential kernel and its boundary form (:448-477)
//   2  exp(-a |x-y|^2)      Gaussian kernel (:388-415)
//   3  erfc(sqrt(a) |x-y|)  1D boundary form of the Gaussian kernel (:418-430: Gamma(1/2, a r^2) = sqrt(pi) erfc(sqrt(a) r))
//   4  exp(-a |x-y|^2)/|x-y|  2D boundary form of the Gaussian kernel (:433-445: Gamma(1, a r^2) = exp(-a r^2)) over the
//                           table of the boundary kernel divided by |x-y|
__device__ __forceinline__ double elem_smooth(int mode, double a, double d2)
{
    switch (mode) {
    case 1: return exp(-a * sqrt(d2));
    case 2: return exp(-a * d2);
    case 3: return erfc(sqrt(a * d2));
    case 4: return exp(-a * d2) * rsqrt(d2);
    default: return 1.;
    }
}

// row of dof slots (sLo in the first cell, sHi in the second; -1 = the dof is not in that cell) of the local matrix of
// the cell pair (lo, hi), lo <= hi.  acc[0..DPE) first-cell slots, acc[DPE..2 DPE) second-cell slots; NOT yet multiplied
// by the volume factor; partial sums of this lane.
template <int DIM, int PORD>
__device__ void elem_pair_row(const DProblem &P, int lo, int hi, int panel, const int *perm1, const int *perm2, int sLo, int sHi,
                              int lane, int nlanes, double *acc, int smode = 0, double sa = 0.)
{
    constexpr int NV = DIM + 1, DPE = ElemDims<DIM, PORD>::DPE;
    double t1[3][2], t2[3][2];
    load_simplex<DIM>(P.simplices, lo, NV, t1);
    load_simplex<DIM>(P.simplices, hi, NV, t2);
    const PowCtx kv(P.pow_int);
#pragma unroll
    for (int k = 0; k < 2 * DPE; k++) acc[k] = 0.;
    if (panel >= 1) {
        const DRule r = P.reg_cell[panel];
        const int n = r.n;
        for (int q = lane; q < n * n; q += nlanes) {
            const int i = q / n, j = q - i * n;
            double lx[NV], ly[NV], px[DPE], py[DPE];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                lx[k] = r.bary[k * n + i];
                ly[k] = r.bary[k * n + j];
                // un-fused, in the reference's order (nodesInGlobalCoords, quadrature.pyx:76-87)
                x0 = PNB_ADD(x0, PNB_MUL(lx[k], t1[k][0]));
                y0 = PNB_ADD(y0, PNB_MUL(ly[k], t2[k][0]));
                if (DIM == 2) {
                    x1 = PNB_ADD(x1, PNB_MUL(lx[k], t1[k][1]));
                    y1 = PNB_ADD(y1, PNB_MUL(ly[k], t2[k][1]));
                }
            }
            double d2 = PNB_MUL(x0 - y0, x0 - y0);
            if (DIM == 2) d2 = PNB_ADD(d2, PNB_MUL(x1 - y1, x1 - y1));
            elem_shape<DIM, PORD>(lx, px);
            elem_shape<DIM, PORD>(ly, py);
            double psiI = 0.;
#pragma unroll
            for (int k = 0; k < DPE; k++) {
                if (k == sLo) psiI += px[k];
                if (k == sHi) psiI -= py[k];
            }
            const double g = (r.w[i] * r.w[j]) * (smode ? kv(d2) * elem_smooth(smode, sa, d2) : kv(d2)) * psiI;
#pragma unroll
            for (int k = 0; k < DPE; k++) {
                acc[k] = fma(g, px[k], acc[k]);
                acc[DPE + k] = fma(-g, py[k], acc[DPE + k]);
            }
        }
    } else {
        // singular pair: rule nodes in barycentric coordinates over the PERMUTED vertices (shared vertices first)
        double s1[3][2], s2[3][2];
#pragma unroll
        for (int k = 0; k < NV; k++) {
#pragma unroll
            for (int m = 0; m < NV; m++) {
                if (perm1[k] == m) { s1[k][0] = t1[m][0]; s1[k][1] = t1[m][1]; }
                if (perm2[k] == m) { s2[k][0] = t2[m][0]; s2[k][1] = t2[m][1]; }
            }
        }
        DRule r;
        if (DIM == 2) r = panel == -3 ? P.q_id : (panel == -2 ? P.q_edge : P.q_vertex);
        else r = panel == -2 ? P.q_id : P.q_vertex;
        const int n = r.n;
        for (int q = lane; q < n; q += nlanes) {
            double lx[NV], ly[NV], px[DPE], py[DPE];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double bx = r.bary[k * n + q], by = r.bary[(NV + k) * n + q];
                // un-fused and left to right as in fractionalLaplacian2D.pyx:858-863
                if (k == 0) {
                    x0 = PNB_MUL(s1[k][0], bx);
                    y0 = PNB_MUL(s2[k][0], by);
                    if (DIM == 2) { x1 = PNB_MUL(s1[k][1], bx); y1 = PNB_MUL(s2[k][1], by); }
                } else {
                    x0 = PNB_ADD(x0, PNB_MUL(s1[k][0], bx));
                    y0 = PNB_ADD(y0, PNB_MUL(s2[k][0], by));
                    if (DIM == 2) { x1 = PNB_ADD(x1, PNB_MUL(s1[k][1], bx)); y1 = PNB_ADD(y1, PNB_MUL(s2[k][1], by)); }
                }
                // back to the cells' own vertex order
#pragma unroll
                for (int m = 0; m < NV; m++) {
                    if (perm1[k] == m) lx[m] = bx;
                    if (perm2[k] == m) ly[m] = by;
                }
            }
            double d2 = PNB_MUL(x0 - y0, x0 - y0);
            if (DIM == 2) d2 = PNB_ADD(d2, PNB_MUL(x1 - y1, x1 - y1));
            elem_shape<DIM, PORD>(lx, px);
            elem_shape<DIM, PORD>(ly, py);
            double psiI = 0.;
#pragma unroll
            for (int k = 0; k < DPE; k++) {
                if (k == sLo) psiI += px[k];
                if (k == sHi) psiI -= py[k];
            }
            const double g = r.w[q] * (smode ? kv(d2) * elem_smooth(smode, sa, d2) : kv(d2)) * psiI;
#pragma unroll
            for (int k = 0; k < DPE; k++) {
                acc[k] = fma(g, px[k], acc[k]);
                acc[DPE + k] = fma(-g, py[k], acc[DPE + k]);
            }
        }
    }
}

// row of dof slot sI of the surface-term local matrix of (cell c1, boundary facet f)
// (eval_distant_boundary, nonlocalOperator_{SCALAR}.pxi:1069-1108; singular: fractionalLaplacian2D.pyx:1356-1407,
// fractionalLaplacian1D.pyx:753-781); NOT yet multiplied by the volume factor
template <int DIM, int PORD>
__device__ void elem_boundary_row(const DProblem &P, int c1, int f, int panel, const int *perm1, const int *perm2, int sI, int lane,
                                  double *acc, int bmode = 0, double ba = 0.)
{
    constexpr int NV = DIM + 1, NF = DIM, DPE = ElemDims<DIM, PORD>::DPE;
    double t1[3][2], t2[3][2];
    load_simplex<DIM>(P.simplices, c1, NV, t1);
    load_simplex<DIM>(P.bsimplices, f, NF, t2);
    double nx = 0., ny = 0.;
    if (DIM == 2) {
        nx = t2[1][1] - t2[0][1];
        ny = t2[0][0] - t2[1][0];
        const double inv = 1. / sqrt(nx * nx + ny * ny);
        nx *= inv;
        ny *= inv;
    }
    const PowCtx kv(DIM == 2 ? P.pow_bnd_unit : P.pow_bnd);
#pragma unroll
    for (int k = 0; k < DPE; k++) acc[k] = 0.;
    if (panel >= 1) {
        const DRule r0 = P.reg_cell[panel], r1 = P.reg_facet[panel];
        const int n0 = r0.n, n1 = r1.n;
        for (int q = lane; q < n0 * n1; q += 32) {
            const int i = q / n1, m = q - i * n1;
            double lx[NV], px[DPE];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                lx[k] = r0.bary[k * n0 + i];
                x0 = PNB_ADD(x0, PNB_MUL(lx[k], t1[k][0]));
                if (DIM == 2) x1 = PNB_ADD(x1, PNB_MUL(lx[k], t1[k][1]));
            }
#pragma unroll
            for (int k = 0; k < NF; k++) {
                const double b = r1.bary[k * n1 + m];
                y0 = PNB_ADD(y0, PNB_MUL(b, t2[k][0]));
                if (DIM == 2) y1 = PNB_ADD(y1, PNB_MUL(b, t2[k][1]));
            }
            const double w0 = y0 - x0, w1 = y1 - x1;
            double d2 = PNB_MUL(w0, w0);
            double nw = 1.;
            if (DIM == 2) {
                d2 = PNB_ADD(d2, PNB_MUL(w1, w1));
                nw = nx * w0 + ny * w1;
            }
            elem_shape<DIM, PORD>(lx, px);
            double pI = 0.;
#pragma unroll
            for (int k = 0; k < DPE; k++)
                if (k == sI) pI = px[k];
            const double g = (r0.w[i] * r1.w[m]) * nw * (bmode ? kv(d2) * elem_smooth(bmode, ba, d2) : kv(d2)) * pI;
#pragma unroll
            for (int k = 0; k < DPE; k++) acc[k] = fma(g, px[k], acc[k]);
        }
    } else {
        double s1[3][2], s2[3][2];
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
            for (int m = 0; m < NV; m++) {
                if (perm1[k] == m) { s1[k][0] = t1[m][0]; s1[k][1] = t1[m][1]; }
                if (k < NF && m < NF && perm2[k] == m) { s2[k][0] = t2[m][0]; s2[k][1] = t2[m][1]; }
            }
        const DRule r = (DIM == 2 && panel == -2) ? P.bq_edge : P.bq_vertex;
        const int n = r.n;
        for (int q = lane; q < n; q += 32) {
            double lx[NV], px[DPE];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double b = r.bary[k * n + q];
                if (k == 0) {
                    x0 = PNB_MUL(s1[k][0], b);
                    if (DIM == 2) x1 = PNB_MUL(s1[k][1], b);
                } else {
                    x0 = PNB_ADD(x0, PNB_MUL(s1[k][0], b));
                    if (DIM == 2) x1 = PNB_ADD(x1, PNB_MUL(s1[k][1], b));
                }
#pragma unroll
                for (int m = 0; m < NV; m++)
                    if (perm1[k] == m) lx[m] = b;
            }
#pragma unroll
            for (int k = 0; k < NF; k++) {
                const double b = r.bary[(NV + k) * n + q];
                if (k == 0) {
                    y0 = PNB_MUL(s2[k][0], b);
                    if (DIM == 2) y1 = PNB_MUL(s2[k][1], b);
                } else {
                    y0 = PNB_ADD(y0, PNB_MUL(s2[k][0], b));
                    if (DIM == 2) y1 = PNB_ADD(y1, PNB_MUL(s2[k][1], b));
                }
            }
            const double w0 = x0 - y0, w1 = x1 - y1;
            double d2 = PNB_MUL(w0, w0);
            double nw = 1.;
            if (DIM == 2) {
                d2 = PNB_ADD(d2, PNB_MUL(w1, w1));
                nw = nx * w0 + ny * w1;
            }
            elem_shape<DIM, PORD>(lx, px);
            double pI = 0.;
#pragma unroll
            for (int k = 0; k < DPE; k++)
                if (k == sI) pI = px[k];
            const double g = r.w[q] * nw * (bmode ? kv(d2) * elem_smooth(bmode, ba, d2) : kv(d2)) * pI;
#pragma unroll
            for (int k = 0; k < DPE; k++) acc[k] = fma(g, px[k], acc[k]);
        }
    }
}

struct ElemJob {
    int N, dpe;
    const int *edofs;       // nc x dpe: cell -> dof (negative: boundary dof)
    const int *dof_ptr;     // N+1: dof -> (cell, slot) list, cells ascending
    const int *dof_cells;   // cell * 8 + slot
    const int *row_order;   // rows by descending number of cells around the dof (vertex dofs before edge dofs): long rows first
    const int *partners;    // all cells in batches of 32 that share no vertex (-1: padding), colour by colour
    int npartners;          // length of `partners` (a multiple of 32)
    int *err;               // [0]: regular order missing in the tables
    int smode, bmode;       // smooth factors of the interior and the boundary kernel (elem_smooth), 0 = none
    double sa, ba;
    int nrows;              // rows of this launch: length of row_order (all N rows, or the rows of one part)
    const int *row_slot;    // several parts (pnb_problem_set_row_part): global row -> row of the local output; nullptr: identity
};

template <int DIM, int PORD>
__global__ void __launch_bounds__(128, 3) elem_rows_kernel(DProblem P, ElemJob J, int zero_exterior, double *__restrict__ A, int64_t ld)
{
    constexpr int NV = DIM + 1, DPE = ElemDims<DIM, PORD>::DPE;
    const int lane = threadIdx.x & 31;
    const int widx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (widx >= J.nrows) return;
    const int I = J.row_order[widx];
    double *row = A + (size_t)(J.row_slot ? J.row_slot[I] : I) * ld;
    for (int j = lane; j < J.N; j += 32) row[j] = 0.;
    __syncwarp();
    for (int t = J.dof_ptr[I]; t < J.dof_ptr[I + 1]; t++) {
        const int c1 = J.dof_cells[t] >> 3, sI1 = J.dof_cells[t] & 7;
        // ---- Omega x Omega: all partner cells, 32 at a time.  The partners of a step share no vertex, hence no dof: the
        // regular pairs of low order are evaluated one per lane and added to the row without conflicts (the entries of
        // the cell c1 itself, common to all lanes, through a fixed butterfly); singular pairs and high orders follow,
        // one after the other, with the lanes over the quadrature nodes
        for (int c20 = 0; c20 < J.npartners; c20 += 32) {
            const int c2 = J.partners[c20 + lane];
            int pan = PNB_IGNORED_PANEL, sI2 = -1;
            int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
            if (c2 >= 0) {
#pragma unroll
                for (int k = 0; k < DPE; k++)
                    if (J.edofs[(size_t)c2 * DPE + k] == I) sI2 = k;
                // a pair of two cells around I is visited from its smaller cell only
                const bool skip = c2 != c1 && sI2 >= 0 && c2 < c1;
                if (!skip) pan = panel_interior(P, min(c1, c2), max(c1, c2), p1, p2);
                if (pan != PNB_IGNORED_PANEL && pan > P.max_order) { atomicMax(J.err, pan); pan = PNB_IGNORED_PANEL; }
            }
            // regular pairs of order <= PNB_ELEM_THREAD_ORDER: one per lane (c2 does not touch c1, so dof I is not in c2)
            const bool mine_far = pan >= 1 && pan <= PNB_ELEM_THREAD_ORDER;
            if (__any_sync(0xffffffffu, mine_far)) {
                double c1side[DPE], c2side[DPE];
#pragma unroll
                for (int k = 0; k < DPE; k++) c1side[k] = c2side[k] = 0.;
                if (mine_far) {
                    const int lo = min(c1, c2), hi = max(c1, c2);
                    double acc[2 * DPE];
                    elem_pair_row<DIM, PORD>(P, lo, hi, pan, p1, p2, lo == c1 ? sI1 : -1, hi == c1 ? sI1 : -1, 0, 1, acc, J.smode, J.sa);
                    const double sc = 2.0 * P.vol[lo] * P.vol[hi];
#pragma unroll
                    for (int k = 0; k < DPE; k++) {
                        c1side[k] = sc * (lo == c1 ? acc[k] : acc[DPE + k]);
                        c2side[k] = sc * (lo == c1 ? acc[DPE + k] : acc[k]);
                    }
#pragma unroll
                    for (int k = 0; k < DPE; k++) {
                        const int d = J.edofs[(size_t)c2 * DPE + k];
                        if (d >= 0) row[d] += c2side[k];
                    }
                }
                warp_allreduce<DPE>(c1side);
                double mine = 0.;
#pragma unroll
                for (int k = 0; k < DPE; k++)
                    if (k == lane) mine = c1side[k];
                if (lane < DPE) {
                    const int d = J.edofs[(size_t)c1 * DPE + lane];
                    if (d >= 0) row[d] += mine;
                }
                __syncwarp();
            }
            unsigned todo = __ballot_sync(0xffffffffu, pan != PNB_IGNORED_PANEL && !mine_far);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int c2s = __shfl_sync(0xffffffffu, c2, src);
                const int pans = __shfl_sync(0xffffffffu, pan, src), sI2s = __shfl_sync(0xffffffffu, sI2, src);
                int q1[3], q2[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    q1[k] = __shfl_sync(0xffffffffu, p1[k], src);
                    q2[k] = __shfl_sync(0xffffffffu, p2[k], src);
                }
                const int lo = min(c1, c2s), hi = max(c1, c2s);
                const int sLo = lo == c1 ? sI1 : sI2s, sHi = hi == c1 ? sI1 : sI2s;
                double acc[2 * DPE];
                elem_pair_row<DIM, PORD>(P, lo, hi, pans, q1, q2, sLo, sHi, lane, 32, acc, J.smode, J.sa);
                warp_allreduce<2 * DPE>(acc);
                // volume factors: vol1 vol2 (nonlocalOperator_{SCALAR}.pxi:756), 4 vol1 vol2 for the singular 2D rules
                // (fractionalLaplacian2D.pyx:851); off-diagonal pairs count twice (nonlocalAssembly_{SCALAR}.pxi:1404-1410)
                const double sc = (lo == hi ? 1.0 : 2.0) * ((pans < 0 && DIM == 2) ? 4.0 : 1.0) * P.vol[lo] * P.vol[hi];
                double mine = 0.;
#pragma unroll
                for (int k = 0; k < DPE; k++)
                    if (k == lane) mine = acc[k];
                if (lane < DPE) {
                    const int d = J.edofs[(size_t)lo * DPE + lane];
                    if (d >= 0) row[d] += sc * mine;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < DPE; k++)
                    if (k == lane) mine = acc[DPE + k];
                if (lane < DPE) {
                    const int d = J.edofs[(size_t)hi * DPE + lane];
                    if (d >= 0) row[d] += sc * mine;
                }
                __syncwarp();
            }
        }
        // ---- Omega x Omega^c: surface terms of the cell with all boundary facets
        if (zero_exterior) {
            for (int f0 = 0; f0 < P.nb; f0 += 32) {
                const int f = f0 + lane;
                int pan = PNB_IGNORED_PANEL;
                int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
                if (f < P.nb) {
                    pan = panel_boundary(P, c1, f, p1, p2);
                    if (pan > P.max_order) { atomicMax(J.err, pan); pan = PNB_IGNORED_PANEL; }
                }
                unsigned todo = __ballot_sync(0xffffffffu, pan != PNB_IGNORED_PANEL);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int fs = f0 + src;
                    const int pans = __shfl_sync(0xffffffffu, pan, src);
                    int q1[3], q2[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        q1[k] = __shfl_sync(0xffffffffu, p1[k], src);
                        q2[k] = __shfl_sync(0xffffffffu, p2[k], src);
                    }
                    double acc[DPE];
                    elem_boundary_row<DIM, PORD>(P, c1, fs, pans, q1, q2, sI1, lane, acc, J.bmode, J.ba);
                    warp_allreduce<DPE>(acc);
                    const double sc = pans >= 1 ? P.vol[c1] * P.bvol[fs] : (DIM == 2 ? -2.0 * P.vol[c1] * P.bvol[fs] : P.vol[c1]);
                    double mine = 0.;
#pragma unroll
                    for (int k = 0; k < DPE; k++)
                        if (k == lane) mine = acc[k];
                    if (lane < DPE) {
                        const int d = J.edofs[(size_t)c1 * DPE + lane];
                        if (d >= 0) row[d] += sc * mine;
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// host side of ElemJob: dof -> (cell, slot) lists, rows by length, partner cells coloured into vertex-disjoint steps of 32;
// everything is uploaded, `dev` receives the device allocations (release with elem_job_free)
static void elem_job_free(std::vector<void *> &dev)
{
    for (void *d : dev) cudaFree(d);
    dev.clear();
}

// rows of part `part` of `nparts` of the row-owner kernels: the rows sorted by descending number of cells around their dof are
// dealt to the parts in turn (equal shares of long and short rows); `local` keeps that order (long rows first), `slot` maps a
// global row to its position among the part's rows in ASCENDING order (-1: not owned)
static void elem_rows_of_part(const std::vector<int> &row_order, int part, int nparts, std::vector<int> &local, std::vector<int> &slot)
{
    local.clear();
    for (size_t k = (size_t)part; k < row_order.size(); k += (size_t)nparts) local.push_back(row_order[k]);
    std::vector<int> sorted(local);
    std::sort(sorted.begin(), sorted.end());
    slot.assign(row_order.size(), -1);
    for (size_t k = 0; k < sorted.size(); k++) slot[sorted[k]] = (int)k;
}

static int elem_row_order(int nc, int dpe, int num_dofs, const int32_t *dofs, std::vector<int> &row_order)
{
    std::vector<int> cnt(num_dofs, 0);
    for (int c = 0; c < nc; c++)
        for (int m = 0; m < dpe; m++) {
            const int d = dofs[(size_t)c * dpe + m];
            if (d >= num_dofs) return fail(PNB_ERR_ARG, "dof index out of range");
            if (d >= 0) cnt[d]++;
        }
    row_order.resize(num_dofs);
    for (int i = 0; i < num_dofs; i++) row_order[i] = i;
    std::stable_sort(row_order.begin(), row_order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
    return 0;
}

extern "C" int pnb_problem_set_row_part(pnb_problem *p, int32_t part, int32_t nparts)
{
    if (!p) return fail(PNB_ERR_ARG, "null argument");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(PNB_ERR_ARG, "part out of range");
    p->row_part = part;
    p->row_nparts = nparts;
    return 0;
}

extern "C" int pnb_element_rows(pnb_problem *p, int dofs_per_element, int num_dofs, const int32_t *dofs, int32_t part, int32_t nparts,
                                int32_t *rows, int32_t *num_rows)
{
    if (!p) return fail(PNB_ERR_ARG, "null argument");
    return pnb_element_rows_host(p->nc, dofs_per_element, num_dofs, dofs, part, nparts, rows, num_rows);
}

extern "C" int pnb_element_rows_host(int32_t num_cells, int dofs_per_element, int num_dofs, const int32_t *dofs, int32_t part,
                                     int32_t nparts, int32_t *rows, int32_t *num_rows)
{
    if (!dofs || !num_rows || num_cells < 0 || dofs_per_element < 1 || num_dofs < 0) return fail(PNB_ERR_ARG, "bad argument");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(PNB_ERR_ARG, "part out of range");
    std::vector<int> order, local, slot;
    const int rc = elem_row_order(num_cells, dofs_per_element, num_dofs, dofs, order);
    if (rc) return rc;
    elem_rows_of_part(order, part, nparts, local, slot);
    *num_rows = (int32_t)local.size();
    if (rows) {
        std::sort(local.begin(), local.end());
        for (size_t k = 0; k < local.size(); k++) rows[k] = local[k];
    }
    return 0;
}

static int elem_job_build(pnb_problem *p, int dpe, int num_dofs, const int32_t *dofs, ElemJob &J, std::vector<void *> &dev)
{
    const int nc = p->nc;
    std::vector<int> dptr(num_dofs + 1, 0), dcells;
    for (int c = 0; c < nc; c++)
        for (int m = 0; m < dpe; m++) {
            const int d = dofs[(size_t)c * dpe + m];
            if (d >= num_dofs) return fail(PNB_ERR_ARG, "dof index out of range");
            if (d >= 0) dptr[d + 1]++;
        }
    for (int i = 0; i < num_dofs; i++) dptr[i + 1] += dptr[i];
    dcells.resize(dptr[num_dofs]);
    {
        std::vector<int> pos(dptr.begin(), dptr.end() - 1);
        for (int c = 0; c < nc; c++)
            for (int m = 0; m < dpe; m++) {
                const int d = dofs[(size_t)c * dpe + m];
                if (d >= 0) dcells[pos[d]++] = c * 8 + m;
            }
    }
    J.N = num_dofs;
    J.dpe = dpe;
    J.smode = J.bmode = 0;
    J.sa = J.ba = 0.;
    std::vector<int> row_order(num_dofs), row_slot;
    for (int i = 0; i < num_dofs; i++) row_order[i] = i;
    std::stable_sort(row_order.begin(), row_order.end(), [&](int a, int b) { return dptr[a + 1] - dptr[a] > dptr[b + 1] - dptr[b]; });
    if (p->row_nparts > 1) {
        std::vector<int> local;
        elem_rows_of_part(row_order, p->row_part, p->row_nparts, local, row_slot);
        row_order.swap(local);
    }
    J.nrows = (int)row_order.size();
    J.row_slot = nullptr;
    // partner cells in steps of 32 that share no vertex: greedy colouring of the cells (two cells are adjacent when they
    // share a vertex), the cells of a colour in ascending order, every colour padded to a multiple of 32
    std::vector<int> partners;
    {
        const int nvc = p->dim + 1, nv = p->P.nv;
        const int *cells = p->h_cells.data();
        std::vector<int> vptr(nv + 1, 0), vcell((size_t)nc * nvc);
        for (int c = 0; c < nc; c++)
            for (int m = 0; m < nvc; m++) vptr[cells[(size_t)c * nvc + m] + 1]++;
        for (int v = 0; v < nv; v++) vptr[v + 1] += vptr[v];
        {
            std::vector<int> pos(vptr.begin(), vptr.end() - 1);
            for (int c = 0; c < nc; c++)
                for (int m = 0; m < nvc; m++) vcell[pos[cells[(size_t)c * nvc + m]]++] = c;
        }
        std::vector<int> color(nc, -1);
        std::vector<std::vector<int>> byc;
        std::vector<char> used;
        for (int c = 0; c < nc; c++) {
            used.assign(byc.size() + 1, 0);
            for (int m = 0; m < nvc; m++) {
                const int v = cells[(size_t)c * nvc + m];
                for (int e = vptr[v]; e < vptr[v + 1]; e++) {
                    const int k = color[vcell[e]];
                    if (k >= 0) used[k] = 1;
                }
            }
            int k = 0;
            while (used[k]) k++;
            if (k == (int)byc.size()) byc.emplace_back();
            color[c] = k;
            byc[k].push_back(c);
        }
        for (auto &l : byc) {
            partners.insert(partners.end(), l.begin(), l.end());
            while (partners.size() % 32) partners.push_back(-1);
        }
    }
    int *d_edofs = nullptr, *d_ptr = nullptr, *d_cells = nullptr, *d_err = nullptr, *d_order = nullptr, *d_partners = nullptr, *d_slot = nullptr;
    if (cudaMalloc(&d_slot, std::max<size_t>(row_slot.size(), 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&d_edofs, (size_t)nc * dpe * sizeof(int)) != cudaSuccess || cudaMalloc(&d_ptr, ((size_t)num_dofs + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&d_cells, std::max<size_t>(dcells.size(), 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&d_err, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&d_order, std::max<size_t>(row_order.size(), 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&d_partners, std::max<size_t>(partners.size(), 1) * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d_edofs); cudaFree(d_ptr); cudaFree(d_cells); cudaFree(d_err); cudaFree(d_order); cudaFree(d_partners); cudaFree(d_slot);
        return fail(PNB_ERR_CUDA, "out of device memory");
    }
    cudaMemcpy(d_edofs, dofs, (size_t)nc * dpe * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_ptr, dptr.data(), dptr.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cells, dcells.data(), dcells.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(d_err, 0, sizeof(int));
    cudaMemcpy(d_order, row_order.data(), row_order.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_partners, partners.data(), partners.size() * sizeof(int), cudaMemcpyHostToDevice);
    J.edofs = d_edofs; J.dof_ptr = d_ptr; J.dof_cells = d_cells; J.err = d_err; J.row_order = d_order;
    J.partners = d_partners; J.npartners = (int)partners.size();
    if (!row_slot.empty()) {
        cudaMemcpy(d_slot, row_slot.data(), row_slot.size() * sizeof(int), cudaMemcpyHostToDevice);
        J.row_slot = d_slot;
    }
    dev = {d_edofs, d_ptr, d_cells, d_err, d_order, d_partners, d_slot};
    return 0;
}

extern "C" int pnb_dense_assemble_element(pnb_problem *p, int polynomial_order, int dofs_per_element, int num_dofs, const int32_t *dofs,
                                          int zero_exterior, double *A_out, int64_t ld_out, int a_on_device)
{
    return pnb_dense_assemble_element_tempered(p, 0., polynomial_order, dofs_per_element, num_dofs, dofs, zero_exterior, A_out, ld_out,
                                               a_on_device);
}

extern "C" int pnb_dense_assemble_element_tempered(pnb_problem *p, double tempered, int polynomial_order, int dofs_per_element,
                                                   int num_dofs, const int32_t *dofs, int zero_exterior, double *A_out, int64_t ld_out,
                                                   int a_on_device)
{
    if (!(tempered >= 0.) || !(tempered < INFINITY)) return fail(PNB_ERR_ARG, "the tempering rate must be finite and >= 0");
    return pnb_dense_assemble_element_smooth(p, tempered != 0. ? PNB_SMOOTH_EXP_R : PNB_SMOOTH_NONE, tempered, PNB_SMOOTH_NONE, 0.,
                                             polynomial_order, dofs_per_element, num_dofs, dofs, zero_exterior, A_out, ld_out, a_on_device);
}

extern "C" int pnb_dense_assemble_element_smooth(pnb_problem *p, int mode, double a, int bmode, double ba, int polynomial_order,
                                                 int dofs_per_element, int num_dofs, const int32_t *dofs, int zero_exterior,
                                                 double *A_out, int64_t ld_out, int a_on_device)
{
    if (!p || !dofs || !A_out) return fail(PNB_ERR_ARG, "null argument");
    if (mode < PNB_SMOOTH_NONE || mode > PNB_SMOOTH_EXP_R2 || bmode < PNB_SMOOTH_NONE || bmode > PNB_SMOOTH_EXP_R2_OVER_R)
        return fail(PNB_ERR_ARG, "unknown smooth factor");
    if ((mode != PNB_SMOOTH_NONE && !(a >= 0. && a < INFINITY)) || (bmode != PNB_SMOOTH_NONE && !(ba >= 0. && ba < INFINITY)))
        return fail(PNB_ERR_ARG, "the rate of a smooth factor must be finite and >= 0");
    if (polynomial_order < 0 || polynomial_order > 3 || (polynomial_order == 3 && p->dim != 1))
        return fail(PNB_ERR_UNSUPPORTED, "elements: P0, P1, P2; P3 on intervals");
    const int dpe = polynomial_order == 0 ? 1 : (polynomial_order == 1 ? p->dim + 1 : (polynomial_order == 2 ? (p->dim == 1 ? 3 : 6) : 4));
    if (dofs_per_element != dpe) return fail(PNB_ERR_ARG, "dofs_per_element does not match the element");
    if (p->finite) return fail(PNB_ERR_UNSUPPORTED, "elements other than P1: infinite horizon only");
    if (!p->h_labels.empty()) return fail(PNB_ERR_UNSUPPORTED, "elements other than P1: constant kernels only");
    if (p->nblocks > 0) return fail(PNB_ERR_UNSUPPORTED, "elements other than P1: no batched blocks");
    if (ld_out < num_dofs) return fail(PNB_ERR_ARG, "leading dimension too small");
    ON_DEVICE(p->device);
    if (num_dofs == 0) return 0;
    // host output: assembled in a device buffer and copied back
    double *A = A_out;
    int64_t ld = ld_out;
    if (!a_on_device && p->row_nparts > 1) return fail(PNB_ERR_UNSUPPORTED, "row parts: device output only");
    if (!a_on_device) {
        ld = num_dofs;
        CK(pool_malloc((void **)&A, (size_t)num_dofs * num_dofs * sizeof(double)));
    }
    ElemJob J;
    std::vector<void *> dev;
    {
        const int rc = elem_job_build(p, dpe, num_dofs, dofs, J, dev);
        if (rc) {
            if (!a_on_device) pool_free(A);
            return rc;
        }
    }
    J.smode = mode; J.sa = a;
    J.bmode = bmode; J.ba = ba;
    const unsigned blocks = (unsigned)std::max<size_t>(((size_t)J.nrows * 32 + 127) / 128, 1);
    if (p->dim == 2) {
        if (polynomial_order == 2) elem_rows_kernel<2, 2><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
        else if (polynomial_order == 1) elem_rows_kernel<2, 1><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
        else elem_rows_kernel<2, 0><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
    } else {
        if (polynomial_order == 3) elem_rows_kernel<1, 3><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
        else if (polynomial_order == 2) elem_rows_kernel<1, 2><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
        else if (polynomial_order == 1) elem_rows_kernel<1, 1><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
        else elem_rows_kernel<1, 0><<<blocks, 128>>>(p->P, J, zero_exterior, A, ld);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, J.err, sizeof(int), cudaMemcpyDeviceToHost);
    elem_job_free(dev);
    if (!a_on_device) {
        if (e == cudaSuccess && herr == 0)
            e = cudaMemcpy2D(A_out, (size_t)ld_out * sizeof(double), A, (size_t)ld * sizeof(double), (size_t)num_dofs * sizeof(double),
                             (size_t)num_dofs, cudaMemcpyDeviceToHost);
        pool_free(A);
    }
    CK(e);
    if (herr > 0) {
        return fail(PNB_ERR_ORDER, "regular quadrature order " + std::to_string(herr) + " exceeds the supplied tables");
    }
    return 0;
}
