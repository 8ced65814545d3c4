// Dense assembly for fractional orders that VARY INSIDE A CELL: s(x, y) = sFun(x), kernel.piecewise == False
// (singleVariableUnsymmetricFractionalOrder, fractionalOrders.pyx:153-183; smoothedLeftRightFractionalOrder :641-645 is
// the driver's `twoDomainNonSym`, nonlocalProblems.py:95).  Row-owner kernel like pnb_element.cuh.
//
// What the reference does on this path, and what is restated here:
//   * unsymmetric local matrices (fractionalLaplacian1D.pyx:549-603, fractionalLaplacian2D.pyx:1127-1184, distant pairs
//     eval_distant_nonsym nonlocalOperator_{SCALAR}.pxi:849-911) for the ordered pair (K1, K2):
//         a(I, J) = vol sum_q w_q [gamma(x_q, y_q) phi_I(x_q) - gamma(y_q, x_q) phi_I(y_q)] [phi_J(x_q) - phi_J(y_q)],
//     both orientations of every pair are visited (nonlocalAssembly_{SCALAR}.pxi:1412-1428).  For a distant pair the
//     tensor rule is the same in both orientations and the two local matrices coincide term by term: it is evaluated once
//     and counted twice.  The singular rules are not symmetric in their two cells: touching pairs are evaluated in both
//     orientations, each with the permutations of its own getProtoPanelType call.
//   * order, scaling constant and kernel at every quadrature node (updateAndEvalFractional, kernelsCy.pyx:596-622;
//     variableFractionalLaplacianScaling.evalPtr, kernelNormalization.pyx:421-440; fracKernelInfinite*, kernelsCy.pyx:159-183):
//         gamma(x, y) = C(s(x)) |x-y|^(-d-2 s(x)),  C(s) = 2^(2s) s Gamma(s + d/2) / (pi^(d/2) Gamma(1 - s)) / 2
//   * per cell pair the singularity -d - 2 max s over the centres and vertices of both cells (evalParamsOnSimplices,
//     kernelsCy.pyx:1826-1850) picks the regular order (getQuadOrder) and a singular rule of its own
//     (getNearQuadRule, cached per value: fractionalLaplacian1D.pyx:452-547).  The host evaluates s at the centres and
//     vertices, ranks the distinct maxima and builds one set of singular tables per value; a pair takes the larger rank.
//   * surface terms: symmetric boundary local matrices with gamma_b(x, y) = C(s(x)) / s(x) |x-y|^(1-d-2 s(x))
//     (kernels.py:151-160), singularity per (cell, facet) pair from the cell and the facet's centre and vertices.
#pragma once

struct VarOrderDev {
    int fun;                        // PNB_ORDERFUN_*
    double sl, sr, r, slope, interface;
    double Cl, Cr;                  // C(sl), C(sr): most nodes lie outside the transition layer
    int nvals;
    const double *vals;             // distinct pair maxima of s, ascending
    const int *cell_val;            // nc: rank of max s over the cell's centre and vertices
    const int *facet_val;           // nb
    const DRule *rules;             // 5 x nvals: identical | edge | vertex | bedge | bvertex
    const PowTab *pt;               // power tables of the two plateaus: interior sl | sr, boundary sl | sr
    const double *vert_s;           // PNB_ORDERFUN_FE: the order at the mesh vertices (P1 function on the assembly mesh)
};

__device__ __forceinline__ double vo_order_fun(const VarOrderDev &V, double x0, double x1);

// order at a point of cell `cellv` (its vertex ids) with barycentric coordinates lam (the cell's own vertex order)
template <int NV> __device__ __forceinline__ double vo_order(const VarOrderDev &V, double x0, double x1, const int *cellv, const double *lam)
{
    if (V.fun == PNB_ORDERFUN_FE) {
        // feFractionalOrder (fractionalOrders.pyx:660-668): sum_k phi_k(x) u[dof_k], lookupExtended.evalPtr :573-586
        double v = 0.;
#pragma unroll
        for (int k = 0; k < NV; k++) v += lam[k] * V.vert_s[cellv[k]];
        return v;
    }
    return vo_order_fun(V, x0, x1);
}

__device__ __forceinline__ double vo_order_fun(const VarOrderDev &V, double x0, double x1)
{
    // smoothStep / linearStep / smoothStepRadial (fractionalOrders.pyx:389-416, 447-470, 497-535)
    double t = x0;
    if (V.fun == PNB_ORDERFUN_SMOOTHSTEP_RADIAL) t = sqrt(x0 * x0 + x1 * x1);
    if (V.fun == PNB_ORDERFUN_CONST) return V.sl;
    if (t < V.interface - V.r) return V.sl;
    if (t > V.interface + V.r) return V.sr;
    if (V.fun == PNB_ORDERFUN_LINEARSTEP) return V.sl + V.slope * (t - V.interface + V.r);
    const double u = (t - V.interface) * V.slope + 0.5;
    return V.sl + (V.sr - V.sl) * (3.0 * (u * u) - 2.0 * (u * u * u));
}

template <int DIM> __device__ __forceinline__ double vo_scaling(const VarOrderDev &V, double s)
{
    if (s == V.sl) return V.Cl;
    if (s == V.sr) return V.Cr;
    const double ipi = DIM == 1 ? 0.56418958354775628 : 0.31830988618379067;    // pi^(-d/2)
    return exp2(2.0 * s) * s * tgamma(s + 0.5 * DIM) * ipi / tgamma(1.0 - s) * 0.5;
}

// kernel value at squared distance d2 for the order s of the first argument.  Most nodes lie on one of the two plateaus of
// the order: there the power comes from the table of that constant order (pnb_device.cuh), elsewhere from tgamma / pow.
// boundary: C(s)/s |x-y|^(1-d-2s), in 2D divided by |x-y| (the normal factor of the surface form is not normalised)
template <int DIM> __device__ __forceinline__ double vo_kernel(const VarOrderDev &V, double d2, double s, bool boundary)
{
    if (s == V.sl || s == V.sr) {
        const PowCtx kv(V.pt + (boundary ? 2 : 0) + (s == V.sl ? 0 : 1));
        return kv(d2);
    }
    const double C = vo_scaling<DIM>(V, s);
    return boundary ? C / s * pow(d2, (DIM == 2 ? -1. : 0.) - s) : C * pow(d2, -0.5 * DIM - s);
}

// getQuadOrder of the unsymmetric local matrices (fractionalLaplacian2D.pyx:915-934, fractionalLaplacian1D.pyx:431-450)
// and of the boundary ones (:1226-1253, :644-669) for the singularity of ONE pair; smax = max(s) of the pair
__device__ inline int vo_quad_order_interior(const DProblem &P, double smax, double h1, double h2, double d)
{
    const double logdh1 = log(d / h1), logdh2 = log(d / h2);
    const double s = fmax(smax, 0.);
    double p1, p2;
    if (P.dim == 2) {
        const double a1 = fabs(log(h1 / P.H0)), a2 = fabs(log(h2 / P.H0)), am = fmax(a1, a2);
        const double num1 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_int, PNB_MUL(s - 1., a2)), am), PNB_MUL(s, logdh2));
        const double num2 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_int, PNB_MUL(s - 1., a1)), am), PNB_MUL(s, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(fmax(logdh1, 0.), 0.4)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(fmax(logdh2, 0.), 0.4)), 2.);
    } else {
        const double a = 2. * s - 1., b = 2. * s;
        const double num1 = PNB_SUB(PNB_ADD(P.c_int, PNB_MUL(a, fabs(log(h2 / P.H0)))), PNB_MUL(b, logdh2));
        const double num2 = PNB_SUB(PNB_ADD(P.c_int, PNB_MUL(a, fabs(log(h1 / P.H0)))), PNB_MUL(b, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(fmax(logdh1, 0.), 0.8)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(fmax(logdh2, 0.), 0.8)), 2.);
    }
    return (int)fmax(p1, p2);
}

__device__ inline int vo_quad_order_boundary(const DProblem &P, double smax, double h1, double h2, double d)
{
    const double logdh1 = fmax(log(d / h1), 0.), logdh2 = fmax(log(d / h2), 0.);
    double p1, p2;
    if (P.dim == 2) {
        const double s = fmax(smax, 0.);          // 0.5 (-bsing - 1), bsing = -1 - 2 s
        const double a1 = fabs(log(h1 / P.H0)), a2 = fabs(log(h2 / P.H0)), am = fmax(a1, a2);
        const double num1 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_bnd, am), PNB_MUL(s - 1., a2)), PNB_MUL(s, logdh2));
        const double num2 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_bnd, am), PNB_MUL(s - 1., a1)), PNB_MUL(s, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(logdh1, 0.35)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(logdh2, 0.35)), 2.);
    } else {
        const double s = fmax(smax - 0.5, 0.);    // 0.5 (-bsing - 1), bsing = -2 s
        const double a = 2. * s - 1., b = 2. * s;
        const double num1 = PNB_SUB(PNB_ADD(P.c_bnd, PNB_MUL(a, fabs(log(h2 / P.H0)))), PNB_MUL(b, log(d / h2)));
        const double num2 = PNB_SUB(PNB_ADD(P.c_bnd, PNB_MUL(a, fabs(log(h1 / P.H0)))), PNB_MUL(b, log(d / h1)));
        p1 = fmax(ceil(num1 / PNB_ADD(logdh1, 0.8)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(logdh2, 0.8)), 2.);
    }
    return (int)fmax(p1, p2);
}

// row of dof slots (sA in the first cell, sB in the second; -1: the dof is not in that cell) of the unsymmetric local matrix
// of the ORDERED cell pair (cA, cB); acc[0..DPE) columns of the first cell, acc[DPE..2 DPE) of the second; not yet
// multiplied by the volume factor; partial sums of this lane.  vidx = rank of the pair's singularity (singular rules).
template <int DIM, int PORD>
__device__ void vo_pair_row(const DProblem &P, const VarOrderDev &V, int cA, int cB, int panel, int vidx, const int *perm1,
                            const int *perm2, int sA, int sB, int lane, int nlanes, double *acc)
{
    constexpr int NV = DIM + 1, DPE = ElemDims<DIM, PORD>::DPE;
    double t1[3][2], t2[3][2];
    load_simplex<DIM>(P.simplices, cA, NV, t1);
    load_simplex<DIM>(P.simplices, cB, NV, t2);
#pragma unroll
    for (int k = 0; k < 2 * DPE; k++) acc[k] = 0.;
    double s1[3][2], s2[3][2];
    DRule r;
    int nq;
    if (panel >= 1) {
        r = P.reg_cell[panel];
        nq = r.n * r.n;
    } else {
#pragma unroll
        for (int k = 0; k < NV; k++) {
#pragma unroll
            for (int m = 0; m < NV; m++) {
                if (perm1[k] == m) { s1[k][0] = t1[m][0]; s1[k][1] = t1[m][1]; }
                if (perm2[k] == m) { s2[k][0] = t2[m][0]; s2[k][1] = t2[m][1]; }
            }
        }
        const int kind = DIM == 2 ? (panel == -3 ? 0 : (panel == -2 ? 1 : 2)) : (panel == -2 ? 0 : 2);
        r = V.rules[kind * V.nvals + vidx];
        nq = r.n;
    }
    const int n = r.n;
    for (int q = lane; q < nq; q += nlanes) {
        double lx[NV], ly[NV], px[DPE], py[DPE];
        double x0 = 0., x1 = 0., y0 = 0., y1 = 0., w;
        if (panel >= 1) {
            const int i = q / n, j = q - i * n;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                lx[k] = r.bary[k * n + i];
                ly[k] = r.bary[k * n + j];
                x0 = PNB_ADD(x0, PNB_MUL(lx[k], t1[k][0]));
                y0 = PNB_ADD(y0, PNB_MUL(ly[k], t2[k][0]));
                if (DIM == 2) {
                    x1 = PNB_ADD(x1, PNB_MUL(lx[k], t1[k][1]));
                    y1 = PNB_ADD(y1, PNB_MUL(ly[k], t2[k][1]));
                }
            }
            w = r.w[i] * r.w[j];
        } else {
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double bx = r.bary[k * n + q], by = r.bary[(NV + k) * n + q];
                if (k == 0) {
                    x0 = PNB_MUL(s1[k][0], bx);
                    y0 = PNB_MUL(s2[k][0], by);
                    if (DIM == 2) { x1 = PNB_MUL(s1[k][1], bx); y1 = PNB_MUL(s2[k][1], by); }
                } else {
                    x0 = PNB_ADD(x0, PNB_MUL(s1[k][0], bx));
                    y0 = PNB_ADD(y0, PNB_MUL(s2[k][0], by));
                    if (DIM == 2) { x1 = PNB_ADD(x1, PNB_MUL(s1[k][1], bx)); y1 = PNB_ADD(y1, PNB_MUL(s2[k][1], by)); }
                }
#pragma unroll
                for (int m = 0; m < NV; m++) {
                    if (perm1[k] == m) lx[m] = bx;
                    if (perm2[k] == m) ly[m] = by;
                }
            }
            w = r.w[q];
        }
        double d2 = PNB_MUL(x0 - y0, x0 - y0);
        if (DIM == 2) d2 = PNB_ADD(d2, PNB_MUL(x1 - y1, x1 - y1));
        elem_shape<DIM, PORD>(lx, px);
        elem_shape<DIM, PORD>(ly, py);
        double pIx = 0., pIy = 0.;
#pragma unroll
        for (int k = 0; k < DPE; k++) {
            if (k == sA) pIx = px[k];
            if (k == sB) pIy = py[k];
        }
        // a kernel value is only needed where the row's shape function does not vanish
        double tI = 0.;
        if (sA >= 0) {
            const double sx = vo_order<NV>(V, x0, x1, P.cells + (size_t)cA * NV, lx);
            tI = vo_kernel<DIM>(V, d2, sx, false) * pIx;
        }
        if (sB >= 0) {
            const double sy = vo_order<NV>(V, y0, y1, P.cells + (size_t)cB * NV, ly);
            tI -= vo_kernel<DIM>(V, d2, sy, false) * pIy;
        }
        tI *= w;
#pragma unroll
        for (int k = 0; k < DPE; k++) {
            acc[k] = fma(tI, px[k], acc[k]);
            acc[DPE + k] = fma(-tI, py[k], acc[DPE + k]);
        }
    }
}

// row of dof slot sI of the surface-term local matrix of (cell c1, boundary facet f), see elem_boundary_row
template <int DIM, int PORD>
__device__ void vo_boundary_row(const DProblem &P, const VarOrderDev &V, int c1, int f, int panel, int vidx, const int *perm1,
                                const int *perm2, int sI, int lane, double *acc)
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
#pragma unroll
    for (int k = 0; k < DPE; k++) acc[k] = 0.;
    double s1[3][2], s2[3][2];
    DRule r0, r1;
    int nq;
    if (panel >= 1) {
        r0 = P.reg_cell[panel];
        r1 = P.reg_facet[panel];
        nq = r0.n * r1.n;
    } else {
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
            for (int m = 0; m < NV; m++) {
                if (perm1[k] == m) { s1[k][0] = t1[m][0]; s1[k][1] = t1[m][1]; }
                if (k < NF && m < NF && perm2[k] == m) { s2[k][0] = t2[m][0]; s2[k][1] = t2[m][1]; }
            }
        r0 = V.rules[((DIM == 2 && panel == -2) ? 3 : 4) * V.nvals + vidx];
        r1 = r0;
        nq = r0.n;
    }
    for (int q = lane; q < nq; q += 32) {
        double lx[NV], px[DPE];
        double x0 = 0., x1 = 0., y0 = 0., y1 = 0., w, w0, w1;
        if (panel >= 1) {
            const int n0 = r0.n, n1 = r1.n;
            const int i = q / n1, m = q - i * n1;
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
            w = r0.w[i] * r1.w[m];
            w0 = y0 - x0;
            w1 = y1 - x1;
        } else {
            const int n = r0.n;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const double b = r0.bary[k * n + q];
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
                const double b = r0.bary[(NV + k) * n + q];
                if (k == 0) {
                    y0 = PNB_MUL(s2[k][0], b);
                    if (DIM == 2) y1 = PNB_MUL(s2[k][1], b);
                } else {
                    y0 = PNB_ADD(y0, PNB_MUL(s2[k][0], b));
                    if (DIM == 2) y1 = PNB_ADD(y1, PNB_MUL(s2[k][1], b));
                }
            }
            w = r0.w[q];
            w0 = x0 - y0;
            w1 = x1 - y1;
        }
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
        // boundary kernel (2D: divided by |x-y|, the normal factor nw is not normalised)
        const double sx = vo_order<NV>(V, x0, x1, P.cells + (size_t)c1 * NV, lx);
        const double g = w * nw * vo_kernel<DIM>(V, d2, sx, true) * pI;
#pragma unroll
        for (int k = 0; k < DPE; k++) acc[k] = fma(g, px[k], acc[k]);
    }
}

// adds acc[0..DPE) (first-cell columns) and acc[DPE..2 DPE) (second-cell columns), scaled, to the row
template <int DPE>
__device__ __forceinline__ void vo_add_row(const ElemJob &J, double *row, int cA, int cB, const double *acc, double sc, int lane)
{
    double mine = 0.;
#pragma unroll
    for (int k = 0; k < DPE; k++)
        if (k == lane) mine = acc[k];
    if (lane < DPE) {
        const int d = J.edofs[(size_t)cA * DPE + lane];
        if (d >= 0) row[d] += sc * mine;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < DPE; k++)
        if (k == lane) mine = acc[DPE + k];
    if (lane < DPE) {
        const int d = J.edofs[(size_t)cB * DPE + lane];
        if (d >= 0) row[d] += sc * mine;
    }
    __syncwarp();
}

template <int DIM, int PORD>
__global__ void __launch_bounds__(128, 2) varorder_rows_kernel(DProblem P, VarOrderDev V, ElemJob J, int zero_exterior,
                                                               double *__restrict__ A, int64_t ld)
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
        const int v1 = V.cell_val[c1];
        for (int c20 = 0; c20 < J.npartners; c20 += 32) {
            const int c2 = J.partners[c20 + lane];
            int pan = PNB_IGNORED_PANEL, sI2 = -1, vidx = v1;
            if (c2 >= 0) {
#pragma unroll
                for (int k = 0; k < DPE; k++)
                    if (J.edofs[(size_t)c2 * DPE + k] == I) sI2 = k;
                // a pair of two cells around I is visited from its smaller cell only
                const bool skip = c2 != c1 && sI2 >= 0 && c2 < c1;
                if (!skip) {
                    vidx = max(v1, V.cell_val[c2]);
                    const int shared = c1 == c2 ? NV : shared_vertices(P.cells + (size_t)c1 * NV, NV, P.cells + (size_t)c2 * NV, NV);
                    if (shared > 0) pan = -shared;
                    else {
                        const double d = center_distance(P.centers + (size_t)c1 * DIM, P.centers + (size_t)c2 * DIM, DIM);
                        // symmetricCells == False: get_h_simplex of both cells (nonlocalOperator_{SCALAR}.pxi:522-530)
                        pan = vo_quad_order_interior(P, V.vals[vidx], P.hcell[min(c1, c2)], P.hcell[max(c1, c2)], d);
                        if (pan > P.max_order) { atomicMax(J.err, pan); pan = PNB_IGNORED_PANEL; }
                    }
                }
            }
            // distant pairs of low order: one per lane, smaller cell first, counted twice (both orientations coincide)
            const bool mine_far = pan >= 1 && pan <= PNB_ELEM_THREAD_ORDER;
            if (__any_sync(0xffffffffu, mine_far)) {
                double c1side[DPE];
#pragma unroll
                for (int k = 0; k < DPE; k++) c1side[k] = 0.;
                if (mine_far) {
                    const int lo = min(c1, c2), hi = max(c1, c2);
                    int id[3] = {0, 1, 2};
                    double acc[2 * DPE];
                    vo_pair_row<DIM, PORD>(P, V, lo, hi, pan, vidx, id, id, lo == c1 ? sI1 : -1, hi == c1 ? sI1 : -1, 0, 1, acc);
                    const double sc = 2.0 * P.vol[lo] * P.vol[hi];
#pragma unroll
                    for (int k = 0; k < DPE; k++) {
                        c1side[k] = sc * (lo == c1 ? acc[k] : acc[DPE + k]);
                        const double c2side = sc * (lo == c1 ? acc[DPE + k] : acc[k]);
                        const int d = J.edofs[(size_t)c2 * DPE + k];
                        if (d >= 0) row[d] += c2side;
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
                const int vs = __shfl_sync(0xffffffffu, vidx, src);
                const int lo = min(c1, c2s), hi = max(c1, c2s);
                const int sLo = lo == c1 ? sI1 : sI2s, sHi = hi == c1 ? sI1 : sI2s;
                double acc[2 * DPE];
                int q1[3] = {0, 1, 2}, q2[3] = {0, 1, 2};
                if (pans >= 1) {
                    vo_pair_row<DIM, PORD>(P, V, lo, hi, pans, vs, q1, q2, sLo, sHi, lane, 32, acc);
                    warp_allreduce<2 * DPE>(acc);
                    vo_add_row<DPE>(J, row, lo, hi, acc, 2.0 * P.vol[lo] * P.vol[hi], lane);
                } else {
                    const double sc = (DIM == 2 ? 4.0 : 1.0) * P.vol[lo] * P.vol[hi];
                    // first visit: (smaller, larger); second visit after swapCells(): (larger, smaller)
                    proto_panel(P.cells + (size_t)lo * NV, NV, P.cells + (size_t)hi * NV, NV, lo == hi, q1, q2);
                    vo_pair_row<DIM, PORD>(P, V, lo, hi, pans, vs, q1, q2, sLo, sHi, lane, 32, acc);
                    warp_allreduce<2 * DPE>(acc);
                    vo_add_row<DPE>(J, row, lo, hi, acc, sc, lane);
                    if (lo != hi) {
                        proto_panel(P.cells + (size_t)hi * NV, NV, P.cells + (size_t)lo * NV, NV, false, q1, q2);
                        vo_pair_row<DIM, PORD>(P, V, hi, lo, pans, vs, q1, q2, sHi, sLo, lane, 32, acc);
                        warp_allreduce<2 * DPE>(acc);
                        vo_add_row<DPE>(J, row, hi, lo, acc, sc, lane);
                    }
                }
            }
        }
        // ---- Omega x Omega^c: surface terms of the cell with all boundary facets
        if (zero_exterior) {
            for (int f0 = 0; f0 < P.nb; f0 += 32) {
                const int f = f0 + lane;
                int pan = PNB_IGNORED_PANEL, vidx = v1;
                int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
                if (f < P.nb) {
                    vidx = max(v1, V.facet_val[f]);
                    pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.bfacets + (size_t)f * DIM, DIM, false, p1, p2);
                    if (pan == 0) {
                        const double d = center_distance(P.centers + (size_t)c1 * DIM, P.bcenters + (size_t)f * DIM, DIM);
                        pan = vo_quad_order_boundary(P, V.vals[vidx], P.hcell[c1], P.bh[f], d);
                    }
                    if (pan > P.max_order) { atomicMax(J.err, pan); pan = PNB_IGNORED_PANEL; }
                }
                unsigned todo = __ballot_sync(0xffffffffu, pan != PNB_IGNORED_PANEL);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int fs = f0 + src;
                    const int pans = __shfl_sync(0xffffffffu, pan, src), vs = __shfl_sync(0xffffffffu, vidx, src);
                    int q1[3], q2[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        q1[k] = __shfl_sync(0xffffffffu, p1[k], src);
                        q2[k] = __shfl_sync(0xffffffffu, p2[k], src);
                    }
                    double acc[DPE];
                    vo_boundary_row<DIM, PORD>(P, V, c1, fs, pans, vs, q1, q2, sI1, lane, acc);
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

extern "C" int pnb_dense_assemble_varorder(pnb_problem *p, const pnb_varorder_t *order, int polynomial_order, int dofs_per_element,
                                           int num_dofs, const int32_t *dofs, int zero_exterior, double *A_out, int64_t ld_out,
                                           int a_on_device)
{
    if (!p || !order || !dofs || !A_out) return fail(PNB_ERR_ARG, "null argument");
    if (polynomial_order < 0 || polynomial_order > 3 || (polynomial_order == 3 && p->dim != 1))
        return fail(PNB_ERR_UNSUPPORTED, "elements: P0, P1, P2; P3 on intervals");
    const int dpe = polynomial_order == 0 ? 1 : (polynomial_order == 1 ? p->dim + 1 : (polynomial_order == 2 ? (p->dim == 1 ? 3 : 6) : 4));
    if (dofs_per_element != dpe) return fail(PNB_ERR_ARG, "dofs_per_element does not match the element");
    if (p->finite) return fail(PNB_ERR_UNSUPPORTED, "orders varying inside a cell: infinite horizon only");
    if (p->nblocks > 0) return fail(PNB_ERR_UNSUPPORTED, "orders varying inside a cell: no batched blocks");
    if (order->fun < PNB_ORDERFUN_CONST || order->fun > PNB_ORDERFUN_FE) return fail(PNB_ERR_ARG, "unknown order function");
    if (order->fun == PNB_ORDERFUN_FE && !order->vertex_values) return fail(PNB_ERR_ARG, "PNB_ORDERFUN_FE needs vertex_values");
    if (order->num_values <= 0 || !order->values || !order->cell_value || !order->identical || !order->vertex || !order->bvertex ||
        (p->dim == 2 && (!order->edge || !order->bedge)) || (p->nb > 0 && !order->bfacet_value))
        return fail(PNB_ERR_ARG, "incomplete order description");
    if (!(order->sl > 0. && order->sl < 1. && order->sr > 0. && order->sr < 1.)) return fail(PNB_ERR_ARG, "orders must lie in (0, 1)");
    if (ld_out < num_dofs) return fail(PNB_ERR_ARG, "leading dimension too small");
    if (!a_on_device && p->row_nparts > 1) return fail(PNB_ERR_UNSUPPORTED, "row parts: device output only");
    const int nvals = order->num_values;
    for (int c = 0; c < p->nc; c++)
        if (order->cell_value[c] < 0 || order->cell_value[c] >= nvals) return fail(PNB_ERR_ARG, "cell_value out of range");
    for (int f = 0; f < p->nb; f++)
        if (order->bfacet_value[f] < 0 || order->bfacet_value[f] >= nvals) return fail(PNB_ERR_ARG, "bfacet_value out of range");
    ON_DEVICE(p->device);
    if (num_dofs == 0) return 0;
    const int dim = p->dim, nvc = dim + 1;
    // all singular tables in one device buffer
    const pnb_rule_t *kinds[5] = {order->identical, order->edge, order->vertex, order->bedge, order->bvertex};
    const int rows[5] = {2 * nvc, 2 * nvc, 2 * nvc, nvc + dim, nvc + dim};
    std::vector<double> pack;
    std::vector<DRule> hr((size_t)5 * nvals);
    std::vector<size_t> off((size_t)5 * nvals, 0);
    for (int k = 0; k < 5; k++)
        for (int v = 0; v < nvals; v++) {
            DRule &r = hr[(size_t)k * nvals + v];
            r.n = 0; r.rows = rows[k]; r.bary = nullptr; r.w = nullptr;
            if (!kinds[k]) continue;
            const pnb_rule_t &s = kinds[k][v];
            if (s.n <= 0) continue;
            if (s.rows != rows[k] || !s.bary || !s.w) return fail(PNB_ERR_ARG, "singular table with the wrong number of rows");
            r.n = s.n;
            off[(size_t)k * nvals + v] = pack.size();
            pack.insert(pack.end(), s.bary, s.bary + (size_t)s.rows * s.n);
            pack.insert(pack.end(), s.w, s.w + s.n);
        }
    double *d_pack = nullptr, *d_vals = nullptr, *d_vs = nullptr;
    PowTab *d_pt = nullptr;
    DRule *d_rules = nullptr;
    int *d_cv = nullptr, *d_fv = nullptr;
    std::vector<void *> dev;
    double *A = A_out;
    int64_t ld = ld_out;
    auto cleanup = [&]() {
        cudaFree(d_pack); cudaFree(d_vals); cudaFree(d_rules); cudaFree(d_cv); cudaFree(d_fv); cudaFree(d_pt); cudaFree(d_vs);
        elem_job_free(dev);
        if (!a_on_device && A != A_out) pool_free(A);
    };
    if (cudaMalloc(&d_pack, std::max<size_t>(pack.size(), 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_vals, (size_t)nvals * sizeof(double)) != cudaSuccess || cudaMalloc(&d_rules, hr.size() * sizeof(DRule)) != cudaSuccess ||
        cudaMalloc(&d_cv, (size_t)p->nc * sizeof(int)) != cudaSuccess || cudaMalloc(&d_fv, std::max<size_t>(p->nb, 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&d_pt, 4 * sizeof(PowTab)) != cudaSuccess ||
        (order->fun == PNB_ORDERFUN_FE && cudaMalloc(&d_vs, (size_t)p->P.nv * sizeof(double)) != cudaSuccess)) {
        cudaGetLastError();
        cleanup();
        return fail(PNB_ERR_CUDA, "out of device memory");
    }
    for (size_t i = 0; i < hr.size(); i++)
        if (hr[i].n > 0) {
            hr[i].bary = d_pack + off[i];
            hr[i].w = d_pack + off[i] + (size_t)hr[i].rows * hr[i].n;
        }
    cudaMemcpy(d_pack, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_vals, order->values, (size_t)nvals * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_rules, hr.data(), hr.size() * sizeof(DRule), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cv, order->cell_value, (size_t)p->nc * sizeof(int), cudaMemcpyHostToDevice);
    if (p->nb > 0) cudaMemcpy(d_fv, order->bfacet_value, (size_t)p->nb * sizeof(int), cudaMemcpyHostToDevice);
    if (d_vs) cudaMemcpy(d_vs, order->vertex_values, (size_t)p->P.nv * sizeof(double), cudaMemcpyHostToDevice);
    VarOrderDev V;
    V.fun = order->fun;
    V.sl = order->sl; V.sr = order->sr; V.r = order->r; V.slope = order->slope; V.interface = order->interface;
    {
        const double ipi = dim == 1 ? 0.56418958354775628 : 0.31830988618379067;
        V.Cl = exp2(2.0 * V.sl) * V.sl * tgamma(V.sl + 0.5 * dim) * ipi / tgamma(1.0 - V.sl) * 0.5;
        V.Cr = exp2(2.0 * V.sr) * V.sr * tgamma(V.sr + 0.5 * dim) * ipi / tgamma(1.0 - V.sr) * 0.5;
    }
    {
        // power tables of the two plateaus, over the exponent window of the problem's own tables
        std::vector<PowTab> tabs(4);
        const double sv[2] = {V.sl, V.sr}, Cv[2] = {V.Cl, V.Cr};
        for (int k = 0; k < 2; k++) {
            build_powtab(&tabs[k], Cv[k], -0.5 * dim - sv[k], p->pow_eoff);
            build_powtab(&tabs[2 + k], Cv[k] / sv[k], (dim == 2 ? -1. : 0.) - sv[k], p->pow_eoff);
        }
        for (auto &t : tabs) t.horizon2 = INFINITY;
        cudaMemcpy(d_pt, tabs.data(), 4 * sizeof(PowTab), cudaMemcpyHostToDevice);
    }
    V.nvals = nvals; V.vals = d_vals; V.cell_val = d_cv; V.facet_val = d_fv; V.rules = d_rules; V.pt = d_pt; V.vert_s = d_vs;
    if (!a_on_device) {
        ld = num_dofs;
        A = nullptr;
        if (pool_malloc((void **)&A, (size_t)num_dofs * num_dofs * sizeof(double)) != cudaSuccess) {
            A = A_out;
            cudaGetLastError();
            cleanup();
            return fail(PNB_ERR_CUDA, "out of device memory");
        }
    }
    ElemJob J;
    {
        const int rc = elem_job_build(p, dpe, num_dofs, dofs, J, dev);
        if (rc) {
            cleanup();
            return rc;
        }
    }
    const unsigned blocks = (unsigned)std::max<size_t>(((size_t)J.nrows * 32 + 127) / 128, 1);
    if (dim == 2) {
        if (polynomial_order == 2) varorder_rows_kernel<2, 2><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
        else if (polynomial_order == 1) varorder_rows_kernel<2, 1><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
        else varorder_rows_kernel<2, 0><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
    } else {
        if (polynomial_order == 3) varorder_rows_kernel<1, 3><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
        else if (polynomial_order == 2) varorder_rows_kernel<1, 2><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
        else if (polynomial_order == 1) varorder_rows_kernel<1, 1><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
        else varorder_rows_kernel<1, 0><<<blocks, 128>>>(p->P, V, J, zero_exterior, A, ld);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, J.err, sizeof(int), cudaMemcpyDeviceToHost);
    if (!a_on_device && e == cudaSuccess && herr == 0)
        e = cudaMemcpy2D(A_out, (size_t)ld_out * sizeof(double), A, (size_t)ld * sizeof(double), (size_t)num_dofs * sizeof(double),
                         (size_t)num_dofs, cudaMemcpyDeviceToHost);
    cleanup();
    CK(e);
    if (herr > 0) {
        return fail(PNB_ERR_ORDER, "regular quadrature order " + std::to_string(herr) + " exceeds the supplied tables");
    }
    return 0;
}
