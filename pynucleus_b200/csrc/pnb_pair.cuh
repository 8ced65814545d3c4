// Per-pair local-matrix evaluators.
//
//  * lane evaluators (warp-cooperative): every lane integrates a strided subset
//    of the quadrature nodes of ONE cell pair; the caller reduces over the
//    lanes with a fixed butterfly (deterministic).  Used for singular panels
//    and for regular panels of order > PNB_FAR_MAX_ORDER.
//  * far evaluator (thread-per-pair): regular panels of low order, the bulk of
//    all pairs, with the 6x6 local matrix factored into its xx / xy / yy
//    blocks so that each kernel value is used in 5 FMAs instead of 63 flops.
#pragma once
#include "pnb_device.cuh"
#include <type_traits>

template <int DIM> struct PairDims {
    static constexpr int NV = DIM + 1;               // vertices (= P1 dofs) per cell
    static constexpr int NL = (2 * NV) * (2 * NV + 1) / 2;  // local entries, interior
    static constexpr int ND = NV * (NV + 1) / 2;     // entries of one cell-diagonal block
    static constexpr int NX = NV * NV;               // entries of the cross block
};

// gamma(x,y) = C |x-y|^(-d-2s)  (kernelsCy.pyx:159-183); boundary kernels :216-240,
// evaluated from d2 = |x-y|^2 with the table-driven power (PowTab).

// Shared-memory copy of a PowTab with the 128 mantissa entries replicated 8 times: lane l reads copy l & 7, i.e. its
// own group of four banks, so that the 16-byte lookups of a quarter warp never collide (a 128-bit shared load is
// served per quarter warp).  With the compact table the random lookups were the largest source of bank conflicts
// of all three pair kernels (ncu, round 1: 2.3e9 conflict cycles per assembly in the unit kernel alone).
#define PNB_POW_REP 8
struct PowTabS {
    double2 IT[128 * PNB_POW_REP];
    double T1[256];
    double coef[8];
    int eoff, pad;
};

__device__ __forceinline__ void powtab_stage(PowTabS *dst, const PowTab *__restrict__ src, int tid, int nthreads)
{
    for (int e = tid; e < 128 * PNB_POW_REP; e += nthreads) dst->IT[e] = src->IT[e / PNB_POW_REP];
    for (int e = tid; e < 256; e += nthreads) dst->T1[e] = src->T1[e];
    if (tid < 8) dst->coef[tid] = src->coef[tid];
    if (tid == 0) { dst->eoff = src->eoff; dst->pad = 0; }
}

// polynomial coefficients in registers, tables wherever they live: REP = 1 compact table (global memory),
// REP = PNB_POW_REP replicated shared-memory copy.  Degree 6: with |r| <= 2^-8 the truncation error is
// binom(e,7) r^7 <= 1e-16 for the exponents that occur (|e| <= 2.5).
template <int REP> struct PowCtxT {
    const double2 *it;
    const double *T1;
    double c0, c1, c2, c3, c4, c5, c6;
    double horizon2 = INFINITY;
    int eoff;          // table offset minus the exponent bias
    __device__ __forceinline__ void init(const double *coef, int eo)
    {
        eoff = eo - 1023;
        c0 = coef[0]; c1 = coef[1]; c2 = coef[2]; c3 = coef[3];
        c4 = coef[4]; c5 = coef[5]; c6 = coef[6];
    }
    __device__ __forceinline__ explicit PowCtxT(const PowTab *tab) : it(tab->IT), T1(tab->T1)
    {
        static_assert(REP == 1, "compact table");
        init(tab->coef, tab->eoff);
        horizon2 = tab->horizon2;
    }
    __device__ __forceinline__ PowCtxT(const PowTabS *tab, int lane) : it(tab->IT + (lane & (PNB_POW_REP - 1))), T1(tab->T1)
    {
        static_assert(REP == PNB_POW_REP, "replicated table");
        init(tab->coef, tab->eoff);
    }
    __device__ __forceinline__ double operator()(double d2) const
    {
        const int hi = __double2hiint(d2), lo = __double2loint(d2);
        const int E = min(max(((hi >> 20) & 0x7ff) + eoff, 0), 255);
        const int idx = REP == 1 ? ((hi >> 13) & 0x7f) : ((hi >> 10) & (0x7f * REP));
        const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
        const double2 iv = it[idx];
        const double r = fma(m, iv.x, -1.0);
        double p = fma(c6, r, c5);
        p = fma(p, r, c4);
        p = fma(p, r, c3);
        p = fma(p, r, c2);
        p = fma(p, r, c1);
        p = fma(p, r, c0);
        const double v = T1[E] * (iv.y * p);
        // finite horizon: indicator of the interaction ball (kernelsCy.pyx:75-114; compact table only -- the unit
        // kernels of the cell-group path serve the infinite horizon)
        return (REP != 1 || d2 <= horizon2) ? v : 0.;
    }
    // N independent powers, written stage by stage: the polynomial is a chain of seven dependent FMAs, and ptxas
    // interleaved at most two such chains when they came from separate calls (ncu, round 2: the dependent FMAs of the
    // unit kernel carried 4-5x the stall samples of the independent ones).  Stage order keeps N chains in flight.
    template <int N> __device__ __forceinline__ void batch(const double *d2, double *g) const
    {
        double r[N], y[N], t[N], p[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int hi = __double2hiint(d2[j]), lo = __double2loint(d2[j]);
            const int E = min(max((int)((unsigned)hi >> 20) + eoff, 0), 255);      // d2 >= 0: no sign bit
            const int idx = REP == 1 ? ((hi >> 13) & 0x7f) : ((hi >> 10) & (0x7f * REP));
            const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
            const double2 iv = it[idx];
            t[j] = T1[E];
            y[j] = iv.y;
            r[j] = fma(m, iv.x, -1.0);
        }
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(c6, r[j], c5);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], c4);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], c3);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], c2);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], c1);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], c0);
#pragma unroll
        for (int j = 0; j < N; j++) g[j] = t[j] * (y[j] * p[j]);
    }
};
typedef PowCtxT<1> PowCtx;
typedef PowCtxT<PNB_POW_REP> PowCtxS;

__device__ __forceinline__ double kernel_value(const PowTab *__restrict__ t, double d2)
{
    const PowCtx ctx(t);
    return ctx(d2);
}

template <int DIM>
__device__ __forceinline__ void load_simplex(const double *base, size_t idx, int nverts, double (*s)[2])
{
    const double *p = base + idx * (size_t)(nverts * DIM);
#pragma unroll
    for (int k = 0; k < DIM + 1; k++)
        if (k < nverts) {
            s[k][0] = p[k * DIM];
            s[k][1] = DIM == 2 ? p[k * DIM + 1] : 0.;
        }
}

// ---------------------------------------------------------------------------
// Regular element pair, lanes over the n x n tensor nodes
// (eval_distant, nonlocalOperator_{SCALAR}.pxi:756-789).  acc[NL] in the
// reference's flattened upper-triangular order; NOT yet multiplied by vol1*vol2.
// ---------------------------------------------------------------------------
template <int DIM>
__device__ void lanes_regular_interior(const DProblem &P, int c1, int c2, int order, int lane, int nlanes, double *acc)
{
    constexpr int NV = PairDims<DIM>::NV, NL = PairDims<DIM>::NL;
    double s1[3][2], s2[3][2];
    load_simplex<DIM>(P.simplices, c1, NV, s1);
    load_simplex<DIM>(P.simplices, c2, NV, s2);
    const DRule r = P.reg_cell[order];
    const int n = r.n;
    const PowCtx kv(P.pow_int);
#pragma unroll
    for (int k = 0; k < NL; k++) acc[k] = 0.;
    for (int q = lane; q < n * n; q += nlanes) {
        const int i = q / n, j = q - i * n;
        double psi[2 * NV];
        double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double bx = r.bary[k * n + i], by = r.bary[k * n + j];
            psi[k] = bx;
            psi[NV + k] = -by;
            // un-fused, in the reference's order (nodesInGlobalCoords, quadrature.pyx:76-87): the node
            // coordinates and hence x-y are then bit-identical to the reference's, which matters for close
            // pairs where |x-y| << |x|
            x0 = PNB_ADD(x0, PNB_MUL(bx, s1[k][0]));
            y0 = PNB_ADD(y0, PNB_MUL(by, s2[k][0]));
            if (DIM == 2) {
                x1 = PNB_ADD(x1, PNB_MUL(bx, s1[k][1]));
                y1 = PNB_ADD(y1, PNB_MUL(by, s2[k][1]));
            }
        }
        double d2 = PNB_MUL(x0 - y0, x0 - y0);
        if (DIM == 2) d2 = PNB_ADD(d2, PNB_MUL(x1 - y1, x1 - y1));
        const double g = (r.w[i] * r.w[j]) * kv(d2);
        int k = 0;
#pragma unroll
        for (int I = 0; I < 2 * NV; I++) {
            const double t = g * psi[I];
#pragma unroll
            for (int J = I; J < 2 * NV; J++) acc[k++] += t * psi[J];
        }
    }
}

// ---------------------------------------------------------------------------
// Singular element pair (fractionalLaplacian2D.pyx:851-891,
// fractionalLaplacian1D.pyx:378-407).  acc holds the upper triangle of the
// (2NV-1)x(2NV-1) matrix over the PSI rows (rows >= 2*NV-common are zero);
// mapping to the reference's local indices through `perm` is done by the
// caller.  NOT yet multiplied by the volume factor.
// ---------------------------------------------------------------------------
template <int DIM>
__device__ void lanes_singular_interior(const DProblem &P, int c1, int c2, int panel, const int *perm1, const int *perm2,
                                        int lane, int nlanes, double *acc)
{
    constexpr int NV = PairDims<DIM>::NV, NR = 2 * NV - 1, NA = NR * (NR + 1) / 2;
    double s1[3][2], s2[3][2], t1[3][2], t2[3][2];
    load_simplex<DIM>(P.simplices, c1, NV, t1);
    load_simplex<DIM>(P.simplices, c2, NV, t2);
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int m = 0; m < NV; m++) {
            if (perm1[k] == m) { s1[k][0] = t1[m][0]; s1[k][1] = t1[m][1]; }
            if (perm2[k] == m) { s2[k][0] = t2[m][0]; s2[k][1] = t2[m][1]; }
        }
    }
    const int common = -panel;
    DRule r;
    if (DIM == 2) r = panel == -3 ? P.q_id : (panel == -2 ? P.q_edge : P.q_vertex);
    else r = panel == -2 ? P.q_id : P.q_vertex;
    const int n = r.n;
    const PowCtx kv(P.pow_int);
#pragma unroll
    for (int k = 0; k < NA; k++) acc[k] = 0.;
    for (int q = lane; q < n; q += nlanes) {
        double psi[NR];
        double bx[NV], by[NV];
        double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            bx[k] = r.bary[k * n + q];
            by[k] = r.bary[(NV + k) * n + q];
            // un-fused and left to right as in fractionalLaplacian2D.pyx:858-863
            if (k == 0) {
                x0 = PNB_MUL(s1[k][0], bx[k]);
                y0 = PNB_MUL(s2[k][0], by[k]);
                if (DIM == 2) { x1 = PNB_MUL(s1[k][1], bx[k]); y1 = PNB_MUL(s2[k][1], by[k]); }
            } else {
                x0 = PNB_ADD(x0, PNB_MUL(s1[k][0], bx[k]));
                y0 = PNB_ADD(y0, PNB_MUL(s2[k][0], by[k]));
                if (DIM == 2) { x1 = PNB_ADD(x1, PNB_MUL(s1[k][1], bx[k])); y1 = PNB_ADD(y1, PNB_MUL(s2[k][1], by[k])); }
            }
        }
        double d2 = PNB_MUL(x0 - y0, x0 - y0);
        if (DIM == 2) d2 = PNB_ADD(d2, PNB_MUL(x1 - y1, x1 - y1));
        const double g = r.w[q] * kv(d2);
        // PSI rows: shared dofs phi(x)-phi(y); then x-only; then y-only
#pragma unroll
        for (int k = 0; k < NR; k++) psi[k] = 0.;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            if (k < common) psi[k] = bx[k] - by[k];
            else {
                psi[k] = bx[k];
#pragma unroll
                for (int m = NV; m < NR; m++)
                    if (m == NV + k - common) psi[m] = -by[k];
            }
        }
        int k = 0;
#pragma unroll
        for (int I = 0; I < NR; I++) {
            const double t = g * psi[I];
#pragma unroll
            for (int J = I; J < NR; J++) acc[k++] += t * psi[J];
        }
    }
}

// ---------------------------------------------------------------------------
// Element x boundary facet.  Regular: eval_distant_boundary
// (nonlocalOperator_{SCALAR}.pxi:1069-1108); singular: fractionalLaplacian2D.pyx:1356-1407,
// fractionalLaplacian1D.pyx:753-781.  acc[ND] over the (permuted, if singular)
// element dofs; NOT yet multiplied by the volume factor.
// ---------------------------------------------------------------------------
template <int DIM>
__device__ void lanes_boundary(const DProblem &P, int c1, int f, int panel, const int *perm1, const int *perm2,
                               int lane, int nlanes, double *acc)
{
    constexpr int NV = PairDims<DIM>::NV, ND = PairDims<DIM>::ND, NF = DIM;
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
    // 2D: the power carries the 1/|x-y| of the unit vector (no reciprocal square root per node pair); 1D: plain kernel
    const PowCtx kv(DIM == 2 ? P.pow_bnd_unit : P.pow_bnd);
#pragma unroll
    for (int k = 0; k < ND; k++) acc[k] = 0.;
    if (panel >= 1) {
        const DRule r0 = P.reg_cell[panel], r1 = P.reg_facet[panel];
        const int n0 = r0.n, n1 = r1.n;
        for (int q = lane; q < n0 * n1; q += nlanes) {
            const int i = q / n1, m = q - i * n1;
            double phi[NV];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                phi[k] = r0.bary[k * n0 + i];
                x0 = PNB_ADD(x0, PNB_MUL(phi[k], t1[k][0]));
                if (DIM == 2) x1 = PNB_ADD(x1, PNB_MUL(phi[k], t1[k][1]));
            }
#pragma unroll
            for (int k = 0; k < NF; k++) {
                const double b = r1.bary[k * n1 + m];
                y0 = PNB_ADD(y0, PNB_MUL(b, t2[k][0]));
                if (DIM == 2) y1 = PNB_ADD(y1, PNB_MUL(b, t2[k][1]));
            }
            double w0 = y0 - x0, w1 = y1 - x1;
            double d2 = PNB_MUL(w0, w0);
            double nw = 1.;
            if (DIM == 2) {
                d2 = PNB_ADD(d2, PNB_MUL(w1, w1));
                nw = nx * w0 + ny * w1;
            }
            const double g = (r0.w[i] * r1.w[m]) * nw * kv(d2);
            int k = 0;
#pragma unroll
            for (int I = 0; I < NV; I++) {
                const double t = g * phi[I];
#pragma unroll
                for (int J = I; J < NV; J++) acc[k++] += t * phi[J];
            }
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
        for (int q = lane; q < n; q += nlanes) {
            double phi[NV];
            double x0 = 0., x1 = 0., y0 = 0., y1 = 0.;
#pragma unroll
            for (int k = 0; k < NV; k++) {
                phi[k] = r.bary[k * n + q];
                if (k == 0) {
                    x0 = PNB_MUL(s1[k][0], phi[k]);
                    if (DIM == 2) x1 = PNB_MUL(s1[k][1], phi[k]);
                } else {
                    x0 = PNB_ADD(x0, PNB_MUL(s1[k][0], phi[k]));
                    if (DIM == 2) x1 = PNB_ADD(x1, PNB_MUL(s1[k][1], phi[k]));
                }
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
            double w0 = x0 - y0, w1 = x1 - y1;
            double d2 = PNB_MUL(w0, w0);
            double nw = 1.;
            if (DIM == 2) {
                d2 = PNB_ADD(d2, PNB_MUL(w1, w1));
                nw = nx * w0 + ny * w1;
            }
            const double g = r.w[q] * nw * kv(d2);
            int k = 0;
#pragma unroll
            for (int I = 0; I < NV; I++) {
                const double t = g * phi[I];
#pragma unroll
                for (int J = I; J < NV; J++) acc[k++] += t * phi[J];
            }
        }
    }
}

// fixed-shape butterfly: every lane ends with the same, order-independent-of-
// scheduling sum
// ---------------------------------------------------------------------------
// Finite horizon (interaction domain = l2 ball of radius delta; ball2_retriangulation, interactionDomains.pyx:866-965).
// Relative position of two simplices from their vertex distances (:875-898).
// ---------------------------------------------------------------------------
#define PNB_REL_INTERACT 0
#define PNB_REL_REMOTE 1
#define PNB_REL_CUT 2
template <int DIM>
__host__ __device__ inline int relative_position(const double (*s1)[2], int n1, const double (*s2)[2], int n2, double horizon2)
{
    double dmin2 = INFINITY, dmax2 = 0.;
    for (int i = 0; i < n1; i++)
        for (int k = 0; k < n2; k++) {
            double d2 = 0.;
            for (int j = 0; j < DIM; j++) { const double t = PNB_SUB(s1[i][j], s2[k][j]); d2 = PNB_ADD(d2, PNB_MUL(t, t)); }
            dmin2 = fmin(dmin2, d2);
            dmax2 = fmax(dmax2, d2);
        }
    if (dmin2 >= horizon2) return PNB_REL_REMOTE;
    if (dmax2 <= horizon2) return PNB_REL_INTERACT;
    return PNB_REL_CUT;
}

// isInside (:900-909)
template <int DIM> __host__ __device__ inline bool ball_inside(const double *x, const double *y, double horizon2)
{
    double d2 = 0.;
    for (int j = 0; j < DIM; j++) { const double t = PNB_SUB(x[j], y[j]); d2 = PNB_ADD(d2, PNB_MUL(t, t)); }
    return d2 <= horizon2;
}

// findIntersections (:911-937): parameters in [0,1] where the segment start -> end of the simplex meets the sphere around
// x.  `out` keeps its old contents when fewer intersections are found, as in the reference.
template <int DIM>
__host__ __device__ inline int ball_intersections(const double *x, const double (*simplex)[2], int start, int end, double horizon2, double *out)
{
    double nn = 0., p = 0., q = 0.;
    for (int k = 0; k < DIM; k++) {
        const double A = PNB_SUB(simplex[end][k], simplex[start][k]);
        const double B = PNB_SUB(simplex[start][k], x[k]);
        nn = PNB_ADD(nn, PNB_MUL(A, A));
        p = PNB_ADD(p, PNB_MUL(A, B));
        q = PNB_ADD(q, PNB_MUL(B, B));
    }
    nn = 1. / nn;
    p = PNB_MUL(p, PNB_MUL(2., nn));
    q = PNB_MUL(PNB_SUB(q, horizon2), nn);
    const double A = PNB_MUL(-p, 0.5);
    const double B = sqrt(PNB_SUB(PNB_MUL(A, A), q));
    int num = 0;
    double c = PNB_SUB(A, B);
    if (c >= 0. && c <= 1.) out[num++] = c;
    c = PNB_ADD(A, B);
    if (c >= 0. && c <= 1.) out[num++] = c;
    return num;
}

// sub-simplices of the outer element: barycentric map lambda -> A lambda + b, volume factor
struct CutOuter {
    double A[3][3][3], b[3][3], vol[3];
    int n;
};
// sub-simplices of the inner element for one outer node: lambda -> A lambda
struct CutInner {
    double A[3][3][3], vol[3];
    int n;
};

// startLoopSubSimplices_Simplex (:406-566) together with nextSubSimplex_Simplex (:63-96)
template <int DIM>
__host__ __device__ inline void cut_outer(const double (*s1)[2], const double (*s2)[2], double horizon2, CutOuter &o)
{
    o.n = 0;
    for (int t = 0; t < 3; t++) {
        o.vol[t] = 0.;
        for (int i = 0; i < 3; i++) { o.b[t][i] = 0.; for (int j = 0; j < 3; j++) o.A[t][i][j] = 0.; }
    }
    if (DIM == 1) {
        const double horizon = sqrt(horizon2);
        const bool lr = s1[0][0] < s2[0][0];
        const double vol1 = fabs(PNB_SUB(s1[0][0], s1[1][0])), inv = 1. / vol1;
        double iv[4];
        iv[0] = PNB_MUL(s1[0][0], inv);
        iv[3] = PNB_MUL(s1[1][0], inv);
        int k0, k1;
        if (lr) {
            iv[1] = PNB_MUL(fmax(s1[0][0], PNB_SUB(s2[0][0], horizon)), inv);
            iv[2] = PNB_MUL(fmin(s1[1][0], PNB_SUB(s2[1][0], horizon)), inv);
            k0 = 1; k1 = 3;
        } else {
            iv[1] = PNB_MUL(fmax(s1[0][0], PNB_ADD(s2[0][0], horizon)), inv);
            iv[2] = PNB_MUL(fmin(s1[1][0], PNB_ADD(s2[1][0], horizon)), inv);
            k0 = 0; k1 = 2;
        }
        for (int k = k0; k < k1; k++) {
            const double l = iv[k], r = iv[k + 1];
            if (!(PNB_SUB(r, l) > 0.)) continue;
            const int t = o.n++;
            o.A[t][0][0] = PNB_SUB(r, l);
            o.A[t][1][1] = PNB_SUB(r, l);
            o.b[t][0] = PNB_SUB(iv[3], r);
            o.b[t][1] = PNB_SUB(l, iv[0]);
            o.vol[t] = PNB_SUB(r, l);
        }
        return;
    }
    bool inIJ[3][3], inI[3];
    int numInside = 0;
    for (int i = 0; i < 3; i++) {
        bool any = false;
        for (int k = 0; k < 3; k++) { inIJ[i][k] = ball_inside<DIM>(s1[i], s2[k], horizon2); any |= inIJ[i][k]; }
        inI[i] = any;
        numInside += any;
    }
    double isec[2] = {0., 0.};
    if (numInside == 1) {
        int in = 0;
        while (!inI[in]) in++;
        const int o1 = (in + 1) % 3, o2 = (in + 2) % 3;
        double c1 = 0., c2 = 0.;
        for (int j = 0; j < 3; j++)
            if (inIJ[in][j]) {
                ball_intersections<DIM>(s2[j], s1, in, o1, horizon2, isec);
                c1 = fmax(c1, isec[0]);
                ball_intersections<DIM>(s2[j], s1, in, o2, horizon2, isec);
                c2 = fmax(c2, isec[0]);
            }
        if (PNB_MUL(c1, c2) > 0.) {
            o.A[0][in][in] = PNB_ADD(c1, c2);
            o.A[0][in][o1] = c2;
            o.A[0][in][o2] = c1;
            o.A[0][o1][o1] = c1;
            o.A[0][o2][o2] = c2;
            o.b[0][in] = PNB_SUB(PNB_SUB(1., c1), c2);
            o.vol[0] = PNB_MUL(c1, c2);
            o.n = 1;
        }
    } else if (numInside == 2) {
        int out = 0;
        while (inI[out]) out++;
        const int i1 = (out + 1) % 3, i2 = (out + 2) % 3;
        double c1 = 1., c2 = 1.;
        for (int j = 0; j < 3; j++) {
            if (inIJ[i1][j]) { ball_intersections<DIM>(s2[j], s1, out, i1, horizon2, isec); c1 = fmin(c1, isec[0]); }
            if (inIJ[i2][j]) { ball_intersections<DIM>(s2[j], s1, out, i2, horizon2, isec); c2 = fmin(c2, isec[0]); }
        }
        // lengths of the two possible cuts; the reference takes them from the second simplex and leaves d2 un-squared
        // (:513-518) -- kept for parity
        double d1 = 0., d2 = 0.;
        for (int k = 0; k < 2; k++) {
            const double t1 = PNB_SUB(PNB_ADD(s2[out][k], PNB_MUL(c1, PNB_SUB(s2[i1][k], s2[out][k]))), s2[i2][k]);
            d1 = PNB_ADD(d1, PNB_MUL(t1, t1));
            d2 = PNB_ADD(d2, PNB_SUB(PNB_ADD(s2[out][k], PNB_MUL(c2, PNB_SUB(s2[i2][k], s2[out][k]))), s2[i1][k]));
        }
        if (d1 < d2) {
            if (PNB_SUB(1., c1) > 0.) {
                const int t = o.n++;
                o.A[t][out][out] = PNB_SUB(1., c1);
                o.A[t][i1][i1] = PNB_SUB(1., c1);
                o.A[t][i1][i2] = -c1;
                o.A[t][i2][i2] = 1.;
                o.b[t][i1] = c1;
                o.vol[t] = PNB_SUB(1., c1);
            }
            if (PNB_MUL(c1, PNB_SUB(1., c2)) > 0.) {
                const int t = o.n++;
                o.A[t][out][out] = PNB_SUB(1., c2);
                o.A[t][i2][i2] = 1.;
                o.A[t][i2][out] = c2;
                o.A[t][out][i1] = PNB_SUB(1., c1);
                o.A[t][i1][i1] = c1;
                o.vol[t] = PNB_MUL(c1, PNB_SUB(1., c2));
            }
        } else {
            if (PNB_SUB(1., c2) > 0.) {
                const int t = o.n++;
                o.A[t][out][out] = PNB_SUB(1., c2);
                o.A[t][i2][i2] = PNB_SUB(1., c2);
                o.A[t][i2][i1] = -c2;
                o.A[t][i1][i1] = 1.;
                o.b[t][i2] = c2;
                o.vol[t] = PNB_SUB(1., c2);
            }
            if (PNB_MUL(c2, PNB_SUB(1., c1)) > 0.) {
                const int t = o.n++;
                o.A[t][out][out] = PNB_SUB(1., c1);
                o.A[t][i1][i1] = 1.;
                o.A[t][i1][out] = c1;
                o.A[t][out][i2] = PNB_SUB(1., c2);
                o.A[t][i2][i2] = c2;
                o.vol[t] = PNB_MUL(c2, PNB_SUB(1., c1));
            }
        }
    } else if (numInside == 3) {
        o.A[0][0][0] = o.A[0][1][1] = o.A[0][2][2] = 1.;
        o.vol[0] = 1.;
        o.n = 1;
    }
    // numInside == 0 cannot happen for a CUT pair (the reference raises NotImplementedError)
}

// startLoopSubSimplices_Node (:568-826) with nextSubSimplex_Node (:101-112); the l2 ball has no special points
template <int DIM>
__host__ __device__ inline void cut_inner(const double *x, const double (*s2)[2], double horizon2, CutInner &o)
{
    o.n = 0;
    for (int t = 0; t < 3; t++) {
        o.vol[t] = 0.;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.A[t][i][j] = 0.;
    }
    bool ind[3] = {false, false, false};
    int numInside = 0;
    for (int j = 0; j < DIM + 1; j++) { ind[j] = ball_inside<DIM>(x, s2[j], horizon2); numInside += ind[j]; }
    double isec[2] = {0., 0.};
    if (DIM == 1) {
        if (numInside == 0) {
            if (ball_intersections<DIM>(x, s2, 0, 1, horizon2, isec) == 2) {
                o.A[0][0][0] = PNB_SUB(1., isec[0]);
                o.A[0][1][0] = isec[0];
                o.A[0][1][1] = isec[1];
                o.A[0][0][1] = PNB_SUB(1., isec[1]);
                o.vol[0] = PNB_SUB(isec[1], isec[0]);
                o.n = 1;
            }
        } else if (numInside == 1) {
            int in = 0;
            while (!ind[in]) in++;
            const int out = (in + 1) % 2;
            ball_intersections<DIM>(x, s2, in, out, horizon2, isec);
            o.A[0][in][in] = 1.;
            o.A[0][out][out] = isec[0];
            o.A[0][in][out] = PNB_SUB(1., isec[0]);
            o.vol[0] = isec[0];
            o.n = 1;
        } else {
            o.A[0][0][0] = o.A[0][1][1] = 1.;
            o.vol[0] = 1.;
            o.n = 1;
        }
        return;
    }
    if (numInside == 1) {
        int in = 0;
        while (!ind[in]) in++;
        const int o1 = (in + 1) % 3, o2 = (in + 2) % 3;
        ball_intersections<DIM>(x, s2, in, o1, horizon2, isec);
        const double c1 = isec[0];
        ball_intersections<DIM>(x, s2, in, o2, horizon2, isec);
        const double c2 = isec[0];
        const int ni = ball_intersections<DIM>(x, s2, o1, o2, horizon2, isec);
        if (ni == 0) {
            o.A[0][in][in] = 1.;
            o.A[0][in][o1] = PNB_SUB(1., c1);
            o.A[0][o1][o1] = c1;
            o.A[0][o2][o2] = c2;
            o.A[0][in][o2] = PNB_SUB(1., c2);
            o.vol[0] = PNB_MUL(c1, c2);
            o.n = 1;
        } else if (ni == 2) {
            o.A[0][in][in] = 1.;
            o.A[0][o1][o1] = c1;
            o.A[0][in][o1] = PNB_SUB(1., c1);
            o.A[0][o2][o2] = isec[0];
            o.A[0][o1][o2] = PNB_SUB(1., isec[0]);
            o.vol[0] = PNB_MUL(c1, isec[0]);
            o.A[1][in][in] = 1.;
            o.A[1][o1][o1] = PNB_SUB(1., isec[0]);
            o.A[1][o2][o1] = isec[0];
            o.A[1][o1][o2] = PNB_SUB(1., isec[1]);
            o.A[1][o2][o2] = isec[1];
            o.vol[1] = PNB_SUB(isec[1], isec[0]);
            o.A[2][in][in] = 1.;
            o.A[2][o1][o1] = PNB_SUB(1., isec[1]);
            o.A[2][o2][o1] = isec[1];
            o.A[2][o2][o2] = c2;
            o.A[2][in][o2] = PNB_SUB(1., c2);
            o.vol[2] = PNB_MUL(c2, PNB_SUB(1., isec[1]));
            o.n = 3;
        } else {
            o.A[0][in][in] = 1.;
            o.A[0][o1][o1] = c1;
            o.A[0][in][o1] = PNB_SUB(1., c1);
            o.A[0][o2][o2] = isec[0];
            o.A[0][o1][o2] = PNB_SUB(1., isec[0]);
            o.vol[0] = PNB_MUL(c1, isec[0]);
            o.A[1][in][in] = 1.;
            o.A[1][o1][o1] = PNB_SUB(1., isec[0]);
            o.A[1][o2][o1] = isec[0];
            o.A[1][o2][o2] = c2;
            o.A[1][in][o2] = PNB_SUB(1., c2);
            o.vol[1] = PNB_MUL(c2, PNB_SUB(1., isec[0]));
            o.n = 2;
        }
    } else if (numInside == 2) {
        int out = 0;
        while (ind[out]) out++;
        const int i1 = (out + 1) % 3, i2 = (out + 2) % 3;
        ball_intersections<DIM>(x, s2, out, i1, horizon2, isec);
        const double c1 = isec[0];
        ball_intersections<DIM>(x, s2, out, i2, horizon2, isec);
        const double c2 = isec[0];
        double d1 = 0., d2 = 0.;
        for (int k = 0; k < DIM; k++) {
            const double t1 = PNB_SUB(s2[i2][k], PNB_ADD(PNB_MUL(c1, s2[i1][k]), PNB_MUL(PNB_SUB(1., c1), s2[out][k])));
            const double t2 = PNB_SUB(s2[i1][k], PNB_ADD(PNB_MUL(c2, s2[i2][k]), PNB_MUL(PNB_SUB(1., c2), s2[out][k])));
            d1 = PNB_ADD(d1, PNB_MUL(t1, t1));
            d2 = PNB_ADD(d2, PNB_MUL(t2, t2));
        }
        if (d1 < d2) {
            o.A[0][i2][i2] = 1.;
            o.A[0][out][out] = PNB_SUB(1., c2);
            o.A[0][i2][out] = c2;
            o.A[0][i1][i1] = c1;
            o.A[0][out][i1] = PNB_SUB(1., c1);
            o.vol[0] = PNB_MUL(c1, PNB_SUB(1., c2));
            o.A[1][i1][i1] = 1.;
            o.A[1][i2][i2] = 1.;
            o.A[1][out][out] = PNB_SUB(1., c1);
            o.A[1][i1][out] = c1;
            o.vol[1] = PNB_SUB(1., c1);
        } else {
            o.A[0][i1][i1] = 1.;
            o.A[0][i2][i2] = c2;
            o.A[0][out][i2] = PNB_SUB(1., c2);
            o.A[0][out][out] = PNB_SUB(1., c1);
            o.A[0][i1][out] = c1;
            o.vol[0] = PNB_MUL(c2, PNB_SUB(1., c1));
            o.A[1][i1][i1] = 1.;
            o.A[1][i2][i2] = 1.;
            o.A[1][out][out] = PNB_SUB(1., c2);
            o.A[1][i2][out] = c2;
            o.vol[1] = PNB_SUB(1., c2);
        }
        o.n = 2;
    } else if (numInside == 3) {
        o.A[0][0][0] = o.A[0][1][1] = o.A[0][2][2] = 1.;
        o.vol[0] = 1.;
        o.n = 1;
    }
    // numInside == 0: no special point for the l2 ball, the intersection (if any) is ignored (:648-667)
}

// Regular element pair whose vertex distances straddle the horizon: cut branch of eval_distant
// (nonlocalOperator_{SCALAR}.pxi:790-847).  The lanes split the (outer sub-simplex, outer node) items; every lane
// re-triangulates the inner element for its outer nodes.  acc[NL] as in lanes_regular_interior (not yet multiplied
// by vol1*vol2).
template <int DIM>
__device__ void lanes_cut_interior(const DProblem &P, int c1, int c2, int order, int lane, int nlanes, double *acc)
{
    constexpr int NV = PairDims<DIM>::NV, NL = PairDims<DIM>::NL;
    double s1[3][2], s2[3][2];
    load_simplex<DIM>(P.simplices, c1, NV, s1);
    load_simplex<DIM>(P.simplices, c2, NV, s2);
    const DRule r = P.reg_cell[order];
    const int n = r.n;
    const PowCtx kv(P.pow_int);
    const double horizon2 = P.horizon2;
#pragma unroll
    for (int k = 0; k < NL; k++) acc[k] = 0.;
    CutOuter co;
    cut_outer<DIM>(s1, s2, horizon2, co);
    CutInner ci;
    for (int q = lane; q < co.n * n; q += nlanes) {
        const int t = q / n, i = q - t * n;
        // transformed outer node (transformQuadratureRule.compute, quadrature.pyx:197-206) and its global coordinates
        double l1[3] = {0., 0., 0.}, x[2] = {0., 0.};
        for (int k = 0; k < NV; k++) {
            double v = co.b[t][k];
            for (int j = 0; j < NV; j++) v = PNB_ADD(v, PNB_MUL(co.A[t][k][j], r.bary[j * n + i]));
            l1[k] = v;
        }
        for (int k = 0; k < NV; k++)
            for (int m = 0; m < DIM; m++) x[m] = PNB_ADD(x[m], PNB_MUL(l1[k], s1[k][m]));
        cut_inner<DIM>(x, s2, horizon2, ci);
        // the outer node is fixed over the inner loops: row sum rs = sum g, tJ = sum g l2, yy = sum g l2 l2^T are
        // accumulated per inner node, the products with l1 formed once per outer node
        double rs = 0., tJ[NV], yy[NV * (NV + 1) / 2];
#pragma unroll
        for (int k = 0; k < NV; k++) tJ[k] = 0.;
#pragma unroll
        for (int k = 0; k < NV * (NV + 1) / 2; k++) yy[k] = 0.;
        for (int u = 0; u < ci.n; u++) {
            const double cc = PNB_MUL(co.vol[t], ci.vol[u]);
            for (int jn = 0; jn < n; jn++) {
                double l2[3] = {0., 0., 0.}, y[2] = {0., 0.};
                for (int k = 0; k < NV; k++) {
                    double v = 0.;
                    for (int j = 0; j < NV; j++) v = PNB_ADD(v, PNB_MUL(ci.A[u][k][j], r.bary[j * n + jn]));
                    l2[k] = v;
                }
                for (int k = 0; k < NV; k++)
                    for (int m = 0; m < DIM; m++) y[m] = PNB_ADD(y[m], PNB_MUL(l2[k], s2[k][m]));
                double d2 = 0.;
                for (int m = 0; m < DIM; m++) { const double w = PNB_SUB(x[m], y[m]); d2 = PNB_ADD(d2, PNB_MUL(w, w)); }
                const double g = (r.w[i] * r.w[jn]) * kv(d2) * cc;
                rs += g;
                int kk = 0;
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    const double ga = g * l2[a];
                    tJ[a] += ga;
#pragma unroll
                    for (int b = a; b < NV; b++) { yy[kk] = fma(ga, l2[b], yy[kk]); kk++; }
                }
            }
        }
        {
            int k = 0;
#pragma unroll
            for (int I = 0; I < 2 * NV; I++)
#pragma unroll
                for (int J = I; J < 2 * NV; J++) {
                    if (J < NV) acc[k] = fma(rs * l1[I], l1[J], acc[k]);
                    else if (I < NV) acc[k] = fma(-l1[I], tJ[J - NV], acc[k]);
                    else acc[k] += yy[tri_idx(NV, I - NV, J - NV)];
                    k++;
                }
        }
    }
}

template <int N> __device__ __forceinline__ void warp_allreduce(double *v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    }
}

// ---------------------------------------------------------------------------
// Far evaluator: regular 2D pair of order 2..5, one thread per pair.  Factored
// form of nonlocalOperator_{SCALAR}.pxi:769-789 with g_ij = gamma(x_i, y_j):
//   xx[I,I'] =  sum_i w_i phi_I(x_i) phi_I'(x_i) r_i,   r_i = sum_j w_j g_ij
//   yy[J,J'] =  sum_j w_j phi_J(y_j) phi_J'(y_j) c_j,   c_j = sum_i w_i g_ij
//   xy[I,J]  = -sum_i w_i phi_I(x_i) sum_j g_ij w_j phi_J(y_j)
// The rule of each order lives in its own __constant__ symbol so that the
// unrolled inner loop uses constant-bank operands; weights are folded into the
// per-node constants (wb = w * bary).  Outputs are NOT scaled by vol1*vol2.
// ---------------------------------------------------------------------------
#ifndef PNB_FAR_UNROLL
#define PNB_FAR_UNROLL 1
#endif
static constexpr int kFarUnroll = PNB_FAR_UNROLL;
struct FarRule {
    int n;
    int pad;
    double bary[3][8];
    double w[8];
    double wb[3][8];   // w[j] * bary[k][j]
    double qq[6][8];   // w[j] * bary[a][j] * bary[b][j], a<=b   (diagonal block of the second cell)
};

// node counts the thread-per-pair evaluator accepts (orders 2..5 of the adopted triangle family)
__host__ __device__ inline int far_expected_nodes(int order) { return order == 2 ? 3 : (order == 3 || order == 4) ? 6 : order == 5 ? 7 : -1; }

// Deliberately ROLLED loops with a runtime node count: one small loop body serves every order.  (Unrolled
// per-order instantiations were measured 2-3x slower: the tile kernel then no longer fits the instruction
// cache and warps stall on instruction fetch.)  R lives in shared memory; all lanes of a warp that work on
// the same order read the same addresses (broadcast).
template <class KV>
__device__ __forceinline__ void far_eval_2d(const FarRule &R, const double (*s1)[2], const double (*s2)[2],
                                            const KV &T, const bool with_d, double *xy, double *xx, double *yy)
{
    const int n = R.n;
#pragma unroll
    for (int k = 0; k < 9; k++) xy[k] = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) xx[k] = yy[k] = 0.;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const double p0 = R.bary[0][i], p1 = R.bary[1][i], p2 = R.bary[2][i];
        const double X0 = p0 * s1[0][0] + p1 * s1[1][0] + p2 * s1[2][0];
        const double X1 = p0 * s1[0][1] + p1 * s1[1][1] + p2 * s1[2][1];
        const double wi = R.w[i];
        double r = 0., t0 = 0., t1 = 0., t2 = 0.;
#pragma unroll kFarUnroll
        for (int j = 0; j < n; j++) {
            const double q0 = R.bary[0][j], q1 = R.bary[1][j], q2 = R.bary[2][j];
            const double a = X0 - (q0 * s2[0][0] + q1 * s2[1][0] + q2 * s2[2][0]);
            const double b = X1 - (q0 * s2[0][1] + q1 * s2[1][1] + q2 * s2[2][1]);
            const double g = T(a * a + b * b);
            t0 = fma(g, R.wb[0][j], t0);
            t1 = fma(g, R.wb[1][j], t1);
            t2 = fma(g, R.wb[2][j], t2);
            if (with_d) {
                r = fma(g, R.w[j], r);
                const double gw = g * wi;
                yy[0] = fma(gw, R.qq[0][j], yy[0]);
                yy[1] = fma(gw, R.qq[1][j], yy[1]);
                yy[2] = fma(gw, R.qq[2][j], yy[2]);
                yy[3] = fma(gw, R.qq[3][j], yy[3]);
                yy[4] = fma(gw, R.qq[4][j], yy[4]);
                yy[5] = fma(gw, R.qq[5][j], yy[5]);
            }
        }
        const double q0 = wi * p0, q1 = wi * p1, q2 = wi * p2;
        if (with_d) {
            const double r0 = r * q0, r1 = r * q1, r2 = r * q2;
            xx[0] += r0 * p0; xx[1] += r0 * p1; xx[2] += r0 * p2;
            xx[3] += r1 * p1; xx[4] += r1 * p2; xx[5] += r2 * p2;
        }
        xy[0] -= q0 * t0; xy[1] -= q0 * t1; xy[2] -= q0 * t2;
        xy[3] -= q1 * t0; xy[4] -= q1 * t1; xy[5] -= q1 * t2;
        xy[6] -= q2 * t0; xy[7] -= q2 * t1; xy[8] -= q2 * t2;
    }
}


// Same evaluation with the node count N known at compile time: the column loop is unrolled (the row loop stays
// rolled, the body is small enough for the instruction cache), so that the nodes of the second cell and the
// column sums c_j = sum_i w_i g_ij live in registers: 19 instead of 31 FP64 operations per node pair, and the
// rule constants are read at fixed shared-memory offsets.
template <int N, class KV>
__device__ __forceinline__ void far_eval_n(const FarRule &R, const double (*s1)[2], const double (*s2)[2], const KV &T,
                                           double *xy, double *xx, double *yy)
{
    double Y0[N], Y1[N], c[N];
    // N = 3: the 12 constants stay in registers
    typename std::conditional<(N > 3), const volatile FarRule &, const FarRule &>::type RV = R;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double q0 = R.bary[0][j], q1 = R.bary[1][j], q2 = R.bary[2][j];
        Y0[j] = q0 * s2[0][0] + q1 * s2[1][0] + q2 * s2[2][0];
        Y1[j] = q0 * s2[0][1] + q1 * s2[1][1] + q2 * s2[2][1];
        c[j] = 0.;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) xy[k] = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) xx[k] = 0.;
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        const double p0 = R.bary[0][i], p1 = R.bary[1][i], p2 = R.bary[2][i];
        const double X0 = p0 * s1[0][0] + p1 * s1[1][0] + p2 * s1[2][0];
        const double X1 = p0 * s1[0][1] + p1 * s1[1][1] + p2 * s1[2][1];
        const double wi = R.w[i];
        double r = 0., t0 = 0., t1 = 0., t2 = 0.;
        double d2[N], g[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const double a = X0 - Y0[j], b = X1 - Y1[j];
            d2[j] = a * a + b * b;
        }
        T.template batch<N>(d2, g);
#pragma unroll
        for (int j = 0; j < N; j++) {
            // the rule constants are re-read from shared memory (broadcast) in every row: hoisted out of the row loop
            // they would not fit into the registers and come back from local memory instead
            t0 = fma(g[j], RV.wb[0][j], t0);
            t1 = fma(g[j], RV.wb[1][j], t1);
            t2 = fma(g[j], RV.wb[2][j], t2);
            c[j] = fma(g[j], wi, c[j]);
        }
        // row sum of the weighted kernel values: the barycentric coordinates of a node add up to one
        r = (t0 + t1) + t2;
        const double q0 = wi * p0, q1 = wi * p1, q2 = wi * p2;
        const double r0 = r * q0, r1 = r * q1, r2 = r * q2;
        xx[0] += r0 * p0; xx[1] += r0 * p1; xx[2] += r0 * p2;
        xx[3] += r1 * p1; xx[4] += r1 * p2; xx[5] += r2 * p2;
        xy[0] -= q0 * t0; xy[1] -= q0 * t1; xy[2] -= q0 * t2;
        xy[3] -= q1 * t0; xy[4] -= q1 * t1; xy[5] -= q1 * t2;
        xy[6] -= q2 * t0; xy[7] -= q2 * t1; xy[8] -= q2 * t2;
    }
#pragma unroll
    for (int e = 0; e < 6; e++) {
        double v = 0.;
#pragma unroll
        for (int j = 0; j < N; j++) v = fma(R.qq[e][j], c[j], v);
        yy[e] = v;
    }
}
