// Device-side data model and per-pair math of the nonlocal assembly path.
//
// Everything here restates (B200-first, not line by line) what the reference
// computes per cell pair; the reference location of each piece is cited at
// the function.  Paths are relative to the reference root.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PNB_FAR_MAX_ORDER 5   // orders handled by the thread-per-pair evaluator
#define PNB_TD 64             // DoFs per output tile
#define PNB_SB 16             // sub-batch: PNB_SB x PNB_SB cell pairs
#define PNB_THREADS 256
#define PNB_ROW_PANELS 12      // row panels of the cell-group path (copy of finished rows overlaps the assembly)
#define PNB_NEAR_THREADS 256    // near evaluator: two CTAs of eight warps per SM (six warps with 170 registers were measured slower)
#define PNB_IGNORED_PANEL (-6)
#define PNB_CUT_FLAG 4096      // added to the regular order of a pair cut by the horizon (near pass of the tile kernel)
#define PNB_NEAR_R 4          // rows per register tile of the near evaluator
#define PNB_NEAR_WARP_POINTS 288   // shared-memory points (double2) per warp of the near evaluator
#define PNB_DER2 7            // double2 per node of the derived rule table (see DProblem::reg_derived)

struct DRule {
    int n;
    int rows;
    const double *bary;  // rows x n
    const double *w;     // n
};

// Table-driven x^e for a fixed exponent e (the kernel singularity):
//   x = 2^E * m,  m in [1,2),  m = m0[idx] * (1+r),  idx = top 7 mantissa bits, |r| <= 2^-8
//   scal * x^e = T1[E] * T2[idx] * sum_k binom(e,k) r^k      (k <= 7)
// T1 = scal * 2^(E e) and T2 = m0^e are rounded from long double on the host;
// max relative error ~5e-16 (measured against 200-bit arithmetic), against
// ~1.3e-16 of libm pow.  T1 covers 256 binary exponents of x; the host places the window (eoff) so that it ends
// above 4 diam^2 of the mesh (pnb_problem_create), i.e. every |x-y|^2 that can occur down to 2^-250 of the domain
// size lies inside.  Arguments outside the window (zero distance, NaN) are clamped to its ends: no branch and no
// call in the evaluation loops (a call there made the compiler keep the loop state in local memory).
struct PowTab {
    double coef[8];
    double T1[256];
    double2 IT[128];   // x = 1/m0 (rounded), y = m0^e for the exact reciprocal of x
    double scal, expo;
    double horizon2;   // finite horizon: the kernel vanishes for |x-y|^2 > horizon2 (+inf otherwise)
    int eoff, pad;     // T1[k] = scal * 2^((k - eoff) e)
};

struct DProblem {
    int dim, nc, nv, N, nb;
    const PowTab *pow_int;     // interior kernel  C |x-y|^(-d-2s)   as a function of |x-y|^2
    const PowTab *pow_bnd;     // boundary kernel
    const PowTab *pow_bnd_unit;    // boundary kernel / |x-y|: gamma_b(x,y) n.(y-x)/|y-x| = pow_bnd_unit(|x-y|^2) n.(y-x)
    const struct FarRule *far_rules;   // PNB_FAR_MAX_ORDER+1 low-order 2D rules for the thread-per-pair evaluator
    const float *lhf;          // nc: (float) log(h)
    const float *ahf;          // nc: (float) |log(h/H0)|
    const float *lhcf, *ahcf;  // nc: the same for get_h_simplex (boundary class)
    const float *lhbf, *ahbf;  // nb: the same for the boundary facets
    const double *simplices;   // nc x (dim+1) x dim     (precomputeSimplices, nonlocalOperator_{SCALAR}.pxi:111-126)
    const double *centers;     // nc x dim
    const int *cells;          // nc x (dim+1)
    const int *dofs;           // nc x (dim+1)
    const double *vol;         // nc
    const double *h;           // nc
    const int *bfacets;        // nb x dim
    const double *bsimplices;  // nb x dim x dim
    const double *bcenters;    // nb x dim
    const double *bvol;        // nb
    const double *bh;          // nb   (get_h_surface_simplex)
    const double *hcell;       // nc   (get_h_simplex, used by the boundary class)
    double s, C, Cb, sing, bsing, expo, bexpo;
    double horizon2;           // fHORIZON2, +inf for the infinite horizon
    double H0, c_int, c_bnd;   // c_* = target-order dependent log constants of getQuadOrder
    DRule q_id, q_edge, q_vertex, bq_edge, bq_vertex;
    int max_order;
    const DRule *reg_cell;
    const DRule *reg_facet;
    // 2D cell rules, per node PNB_DER2 double2: (w, w b0) (w b1, w b2) (q0,q1) (q2,q3) (q4,q5) with q = w b_a b_b (a<=b),
    // (b0,b1) (b2,0); rule o starts at node reg_doff[o]
    const double *reg_derived;
    const int *reg_doff;
    int reg_nmax;              // largest node count of a cell rule
    const int4 *reg_grid;      // per order: lane groups of the near evaluator (lanes per item W, items per warp K, row tiles per lane, 0)
    // piecewise constant variable kernels: this problem instance takes the pairs of one class (see pnb_kernel_t)
    const unsigned char *labels;   // nc, nullptr = constant kernel
    const unsigned char *blabels;  // nb
    int active_class;
    unsigned int pair_class[4];    // packed: byte l2 of word l1
    unsigned int bpair_class[4];   // (cell label, boundary facet label)
    int pair_orientation;          // singular pairs: 0 smaller cell index first, 1 larger first
    int pair_filter;               // 1: touching (singular) pairs only
};

// does the pair of labels belong to this problem instance?
__host__ __device__ inline bool pnb_class_active(const DProblem &P, int l1, int l2)
{
    return (int)((P.pair_class[l1 & 3] >> (8 * (l2 & 3))) & 0xFF) == P.active_class;
}

__host__ __device__ inline bool pnb_bclass_active(const DProblem &P, int lc, int lf)
{
    return (int)((P.bpair_class[lc & 3] >> (8 * (lf & 3))) & 0xFF) == P.active_class;
}

__host__ __device__ inline int tri_idx(int n, int i, int j) { return n * i - ((i * (i + 1)) >> 1) + j; }

#ifdef __CUDA_ARCH__
#define PNB_MUL(a, b) __dmul_rn((a), (b))
#define PNB_ADD(a, b) __dadd_rn((a), (b))
#define PNB_SUB(a, b) __dsub_rn((a), (b))
#else
#define PNB_MUL(a, b) ((a) * (b))
#define PNB_ADD(a, b) ((a) + (b))
#define PNB_SUB(a, b) ((a) - (b))
#endif

// Shared-vertex classification with the reference's first-match permutation
// choice (getProtoPanelType, nonlocalOperator_{SCALAR}.pxi:280-378).
// Returns -(#shared vertices), or -n1 for identical cells.
__host__ __device__ inline int proto_panel(const int *v1, int n1, const int *v2, int n2, bool identical,
                                           int *perm1, int *perm2)
{
    int common = 0;
    unsigned m1 = 0, m2 = 0;
    if (identical) {
        for (int k = 0; k < n1; k++) perm1[k] = k;
        for (int k = 0; k < n2; k++) perm2[k] = k;
        return -n1;
    }
    for (int a = 0; a < n1; a++) {
        for (int b = 0; b < n2; b++) {
            if (m2 & (1u << b)) continue;
            if (v1[a] == v2[b]) {
                perm1[common] = a;
                perm2[common] = b;
                m1 |= 1u << a;
                m2 |= 1u << b;
                common++;
                break;
            }
        }
    }
    if (common == 0) {
        for (int k = 0; k < n1; k++) perm1[k] = k;
        for (int k = 0; k < n2; k++) perm2[k] = k;
        return 0;
    }
    int a = 0;
    for (int k = common; k < n1; k++) {
        while (m1 & (1u << a)) a++;
        perm1[k] = a;
        m1 |= 1u << a;
    }
    int b = 0;
    for (int k = common; k < n2; k++) {
        while (m2 & (1u << b)) b++;
        perm2[k] = b;
        m2 |= 1u << b;
    }
    return -common;
}

// Number of shared vertices only (no permutations).
__host__ __device__ inline int shared_vertices(const int *v1, int n1, const int *v2, int n2)
{
    int common = 0;
    unsigned m2 = 0;
    for (int a = 0; a < n1; a++)
        for (int b = 0; b < n2; b++) {
            if (m2 & (1u << b)) continue;
            if (v1[a] == v2[b]) { m2 |= 1u << b; common++; break; }
        }
    return common;
}

// getQuadOrder of the interior local matrix
// (fractionalLaplacian2D.pyx:622-642, fractionalLaplacian1D.pyx:234-253).
// Products and sums are kept un-fused so that the value handed to ceil() is
// rounded exactly as in the reference's scalar code.
__host__ __device__ inline int quad_order_interior(const DProblem &P, double h1, double h2, double d)
{
    double logdh1 = log(d / h1), logdh2 = log(d / h2);
    double p1, p2;
    if (P.dim == 2) {
        double logh1H0 = fabs(log(h1 / P.H0)), logh2H0 = fabs(log(h2 / P.H0));
        double loghminH0 = fmax(logh1H0, logh2H0);
        double s = fmax(-0.5 * (P.sing + 2), 0.);
        double num1 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_int, PNB_MUL(s - 1., logh2H0)), loghminH0), PNB_MUL(s, logdh2));
        double num2 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_int, PNB_MUL(s - 1., logh1H0)), loghminH0), PNB_MUL(s, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(fmax(logdh1, 0.), 0.4)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(fmax(logdh2, 0.), 0.4)), 2.);
    } else {
        double s = fmax(-0.5 * (P.sing + 1), 0.);
        double a = 2. * s - 1., b = 2. * s;
        double num1 = PNB_SUB(PNB_ADD(P.c_int, PNB_MUL(a, fabs(log(h2 / P.H0)))), PNB_MUL(b, logdh2));
        double num2 = PNB_SUB(PNB_ADD(P.c_int, PNB_MUL(a, fabs(log(h1 / P.H0)))), PNB_MUL(b, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(fmax(logdh1, 0.), 0.8)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(fmax(logdh2, 0.), 0.8)), 2.);
    }
    return (int)fmax(p1, p2);
}

// getQuadOrder of the boundary local matrix, infinite horizon
// (fractionalLaplacian2D.pyx:1226-1253, fractionalLaplacian1D.pyx:644-669).
__host__ __device__ inline int quad_order_boundary(const DProblem &P, double h1, double h2, double d)
{
    double p1, p2;
    double logdh1 = fmax(log(d / h1), 0.), logdh2 = fmax(log(d / h2), 0.);
    if (P.dim == 2) {
        double logh1H0 = fabs(log(h1 / P.H0)), logh2H0 = fabs(log(h2 / P.H0));
        double loghminH0 = fmax(logh1H0, logh2H0);
        double s = fmax(0.5 * (-P.bsing - 1.), 0.);
        double num1 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_bnd, loghminH0), PNB_MUL(s - 1., logh2H0)), PNB_MUL(s, logdh2));
        double num2 = PNB_SUB(PNB_ADD(PNB_ADD(P.c_bnd, loghminH0), PNB_MUL(s - 1., logh1H0)), PNB_MUL(s, logdh1));
        p1 = fmax(ceil(num1 / PNB_ADD(fmax(logdh1, 0.), 0.35)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(fmax(logdh2, 0.), 0.35)), 2.);
    } else {
        double s = fmax(0.5 * (-P.bsing - 1.), 0.);
        double a = 2. * s - 1., b = 2. * s;
        double num1 = PNB_SUB(PNB_ADD(P.c_bnd, PNB_MUL(a, fabs(log(h2 / P.H0)))), PNB_MUL(b, log(d / h2)));
        double num2 = PNB_SUB(PNB_ADD(P.c_bnd, PNB_MUL(a, fabs(log(h1 / P.H0)))), PNB_MUL(b, log(d / h1)));
        p1 = fmax(ceil(num1 / PNB_ADD(logdh1, 0.8)), 2.);
        p2 = fmax(ceil(num2 / PNB_ADD(logdh2, 0.8)), 2.);
    }
    return (int)fmax(p1, p2);
}

// computeCenterDistance (nonlocalOperator_{SCALAR}.pxi:380-386), un-fused.
__host__ __device__ inline double center_distance(const double *c1, const double *c2, int dim)
{
    double d2 = 0.;
    for (int j = 0; j < dim; j++) {
        double t = PNB_SUB(c1[j], c2[j]);
        d2 = PNB_ADD(d2, PNB_MUL(t, t));
    }
    return sqrt(d2);
}

// ball2_retriangulation.getRelativePosition (interactionDomains.pyx:875-898) of two cells:
// 0 INTERACT (all vertex distances <= horizon), 1 REMOTE (all >= horizon), 2 CUT
__host__ __device__ inline int pair_relative_position(const DProblem &P, int c1, int c2)
{
    const int nvc = P.dim + 1;
    double dmin2 = INFINITY, dmax2 = 0.;
    for (int i = 0; i < nvc; i++)
        for (int k = 0; k < nvc; k++) {
            double d2 = 0.;
            for (int j = 0; j < P.dim; j++) {
                const double t = PNB_SUB(P.simplices[((size_t)c1 * nvc + i) * P.dim + j], P.simplices[((size_t)c2 * nvc + k) * P.dim + j]);
                d2 = PNB_ADD(d2, PNB_MUL(t, t));
            }
            dmin2 = fmin(dmin2, d2);
            dmax2 = fmax(dmax2, d2);
        }
    if (dmin2 >= P.horizon2) return 1;
    if (dmax2 <= P.horizon2) return 0;
    return 2;
}

// getPanelType for an element pair c1<=c2 (nonlocalOperator_{SCALAR}.pxi:493-540).
__host__ __device__ inline int panel_interior(const DProblem &P, int c1, int c2, int *perm1, int *perm2)
{
    const int nvc = P.dim + 1;
    if (c1 > c2) return PNB_IGNORED_PANEL;
    int panel = proto_panel(P.cells + (size_t)c1 * nvc, nvc, P.cells + (size_t)c2 * nvc, nvc, c1 == c2, perm1, perm2);
    if (panel == 0) {
        // finite horizon: pairs entirely outside each other's interaction ball are ignored (:516)
        if (P.horizon2 < INFINITY && pair_relative_position(P, c1, c2) == 1) return PNB_IGNORED_PANEL;
        double d = center_distance(P.centers + (size_t)c1 * P.dim, P.centers + (size_t)c2 * P.dim, P.dim);
        panel = quad_order_interior(P, P.h[c1], P.h[c2], d);
    }
    return panel;
}

// getPanelType for an element x boundary facet pair.
__host__ __device__ inline int panel_boundary(const DProblem &P, int c1, int f, int *perm1, int *perm2)
{
    const int nvc = P.dim + 1, nvf = P.dim;
    int panel = proto_panel(P.cells + (size_t)c1 * nvc, nvc, P.bfacets + (size_t)f * nvf, nvf, false, perm1, perm2);
    if (panel == 0) {
        double d = center_distance(P.centers + (size_t)c1 * P.dim, P.bcenters + (size_t)f * P.dim, P.dim);
        panel = quad_order_boundary(P, P.hcell[c1], P.bh[f], d);
    }
    return panel;
}
