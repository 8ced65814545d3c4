/* ORACLE -- test infrastructure only.  Never linked, imported or executed by
 * the product path (pynucleus_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs may use it, as the checker.
 *
 * Plain-C restatement of the reference's nonlocal dense assembly for
 * symmetric kernels with constant parameters, P1 elements, 1D and 2D:
 *
 *   pair classification     nl/PyNucleus_nl/nonlocalOperator_{SCALAR}.pxi:280-378
 *   panel / order choice    nonlocalOperator_{SCALAR}.pxi:493-540,
 *                           fractionalLaplacian2D.pyx:622-642, 1226-1253
 *                           fractionalLaplacian1D.pyx:234-253, 644-669
 *   singular local matrix   fractionalLaplacian2D.pyx:823-891, 1324-1407
 *                           fractionalLaplacian1D.pyx:349-407, 719-783
 *   regular local matrix    nonlocalOperator_{SCALAR}.pxi:722-789, 1022-1108
 *   kernels                 nl/PyNucleus_nl/kernelsCy.pyx:159-240
 *   scatter                 nl/PyNucleus_nl/nonlocalAssembly_{SCALAR}.pxi:138-221
 *   loops                   nonlocalAssembly_{SCALAR}.pxi:1386-1448
 *
 * Pinned against the reference itself: tests/golden/*.npz hold panel types,
 * permutations, local matrices and assembled matrices produced by the
 * stub-built reference (oracle/refbuild), see tests/test_oracle_golden.py.
 *
 * Quadrature tables come from oracle/tables.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IGNORED (-6)
#define MAXV 3
#define MAX_ORDER 256

typedef struct {
    int n;
    const double *bary; /* rows x n, row-major */
    const double *w;
} orc_rule;

typedef struct {
    int dim;            /* 1 or 2 */
    int nv, nc;
    const double *vertices; /* nv x dim */
    const int32_t *cells;   /* nc x (dim+1) */
    const double *vol;      /* nc */
    const double *h;        /* nc */
    const int32_t *dofs;    /* nc x (dim+1), <0 = boundary dof */
    int num_dofs;
    int nb;
    const int32_t *bfacets; /* nb x dim */
    double H0;
    /* kernel: C |x-y|^(-dim-2s); boundary kernel: Cb |x-y|^(-(dim-1)-2s) */
    double s, C, Cb, singularity, bsingularity;
    double target_order, btarget_order;
    /* singular tables */
    orc_rule qr_face, qr_edge, qr_vertex, bqr_edge, bqr_vertex;
    /* regular tables indexed by order */
    int max_order;
    const orc_rule *reg_cell;   /* rule on a cell       */
    const orc_rule *reg_facet;  /* rule on a boundary facet */
    /* num_dofs entering getQuadOrder; 0 = num_dofs.  Two DoFMaps (NA.pxi:1366-1378): the local matrices keep the
     * count of the first map while the assembly runs over the combined map */
    int order_num_dofs;
    /* piecewise constant variable kernels (kernel.evalParams at the cell centres, NO.pxi:509-513): the kernel
     * parameters of this problem hold for the pairs whose label pair maps to active_class; all other pairs belong to
     * another problem instance (the operator is the sum over the classes).  labels == NULL: constant kernel */
    const unsigned char *labels;   /* nc */
    const unsigned char *blabels;  /* nb */
    int active_class;
    unsigned char pair_class[16];  /* [label of cell c1 * 4 + label of cell c2], c1 <= c2 */
    unsigned char bpair_class[16]; /* [cell label * 4 + boundary facet label] */
    /* Unsymmetric piecewise orders: the reference visits (c1,c2) and (c2,c1) (NA.pxi:1412-1428); pair_orientation 1
     * evaluates the second visit, i.e. the touching pair with the cell of the larger index as first cell of the rule */
    int pair_orientation;
    /* tempered fractional kernel (temperedFracKernelInfinite*, kernelsCy.pyx:186-213): interior kernel times
     * exp(-tempered |x-y|); the boundary kernel stays the plain power law (getBoundaryKernel, kernelsCy.pyx:2011-2020,
     * does not hand `tempered` on).  0 = not tempered */
    double tempered;
    /* integrable kernels on the full space (gaussianKernel*, exponentialKernel, kernelsCy.pyx:388-477): C f(|x-y|) with
     * smode 1: f = exp(-sa r), 2: f = exp(-sa r^2); boundary forms Cb fb(|x-y|) with bmode 1: exp(-ba r),
     * 3: erfc(sqrt(ba) r), 4: exp(-ba r^2)/r.  0 = power-law kernels */
    int smode, bmode;
    double sa, ba;
} orc_problem;
#define ORDN(P) ((P)->order_num_dofs > 0 ? (P)->order_num_dofs : (P)->num_dofs)


/* ---------------------------------------------------------------------- */
/* classification: shared-vertex count and vertex permutations             */
/* ---------------------------------------------------------------------- */
static int proto_panel(const int32_t *v1, int n1, const int32_t *v2, int n2,
                       int identical, int *perm1, int *perm2)
{
    int k, a, b, common = 0;
    unsigned m1 = 0, m2 = 0;
    if (identical) {
        for (k = 0; k < n1; k++) perm1[k] = k;
        for (k = 0; k < n2; k++) perm2[k] = k;
        return -n1;
    }
    for (a = 0; a < n1; a++) {
        for (b = 0; b < n2; b++) {
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
        for (k = 0; k < n1; k++) perm1[k] = k;
        for (k = 0; k < n2; k++) perm2[k] = k;
        return 0;
    }
    a = 0;
    for (k = common; k < n1; k++) {
        while (m1 & (1u << a)) a++;
        perm1[k] = a;
        m1 |= 1u << a;
    }
    b = 0;
    for (k = common; k < n2; k++) {
        while (m2 & (1u << b)) b++;
        perm2[k] = b;
        m2 |= 1u << b;
    }
    return -common;
}

static double maxd(double a, double b) { return a > b ? a : b; }

/* fractionalLaplacian2D.pyx:622-642 and fractionalLaplacian1D.pyx:234-253 */
static int quad_order_interior(const orc_problem *P, double h1, double h2, double d)
{
    double logdh1 = log(d / h1), logdh2 = log(d / h2);
    double p1, p2;
    if (P->dim == 2) {
        double c = (0.5 * P->target_order + 0.5) * log(ORDN(P) * (P->H0 * P->H0));
        double logh1H0 = fabs(log(h1 / P->H0)), logh2H0 = fabs(log(h2 / P->H0));
        double loghminH0 = maxd(logh1H0, logh2H0);
        double s = maxd(-0.5 * (P->singularity + 2), 0.);
        p1 = maxd(ceil((c + (s - 1.) * logh2H0 + loghminH0 - s * logdh2) / (maxd(logdh1, 0) + 0.4)), 2);
        p2 = maxd(ceil((c + (s - 1.) * logh1H0 + loghminH0 - s * logdh1) / (maxd(logdh2, 0) + 0.4)), 2);
    } else {
        double s = maxd(-0.5 * (P->singularity + 1), 0.);
        double c = (P->target_order + 2.) * log(ORDN(P) * P->H0);
        p1 = maxd(ceil((c + (2. * s - 1.) * fabs(log(h2 / P->H0)) - 2. * s * logdh2) / (maxd(logdh1, 0) + 0.8)), 2);
        p2 = maxd(ceil((c + (2. * s - 1.) * fabs(log(h1 / P->H0)) - 2. * s * logdh1) / (maxd(logdh2, 0) + 0.8)), 2);
    }
    return (int)maxd(p1, p2);
}

/* fractionalLaplacian2D.pyx:1226-1253 and fractionalLaplacian1D.pyx:644-669
 * (infinite horizon: the "*3" branch for cut elements never triggers) */
static int quad_order_boundary(const orc_problem *P, double h1, double h2, double d)
{
    double p1, p2;
    if (P->dim == 2) {
        double logdh1 = maxd(log(d / h1), 0.), logdh2 = maxd(log(d / h2), 0.);
        double logh1H0 = fabs(log(h1 / P->H0)), logh2H0 = fabs(log(h2 / P->H0));
        double loghminH0 = maxd(logh1H0, logh2H0);
        double s = maxd(0.5 * (-P->bsingularity - 1.), 0.);
        double c = (0.5 * P->btarget_order + 0.25) * log(ORDN(P) * (P->H0 * P->H0));
        p1 = maxd(ceil((c + loghminH0 + (s - 1.) * logh2H0 - s * logdh2) / (maxd(logdh1, 0) + 0.35)), 2);
        p2 = maxd(ceil((c + loghminH0 + (s - 1.) * logh1H0 - s * logdh1) / (maxd(logdh2, 0) + 0.35)), 2);
    } else {
        double logdh1 = maxd(log(d / h1), 0.), logdh2 = maxd(log(d / h2), 0.);
        double s = maxd(0.5 * (-P->bsingularity - 1.), 0.);
        double c = (P->btarget_order + 1.) * log(ORDN(P) * P->H0);
        p1 = maxd(ceil((c + (2. * s - 1.) * fabs(log(h2 / P->H0)) - 2. * s * log(d / h2)) / (logdh1 + 0.8)), 2);
        p2 = maxd(ceil((c + (2. * s - 1.) * fabs(log(h1 / P->H0)) - 2. * s * log(d / h1)) / (logdh2 + 0.8)), 2);
    }
    return (int)maxd(p1, p2);
}

static void get_simplex(const orc_problem *P, const int32_t *verts, int n, double sx[MAXV][2], double c[2])
{
    int k, j;
    c[0] = c[1] = 0.;
    for (k = 0; k < n; k++)
        for (j = 0; j < P->dim; j++) {
            sx[k][j] = P->vertices[(size_t)verts[k] * P->dim + j];
            c[j] += sx[k][j];
        }
    for (j = 0; j < P->dim; j++) c[j] *= 1. / n;
}

/* interior kernel C*(d^2)^(-dim/2-s), kernelsCy.pyx:159-183 */
static double kernel_interior(const orc_problem *P, const double *x, const double *y)
{
    double d2 = (x[0] - y[0]) * (x[0] - y[0]);
    if (P->smode) {
        if (P->dim == 2) d2 += (x[1] - y[1]) * (x[1] - y[1]);
        return P->C * (P->smode == 1 ? exp(-P->sa * sqrt(d2)) : exp(-d2 * P->sa));
    }
    if (P->dim == 2) {
        d2 += (x[1] - y[1]) * (x[1] - y[1]);
        if (P->tempered != 0.) return P->C * pow(d2, -1. - P->s) * exp(-P->tempered * sqrt(d2));
        return P->C * pow(d2, -1. - P->s);
    }
    if (P->tempered != 0.) return P->C * pow(d2, -0.5 - P->s) * exp(-P->tempered * sqrt(d2));
    return P->C * pow(d2, -0.5 - P->s);
}

/* boundary kernel, kernelsCy.pyx:216-240 */
static double kernel_boundary(const orc_problem *P, const double *x, const double *y)
{
    double d2 = (x[0] - y[0]) * (x[0] - y[0]);
    if (P->bmode) {
        if (P->dim == 2) d2 += (x[1] - y[1]) * (x[1] - y[1]);
        if (P->bmode == 1) return P->Cb * exp(-P->ba * sqrt(d2));
        if (P->bmode == 3) return P->Cb * erfc(sqrt(P->ba * d2));
        return P->Cb * exp(-P->ba * d2) / sqrt(d2);
    }
    if (P->dim == 2) {
        d2 += (x[1] - y[1]) * (x[1] - y[1]);
        return P->Cb * pow(d2, -0.5 - P->s);
    }
    return P->Cb * pow(d2, -P->s);
}

/* ---------------------------------------------------------------------- */
/* panel type of a cell pair (symmetric cells: c1 <= c2)                   */
/* ---------------------------------------------------------------------- */
static int orc_panel_interior_any(const orc_problem *P, int c1, int c2, int *perm1, int *perm2);
int orc_panel_interior(const orc_problem *P, int c1, int c2, int *perm1, int *perm2)
{
    if (c1 > c2) return IGNORED;
    return orc_panel_interior_any(P, c1, c2, perm1, perm2);
}

/* unsymmetric cells (NO.pxi:296: no c1 > c2 test) */
static int orc_panel_interior_any(const orc_problem *P, int c1, int c2, int *perm1, int *perm2)
{
    int nvc = P->dim + 1, panel;
    panel = proto_panel(P->cells + (size_t)c1 * nvc, nvc, P->cells + (size_t)c2 * nvc, nvc, c1 == c2, perm1, perm2);
    if (panel == 0) {
        double s1[MAXV][2], s2[MAXV][2], m1[2], m2[2], d2 = 0.;
        int j;
        get_simplex(P, P->cells + (size_t)c1 * nvc, nvc, s1, m1);
        get_simplex(P, P->cells + (size_t)c2 * nvc, nvc, s2, m2);
        for (j = 0; j < P->dim; j++) d2 += (m1[j] - m2[j]) * (m1[j] - m2[j]);
        panel = quad_order_interior(P, P->h[c1], P->h[c2], sqrt(d2));
    }
    return panel;
}

static double cell_h(const orc_problem *P, double sx[MAXV][2])
{
    double hmax = 0.;
    int i, j;
    if (P->dim == 1) return fabs(sx[1][0] - sx[0][0]);
    for (i = 0; i < 2; i++)
        for (j = i + 1; j < 3; j++) {
            double h2 = (sx[j][0] - sx[i][0]) * (sx[j][0] - sx[i][0]) + (sx[j][1] - sx[i][1]) * (sx[j][1] - sx[i][1]);
            hmax = maxd(hmax, h2);
        }
    return sqrt(hmax);
}

/* get_h_surface_simplex, nonlocalOperator.pyx:120-121, 162-169 */
static double facet_h(const orc_problem *P, double sx[MAXV][2])
{
    if (P->dim == 1) return 1.;
    return sqrt((sx[0][0] - sx[1][0]) * (sx[0][0] - sx[1][0]) + (sx[0][1] - sx[1][1]) * (sx[0][1] - sx[1][1]));
}

int orc_panel_boundary(const orc_problem *P, int c1, int f, int *perm1, int *perm2)
{
    int nvc = P->dim + 1, nvf = P->dim, panel;
    panel = proto_panel(P->cells + (size_t)c1 * nvc, nvc, P->bfacets + (size_t)f * nvf, nvf, 0, perm1, perm2);
    if (panel == 0) {
        double s1[MAXV][2], s2[MAXV][2], m1[2], m2[2], d2 = 0., h1, h2;
        int j;
        get_simplex(P, P->cells + (size_t)c1 * nvc, nvc, s1, m1);
        get_simplex(P, P->bfacets + (size_t)f * nvf, nvf, s2, m2);
        for (j = 0; j < P->dim; j++) d2 += (m1[j] - m2[j]) * (m1[j] - m2[j]);
        /* symmetricCells is False for the boundary class, so h1 comes from
         * get_h_simplex (nonlocalOperator.pyx:114-118, 152-160) */
        h1 = cell_h(P, s1);
        h2 = facet_h(P, s2);
        panel = quad_order_boundary(P, h1, h2, sqrt(d2));
    }
    return panel;
}

/* ---------------------------------------------------------------------- */
/* local matrices                                                          */
/* ---------------------------------------------------------------------- */
static inline int tri_index(int n, int i, int j) /* i<=j, n x n upper triangle */
{
    return n * i - (i * (i + 1) >> 1) + j;
}

/* element x element, any panel. contrib has (2*dpe)(2*dpe+1)/2 entries */
void orc_local_interior(const orc_problem *P, int c1, int c2, int panel, const int *perm1, const int *perm2, double *contrib)
{
    int nvc = P->dim + 1, dpe = nvc, nloc = (2 * dpe) * (2 * dpe + 1) / 2;
    double s1[MAXV][2], s2[MAXV][2], m1[2], m2[2];
    double vol1 = P->vol[c1], vol2 = P->vol[c2];
    int k, I, J, m, j;
    get_simplex(P, P->cells + (size_t)c1 * nvc, nvc, s1, m1);
    get_simplex(P, P->cells + (size_t)c2 * nvc, nvc, s2, m2);
    memset(contrib, 0, sizeof(double) * nloc);
    if (panel >= 1) {
        /* regular pair: nonlocalOperator_{SCALAR}.pxi:756-789 */
        const orc_rule *r = &P->reg_cell[panel];
        int n = r->n, i;
        double *x = malloc(sizeof(double) * 2 * n), *y = malloc(sizeof(double) * 2 * n);
        double *temp = malloc(sizeof(double) * (size_t)n * n);
        double vol = vol1 * vol2;
        for (i = 0; i < 2 * n; i++) x[i] = y[i] = 0.;
        for (k = 0; k < nvc; k++)
            for (i = 0; i < n; i++)
                for (j = 0; j < P->dim; j++) {
                    x[2 * i + j] += r->bary[k * n + i] * s1[k][j];
                    y[2 * i + j] += r->bary[k * n + i] * s2[k][j];
                }
        for (i = 0; i < n; i++)
            for (j = 0; j < n; j++)
                temp[(size_t)i * n + j] = (r->w[i] * r->w[j]) * kernel_interior(P, x + 2 * i, y + 2 * j);
        k = 0;
        for (I = 0; I < 2 * dpe; I++)
            for (J = I; J < 2 * dpe; J++) {
                double val = 0.;
                for (i = 0; i < n; i++)
                    for (j = 0; j < n; j++) {
                        double pI = I < dpe ? r->bary[I * n + i] : -r->bary[(I - dpe) * n + j];
                        double pJ = J < dpe ? r->bary[J * n + i] : -r->bary[(J - dpe) * n + j];
                        val += temp[(size_t)i * n + j] * pI * pJ;
                    }
                contrib[k++] = val * vol;
            }
        free(x); free(y); free(temp);
        return;
    }
    {
        /* singular pair */
        const orc_rule *r;
        int common = -panel, rows = 2 * dpe - common, n;
        int perm[2 * MAXV];
        double vol, *temp, *PSI;
        if (P->dim == 2) {
            r = panel == -3 ? &P->qr_face : (panel == -2 ? &P->qr_edge : &P->qr_vertex);
            vol = 4.0 * vol1 * vol2;
        } else {
            r = panel == -2 ? &P->qr_face : &P->qr_vertex;
            vol = vol1 * vol2;
        }
        n = r->n;
        for (k = 0; k < dpe; k++) perm[k] = perm1[k];
        for (k = common; k < dpe; k++) perm[dpe + k - common] = dpe + perm2[k];
        temp = malloc(sizeof(double) * n);
        PSI = malloc(sizeof(double) * (size_t)rows * n);
        for (m = 0; m < n; m++) {
            double x[2] = {0., 0.}, y[2] = {0., 0.};
            for (j = 0; j < P->dim; j++) {
                for (k = 0; k < nvc; k++) {
                    x[j] += s1[perm1[k]][j] * r->bary[k * n + m];
                    y[j] += s2[perm2[k]][j] * r->bary[(nvc + k) * n + m];
                }
            }
            temp[m] = r->w[m] * kernel_interior(P, x, y);
            for (k = 0; k < common; k++) PSI[(size_t)k * n + m] = r->bary[k * n + m] - r->bary[(nvc + k) * n + m];
            for (k = common; k < dpe; k++) {
                PSI[(size_t)k * n + m] = r->bary[k * n + m];
                PSI[(size_t)(dpe + k - common) * n + m] = -r->bary[(nvc + k) * n + m];
            }
        }
        for (I = 0; I < rows; I++) {
            int i = perm[I];
            for (J = I; J < rows; J++) {
                int jj = perm[J];
                double val = 0.;
                k = jj < i ? tri_index(2 * dpe, jj, i) : tri_index(2 * dpe, i, jj);
                for (m = 0; m < n; m++) val += temp[m] * PSI[(size_t)I * n + m] * PSI[(size_t)J * n + m];
                contrib[k] = val * vol;
            }
        }
        free(temp); free(PSI);
    }
}

/* element x boundary facet. contrib has dpe(dpe+1)/2 entries */
void orc_local_boundary(const orc_problem *P, int c1, int f, int panel, const int *perm1, const int *perm2, double *contrib)
{
    int nvc = P->dim + 1, nvf = P->dim, dpe = nvc, nloc = dpe * (dpe + 1) / 2;
    double s1[MAXV][2], s2[MAXV][2], m1[2], m2[2], nrm[2] = {0., 0.};
    double vol1 = P->vol[c1], vol2;
    int k, I, J, m, j, i;
    get_simplex(P, P->cells + (size_t)c1 * nvc, nvc, s1, m1);
    get_simplex(P, P->bfacets + (size_t)f * nvf, nvf, s2, m2);
    memset(contrib, 0, sizeof(double) * nloc);
    if (P->dim == 2) {
        double inv;
        vol2 = facet_h(P, s2);
        nrm[0] = s2[1][1] - s2[0][1];
        nrm[1] = s2[0][0] - s2[1][0];
        inv = 1. / sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1]);
        nrm[0] *= inv;
        nrm[1] *= inv;
    } else {
        vol2 = 1.;
    }
    if (panel >= 1) {
        /* nonlocalOperator_{SCALAR}.pxi:1069-1108 */
        const orc_rule *r0 = &P->reg_cell[panel], *r1 = &P->reg_facet[panel];
        int n0 = r0->n, n1 = r1->n;
        double *x = calloc(2 * n0, sizeof(double)), *y = calloc(2 * n1, sizeof(double));
        double *temp = malloc(sizeof(double) * (size_t)n0 * n1);
        double vol = vol1 * vol2;
        for (k = 0; k < nvc; k++)
            for (i = 0; i < n0; i++)
                for (j = 0; j < P->dim; j++) x[2 * i + j] += r0->bary[k * n0 + i] * s1[k][j];
        for (k = 0; k < nvf; k++)
            for (i = 0; i < n1; i++)
                for (j = 0; j < P->dim; j++) y[2 * i + j] += r1->bary[k * n1 + i] * s2[k][j];
        for (k = 0; k < n0; k++)
            for (m = 0; m < n1; m++) {
                double nw = 1.;
                if (P->dim == 2) {
                    double w0 = y[2 * m] - x[2 * k], w1 = y[2 * m + 1] - x[2 * k + 1];
                    double normW = 1. / sqrt(w0 * w0 + w1 * w1);
                    w0 *= normW;
                    w1 *= normW;
                    nw = nrm[0] * w0 + nrm[1] * w1;
                }
                temp[(size_t)k * n1 + m] = (r0->w[k] * r1->w[m]) * nw * kernel_boundary(P, x + 2 * k, y + 2 * m);
            }
        k = 0;
        for (I = 0; I < dpe; I++)
            for (J = I; J < dpe; J++) {
                double val = 0.;
                for (i = 0; i < n0; i++)
                    for (m = 0; m < n1; m++)
                        val += temp[(size_t)i * n1 + m] * r0->bary[I * n0 + i] * r0->bary[J * n0 + i];
                contrib[k++] = val * vol;
            }
        free(x); free(y); free(temp);
        return;
    }
    {
        const orc_rule *r = (P->dim == 2 && panel == -2) ? &P->bqr_edge : &P->bqr_vertex;
        int n = r->n;
        double vol = P->dim == 2 ? -2.0 * vol1 * vol2 : vol1;
        double *temp = malloc(sizeof(double) * n);
        for (m = 0; m < n; m++) {
            double x[2] = {0., 0.}, y[2] = {0., 0.}, nw = 1.;
            for (j = 0; j < P->dim; j++) {
                for (k = 0; k < nvc; k++) x[j] += s1[perm1[k]][j] * r->bary[k * n + m];
                for (k = 0; k < nvf; k++) y[j] += s2[perm2[k]][j] * r->bary[(nvc + k) * n + m];
            }
            if (P->dim == 2) {
                double w0 = x[0] - y[0], w1 = x[1] - y[1];
                double normW = 1. / sqrt(w0 * w0 + w1 * w1);
                w0 *= normW;
                w1 *= normW;
                nw = nrm[0] * w0 + nrm[1] * w1;
            }
            temp[m] = r->w[m] * nw * kernel_boundary(P, x, y);
        }
        for (I = 0; I < dpe; I++) {
            int ii = perm1[I];
            for (J = I; J < dpe; J++) {
                int jj = perm1[J];
                double val = 0.;
                k = jj < ii ? tri_index(dpe, jj, ii) : tri_index(dpe, ii, jj);
                for (m = 0; m < n; m++) val += temp[m] * r->bary[I * n + m] * r->bary[J * n + m];
                contrib[k] = val * vol;
            }
        }
        free(temp);
    }
}

/* ---------------------------------------------------------------------- */
/* batch entry points used by the tests                                    */
/* ---------------------------------------------------------------------- */
void orc_pairs(const orc_problem *P, int np, const int32_t *pairs, int32_t *panels,
               int32_t *perm1, int32_t *perm2, double *contribs)
{
    int n, nvc = P->dim + 1, nloc = (2 * nvc) * (2 * nvc + 1) / 2, k;
#pragma omp parallel for schedule(dynamic, 16) private(k)
    for (n = 0; n < np; n++) {
        int p1[MAXV], p2[MAXV];
        int panel = orc_panel_interior(P, pairs[2 * n], pairs[2 * n + 1], p1, p2);
        panels[n] = panel;
        for (k = 0; k < nvc; k++) {
            perm1[n * nvc + k] = p1[k];
            perm2[n * nvc + k] = p2[k];
        }
        if (contribs && panel != IGNORED) orc_local_interior(P, pairs[2 * n], pairs[2 * n + 1], panel, p1, p2, contribs + (size_t)n * nloc);
    }
}

void orc_boundary_pairs(const orc_problem *P, int np, const int32_t *pairs, int32_t *panels, double *contribs)
{
    int n, nvc = P->dim + 1, nloc = nvc * (nvc + 1) / 2;
#pragma omp parallel for schedule(dynamic, 16)
    for (n = 0; n < np; n++) {
        int p1[MAXV], p2[MAXV];
        int panel = orc_panel_boundary(P, pairs[2 * n], pairs[2 * n + 1], p1, p2);
        panels[n] = panel;
        if (contribs) orc_local_boundary(P, pairs[2 * n], pairs[2 * n + 1], panel, p1, p2, contribs + (size_t)n * nloc);
    }
}

/* max regular order over cells [start,end) x [c1, nc) and over the facets */
int orc_max_order(const orc_problem *P, int start, int end, int zero_exterior)
{
    int c1, best = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(max : best)
    for (c1 = start; c1 < end; c1++) {
        int p1[MAXV], p2[MAXV], c2, f, p;
        for (c2 = c1; c2 < P->nc; c2++) {
            p = orc_panel_interior(P, c1, c2, p1, p2);
            if (p > best) best = p;
        }
        if (zero_exterior)
            for (f = 0; f < P->nb; f++) {
                p = orc_panel_boundary(P, c1, f, p1, p2);
                if (p > best) best = p;
            }
    }
    return best;
}

/* histogram of panel types for cells [start,end): hist[0..2] = identical,
 * common edge (2D) / vertex..., hist[3+p] = regular order p (p < MAX_ORDER) */
void orc_histogram(const orc_problem *P, int start, int end, int64_t *hist)
{
    int c1;
#pragma omp parallel
    {
        int64_t *loc = calloc(4 + MAX_ORDER, sizeof(int64_t));
        int k;
#pragma omp for schedule(dynamic, 8)
        for (c1 = start; c1 < end; c1++) {
            int p1[MAXV], p2[MAXV], c2;
            for (c2 = c1; c2 < P->nc; c2++) {
                int p = orc_panel_interior(P, c1, c2, p1, p2);
                if (p >= -3 && p < MAX_ORDER) loc[3 + p]++;
            }
        }
#pragma omp critical
        for (k = 0; k < 4 + MAX_ORDER; k++) hist[k] += loc[k];
        free(loc);
    }
}

/* nonlocalAssembly_{SCALAR}.pxi:204-221 */
static void scatter_sym(double *A, size_t ld, size_t wrap, const int32_t *ld_dofs, int n, const double *contrib, double fac)
{
    int p, q, k = 0;
    for (p = 0; p < n; p++) {
        int I = ld_dofs[p];
        if (I >= 0) {
            size_t a = (size_t)I * ld + I;
            A[wrap ? a & (wrap - 1) : a] += fac * contrib[k];
            k++;
            for (q = p + 1; q < n; q++) {
                int J = ld_dofs[q];
                if (J >= 0) {
                    size_t b = (size_t)I * ld + J, c = (size_t)J * ld + I;
                    A[wrap ? b & (wrap - 1) : b] += fac * contrib[k];
                    A[wrap ? c & (wrap - 1) : c] += fac * contrib[k];
                }
                k++;
            }
        } else {
            k += n - p;
        }
    }
}

/* Dense assembly of the cell slice [start,end) x [c1, nc) -- the work of one
 * MPI rank in the reference (nonlocalAssembly_{SCALAR}.pxi:1280-1285, 1386-1448).
 * A is num_dofs x num_dofs, zero-initialised by the caller; partial results
 * of different slices add up (the reference's Allreduce, :1450).
 *
 * wrap != 0: throughput-sampling mode for the CPU baseline: A has `wrap`
 * (power of two) entries and flat indices are wrapped into it, so that a slice
 * of a problem whose full matrix does not fit host memory can still be timed
 * with the scatter's read-modify-write traffic in place.  The values are then
 * meaningless.
 *
 * Threads: the c1 loop is split over OpenMP threads; each thread owns a
 * private copy of A when nthreads_private != 0 (summed at the end), else
 * updates are done with atomics.  Returns number of evaluated pairs. */
int64_t orc_dense(const orc_problem *P, int start, int end, int zero_exterior, double *A, size_t wrap, int use_atomic)
{
    int nvc = P->dim + 1;
    size_t ld = (size_t)P->num_dofs;
    size_t total = wrap ? wrap : ld * ld;
    int64_t npairs = 0;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if (nthreads == 1) use_atomic = 0;
#pragma omp parallel reduction(+ : npairs)
    {
        double *Aloc = A;
        double contrib[21], bcontrib[6];
        int c1;
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        if (nthreads > 1 && !use_atomic && tid > 0) Aloc = calloc(total, sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (c1 = start; c1 < end; c1++) {
            int p1[MAXV], p2[MAXV], c2, f, k;
            int32_t ldofs[2 * MAXV];
            for (c2 = c1; c2 < P->nc; c2++) {
                int skip = 1, panel;
                for (k = 0; k < nvc; k++) {
                    ldofs[k] = P->dofs[(size_t)c1 * nvc + k];
                    ldofs[nvc + k] = P->dofs[(size_t)c2 * nvc + k];
                    skip = skip && ldofs[k] < 0 && ldofs[nvc + k] < 0;
                }
                if (skip) continue;
                if (P->labels && P->pair_class[P->labels[c1] * 4 + P->labels[c2]] != P->active_class) continue;
                if (P->labels && P->pair_orientation && c1 != c2) {
                    /* second visit of the unsymmetric loop: (c2, c1) */
                    for (k = 0; k < nvc; k++) {
                        ldofs[k] = P->dofs[(size_t)c2 * nvc + k];
                        ldofs[nvc + k] = P->dofs[(size_t)c1 * nvc + k];
                    }
                    panel = orc_panel_interior_any(P, c2, c1, p1, p2);
                    if (panel == IGNORED) continue;
                    orc_local_interior(P, c2, c1, panel, p1, p2, contrib);
                } else {
                panel = orc_panel_interior(P, c1, c2, p1, p2);
                if (panel == IGNORED) continue;
                orc_local_interior(P, c1, c2, panel, p1, p2, contrib);
                }
                npairs++;
                if (use_atomic) {
                    /* same scatter, atomic adds */
                    int p, q, kk = 0, n = 2 * nvc;
                    double fac = c1 == c2 ? 1. : 2.;
                    for (p = 0; p < n; p++) {
                        int I = ldofs[p];
                        if (I >= 0) {
                            size_t a = (size_t)I * ld + I;
                            if (wrap) a &= wrap - 1;
#pragma omp atomic
                            A[a] += fac * contrib[kk];
                            kk++;
                            for (q = p + 1; q < n; q++) {
                                int J = ldofs[q];
                                if (J >= 0) {
                                    size_t b = (size_t)I * ld + J, c = (size_t)J * ld + I;
                                    if (wrap) { b &= wrap - 1; c &= wrap - 1; }
#pragma omp atomic
                                    A[b] += fac * contrib[kk];
#pragma omp atomic
                                    A[c] += fac * contrib[kk];
                                }
                                kk++;
                            }
                        } else kk += n - p;
                    }
                } else {
                    scatter_sym(Aloc, ld, wrap, ldofs, 2 * nvc, contrib, c1 == c2 ? 1. : 2.);
                }
            }
            if (zero_exterior) {
                for (k = 0; k < nvc; k++) ldofs[k] = P->dofs[(size_t)c1 * nvc + k];
                for (f = 0; f < P->nb; f++) {
                    int panel;
                    if (P->labels && P->bpair_class[P->labels[c1] * 4 + P->blabels[f]] != P->active_class) continue;
                    panel = orc_panel_boundary(P, c1, f, p1, p2);
                    orc_local_boundary(P, c1, f, panel, p1, p2, bcontrib);
                    if (use_atomic) {
                        int p, q, kk = 0;
                        for (p = 0; p < nvc; p++) {
                            int I = ldofs[p];
                            if (I >= 0) {
                                size_t a = (size_t)I * ld + I;
                                if (wrap) a &= wrap - 1;
#pragma omp atomic
                                A[a] += bcontrib[kk];
                                kk++;
                                for (q = p + 1; q < nvc; q++) {
                                    int J = ldofs[q];
                                    if (J >= 0) {
                                        size_t b = (size_t)I * ld + J, c = (size_t)J * ld + I;
                                        if (wrap) { b &= wrap - 1; c &= wrap - 1; }
#pragma omp atomic
                                        A[b] += bcontrib[kk];
#pragma omp atomic
                                        A[c] += bcontrib[kk];
                                    }
                                    kk++;
                                }
                            } else kk += nvc - p;
                        }
                    } else {
                        scatter_sym(Aloc, ld, wrap, ldofs, nvc, bcontrib, 1.);
                    }
                }
            }
        }
        if (Aloc != A) {
            size_t i;
#pragma omp critical
            for (i = 0; i < total; i++) A[i] += Aloc[i];
            free(Aloc);
        }
    }
    return npairs;
}

/* mesh.hVector / mesh.h / mesh.hmin, hdeltaCy (fem/PyNucleus_fem/meshCy.pyx:1654-1732): per cell the longest edge
 * (hVec), over the mesh the longest (h) and the SHORTEST edge (hmin; :1724 takes the min over ALL edges).  An edge
 * length is sqrt(mydot(e, e)); mydot is the BLAS ddot (base/PyNucleus_base/opt_true_blas.pxi:125-141) and the
 * OpenBLAS behind scipy accumulates it with fused multiply-adds: e1*e1 + fl(e0*e0) rounded once.  Pinned against the
 * hVector / hmin arrays of every fixture in tests/golden (tests/test_oracle_golden.py). */
void orc_edge_lengths(int dim, int nc, const double *vertices, const int32_t *cells, double *h, double *hmax_out, double *hmin_out)
{
    double hmax = 0., hmin = 100.;
    for (int c = 0; c < nc; c++) {
        double hl = 0.;
        if (dim == 1) {
            hl = fabs(vertices[cells[2 * (size_t)c + 1]] - vertices[cells[2 * (size_t)c]]);
            if (hl < hmin) hmin = hl;
        } else {
            const double *v0 = vertices + 2 * (size_t)cells[3 * (size_t)c], *v1 = vertices + 2 * (size_t)cells[3 * (size_t)c + 1],
                         *v2 = vertices + 2 * (size_t)cells[3 * (size_t)c + 2];
            const double e[3][2] = {{v2[0] - v1[0], v2[1] - v1[1]}, {v2[0] - v0[0], v2[1] - v0[1]}, {v1[0] - v0[0], v1[1] - v0[1]}};
            for (int j = 0; j < 3; j++) {
                const double hS = sqrt(fma(e[j][1], e[j][1], e[j][0] * e[j][0]));
                if (hS < hmin) hmin = hS;
                if (hS > hl) hl = hS;
            }
        }
        if (hl > hmax) hmax = hl;
        h[c] = hl;
    }
    *hmax_out = hmax;
    *hmin_out = hmin;
}
