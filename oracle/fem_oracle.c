/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call this file.  The product
 * path (inmost-fem_b200/) never does.
 *
 * fem_oracle.c: plain-C restatement of the reference's element-matrix algorithm
 *     A_r = V_r^T * (w_n |T_r| D U_r),   r < f fused tetrahedra
 * i.e. Ani::fem3Dtet<OpA,OpB,Traits>  (anifem++/fem/operations/int_tet.inl:30-57) ->
 *      internalFem3DtetGeomInit        (anifem++/fem/operations/core.inl:217-275) ->
 *      internalFem3Dtet                (anifem++/fem/operations/core.inl:277-367).
 * Parity status: PINNED at element level -- checked against the reference's own golden tables
 * (tests/fem/operations/int_tet_test.cpp:230-244, :334-347, :455; predefined_spaces_test.cpp) and
 * against the reference itself compiled here (oracle/_ref), see tests/test_oracle_*.py.
 *
 * Layouts (all column-major, FP64), identical to the reference:
 *   XYk : 3 x f                              (geometry.h:108-122)
 *   U   : U[k + dim*(n + q*(i + nfa*r))]     (spaces/poly_2.h:94)
 *   A   : A[ib + nfB*(ia + nfA*r)]           (core.inl:47, core.h:41)  rows = test (OpB), cols = trial (OpA)
 *   D   : user-callback layout, col-major (jdim x idim): D[k + jdim*j] = K(k,j)   (diff_tensor.h:73-83)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../inmost-fem_b200/csrc/tet_quadrature.inc"
#include "../inmost-fem_b200/csrc/tri_quadrature.inc"

enum { OP_IDEN = 1, OP_GRAD = 2, OP_DIV = 3 };                         /* operators.h:36-44 */
enum { FEM_P0 = 1, FEM_P1 = 2, FEM_P2 = 3, FEM_P3 = 4 };               /* operators.h:24-34 */
enum { T_NULL = 1, T_SCALAR = 2, T_SYMMETRIC = 3, T_GENERAL = 4 };     /* diff_tensor.h:17-22 */
enum { L_CONST = 0, L_PER_TET = 1, L_PER_POINT = 2 };

static int base_nf(int fem) { return fem == FEM_P0 ? 1 : fem == FEM_P1 ? 4 : fem == FEM_P2 ? 10 : fem == FEM_P3 ? 20 : -1; }

/* Nfa / Dim of Operator<op, FemFix|FemVec>   (operators.h:127-131, :320-324; spaces/poly_*.h) */
int orc_op_dims(int op, int fem, int vec, int* nfa, int* dim) {
    int nf = base_nf(fem);
    if (nf < 0 || (vec != 1 && vec != 3)) return -3;
    if (op == OP_IDEN) { *nfa = vec * nf; *dim = vec; return 0; }
    if (op == OP_GRAD) { *nfa = vec * nf; *dim = 3 * vec; return 0; }
    if (op == OP_DIV && vec == 3 && fem != FEM_P0) { *nfa = 3 * nf; *dim = 1; return 0; }
    return -3;
}

int orc_tet_quadrature(int order, const double** p, const double** w) {
    if (order < 0 || order > AFB_TETQ_MAX_ORDER) return -1;   /* quadrature_formulas.cpp:1498-1499 */
    *p = AFB_TETQ_P + 4 * AFB_TETQ_OFFS[order];
    *w = AFB_TETQ_W + AFB_TETQ_OFFS[order];
    return AFB_TETQ_NPTS[order];
}

/* geometry.h:25-45 */
static double inverse3x3(const double* m, double* inv) {
#define M(I, J) m[(I) + 3 * (J)]
    double da = M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1);
    double db = M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2);
    double dc = M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0);
    double det = M(0, 0) * da + M(0, 1) * db + M(0, 2) * dc;
    inv[0 + 3 * 0] = da / det;
    inv[1 + 3 * 0] = db / det;
    inv[2 + 3 * 0] = dc / det;
    for (int i = 1; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            inv[j + 3 * i] = (M((i + 1) % 3, (j + 1) % 3) * M((i + 2) % 3, (j + 2) % 3) -
                              M((i + 1) % 3, (j + 2) % 3) * M((i + 2) % 3, (j + 1) % 3)) / det;
#undef M
    return det;
}

/* basis values phi[n + q*i] at the barycentric points XYL[4*n..]
 * (poly_0.h:77, poly_1.h:31-33, poly_2.h:35-40, poly_3.h:34-47) */
static void base_values(int fem, int q, const double* XYL, double* phi) {
    static const int I6[6] = {0, 0, 0, 1, 1, 2}, J6[6] = {1, 2, 3, 2, 3, 3};
    static const int IPF[16] = {0, 1, 2, 3, 1, 2, 3, 0, 0, 2, 3, 1, 0, 1, 3, 2};
    for (int n = 0; n < q; ++n) {
        const double* l = XYL + 4 * n;
        switch (fem) {
            case FEM_P0: phi[n] = 1; break;
            case FEM_P1: for (int i = 0; i < 4; ++i) phi[n + q * i] = l[i]; break;
            case FEM_P2:
                for (int i = 0; i < 4; ++i) phi[n + q * i] = l[i] * (2 * l[i] - 1);
                for (int e = 0; e < 6; ++e) phi[n + q * (4 + e)] = 4 * l[I6[e]] * l[J6[e]];
                break;
            case FEM_P3:
                for (int i = 0; i < 4; ++i) phi[n + q * i] = l[i] * (3 * l[i] - 1) * (3 * l[i] - 2) / 2;
                for (int e = 0; e < 6; ++e) {
                    double s1 = l[I6[e]], s2 = l[J6[e]];
                    phi[n + q * (4 + 2 * e)] = s1 * (3 * s1 - 1) * s2 * 4.5;
                    phi[n + q * (4 + 2 * e + 1)] = s1 * (3 * s2 - 1) * s2 * 4.5;
                }
                for (int fc = 0; fc < 4; ++fc)
                    phi[n + q * (16 + fc)] = 27 * l[IPF[4 * fc]] * l[IPF[4 * fc + 1]] * l[IPF[4 * fc + 2]];
                break;
        }
    }
}

/* reference-cell gradients G[d + 3*(i + nf*n)] = d phi_i / d(x^) on the unit tet, where
 * grad lambda = GRAD_P1 (poly_1.h:107-118 in reference coordinates; poly_2.h:72-88; poly_3.h:83-108) */
static void base_ref_grads(int fem, int q, const double* XYL, double* G) {
    static const double GP1[12] = {-1, -1, -1, 1, 0, 0, 0, 1, 0, 0, 0, 1};
    static const int I6[6] = {0, 0, 0, 1, 1, 2}, J6[6] = {1, 2, 3, 2, 3, 3};
    static const int IPF[16] = {0, 1, 2, 3, 1, 2, 3, 0, 0, 2, 3, 1, 0, 1, 3, 2};
    int nf = base_nf(fem);
    for (int n = 0; n < q; ++n) {
        const double* l = XYL + 4 * n;
        double* g = G + 3 * nf * n;
        switch (fem) {
            case FEM_P0: g[0] = g[1] = g[2] = 0; break;
            case FEM_P1: for (int i = 0; i < 12; ++i) g[i] = GP1[i]; break;
            case FEM_P2:
                for (int i = 0; i < 4; ++i)
                    for (int d = 0; d < 3; ++d) g[d + 3 * i] = GP1[d + 3 * i] * (4 * l[i] - 1);
                for (int e = 0; e < 6; ++e) {
                    double s1 = l[I6[e]], s2 = l[J6[e]];
                    for (int d = 0; d < 3; ++d) g[d + 3 * (4 + e)] = 4 * (GP1[d + 3 * J6[e]] * s1 + GP1[d + 3 * I6[e]] * s2);
                }
                break;
            case FEM_P3:
                for (int i = 0; i < 4; ++i) {
                    double s1 = l[i];
                    for (int d = 0; d < 3; ++d) g[d + 3 * i] = GP1[d + 3 * i] * ((13.5 * s1 - 9) * s1 + 1);
                }
                for (int e = 0; e < 6; ++e) {
                    double s1 = l[I6[e]], s2 = l[J6[e]];
                    for (int d = 0; d < 3; ++d) {
                        g[d + 3 * (4 + 2 * e)] = (GP1[d + 3 * I6[e]] * (6 * s1 - 1) * s2 + GP1[d + 3 * J6[e]] * (3 * s1 - 1) * s1) * 4.5;
                        g[d + 3 * (4 + 2 * e + 1)] = (GP1[d + 3 * J6[e]] * (6 * s2 - 1) * s1 + GP1[d + 3 * I6[e]] * (3 * s2 - 1) * s2) * 4.5;
                    }
                }
                for (int fc = 0; fc < 4; ++fc) {
                    int a = IPF[4 * fc], b = IPF[4 * fc + 1], c = IPF[4 * fc + 2];
                    double s1 = l[a], s2 = l[b], s3 = l[c];
                    for (int d = 0; d < 3; ++d)
                        g[d + 3 * (16 + fc)] = 27 * (GP1[d + 3 * a] * s2 * s3 + s1 * GP1[d + 3 * b] * s3 + s1 * s2 * GP1[d + 3 * c]);
                }
                break;
        }
    }
}

/* Dense U[k + dim*(n + q*i)] of Operator<op, FemFix|FemVec> on ONE tet with inverse Jacobian PSI.
 * Vector spaces: block-diagonal expansion of the scalar table (operators.h:140-152);
 * DIV: U(n, i + nf*k) = dphi_i/dx_k (operators.h:337-346). */
static void apply_op(int op, int fem, int vec, int q, const double* XYL, const double* PSI, double* U,
                     double* scratch /* >= 4*q*nf */) {
    int nf = base_nf(fem), nfa, dim;
    orc_op_dims(op, fem, vec, &nfa, &dim);
    memset(U, 0, sizeof(double) * (size_t)dim * q * nfa);
    if (op == OP_IDEN) {
        double* phi = scratch;
        base_values(fem, q, XYL, phi);
        for (int c = 0; c < vec; ++c)
            for (int i = 0; i < nf; ++i)
                for (int n = 0; n < q; ++n) U[c + dim * (n + q * (i + nf * c))] = phi[n + q * i];
        return;
    }
    /* physical gradients gp[k + 3*(n + q*i)] = sum_j PSI[j + 3k] * G[j + 3*(i + nf*n)]  (poly_2.h:89-97) */
    double* G = scratch;
    double* gp = scratch + 3 * q * nf;
    base_ref_grads(fem, q, XYL, G);
    for (int i = 0; i < nf; ++i)
        for (int n = 0; n < q; ++n)
            for (int k = 0; k < 3; ++k) {
                double s = 0;
                for (int j = 0; j < 3; ++j) s += PSI[j + 3 * k] * G[j + 3 * (i + nf * n)];
                gp[k + 3 * (n + q * i)] = (fem == FEM_P0) ? 0.0 : s;
            }
    if (op == OP_GRAD) {
        for (int c = 0; c < vec; ++c)
            for (int i = 0; i < nf; ++i)
                for (int n = 0; n < q; ++n)
                    for (int k = 0; k < 3; ++k) U[3 * c + k + dim * (n + q * (i + nf * c))] = gp[k + 3 * (n + q * i)];
    } else { /* DIV of a 3-vector */
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < nf; ++i)
                for (int n = 0; n < q; ++n) U[n + q * (i + nf * k)] = gp[k + 3 * (n + q * i)];
    }
}

typedef struct orc_form {
    int opA, femA, vecA; /* trial: columns of A */
    int opB, femB, vecB; /* test:  rows of A */
    int quad_order;
    int tensor_type;   /* T_* */
    int tensor_layout; /* L_* */
    const double* D;   /* user layout, see header comment */
} orc_form;

/* physical quadrature points XYG[k + 3*(n + q*r)]  (core.inl:249-269) */
int orc_quad_points(int order, long f, const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* XYG) {
    const double *p, *w;
    int q = orc_tet_quadrature(order, &p, &w);
    if (q < 0) return -1;
    for (long r = 0; r < f; ++r)
        for (int n = 0; n < q; ++n)
            for (int k = 0; k < 3; ++k) {
                double s = XY0[k + 3 * r];
                const double* X[3] = {XY1, XY2, XY3};
                for (int l = 0; l < 3; ++l) s += p[l + 1 + 4 * n] * (X[l][k + 3 * r] - XY0[k + 3 * r]);
                XYG[k + 3 * (n + q * r)] = s;
            }
    return q;
}

/* triangle rule of fem3Dface: p[3*q] barycentric, w[q] (quadrature_formulas.cpp:109-516) */
int orc_tri_quadrature(int order, const double** p, const double** w) {
    if (order < 0 || order > AFB_TRIQ_MAX_ORDER) return -1;
    *p = AFB_TRIQ_P + 3 * AFB_TRIQ_OFFS[order];
    *w = AFB_TRIQ_W + AFB_TRIQ_OFFS[order];
    return AFB_TRIQ_NPTS[order];
}

/* area of the triangle p0 p1 p2 (fem/geometry.h:67-72) */
static double tri_area(const double* p0, const double* p1, const double* p2) {
    double a[3] = {p0[0] - p2[0], p0[1] - p2[1], p0[2] - p2[2]}, b[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double c[3] = {a[1] * b[2] - a[2] * b[1], -a[0] * b[2] + a[2] * b[0], a[0] * b[1] - a[1] * b[0]};
    return sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / 2;
}

/* Element matrices for f tets: volume integrals (face == NULL, fem3Dtet) or surface integrals over face face[r] of tet r
 * (fem3Dface, fem/operations/int_face.inl:160-199: the triangle rule is lifted to the face {face, face+1, face+2 mod 4} with
 * zero barycentric weight on vertex face+3, and the measure is the face area).  Returns 0, or <0 on bad arguments (-3
 * unsupported space/operator, -5 identity/scalar tensor with incompatible operator dimensions: diff_tensor.h:315-317). */
static int fem3d_core(const orc_form* fm, long f, const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A,
                      const int* face) {
    int nfa, idim, nfb, jdim;
    if (orc_op_dims(fm->opA, fm->femA, fm->vecA, &nfa, &idim)) return -3;
    if (orc_op_dims(fm->opB, fm->femB, fm->vecB, &nfb, &jdim)) return -3;
    const double *XYL, *W;
    int q;
    double* XYLf = NULL;
    if (face) {
        q = orc_tri_quadrature(fm->quad_order, &XYL, &W);
        if (q < 0) return -1;
        XYLf = (double*)malloc(sizeof(double) * 4 * (size_t)q);
    } else q = orc_tet_quadrature(fm->quad_order, &XYL, &W);
    if (q < 0) return -1;
    const double* XYLt = XYL;   /* triangle points (face mode) */
    int tt = fm->tensor_type;
    if ((tt == T_NULL || tt == T_SCALAR) && jdim != idim && (nfa != 1 || idim != 1)) return -5;
    int same = (fm->opA == fm->opB && fm->femA == fm->femB && fm->vecA == fm->vecB);
    int nfmax = 20;
    double* U = (double*)malloc(sizeof(double) * (size_t)idim * q * nfa);
    double* V = same ? U : (double*)malloc(sizeof(double) * (size_t)jdim * q * nfb);
    double* DU = (double*)malloc(sizeof(double) * (size_t)jdim * q * nfa);
    double* scratch = (double*)malloc(sizeof(double) * (size_t)8 * q * nfmax);
    long dlen = (tt == T_SCALAR) ? 1 : (tt == T_NULL ? 0 : (long)idim * jdim);
    for (long r = 0; r < f; ++r) {
        /* core.inl:231-242 */
        double XYP[9], PSI[9];
        for (int i = 0; i < 3; ++i) {
            XYP[i + 0] = XY1[i + 3 * r] - XY0[i + 3 * r];
            XYP[i + 3] = XY2[i + 3 * r] - XY0[i + 3 * r];
            XYP[i + 6] = XY3[i + 3 * r] - XY0[i + 3 * r];
        }
        double det = inverse3x3(XYP, PSI);
        double vol = fabs(det) / 6;
        if (face) {
            const int fc = face[r];
            if (fc < 0 || fc > 3) { free(U); if (!same) free(V); free(DU); free(scratch); free(XYLf); return -7; }
            for (int n = 0; n < q; ++n) {   /* int_face.inl:175-180 */
                XYLf[4 * n + fc] = XYLt[3 * n + 0];
                XYLf[4 * n + (fc + 1) % 4] = XYLt[3 * n + 1];
                XYLf[4 * n + (fc + 2) % 4] = XYLt[3 * n + 2];
                XYLf[4 * n + (fc + 3) % 4] = 0;
            }
            XYL = XYLf;
            /* vertices relative to P0 like mem.XYP (core.inl:231-240); int_face.inl:184-187 */
            double X[12] = {0, 0, 0, XYP[0], XYP[1], XYP[2], XYP[3], XYP[4], XYP[5], XYP[6], XYP[7], XYP[8]};
            vol = tri_area(X + 3 * fc, X + 3 * ((fc + 1) % 4), X + 3 * ((fc + 2) % 4));
        }
        apply_op(fm->opA, fm->femA, fm->vecA, q, XYL, PSI, U, scratch);
        if (!same) apply_op(fm->opB, fm->femB, fm->vecB, q, XYL, PSI, V, scratch);
        /* DU = w_n |T| D U  (diff_tensor.h:498-549 PerPoint path) */
        for (int i = 0; i < nfa; ++i)
            for (int n = 0; n < q; ++n) {
                const double* Dn = fm->D;
                if (fm->tensor_layout == L_PER_TET) Dn += dlen * r;
                if (fm->tensor_layout == L_PER_POINT) Dn += dlen * (n + (long)q * r);
                double wg = W[n];
                const double* u = U + idim * (n + q * i);
                double* du = DU + jdim * (n + q * i);
                if (tt == T_GENERAL || tt == T_SYMMETRIC) {
                    for (int k = 0; k < jdim; ++k) {
                        double s = 0;
                        for (int j = 0; j < idim; ++j) s += Dn[k + jdim * j] * u[j];
                        du[k] = wg * vol * s;
                    }
                } else {
                    double sc = (tt == T_SCALAR) ? Dn[0] : 1.0;
                    if (jdim == idim) for (int k = 0; k < jdim; ++k) du[k] = wg * vol * (sc * u[k]);
                    else for (int k = 0; k < jdim; ++k) du[k] = wg * vol * (sc * u[0]); /* P0 broadcast, diff_tensor.h:333-338 */
                }
            }
        /* A = V^T DU  (core.inl:38-47) */
        double* Ar = A + (long)nfa * nfb * r;
        int nrow = q * jdim;
        for (int ia = 0; ia < nfa; ++ia)
            for (int ib = 0; ib < nfb; ++ib) {
                double s = 0;
                for (int j = 0; j < nrow; ++j) s += DU[j + nrow * ia] * V[j + nrow * ib];
                Ar[ib + nfb * ia] = s;
            }
    }
    free(U); if (!same) free(V); free(DU); free(scratch); free(XYLf);
    return 0;
}

int orc_fem3dtet(const orc_form* fm, long f, const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A) {
    return fem3d_core(fm, f, XY0, XY1, XY2, XY3, A, NULL);
}

/* fem3Dface for f tets, face[r] in 0..3 = face {r, r+1, r+2 mod 4} of the tet (int_face.h:49-55) */
int orc_fem3dface(const orc_form* fm, long f, const int* face, const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                  double* A) {
    if (!face) return -7;
    return fem3d_core(fm, f, XY0, XY1, XY2, XY3, A, face);
}

/* U table of one operator on f tets with an explicit rule; layout U[k + dim*(n + q*(i + nfa*r))] */
int orc_operator_apply(int op, int fem, int vec, int q, const double* XYL, long f,
                       const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* Uout) {
    int nfa, dim;
    if (orc_op_dims(op, fem, vec, &nfa, &dim)) return -3;
    double* scratch = (double*)malloc(sizeof(double) * (size_t)8 * q * 20);
    for (long r = 0; r < f; ++r) {
        double XYP[9], PSI[9];
        for (int i = 0; i < 3; ++i) {
            XYP[i + 0] = XY1[i + 3 * r] - XY0[i + 3 * r];
            XYP[i + 3] = XY2[i + 3 * r] - XY0[i + 3 * r];
            XYP[i + 6] = XY3[i + 3 * r] - XY0[i + 3 * r];
        }
        inverse3x3(XYP, PSI);
        apply_op(op, fem, vec, q, XYL, PSI, Uout + (long)dim * q * nfa * r, scratch);
    }
    free(scratch);
    return 0;
}

/* ---- global scatter (restating inmost_interface/assembler.inl:397-425 on a pre-built CSR) ----
 * For element e: rows[i] / cols[j] are signed codes sign*(id+1), 0 = skip (assembler.inl:49-55).
 * matrix[r][c] += s_r*s_c*A[j*nRows+i] if |A| > drop_val, located by binary search in the sorted row
 * (the "is_mtx_include_template" branch :428-438); rhs[r] += s_r*F[i] (:407). Returns -1 on NaN/Inf (:419-424). */
int orc_scatter_csr(long ne, int nrow, int ncol, const long* rowcode, const long* colcode,
                    const double* A /* nrow x ncol x ne col-major */, const double* F /* nrow x ne or NULL */,
                    long row_begin, const long* rowptr, const int* colind, double* val, double* rhs, double drop_val) {
    int status = 0;
    for (long e = 0; e < ne; ++e) {
        const long* rc = rowcode + (long)nrow * e;
        const long* cc = colcode + (long)ncol * e;
        const double* Ae = A ? A + (long)nrow * ncol * e : NULL;
        for (int i = 0; i < nrow; ++i) {
            if (rc[i] == 0) continue;
            long rid = labs(rc[i]) - 1; int rs = rc[i] < 0 ? -1 : 1;
            if (F && rhs) {
                rhs[rid - row_begin] += rs * F[i + (long)nrow * e];
                if (!isfinite(F[i + (long)nrow * e])) status = -1;
            }
            if (!Ae) continue;
            long b = rowptr[rid - row_begin], en = rowptr[rid - row_begin + 1];
            for (int j = 0; j < ncol; ++j) {
                long cid = labs(cc[j]) - 1; int cs = cc[j] < 0 ? -1 : 1;
                double a = Ae[i + (long)nrow * j];
                if (!isfinite(a)) { status = -1; continue; }
                if (!(fabs(a) > drop_val)) continue;
                long lo = b, hi = en;
                while (lo < hi) { long mid = (lo + hi) / 2; if (colind[mid] < cid) lo = mid + 1; else hi = mid; }
                if (lo < en && colind[lo] == cid) val[lo] += rs * cs * a;
                else status = status ? status : -7; /* pattern does not include the entry */
            }
        }
    }
    return status;
}
