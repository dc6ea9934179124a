// Host-side tables of the element path: quadrature rules and reference basis tables.
//
// Replaces (reference, AniFem++): tetrahedron_quadrature_formulas (fem/quadrature_formulas.cpp:526-1503),
// the value/gradient parts of Operator<IDEN|GRAD, FemFix<FEM_P0..P3>>::apply that do not depend on the
// element (fem/spaces/poly_0.h:62-82, poly_1.h:24-40, poly_2.h:35-40,72-88, poly_3.h:34-47,83-108) and
// Operator<>::Nfa/Dim (fem/operators.h:127-131,320-324).
//
// Design: instead of one hand-written formula per space, every P_k space is the Lagrange basis on the
// principal lattice of order k.  A basis function is named by a multi-index a = (a0,a1,a2,a3), |a| = k,
// and  phi_a(l) = prod_i prod_{m < a_i} (k*l_i - m)/(m + 1).  The multi-index lists below reproduce the
// reference's local dof order (vertices; edges 01,02,03,12,13,23 with the pair ordered towards the first
// endpoint first; faces 012,123,023,013: fem/fem_space.h:27-69, poly_3.h:31-47).  The tables are evaluated
// once per (space, rule) on the host and uploaded; the kernels only see phi[n][i] and the reference-cell
// gradients G[n][i][3] (d/dx^ on the unit tetrahedron, grad l0 = (-1,-1,-1), grad l_m = e_m).
#include <cstring>

#include "afb_internal.h"
#include "tet_quadrature.inc"
#include "tri_quadrature.inc"

namespace afb {

namespace {
struct Lattice {
    int k, nf;
    unsigned char a[AFB_MAX_BASE_NF][4];
};

const Lattice L0 = {0, 1, {{0, 0, 0, 0}}};
const Lattice L1 = {1, 4, {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
const Lattice L2 = {2, 10, {{2, 0, 0, 0}, {0, 2, 0, 0}, {0, 0, 2, 0}, {0, 0, 0, 2},
                            {1, 1, 0, 0}, {1, 0, 1, 0}, {1, 0, 0, 1}, {0, 1, 1, 0}, {0, 1, 0, 1}, {0, 0, 1, 1}}};
const Lattice L3 = {3, 20, {{3, 0, 0, 0}, {0, 3, 0, 0}, {0, 0, 3, 0}, {0, 0, 0, 3},
                            {2, 1, 0, 0}, {1, 2, 0, 0}, {2, 0, 1, 0}, {1, 0, 2, 0}, {2, 0, 0, 1}, {1, 0, 0, 2},
                            {0, 2, 1, 0}, {0, 1, 2, 0}, {0, 2, 0, 1}, {0, 1, 0, 2}, {0, 0, 2, 1}, {0, 0, 1, 2},
                            {1, 1, 1, 0}, {0, 1, 1, 1}, {1, 0, 1, 1}, {1, 1, 0, 1}}};

const Lattice* lattice(int fem) {
    switch (fem) {
        case AFB_FEM_P0: return &L0;
        case AFB_FEM_P1: return &L1;
        case AFB_FEM_P2: return &L2;
        case AFB_FEM_P3: return &L3;
    }
    return nullptr;
}

// factor_i(l) = prod_{m < a} (k*l - m)/(m+1) and its derivative w.r.t. l
inline void lattice_factor(int k, int a, double l, double* val, double* der) {
    double v = 1, d = 0;
    for (int m = 0; m < a; ++m) {
        const double t = (k * l - m) / (m + 1), dt = double(k) / (m + 1);
        d = d * t + v * dt;
        v = v * t;
    }
    *val = v;
    *der = d;
}
}  // namespace

int resolve_op(int op, int fem, int vec, OpInfo* o) {
    const Lattice* L = lattice(fem);
    if (!L || (vec != 1 && vec != 3)) return -3;
    o->op = op; o->fem = fem; o->vec = vec; o->nf_base = L->nf;
    switch (op) {
        case AFB_IDEN: o->nfa = vec * L->nf; o->dim = vec; o->dim_base = 1; return 0;
        case AFB_GRAD: o->nfa = vec * L->nf; o->dim = 3 * vec; o->dim_base = 3; return 0;
        case AFB_DIV:
            if (vec != 3 || fem == AFB_FEM_P0) return -3;
            o->nfa = 3 * L->nf; o->dim = 1; o->dim_base = 3; return 0;
    }
    return -3;
}

void basis_values(int fem, int q, const double* XYL, double* phi) {
    const Lattice* L = lattice(fem);
    for (int n = 0; n < q; ++n)
        for (int i = 0; i < L->nf; ++i) {
            double v = 1;
            for (int c = 0; c < 4; ++c) {
                double f, d;
                lattice_factor(L->k, L->a[i][c], XYL[4 * n + c], &f, &d);
                v *= f;
            }
            phi[n * L->nf + i] = v;
        }
}

void basis_ref_grads(int fem, int q, const double* XYL, double* G) {
    const Lattice* L = lattice(fem);
    for (int n = 0; n < q; ++n)
        for (int i = 0; i < L->nf; ++i) {
            double f[4], d[4];
            for (int c = 0; c < 4; ++c) lattice_factor(L->k, L->a[i][c], XYL[4 * n + c], &f[c], &d[c]);
            double dl[4];  // d phi / d lambda_c
            for (int c = 0; c < 4; ++c) {
                double p = d[c];
                for (int c2 = 0; c2 < 4; ++c2) if (c2 != c) p *= f[c2];
                dl[c] = p;
            }
            for (int dd = 0; dd < 3; ++dd) G[(n * L->nf + i) * 3 + dd] = dl[dd + 1] - dl[0];
        }
}

int tet_rule(int order, const double** p, const double** w) {
    if (order < 0 || order > AFB_TETQ_MAX_ORDER) return -1;
    *p = AFB_TETQ_P + 4 * AFB_TETQ_OFFS[order];
    *w = AFB_TETQ_W + AFB_TETQ_OFFS[order];
    return AFB_TETQ_NPTS[order];
}

// triangle rule behind fem3Dface (reference: triangle_quadrature_formulas, fem/quadrature_formulas.cpp:109-516)
int tri_rule(int order, const double** p, const double** w) {
    if (order < 0 || order > AFB_TRIQ_MAX_ORDER) return -1;
    *p = AFB_TRIQ_P + 3 * AFB_TRIQ_OFFS[order];
    *w = AFB_TRIQ_W + AFB_TRIQ_OFFS[order];
    return AFB_TRIQ_NPTS[order];
}

}  // namespace afb

extern "C" int afb_tri_quadrature(int order, double* p, double* w, int capacity) {
    const double *pp, *ww;
    int q = afb::tri_rule(order, &pp, &ww);
    if (q < 0) return -7;
    if (p && w) {
        if (capacity < q) return -7;
        std::memcpy(p, pp, sizeof(double) * 3 * q);
        std::memcpy(w, ww, sizeof(double) * q);
    }
    return q;
}

extern "C" int afb_op_dims(int op, int fem, int vec, int* nfa, int* dim) {
    afb::OpInfo o;
    int rc = afb::resolve_op(op, fem, vec, &o);
    if (rc) return rc;
    if (nfa) *nfa = o.nfa;
    if (dim) *dim = o.dim;
    return 0;
}

extern "C" int afb_tet_quadrature(int order, double* p, double* w, int capacity) {
    const double *pp, *ww;
    int q = afb::tet_rule(order, &pp, &ww);
    if (q < 0) return -7;
    if (p && w) {
        if (capacity < q) return -7;
        std::memcpy(p, pp, sizeof(double) * 4 * q);
        std::memcpy(w, ww, sizeof(double) * q);
    }
    return q;
}
