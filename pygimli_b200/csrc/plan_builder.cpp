// plan_builder.cpp -- compiled host-side set-up of the B200 ERT path: mesh + data container -> pgb200_plan.
//
// What ModellingBase::setMesh / setData (core/src/modellingbase.h:68-99) trigger in the reference, restated for flat
// arrays: sparsity pattern (sparsematrix.h:966-1032), electrode matching (bert/dcfemmodelling.cpp:845-940,
// bert/electrode.cpp:102-287), wavenumbers (bert/bertMisc.cpp:36-129, numericbase.cpp:51-150), mixed-boundary
// coefficients (dcfemmodelling.cpp:243-299, :430-506), Dirichlet rows (:141-161), background prolongation
// (modellingbase.cpp:401-497, mesh.cpp:2247-2316), Jacobian column segments (bert/bertJacobian.cpp:280-299), plus what
// only the GPU path needs: the space-filling-curve node order, cell colours and the aggregation hierarchy of the
// multilevel preconditioner.  pygimli_b200/host_setup.py is the numpy twin the tests compare this file with.
// Host only (g++, OpenMP); no CUDA here.
#include "../../include/pgb200_ert.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace {

struct StageTimer {
    bool on; std::chrono::steady_clock::time_point t;
    StageTimer() : on(getenv("PGB200_VERBOSE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) { if (!on) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "[plan_build] %-28s %.3f s\n", what, std::chrono::duration<double>(n - t).count()); t = n; }
};

constexpr double TOLERANCE = 1e-12;
constexpr int MARKER_NODE_ELECTRODE = -99, MARKER_NODE_REFERENCE = -999, MARKER_NODE_CALIBRATION = -1000;
constexpr int MARKER_BOUND_NEUMANN = -1, MARKER_BOUND_MIXED = -2, MARKER_BOUND_DIRICHLET = -3;

typedef std::vector<int> IVec;
typedef std::vector<double> DVec;

// ---- Bessel functions: Abramowitz & Stegun 9.8.1-9.8.8, as numericbase.h:80-180 ------------------------------------
double bessel_i0(double x) {
    const double ax = std::fabs(x);
    if (ax < 3.75) {
        double y = x / 3.75; y *= y;
        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    const double y = 3.75 / ax;
    return (std::exp(ax) / std::sqrt(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
           y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}
double bessel_i1(double x) {
    const double ax = std::fabs(x);
    double res;
    if (ax < 3.75) {
        double y = x / 3.75; y *= y;
        res = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
    } else {
        const double y = 3.75 / ax;
        double r = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
        r = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * r))));
        res = r * (std::exp(ax) / std::sqrt(ax));
    }
    return x < 0.0 ? -res : res;
}
double bessel_k0(double x) {
    if (x <= 2.0) {
        const double y = x * x / 4.0;
        return (-std::log(x / 2.0) * bessel_i0(x)) + (-0.57721566 + y * (0.42278420 + y * (0.23069756 + y * (0.3488590e-1 +
               y * (0.262698e-2 + y * (0.10750e-3 + y * 0.74e-5))))));
    }
    const double y = 2.0 / x;
    return (std::exp(-x) / std::sqrt(x)) * (1.25331414 + y * (-0.7832358e-1 + y * (0.2189568e-1 + y * (-0.1062446e-1 +
           y * (0.587872e-2 + y * (-0.251540e-2 + y * 0.53208e-3))))));
}
double bessel_k1(double x) {
    if (x <= 2.0) {
        const double y = x * x / 4.0;
        return (std::log(x / 2.0) * bessel_i1(x)) + (1.0 / x) * (1.0 + y * (0.15443144 + y * (-0.67278579 + y * (-0.18156897 +
               y * (-0.1919402e-1 + y * (-0.110404e-2 + y * (-0.4686e-4)))))));
    }
    const double y = 2.0 / x;
    return (std::exp(-x) / std::sqrt(x)) * (1.25331414 + y * (0.23498619 + y * (-0.3655620e-1 + y * (0.1504268e-1 +
           y * (-0.780353e-2 + y * (0.325614e-2 + y * (-0.68245e-3)))))));
}

// ---- Gauss rules exactly as the reference iterates them (numericbase.cpp:51-150) -----------------------------------
void gauss_legendre(double x1, double x2, int n, DVec &x, DVec &w) {
    x.assign(n, 0.0); w.assign(n, 0.0);
    const double eps = 3.0e-6, m = (n + 1.0) / 2.0, xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
    for (int i = 1; i <= m; i++) {
        double z = std::cos(M_PI * (i - 0.25) / (n + 0.5)), z1 = z + 2.0 * eps, pp = 0.0;
        while (std::fabs(z - z1) > eps) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / (double)j; }
            pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
            z1 = z;
            z = z1 - p1 / pp;
        }
        x[i - 1] = xm - xl * z; x[n - i] = xm + xl * z;
        w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp); w[n - i] = w[i - 1];
    }
}
void gauss_laguerre(int n, DVec &x, DVec &w) {
    x.assign(n, 0.0); w.assign(n, 0.0);
    const double eps = 3.0e-11;
    double z = 0.0;
    for (int i = 1; i <= n; i++) {
        if (i == 1) z = 3.0 / (1.0 + 2.4 * n);
        else if (i == 2) z = z + 15.0 / (1.0 + 2.5 * n);
        else { const int ai = i - 2; z = z + (1.0 + 2.55 * ai) / (1.9 * ai) * (z - x[ai - 1]); }
        double pp = 0.0, p2 = 0.0;
        for (int it = 0; it < 20; it++) {
            double p1 = 1.0; p2 = 0.0;
            for (int j = 1; j <= n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1 - z) * p2 - (j - 1) * p3) / j; }
            pp = n * (p1 - p2) / z;
            const double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs(z - z1) <= eps) break;
        }
        x[i - 1] = z; w[i - 1] = -1.0 / (pp * n * p2);
    }
}
void kwave_from_range(double rmin, double rmax, int nleg, int nlag, DVec &k, DVec &w) {
    (void)rmax;
    const double k0 = 1.0 / (2.0 * rmin);
    DVec x, ww;
    gauss_legendre(0.0, 1.0, nleg, x, ww);
    k.clear(); w.clear();
    for (int i = 0; i < nleg; i++) { k.push_back(k0 * x[i] * x[i]); w.push_back(2.0 * k0 * x[i] * ww[i] / M_PI); }
    gauss_laguerre(nlag, x, ww);
    for (int i = 0; i < nlag; i++) { k.push_back(k0 * (x[i] + 1.0)); w.push_back(k0 * std::exp(x[i]) * ww[i] / M_PI); }
}
// bertMisc.cpp:36-129
std::string init_kwave_list(int dim, int ne, const double *sens, DVec &k, DVec &w) {
    if (dim == 3) { k.assign(1, 0.0); w.assign(1, 1.0); return ""; }
    if (ne < 2) return "need at least two sensors to initialise the wavenumber list";
    double dmin = 1e300, dmax = 0.0;
    for (int i = 0; i < ne; i++)
        for (int j = i + 1; j < ne; j++) {
            double s = 0.0;
            for (int d = 0; d < 3; d++) { const double t = sens[3 * i + d] - sens[3 * j + d]; s += t * t; }
            const double r = std::sqrt(s);
            dmin = std::min(dmin, r); dmax = std::max(dmax, r);
        }
    const double rmin = dmin / 2.0, rmax = dmax * 2.0;
    const int nleg = std::max((int)std::floor(6.0 * std::log10(rmax / rmin)), 4);
    kwave_from_range(rmin, rmax, nleg, 4, k, w);
    return "";
}

// unit mass matrices of the boundary faces (elementmatrix.cpp:94-136): int N_i N_j over a face of unit size
void unit_mass_face(int dim, int nlb, DVec &U) {
    U.assign((size_t)nlb * nlb, 0.0);
    auto at = [&](int i, int j) -> double & { return U[(size_t)i * nlb + j]; };
    if (dim == 2 && nlb == 2) { at(0, 0) = at(1, 1) = 2.0 / 6.0; at(0, 1) = at(1, 0) = 1.0 / 6.0; }
    else if (dim == 2 && nlb == 3) {
        const double m[3][3] = {{4.0, -1.0, 2.0}, {-1.0, 4.0, 2.0}, {2.0, 2.0, 16.0}};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) at(i, j) = m[i][j] / 30.0;
    } else if (dim == 3 && nlb == 3) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) at(i, j) = (1.0 + (i == j ? 1.0 : 0.0)) / 12.0;
    } else {   // tri6: corners 0,1,2; mids (0-1),(1-2),(2-0)
        double M[6][6] = {};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[i][j] = (i == j) ? 6.0 : -1.0;
        for (int i = 3; i < 6; i++) for (int j = 3; j < 6; j++) M[i][j] = (i == j) ? 32.0 : 16.0;
        const int opp[3] = {4, 5, 3};
        for (int v = 0; v < 3; v++) { M[v][opp[v]] = -4.0; M[opp[v]][v] = -4.0; }
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) at(i, j) = M[i][j] / 180.0;
    }
}

struct Face { double c[3], n[3], size; };
// centre, unit normal and size of a straight face from its corner nodes (first `dim` nodes)
Face face_geometry(int dim, const double *pos, const int *nodes) {
    Face f{};
    for (int d = 0; d < 3; d++) { double s = 0.0; for (int v = 0; v < dim; v++) s += pos[3 * (size_t)nodes[v] + d]; f.c[d] = s / (double)dim; }
    const double *p0 = pos + 3 * (size_t)nodes[0], *p1 = pos + 3 * (size_t)nodes[1];
    if (dim == 2) {
        const double tx = p1[0] - p0[0], ty = p1[1] - p0[1];
        f.size = std::sqrt(tx * tx + ty * ty);
        f.n[0] = ty / f.size; f.n[1] = -tx / f.size; f.n[2] = 0.0 / f.size;
    } else {
        const double *p2 = pos + 3 * (size_t)nodes[2];
        const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const double nv[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const double nn = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        f.size = 0.5 * nn;
        for (int d = 0; d < 3; d++) f.n[d] = nv[d] / nn;
    }
    return f;
}

// dcfemmodelling.cpp:430-506: mirror plane hard-wired at z (3-D) / y (2.5-D) = 0
double mixed_bc_beta(const Face &f, const double *source, double k) {
    const int dimc = k > 0 ? 1 : 2;
    double smir[3] = {source[0], source[1], source[2]};
    smir[dimc] = -smir[dimc];
    double r[3], rm[3];
    for (int d = 0; d < 3; d++) { r[d] = source[d] - f.c[d]; rm[d] = smir[d] - f.c[d]; }
    const double ra = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]), rma = std::sqrt(rm[0] * rm[0] + rm[1] * rm[1] + rm[2] * rm[2]);
    const double rn = std::fabs(r[0] * f.n[0] + r[1] * f.n[1] + r[2] * f.n[2]), rmn = std::fabs(rm[0] * f.n[0] + rm[1] * f.n[1] + rm[2] * f.n[2]);
    if (k == 0) return ((rma * rma) * rn / ra + (ra * ra) * rmn / rma) / (rma * ra * (ra + rma));
    const double k0a = bessel_k0(ra * k), k0m = bessel_k0(rma * k);
    if (std::fabs(k0a) < TOLERANCE || std::fabs(k0m) < TOLERANCE) return 0.0;
    return k * (rn / ra * bessel_k1(ra * k) + rmn / rma * bessel_k1(rma * k)) / (k0a + k0m);
}

// Lagrange shape functions at barycentric coordinates L (P1/P2 simplices)
void shape_functions(int nloc, int dim, const double *L, double *out) {
    if (nloc == dim + 1) { for (int i = 0; i <= dim; i++) out[i] = L[i]; return; }
    static const int e2[3][2] = {{0, 1}, {1, 2}, {2, 0}}, e3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {2, 3}, {3, 1}};
    for (int i = 0; i <= dim; i++) out[i] = L[i] * (2.0 * L[i] - 1.0);
    const int ne = dim == 2 ? 3 : 6;
    for (int e = 0; e < ne; e++) { const int a = dim == 2 ? e2[e][0] : e3[e][0], b = dim == 2 ? e2[e][1] : e3[e][1]; out[dim + 1 + e] = 4.0 * L[a] * L[b]; }
}

inline uint64_t spread_bits(uint64_t v, int dim, int bits) {
    uint64_t out = 0;
    for (int b = 0; b < bits; b++) out |= ((v >> b) & 1ull) << (dim * b);
    return out;
}

} // namespace

// =====================================================================================================================
struct pgb200_built_plan {
    pgb200_plan view{};
    std::string err;
    // geometry in the internal numbering
    int dim = 0, nloc = 0, N = 0, C = 0, nE = 0, nK = 0, M = 0, D = 0, nlb = 0, n_bounds = 0;
    long long nnz = 0;
    int topography = 0, neumann_domain = 0, has_background = 0, k_missing = 0, ref_node = -1, ref_last = 0;
    double surface_z = 0.0;
    IVec node_perm, node_inv;                     // perm[new] = old, inv[old] = new
    DVec pos; IVec node_marker, cells, cell_marker, bounds, bound_marker;
    IVec rowptr, colidx, diag_pos, ref_rowptr, ref_colidx; std::vector<long long> ref_slot;
    IVec color_ptr, color_order, cells_col, pos_col; int n_colors = 0;
    DVec kv, kw;
    IVec bc_slot, bc_ptr, bc_owner; DVec bc_coef;
    IVec dir_zero, dir_diag, dir_nodes;
    DVec el_pos, sing_val, pick_w, min_radius; IVec el_node, el_node_ref, el_cell, sing_node, pick_ptr, pick_idx, src_cell_ptr, src_cells;
    double source_center[3] = {0, 0, 0};
    IVec pro_level_ptr, pro_cells, pro_nb; DVec pro_w; int pro_nf = 0;
    IVec jac_cells, jac_col_ptr;
    IVec abmn; DVec kfac, sensors;
    // aggregation hierarchy (filled by pgb200_plan_build_hierarchy)
    struct Level { int n = 0; long long nnz = 0; IVec rowptr, colidx, diag_pos, gal_ptr, gal_idx, agg, mem_ptr, mem_idx; };
    std::vector<Level> levels; std::vector<pgb200_amg_level> level_views;
};

namespace {

thread_local std::string g_plan_err;

// CSR slot of (row, col); -1 if absent
inline int csr_find(const IVec &rowptr, const IVec &colidx, int row, int col) {
    const int *b = colidx.data() + rowptr[row], *e = colidx.data() + rowptr[row + 1];
    const int *p = std::lower_bound(b, e, col);
    return (p != e && *p == col) ? (int)(p - colidx.data()) : -1;
}

void coarsen(const IVec &rp, const IVec &ci, const IVec &agg, int nc, IVec &crp, IVec &cci, IVec &gal_ptr, IVec &gal_idx) {
    // pattern of P^T A P for piecewise-constant P and the gather lists; identical to amg_setup._coarsen: entries ordered by
    // (coarse row, coarse column), fine slots of one coarse entry ascending
    const int n = (int)rp.size() - 1;
    IVec mem_ptr(nc + 1, 0), mem(n);
    for (int i = 0; i < n; i++) mem_ptr[agg[i] + 1]++;
    for (int I = 0; I < nc; I++) mem_ptr[I + 1] += mem_ptr[I];
    { IVec fill(mem_ptr.begin(), mem_ptr.end() - 1); for (int i = 0; i < n; i++) mem[fill[agg[i]]++] = i; }
    crp.assign(nc + 1, 0); cci.clear(); gal_ptr.clear(); gal_idx.resize(ci.size());
    std::vector<std::pair<int, int>> ent;      // (coarse column, fine slot)
    size_t gpos = 0;
    for (int I = 0; I < nc; I++) {
        ent.clear();
        for (int q = mem_ptr[I]; q < mem_ptr[I + 1]; q++) { const int i = mem[q]; for (int p = rp[i]; p < rp[i + 1]; p++) ent.push_back({agg[ci[p]], p}); }
        std::sort(ent.begin(), ent.end());
        for (size_t e = 0; e < ent.size(); e++) {
            if (e == 0 || ent[e].first != ent[e - 1].first) { cci.push_back(ent[e].first); gal_ptr.push_back((int)gpos); }
            gal_idx[gpos++] = ent[e].second;
        }
        crp[I + 1] = (int)cci.size();
    }
    gal_ptr.push_back((int)gpos);
}

} // namespace

extern "C" {

const char *pgb200_plan_error(void) { return g_plan_err.c_str(); }

int pgb200_plan_free(pgb200_built_plan *P) { delete P; return 0; }
const pgb200_plan *pgb200_plan_view(const pgb200_built_plan *P) { return P ? &P->view : nullptr; }

// named arrays of a built plan (host pointers owned by the plan): type 0 = int32, 1 = float64, 2 = int64
int pgb200_plan_array(const pgb200_built_plan *P, const char *name, const void **ptr, long long *count, int *type) {
    if (!P || !name || !ptr || !count || !type) { g_plan_err = "null argument"; return 1; }
    const std::string s(name);
#define IARR(nm, v) if (s == nm) { *ptr = (v).data(); *count = (long long)(v).size(); *type = 0; return 0; }
#define DARR(nm, v) if (s == nm) { *ptr = (v).data(); *count = (long long)(v).size(); *type = 1; return 0; }
    IARR("node_perm", P->node_perm) IARR("node_inv", P->node_inv) DARR("pos", P->pos) IARR("cells", P->cells) IARR("cell_marker", P->cell_marker)
    IARR("node_marker", P->node_marker) IARR("bounds", P->bounds) IARR("bound_marker", P->bound_marker)
    IARR("rowptr", P->rowptr) IARR("colidx", P->colidx) IARR("diag_pos", P->diag_pos) IARR("ref_rowptr", P->ref_rowptr) IARR("ref_colidx", P->ref_colidx)
    IARR("color_ptr", P->color_ptr) IARR("color_order", P->color_order) IARR("cells_col", P->cells_col) IARR("pos_col", P->pos_col)
    DARR("k", P->kv) DARR("w", P->kw) IARR("bc_slot", P->bc_slot) IARR("bc_ptr", P->bc_ptr) IARR("bc_owner", P->bc_owner) DARR("bc_coef", P->bc_coef)
    IARR("dir_zero_slots", P->dir_zero) IARR("dir_diag_slots", P->dir_diag) IARR("dir_nodes", P->dir_nodes)
    DARR("el_pos", P->el_pos) DARR("sing_val", P->sing_val) DARR("pick_w", P->pick_w) DARR("min_radius", P->min_radius)
    IARR("el_node", P->el_node) IARR("el_node_ref", P->el_node_ref) IARR("el_cell", P->el_cell) IARR("sing_node", P->sing_node)
    IARR("pick_ptr", P->pick_ptr) IARR("pick_idx", P->pick_idx) IARR("src_cell_ptr", P->src_cell_ptr) IARR("src_cells", P->src_cells)
    IARR("pro_level_ptr", P->pro_level_ptr) IARR("pro_cells", P->pro_cells) IARR("pro_nb", P->pro_nb) DARR("pro_w", P->pro_w)
    IARR("jac_cells", P->jac_cells) IARR("jac_col_ptr", P->jac_col_ptr) IARR("abmn", P->abmn) DARR("k_fac", P->kfac)
#undef IARR
#undef DARR
    if (s == "ref_slot") { *ptr = P->ref_slot.data(); *count = (long long)P->ref_slot.size(); *type = 2; return 0; }
    if (s.rfind("level", 0) == 0) {        // "level<l>.<field>"
        const size_t dot = s.find('.');
        if (dot != std::string::npos) {
            const int l = atoi(s.substr(5, dot - 5).c_str());
            if (l >= 0 && l < (int)P->levels.size()) {
                const auto &L = P->levels[l];
                const std::string f = s.substr(dot + 1);
#define LARR(nm, v) if (f == nm) { *ptr = (v).data(); *count = (long long)(v).size(); *type = 0; return 0; }
                LARR("rowptr", L.rowptr) LARR("colidx", L.colidx) LARR("diag_pos", L.diag_pos) LARR("gal_ptr", L.gal_ptr) LARR("gal_idx", L.gal_idx)
                LARR("agg", L.agg) LARR("mem_ptr", L.mem_ptr) LARR("mem_idx", L.mem_idx)
#undef LARR
            }
        }
    }
    g_plan_err = "unknown plan array: " + s;
    return 1;
}
// scalars: "N" "C" "nnz" "nE" "nK" "M" "D" "dim" "nloc" "topography" "neumann_domain" "has_background" "k_missing" "n_colors" "n_levels"
// "surface_z" "pro_nf"
int pgb200_plan_scalar(const pgb200_built_plan *P, const char *name, double *out) {
    if (!P || !name || !out) { g_plan_err = "null argument"; return 1; }
    const std::string s(name);
    if (s == "N") *out = P->N; else if (s == "C") *out = P->C; else if (s == "nnz") *out = (double)P->nnz; else if (s == "nE") *out = P->nE;
    else if (s == "nK") *out = P->nK; else if (s == "M") *out = P->M; else if (s == "D") *out = P->D; else if (s == "dim") *out = P->dim;
    else if (s == "nloc") *out = P->nloc; else if (s == "topography") *out = P->topography; else if (s == "neumann_domain") *out = P->neumann_domain; else if (s == "ref_node") *out = P->ref_node; else if (s == "ref_last") *out = P->ref_last;
    else if (s == "has_background") *out = P->has_background; else if (s == "k_missing") *out = P->k_missing; else if (s == "n_colors") *out = P->n_colors;
    else if (s == "n_levels") *out = (double)P->levels.size(); else if (s == "surface_z") *out = P->surface_z; else if (s == "pro_nf") *out = P->pro_nf;
    else { g_plan_err = "unknown plan scalar: " + s; return 1; }
    return 0;
}

int pgb200_plan_build(const pgb200_mesh_in *mi, const pgb200_scheme_in *si, int sr, int n_k_user, const double *k_user,
                      const double *w_user, pgb200_built_plan **out) {
    if (!mi || !si || !out) { g_plan_err = "null argument"; return 1; }
    *out = nullptr;
    pgb200_built_plan *Pp = new pgb200_built_plan();
    pgb200_built_plan &P = *Pp;
#define PLAN_FAIL(msg) do { g_plan_err = (msg); delete Pp; return 1; } while (0)
    const int dim = mi->dim, nloc = mi->nloc, N = mi->n_nodes, C = mi->n_cells, nB = mi->n_bounds, nlb = mi->nlb;
    if (!((dim == 2 && (nloc == 3 || nloc == 6)) || (dim == 3 && (nloc == 4 || nloc == 10)))) PLAN_FAIL("unsupported cell type (need Tri3/Tri6/Tet4/Tet10)");
    if (N <= 0 || C <= 0) PLAN_FAIL("Found no mesh, so cannot calculate a response.");
    P.dim = dim; P.nloc = nloc; P.N = N; P.C = C; P.nlb = nlb; P.n_bounds = nB;
    const int nvert = dim + 1;
    StageTimer tm;

    // ---- internal node order: Morton curve of rank-quantised coordinates ------------------------------------------
    {
        const int bits = dim == 3 ? 10 : 15;
        std::vector<uint64_t> code((size_t)N, 0);
        std::vector<double> u((size_t)N);
        for (int ax = 0; ax < dim; ax++) {
            for (int i = 0; i < N; i++) u[i] = mi->pos[3 * (size_t)i + ax];
            std::vector<double> us(u);
            std::sort(us.begin(), us.end());
            us.erase(std::unique(us.begin(), us.end()), us.end());
            // structured meshes (few distinct coordinates per axis): the rank itself, so that 2^d consecutive nodes of the
            // curve are the corners of one grid cell -- the 8-row groups of k_spmm_mma then share most of their columns
            const bool raw = us.size() <= ((size_t)1 << bits);
            const double factor = (double)(1 << bits) / (double)std::max<size_t>(1, us.size());
            for (int i = 0; i < N; i++) {
                const uint64_t rank = (uint64_t)(std::lower_bound(us.begin(), us.end(), u[i]) - us.begin());
                const uint64_t q = raw ? rank : (uint64_t)((double)rank * factor);
                code[i] |= spread_bits(q, dim, bits) << ax;
            }
        }
        P.node_perm.resize(N); P.node_inv.resize(N);
        std::iota(P.node_perm.begin(), P.node_perm.end(), 0);
        std::stable_sort(P.node_perm.begin(), P.node_perm.end(), [&](int a, int b) { return code[a] < code[b]; });
        for (int i = 0; i < N; i++) P.node_inv[P.node_perm[i]] = i;
    }
    const IVec &perm = P.node_perm, &inv = P.node_inv;
    P.pos.resize(3 * (size_t)N); P.node_marker.resize(N);
    for (int i = 0; i < N; i++) { for (int d = 0; d < 3; d++) P.pos[3 * (size_t)i + d] = mi->pos[3 * (size_t)perm[i] + d]; P.node_marker[i] = mi->node_marker[perm[i]]; }
    P.cells.resize((size_t)C * nloc);
    for (size_t x = 0; x < (size_t)C * nloc; x++) { const int v = mi->cells[x]; if (v < 0 || v >= N) PLAN_FAIL("cell node index out of range"); P.cells[x] = inv[v]; }
    P.cell_marker.assign(mi->cell_marker, mi->cell_marker + C);
    P.bounds.resize((size_t)nB * nlb);
    for (size_t x = 0; x < (size_t)nB * nlb; x++) P.bounds[x] = inv[mi->bounds[x]];
    P.bound_marker.assign(mi->bound_marker, mi->bound_marker + nB);
    const double *pos = P.pos.data();
    const int *cells = P.cells.data();

    tm.lap("node order + renumber");
    // ---- boundary classification / topography (dcfemmodelling.cpp:725-765) -----------------------------------------
    bool any_mixed_dir = false;
    for (int b = 0; b < nB; b++) if (P.bound_marker[b] == MARKER_BOUND_MIXED || P.bound_marker[b] == MARKER_BOUND_DIRICHLET) any_mixed_dir = true;
    bool neumann_domain = !any_mixed_dir, topography = false, have_surf = false;
    double surface_z = -1.7976931348623157e308;
    for (int b = 0; b < nB; b++) {
        if (P.bound_marker[b] != MARKER_BOUND_NEUMANN) continue;
        double s = 0.0;
        for (int v = 0; v < nlb; v++) s += pos[3 * (size_t)P.bounds[(size_t)b * nlb + v] + (dim - 1)];
        const double cz = s / (double)nlb;
        if (!have_surf) { surface_z = cz; have_surf = true; }
        else if (cz != surface_z) topography = true;
    }
    if (neumann_domain) { topography = true; if (dim == 2) neumann_domain = false; }
    P.topography = topography ? 1 : 0; P.surface_z = surface_z; P.neumann_domain = neumann_domain ? 1 : 0;
    // reference-electrode node (-999): the first one in the reference's node order (:1009-1015); a pure-Neumann domain without
    // one takes the last electrode as current reference (:1054-1064)
    for (int R = 0; R < N && P.ref_node < 0; R++) if (mi->node_marker[R] == MARKER_NODE_REFERENCE) P.ref_node = inv[R];
    if (neumann_domain && P.ref_node < 0) P.ref_last = 1;

    // ---- node -> cells incidence --------------------------------------------------------------------------------------
    IVec nc_ptr(N + 1, 0), nc_cells((size_t)C * nloc);
    for (size_t x = 0; x < (size_t)C * nloc; x++) nc_ptr[cells[x] + 1]++;
    for (int i = 0; i < N; i++) nc_ptr[i + 1] += nc_ptr[i];
    { IVec fill(nc_ptr.begin(), nc_ptr.end() - 1); for (int c = 0; c < C; c++) for (int j = 0; j < nloc; j++) nc_cells[fill[cells[(size_t)c * nloc + j]]++] = c; }

    tm.lap("boundary + incidence");
    // ---- CSR pattern (sparsematrix.h:966-1032): union of all node pairs per cell, columns ascending -----------------
    P.rowptr.assign(N + 1, 0);
    {
        std::vector<IVec> rows((size_t)N);
#pragma omp parallel for schedule(dynamic, 512)
        for (int i = 0; i < N; i++) {
            IVec &r = rows[i];
            for (int q = nc_ptr[i]; q < nc_ptr[i + 1]; q++) { const int c = nc_cells[q]; for (int j = 0; j < nloc; j++) r.push_back(cells[(size_t)c * nloc + j]); }
            std::sort(r.begin(), r.end());
            r.erase(std::unique(r.begin(), r.end()), r.end());
        }
        long long tot = 0;
        for (int i = 0; i < N; i++) { tot += (long long)rows[i].size(); if (tot >= 2147483647LL) PLAN_FAIL("pattern exceeds int32 index range"); P.rowptr[i + 1] = (int)tot; }
        P.colidx.resize((size_t)tot);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) std::copy(rows[i].begin(), rows[i].end(), P.colidx.begin() + P.rowptr[i]);
        P.nnz = tot;
    }
    const IVec &rowptr = P.rowptr, &colidx = P.colidx;
    tm.lap("pattern");
    // per-cell scatter map
    IVec cpos((size_t)C * nloc * nloc);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++)
        for (int i = 0; i < nloc; i++)
            for (int j = 0; j < nloc; j++) cpos[((size_t)c * nloc + i) * nloc + j] = csr_find(rowptr, colidx, cells[(size_t)c * nloc + i], cells[(size_t)c * nloc + j]);
    tm.lap("scatter map");
    // colours
    {
        IVec color(C);
        const int ncol = pgb200_color_cells(C, nloc, cells, N, color.data());
        if (ncol <= 0) PLAN_FAIL(std::string("colouring failed: ") + pgb200_last_error());
        P.n_colors = ncol;
        P.color_ptr.assign(ncol + 1, 0);
        for (int c = 0; c < C; c++) P.color_ptr[color[c] + 1]++;
        for (int k = 0; k < ncol; k++) P.color_ptr[k + 1] += P.color_ptr[k];
        P.color_order.resize(C);
        { IVec fill(P.color_ptr.begin(), P.color_ptr.end() - 1); for (int c = 0; c < C; c++) P.color_order[fill[color[c]]++] = c; }
        P.cells_col.resize((size_t)nloc * C); P.pos_col.resize((size_t)nloc * nloc * C);
#pragma omp parallel for schedule(static)
        for (int s = 0; s < C; s++) {
            const int c = P.color_order[s];
            for (int j = 0; j < nloc; j++) P.cells_col[(size_t)j * C + s] = cells[(size_t)c * nloc + j];
            for (int x = 0; x < nloc * nloc; x++) P.pos_col[(size_t)x * C + s] = cpos[(size_t)c * nloc * nloc + x];
        }
    }
    tm.lap("colours");
    P.diag_pos.resize(N);
    for (int i = 0; i < N; i++) { P.diag_pos[i] = csr_find(rowptr, colidx, i, i); if (P.diag_pos[i] < 0) PLAN_FAIL("matrix row without diagonal entry"); }
    // the reference's pattern (original numbering) and the slot correspondence
    {
        P.ref_rowptr.assign(N + 1, 0); P.ref_colidx.resize((size_t)P.nnz); P.ref_slot.resize((size_t)P.nnz);
        for (int R = 0; R < N; R++) { const int i = inv[R]; P.ref_rowptr[R + 1] = P.ref_rowptr[R] + (rowptr[i + 1] - rowptr[i]); }
#pragma omp parallel for schedule(dynamic, 512)
        for (int R = 0; R < N; R++) {
            const int i = inv[R];
            std::vector<std::pair<int, int>> e;
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) e.push_back({perm[colidx[p]], p});
            std::sort(e.begin(), e.end());
            for (size_t x = 0; x < e.size(); x++) { P.ref_colidx[(size_t)P.ref_rowptr[R] + x] = e[x].first; P.ref_slot[(size_t)P.ref_rowptr[R] + x] = e[x].second; }
        }
    }

    tm.lap("diag + reference pattern");
    // ---- electrodes (dcfemmodelling.cpp:845-940) -------------------------------------------------------------------------
    const int nE = si->n_elec, D = si->n_data;
    if (nE <= 0) PLAN_FAIL("no response without data container");
    P.nE = nE; P.D = D;
    P.sensors.assign(si->sensors, si->sensors + 3 * (size_t)nE);
    DVec sens(P.sensors);
    if (dim == 2) {
        double zmin = 1e300, zmax = -1e300, ymin = 1e300, ymax = -1e300, zabs = 0.0, yabs = 0.0;
        for (int i = 0; i < nE; i++) {
            zmin = std::min(zmin, sens[3 * i + 2]); zmax = std::max(zmax, sens[3 * i + 2]); zabs = std::max(zabs, std::fabs(sens[3 * i + 2]));
            ymin = std::min(ymin, sens[3 * i + 1]); ymax = std::max(ymax, sens[3 * i + 1]); yabs = std::max(yabs, std::fabs(sens[3 * i + 1]));
        }
        const bool zvar = (zmax - zmin) > 0 || zabs > 0, yflat = (ymax - ymin) == 0 && yabs < 1e-8;
        if (zvar && yflat) for (int i = 0; i < nE; i++) std::swap(sens[3 * i + 1], sens[3 * i + 2]);
    }
    IVec src_nodes;                         // electrode-node candidates in the REFERENCE's node order
    for (int R = 0; R < N; R++) if (mi->node_marker[R] == MARKER_NODE_ELECTRODE) src_nodes.push_back(inv[R]);
    P.el_node.assign(nE, -1); P.el_cell.assign(nE, -1); P.sing_node.assign(nE, -1); P.el_pos.assign(3 * (size_t)nE, 0.0);
    P.pick_ptr.assign(1, 0);
    for (int i = 0; i < nE; i++) {
        int hit = -1;
        for (size_t t = 0; t < src_nodes.size(); t++) {
            const int n = src_nodes[t];
            double s = 0.0;
            for (int d = 0; d < 3; d++) { const double dd = sens[3 * i + d] - pos[3 * (size_t)n + d]; s += dd * dd; }
            if (std::sqrt(s) < 0.01) { hit = (int)t; break; }
        }
        if (hit >= 0) {
            const int n = src_nodes[hit];
            src_nodes.erase(src_nodes.begin() + hit);
            P.el_node[i] = n; P.sing_node[i] = n;
            for (int d = 0; d < 3; d++) P.el_pos[3 * i + d] = pos[3 * (size_t)n + d];
            P.pick_idx.push_back(n); P.pick_w.push_back(1.0);
        } else {
            // free electrode: containing cell + barycentric coordinates (brute force, set-up only)
            int best = -1; double best_min = -1e300, bestL[4] = {0, 0, 0, 0};
            for (int c = 0; c < C; c++) {
                const int *cn = cells + (size_t)c * nloc;
                const double *v0 = pos + 3 * (size_t)cn[0];
                double T[3][3], rhs[3], lam[3];
                for (int a = 0; a < dim; a++) { rhs[a] = sens[3 * i + a] - v0[a]; for (int b = 0; b < dim; b++) T[a][b] = pos[3 * (size_t)cn[b + 1] + a] - v0[a]; }
                if (dim == 2) {
                    const double det = T[0][0] * T[1][1] - T[0][1] * T[1][0];
                    lam[0] = (rhs[0] * T[1][1] - T[0][1] * rhs[1]) / det; lam[1] = (T[0][0] * rhs[1] - rhs[0] * T[1][0]) / det; lam[2] = 0.0;
                } else {
                    const double det = T[0][0] * (T[1][1] * T[2][2] - T[1][2] * T[2][1]) - T[0][1] * (T[1][0] * T[2][2] - T[1][2] * T[2][0]) + T[0][2] * (T[1][0] * T[2][1] - T[1][1] * T[2][0]);
                    lam[0] = (rhs[0] * (T[1][1] * T[2][2] - T[1][2] * T[2][1]) - T[0][1] * (rhs[1] * T[2][2] - T[1][2] * rhs[2]) + T[0][2] * (rhs[1] * T[2][1] - T[1][1] * rhs[2])) / det;
                    lam[1] = (T[0][0] * (rhs[1] * T[2][2] - T[1][2] * rhs[2]) - rhs[0] * (T[1][0] * T[2][2] - T[1][2] * T[2][0]) + T[0][2] * (T[1][0] * rhs[2] - rhs[1] * T[2][0])) / det;
                    lam[2] = (T[0][0] * (T[1][1] * rhs[2] - rhs[1] * T[2][1]) - T[0][1] * (T[1][0] * rhs[2] - rhs[1] * T[2][0]) + rhs[0] * (T[1][0] * T[2][1] - T[1][1] * T[2][0])) / det;
                }
                double L[4] = {1.0, 0, 0, 0};
                for (int a = 0; a < dim; a++) { L[a + 1] = lam[a]; L[0] -= lam[a]; }
                double mn = 1e300; bool ok = true;
                for (int a = 0; a <= dim; a++) { if (!(L[a] >= -1e-10)) ok = false; mn = std::min(mn, L[a]); }
                if (ok && mn > best_min) { best_min = mn; best = c; for (int a = 0; a < 4; a++) bestL[a] = L[a]; }
            }
            if (best < 0) PLAN_FAIL("There is a requested electrode that does not match the given mesh.");
            P.el_cell[i] = best;
            for (int d = 0; d < 3; d++) P.el_pos[3 * i + d] = sens[3 * i + d];
            double sf[10];
            shape_functions(nloc, dim, bestL, sf);
            const int *cn = cells + (size_t)best * nloc;
            for (int j = 0; j < nloc; j++) {
                P.pick_idx.push_back(cn[j]); P.pick_w.push_back(sf[j]);
                double s = 0.0;
                for (int d = 0; d < 3; d++) { const double dd = pos[3 * (size_t)cn[j] + d] - sens[3 * i + d]; s += dd * dd; }
                if (std::sqrt(s) < 1e-4) P.sing_node[i] = cn[j];       // the last near node wins, as in the numpy twin
            }
        }
        P.pick_ptr.push_back((int)P.pick_idx.size());
    }
    P.el_node_ref.assign(nE, -1);
    for (int i = 0; i < nE; i++) if (P.el_node[i] >= 0) P.el_node_ref[i] = perm[P.el_node[i]];
    for (int d = 0; d < 3; d++) { double s = 0.0; for (int i = 0; i < nE; i++) s += P.el_pos[3 * i + d]; P.source_center[d] = s / (double)nE; }
    // cells around the electrodes (rho at the source, electrode.cpp:102-120, :252-268) and the singular-patch radius
    auto cells_of = [&](int node) { IVec cs(nc_cells.begin() + nc_ptr[node], nc_cells.begin() + nc_ptr[node + 1]); std::sort(cs.begin(), cs.end()); cs.erase(std::unique(cs.begin(), cs.end()), cs.end()); return cs; };
    P.src_cell_ptr.assign(1, 0); P.min_radius.assign(nE, 0.0);
    for (int i = 0; i < nE; i++) {
        if (P.el_node[i] >= 0) { const IVec cs = cells_of(P.el_node[i]); P.src_cells.insert(P.src_cells.end(), cs.begin(), cs.end()); }
        else P.src_cells.push_back(P.el_cell[i]);
        P.src_cell_ptr.push_back((int)P.src_cells.size());
        if (P.sing_node[i] >= 0) {
            const int sn = P.sing_node[i];
            double mn = 1e300;
            for (int c : cells_of(sn))
                for (int j = 0; j < nloc; j++) {
                    const int n = cells[(size_t)c * nloc + j];
                    if (n == sn) continue;
                    double s = 0.0;
                    for (int d = 0; d < 3; d++) { const double dd = pos[3 * (size_t)n + d] - pos[3 * (size_t)sn + d]; s += dd * dd; }
                    mn = std::min(mn, std::sqrt(s));
                }
            P.min_radius[i] = mn;
        }
    }

    tm.lap("electrodes");
    // ---- wavenumbers ------------------------------------------------------------------------------------------------------
    if (k_user && w_user && n_k_user > 0) { P.kv.assign(k_user, k_user + n_k_user); P.kw.assign(w_user, w_user + n_k_user); }
    else { const std::string e = init_kwave_list(dim, nE, si->sensors, P.kv, P.kw); if (!e.empty()) PLAN_FAIL(e); }
    const int nK = (int)P.kv.size();
    P.nK = nK;
    // singular-value patch per (electrode, k) (electrode.cpp:154-189 with scale = 0)
    P.sing_val.assign((size_t)nK * nE, 0.0);
    for (int kk = 0; kk < nK; kk++)
        for (int i = 0; i < nE; i++) {
            if (P.sing_node[i] < 0) continue;
            const double kv = P.kv[kk];
            P.sing_val[(size_t)kk * nE + i] = kv > 0.0 ? bessel_k0(P.min_radius[i] / 6.0 * kv) / M_PI : 1.0 / (2.0 * M_PI * P.min_radius[i] / 2.0);
        }

    tm.lap("wavenumbers");
    // ---- boundary faces: owner cells, mixed-BC coefficient table, Dirichlet nodes ----------------------------------------
    IVec owner(nB, -1);
    {
        // owner = the cell that contains all corner nodes of the face
        for (int b = 0; b < nB; b++) {
            const int *fb = P.bounds.data() + (size_t)b * nlb;
            for (int q = nc_ptr[fb[0]]; q < nc_ptr[fb[0] + 1] && owner[b] < 0; q++) {
                const int c = nc_cells[q];
                bool all = true;
                for (int v = 1; v < dim && all; v++) { bool f = false; for (int j = 0; j < nvert; j++) if (cells[(size_t)c * nloc + j] == fb[v]) f = true; all = f; }
                bool first = false; for (int j = 0; j < nvert; j++) if (cells[(size_t)c * nloc + j] == fb[0]) first = true;
                if (all && first) owner[b] = c;
            }
            if (owner[b] < 0) PLAN_FAIL("boundary face without an adjacent cell");
        }
    }
    {
        struct Ent { int slot, owner; long long src; };
        std::vector<Ent> ent; std::vector<Face> faces; IVec fidx;
        DVec U; unit_mass_face(dim, nlb, U);
        for (int b = 0; b < nB; b++) {
            if (P.bound_marker[b] != MARKER_BOUND_MIXED) continue;
            const int *fb = P.bounds.data() + (size_t)b * nlb;
            const int f = (int)faces.size();
            faces.push_back(face_geometry(dim, pos, fb));
            for (int i = 0; i < nlb; i++)
                for (int j = 0; j < nlb; j++) {
                    const int slot = csr_find(rowptr, colidx, fb[i], fb[j]);
                    if (slot < 0) PLAN_FAIL("requested entry not in the sparsity pattern");
                    ent.push_back({slot, owner[b], (long long)f * nlb * nlb + i * nlb + j});
                }
        }
        std::stable_sort(ent.begin(), ent.end(), [](const Ent &a, const Ent &b) { return a.slot < b.slot; });
        const size_t ne = ent.size();
        P.bc_coef.assign((size_t)nK * ne, 0.0); P.bc_owner.resize(ne);
        P.bc_ptr.clear();
        for (size_t e = 0; e < ne; e++) {
            if (e == 0 || ent[e].slot != ent[e - 1].slot) { P.bc_slot.push_back(ent[e].slot); P.bc_ptr.push_back((int)e); }
            P.bc_owner[e] = ent[e].owner;
        }
        P.bc_ptr.push_back((int)ne);
        for (int kk = 0; kk < nK; kk++) {
            DVec bs(faces.size());
            for (size_t f = 0; f < faces.size(); f++) bs[f] = mixed_bc_beta(faces[f], P.source_center, P.kv[kk]) * faces[f].size;
            for (size_t e = 0; e < ne; e++) { const long long f = ent[e].src / (nlb * nlb), ij = ent[e].src % (nlb * nlb); P.bc_coef[(size_t)kk * ne + e] = bs[(size_t)f] * U[(size_t)ij]; }
        }
    }
    {
        std::vector<char> isd((size_t)N, 0);
        for (int b = 0; b < nB; b++) if (P.bound_marker[b] == MARKER_BOUND_DIRICHLET) for (int v = 0; v < nlb; v++) isd[P.bounds[(size_t)b * nlb + v]] = 1;
        if (P.neumann_domain) {
            // calibration nodes (-1000) pin the potential of a pure-Neumann domain; without one the reference takes its node 0
            // (:1044-1052).  On other domains calibration nodes are ignored (:1066-1070).
            bool any = false;
            for (int i = 0; i < N; i++) if (P.node_marker[i] == MARKER_NODE_CALIBRATION) { isd[i] = 1; any = true; }
            if (!any) isd[inv[0]] = 1;
        }
        for (int i = 0; i < N; i++) if (isd[i]) { P.dir_nodes.push_back(i); P.dir_diag.push_back(P.diag_pos[i]); }
        if (!P.dir_nodes.empty())
            for (int i = 0; i < N; i++) for (int p = rowptr[i]; p < rowptr[i + 1]; p++) if (isd[i] || isd[colidx[p]]) P.dir_zero.push_back(p);
    }

    tm.lap("boundary tables");
    // ---- model mapping (modellingbase.cpp:401-497, mesh.cpp:2247-2316) ---------------------------------------------------
    int maxm = -1;
    for (int c = 0; c < C; c++) { if (P.cell_marker[c] <= -1000000) PLAN_FAIL("fixed-value regions are not supported on the B200 path"); maxm = std::max(maxm, P.cell_marker[c]); }
    const int M = maxm >= 0 ? maxm + 1 : 0;
    P.M = M;
    bool has_bg = false;
    for (int c = 0; c < C; c++) if (P.cell_marker[c] < 0) has_bg = true;
    P.has_background = has_bg ? 1 : 0;
    P.pro_level_ptr.assign(1, 0);
    if (has_bg) {
        static const int fl2[3][2] = {{0, 1}, {1, 2}, {2, 0}}, fl3[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};
        const int nf = dim == 2 ? 3 : 4;
        P.pro_nf = nf;
        // face-sharing neighbours through the node -> cells incidence (each face: the other cell that holds all its corners)
        IVec nb((size_t)C * nf, -1);
#pragma omp parallel for schedule(static)
        for (int c = 0; c < C; c++)
            for (int f = 0; f < nf; f++) {
                int v[3];
                for (int a = 0; a < dim; a++) v[a] = cells[(size_t)c * nloc + (dim == 2 ? fl2[f][a] : fl3[f][a])];
                for (int q = nc_ptr[v[0]]; q < nc_ptr[v[0] + 1]; q++) {
                    const int c2 = nc_cells[q];
                    if (c2 == c) continue;
                    bool all = true;
                    for (int a = 1; a < dim && all; a++) { bool fnd = false; for (int j = 0; j < nvert; j++) if (cells[(size_t)c2 * nloc + j] == v[a]) fnd = true; all = fnd; }
                    bool first = false; for (int j = 0; j < nvert; j++) if (cells[(size_t)c2 * nloc + j] == v[0]) first = true;
                    if (all && first) { nb[(size_t)c * nf + f] = c2; break; }
                }
            }
        DVec zw((size_t)C * nf);
        const double xy[3] = {1.0, dim == 3 ? 1.0 : 0.0, 0.0};
#pragma omp parallel for schedule(static)
        for (int c = 0; c < C; c++)
            for (int f = 0; f < nf; f++) {
                int v[3];
                for (int a = 0; a < dim; a++) v[a] = cells[(size_t)c * nloc + (dim == 2 ? fl2[f][a] : fl3[f][a])];
                const Face fg = face_geometry(dim, pos, v);
                const double a0 = fg.n[0] * xy[0], a1 = fg.n[1] * xy[1], a2 = fg.n[2] * xy[2];
                zw[(size_t)c * nf + f] = std::sqrt(a0 * a0 + a1 * a1 + a2 * a2) + 1e-6;
            }
        IVec lvl(C);
        for (int c = 0; c < C; c++) lvl[c] = P.cell_marker[c] < 0 ? -1 : 0;
        // level l = the empty cells with a neighbour filled at a level < l (mesh.cpp:2276-2306, one recursion per level);
        // frontier search: the candidates of a level are the empty neighbours of the cells filled one level before
        int cur = 0, remaining = 0;
        std::vector<char> queued((size_t)C, 0);
        IVec now;
        for (int c = 0; c < C; c++) {
            if (lvl[c] >= 0) continue;
            remaining++;
            for (int f = 0; f < nf; f++) { const int n = nb[(size_t)c * nf + f]; if (n >= 0 && lvl[n] >= 0) { now.push_back(c); queued[c] = 1; break; } }
        }
        while (remaining > 0) {
            if (now.empty()) PLAN_FAIL("cannot fill empty cells: disconnected background region");
            std::sort(now.begin(), now.end());
            for (int c : now) {
                double wsum = 0.0, wv[4]; int nv[4];
                for (int f = 0; f < nf; f++) {
                    const int n = nb[(size_t)c * nf + f];
                    const bool ok = n >= 0 && lvl[n] >= 0 && lvl[n] <= cur;
                    wv[f] = ok ? zw[(size_t)c * nf + f] : 0.0; nv[f] = ok ? n : 0;
                    wsum += wv[f];
                }
                P.pro_cells.push_back(c);
                for (int f = 0; f < nf; f++) { P.pro_nb.push_back(nv[f]); P.pro_w.push_back(wv[f] / wsum); }
            }
            cur++;
            for (int c : now) lvl[c] = cur;
            remaining -= (int)now.size();
            P.pro_level_ptr.push_back((int)P.pro_cells.size());
            IVec next;
            for (int c : now)
                for (int f = 0; f < nf; f++) { const int n = nb[(size_t)c * nf + f]; if (n >= 0 && lvl[n] < 0 && !queued[n]) { queued[n] = 1; next.push_back(n); } }
            now.swap(next);
        }
    }

    // ---- Jacobian columns: cells with marker >= 0 sorted by marker (bertJacobian.cpp:298-299) --------------------------------
    P.jac_col_ptr.assign(M + 1, 0);
    for (int c = 0; c < C; c++) if (P.cell_marker[c] >= 0) P.jac_col_ptr[P.cell_marker[c] + 1]++;
    for (int m = 0; m < M; m++) P.jac_col_ptr[m + 1] += P.jac_col_ptr[m];
    P.jac_cells.resize(P.jac_col_ptr[M]);
    { IVec fill(P.jac_col_ptr.begin(), P.jac_col_ptr.end() - 1); for (int c = 0; c < C; c++) if (P.cell_marker[c] >= 0) P.jac_cells[fill[P.cell_marker[c]]++] = c; }

    // ---- data ----------------------------------------------------------------------------------------------------------------
    P.abmn.assign(si->abmn, si->abmn + 4 * (size_t)D);
    for (int x = 0; x < 4 * D; x++) if (P.abmn[x] >= nE || P.abmn[x] < -1) PLAN_FAIL("electrode index out of range in the data");
    bool have_k = si->k_fac != nullptr;
    if (have_k) { double mn = 1e300; for (int d = 0; d < D; d++) mn = std::min(mn, std::fabs(si->k_fac[d])); if (D > 0 && mn < TOLERANCE) have_k = false; }
    if (have_k) P.kfac.assign(si->k_fac, si->k_fac + D);
    else if (!topography) {
        // analytic flat-earth factors (bertMisc.cpp:131-176, :186-214); in 2-D a non-zero y is moved to z first
        DVec s(P.sensors);
        if (dim == 2) for (int i = 0; i < nE; i++) if (s[3 * i + 1] != 0.0) { s[3 * i + 2] = s[3 * i + 1]; s[3 * i + 1] = 0.0; }
        auto u = [&](int src, int rec) -> double {
            if (src < 0 || rec < 0) return 0.0;
            const double *p = &s[3 * (size_t)rec], *q = &s[3 * (size_t)src];
            const double r = std::sqrt((p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]));
            const double rm = std::sqrt((p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] + q[2]) * (p[2] + q[2]));
            return r < 1e-12 ? 1.0 : (1.0 / r + 1.0 / rm) / (4.0 * M_PI);
        };
        P.kfac.resize(D);
        for (int d = 0; d < D; d++) { const int a = P.abmn[4 * d], b = P.abmn[4 * d + 1], m = P.abmn[4 * d + 2], n = P.abmn[4 * d + 3]; P.kfac[d] = 1.0 / (u(a, m) - u(b, m) - u(a, n) + u(b, n)); }
    } else { P.kfac.assign(D, 0.0); P.k_missing = 1; }

    tm.lap("jacobian columns + data");
    // ---- the flat view ----------------------------------------------------------------------------------------------------------
    pgb200_plan &V = P.view;
    V.dim = dim; V.nloc = nloc; V.n_nodes = N; V.n_cells = C; V.nnz = (int)P.nnz; V.n_elec = nE; V.n_k = nK; V.n_model = M; V.n_data = D; V.sr = sr ? 1 : 0;
    V.fullspace = surface_z <= -1e300 ? 1 : 0; V.surface_z = V.fullspace ? 0.0 : surface_z;
    V.pos = P.pos.data(); V.cells = P.cells.data(); V.cell_marker = P.cell_marker.data(); V.rowptr = P.rowptr.data(); V.colidx = P.colidx.data(); V.diag_pos = P.diag_pos.data();
    V.n_colors = P.n_colors; V.color_ptr = P.color_ptr.data(); V.color_order = P.color_order.data(); V.cells_col = P.cells_col.data(); V.pos_col = P.pos_col.data();
    V.k_values = P.kv.data(); V.k_weights = P.kw.data();
    V.n_bc_slots = (int)P.bc_slot.size(); V.n_bc_entries = (int)P.bc_owner.size(); V.bc_slot = P.bc_slot.data(); V.bc_ptr = P.bc_ptr.data(); V.bc_owner = P.bc_owner.data(); V.bc_coef = P.bc_coef.data();
    V.n_dir_zero = (int)P.dir_zero.size(); V.n_dir_nodes = (int)P.dir_nodes.size(); V.dir_zero_slots = P.dir_zero.data(); V.dir_diag_slots = P.dir_diag.data(); V.dir_nodes = P.dir_nodes.data();
    V.el_pos = P.el_pos.data(); V.sing_node = P.sing_node.data(); V.sing_val = P.sing_val.data(); V.pick_ptr = P.pick_ptr.data(); V.pick_idx = P.pick_idx.data(); V.pick_w = P.pick_w.data();
    V.src_cell_ptr = P.src_cell_ptr.data(); V.src_cells = P.src_cells.data();
    V.n_pro_levels = (int)P.pro_level_ptr.size() - 1; V.pro_nf = V.n_pro_levels ? P.pro_nf : 0; V.pro_level_ptr = P.pro_level_ptr.data(); V.pro_cells = P.pro_cells.data();
    V.pro_nb = P.pro_nb.data(); V.pro_w = P.pro_w.data();
    V.n_jac_cells = (int)P.jac_cells.size(); V.jac_cells = P.jac_cells.data(); V.jac_col_ptr = P.jac_col_ptr.data();
    V.abmn = P.abmn.data(); V.k_fac = P.kfac.data(); V.topography = P.topography; V.ref_node = P.ref_node; V.ref_last = P.ref_last;
#undef PLAN_FAIL
    *out = Pp;
    return 0;
}

// Aggregation hierarchy of the multilevel preconditioner from the rho = 1 matrix values of the smallest wavenumber
// (geometry only; twin of pygimli_b200/amg_setup.build_hierarchy): `passes` rounds of strength-thresholded pairwise matching
// per level, piecewise-constant transfer, Galerkin gather lists.  Returns the number of levels (< 0 on failure).
int pgb200_plan_build_hierarchy(pgb200_built_plan *P, const double *vals1, double theta, int passes, int min_size, int max_levels) {
    if (!P || !vals1) { g_plan_err = "null argument"; return -1; }
    P->levels.clear(); P->level_views.clear();
    IVec rp(P->rowptr), ci(P->colidx);
    DVec v(vals1, vals1 + P->nnz);
    while ((int)P->levels.size() < max_levels) {
        const int n = (int)rp.size() - 1;
        if (n <= min_size) break;
        IVec agg(n);
        std::iota(agg.begin(), agg.end(), 0);
        IVec rp_p(rp), ci_p(ci); DVec v_p(v);
        int nc = n;
        for (int ps = 0; ps < passes; ps++) {
            const int np = (int)rp_p.size() - 1;
            IVec a(np);
            const int na = pgb200_pairwise_aggregate(np, rp_p.data(), ci_p.data(), v_p.data(), nullptr, theta, a.data());
            if (na < 0) { g_plan_err = pgb200_last_error(); return -1; }
            IVec crp, cci, gp, gi;
            coarsen(rp_p, ci_p, a, na, crp, cci, gp, gi);
            DVec vc(cci.size(), 0.0);
            for (size_t s = 0; s < cci.size(); s++) { double acc = 0.0; for (int q = gp[s]; q < gp[s + 1]; q++) acc += v_p[gi[q]]; vc[s] = acc; }
            rp_p.swap(crp); ci_p.swap(cci); v_p.swap(vc);
            for (int i = 0; i < n; i++) agg[i] = a[agg[i]];
            nc = na;
        }
        if (nc > 0.7 * n) break;
        pgb200_built_plan::Level L;
        coarsen(rp, ci, agg, nc, L.rowptr, L.colidx, L.gal_ptr, L.gal_idx);
        L.n = nc; L.nnz = (long long)L.colidx.size(); L.agg = agg;
        L.diag_pos.resize(nc);
        for (int I = 0; I < nc; I++) { L.diag_pos[I] = csr_find(L.rowptr, L.colidx, I, I); if (L.diag_pos[I] < 0) { g_plan_err = "every row needs a diagonal entry"; return -1; } }
        L.mem_ptr.assign(nc + 1, 0); L.mem_idx.resize(n);
        for (int i = 0; i < n; i++) L.mem_ptr[agg[i] + 1]++;
        for (int I = 0; I < nc; I++) L.mem_ptr[I + 1] += L.mem_ptr[I];
        { IVec fill(L.mem_ptr.begin(), L.mem_ptr.end() - 1); for (int i = 0; i < n; i++) L.mem_idx[fill[agg[i]]++] = i; }
        DVec vc(L.colidx.size(), 0.0);
        for (size_t s = 0; s < L.colidx.size(); s++) { double acc = 0.0; for (int q = L.gal_ptr[s]; q < L.gal_ptr[s + 1]; q++) acc += v[L.gal_idx[q]]; vc[s] = acc; }
        rp = L.rowptr; ci = L.colidx; v.swap(vc);
        P->levels.push_back(std::move(L));
    }
    for (auto &L : P->levels) {
        pgb200_amg_level a{};
        a.n = L.n; a.nnz = (int)L.nnz; a.rowptr = L.rowptr.data(); a.colidx = L.colidx.data(); a.diag_pos = L.diag_pos.data(); a.gal_ptr = L.gal_ptr.data();
        a.gal_idx = L.gal_idx.data(); a.agg = L.agg.data(); a.mem_ptr = L.mem_ptr.data(); a.mem_idx = L.mem_idx.data();
        P->level_views.push_back(a);
    }
    return (int)P->levels.size();
}
const pgb200_amg_level *pgb200_plan_levels(const pgb200_built_plan *P) { return (P && !P->level_views.empty()) ? P->level_views.data() : nullptr; }

} // extern "C"
