// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// A thin extern "C" shim around the UNMODIFIED reference C++ (libgimli sources
// compiled where they lie under /root/reference/core/src, see oracle/Makefile).
// It lets tests/ and bench.py's CPU arm drive the reference's own
// DCSRMultiElectrodeModelling (core/src/bert/dcfemmodelling.cpp:2152 calculateK,
// :1085 response, :1446 createJacobian; core/src/bert/bertJacobian.cpp:267
// createSensitivityCol) on flat mesh/scheme arrays and read back every
// intermediate the CUDA path is compared against.
//
// Nothing here restates reference arithmetic: all numbers come from reference
// code.  The only non-reference arithmetic is the linear solver, because
// SuiteSparse/CHOLMOD is not installed in this image: the reference's own
// injection seam SolverWrapper (core/src/solverWrapper.h:26-54, used through
// DCMultiElectrodeModelling::setSolver, dcfemmodelling.h:266) receives either
//   (a) a callback pair into Python (scipy SuperLU direct solve + refinement), or
//   (b) a built-in Jacobi-PCG run to 1e-14 relative residual.
#include <gimli.h>
#include <mesh.h>
#include <meshentities.h>
#include <node.h>
#include <shape.h>
#include <sparsematrix.h>
#include <solverWrapper.h>
#include <matrix.h>
#include <interpolate.h>
#include <stopwatch.h>
#include <bert/bert.h>
#include <bert/dcfemmodelling.h>
#include <bert/bertDataContainer.h>
#include <bert/bertMisc.h>
#include <bert/bertJacobian.h>
#include <bert/electrode.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

using namespace GIMLI;

// defined (non-static) at core/src/bert/dcfemmodelling.cpp:301 but not declared in its header
namespace GIMLI {
void dcfemBoundaryAssembleStiffnessMatrix(RSparseMatrix & S, const Mesh & mesh,
                                          const RVector3 & source, double k);
}

extern "C" {
typedef void (*set_matrix_cb)(int n, int nnz, const int * rowptr, const int * colidx, const double * vals);
typedef void (*solve_cb)(int n, const double * rhs, double * sol);
// complex variants: values / vectors are interleaved (re, im) doubles, i.e. std::complex<double> / numpy complex128
typedef void (*set_matrix_c_cb)(int n, int nnz, const int * rowptr, const int * colidx, const double * vals);
typedef void (*solve_c_cb)(int n, const double * rhs, double * sol);
}

namespace {

class InjectedSolver : public SolverWrapper {
public:
    InjectedSolver() : SolverWrapper(false), S_(0), setCb_(0), solveCb_(0),
        nSetMatrix(0), nSolve(0), tSetMatrix(0.0), tSolve(0.0), tol_(1e-14), iters_(0) { name_ = "oracleInjected"; }

    virtual void setMatrix(const RSparseMatrix & S){
        Stopwatch sw(true);
        S_ = &S;
        if (setCb_){
            setCb_((int)S.rows(), (int)S.nVals(), &S.vecColPtr()[0], &S.vecRowIdx()[0], &S.vecVals()[0]);
        } else {
            Index n = S.rows(); diag_.resize(n);
            for (Index i = 0; i < n; i++) diag_[i] = S.getVal(i, i);
        }
        nSetMatrix++; tSetMatrix += sw.duration();
    }
    // complex resistivity (dcfemmodelling.cpp:1843-1866 with ValueType = Complex): only through the callbacks
    virtual void setMatrix(const CSparseMatrix & S){
        if (!setCbC_) throwError("oracle driver: complex solve needs the complex callbacks");
        nC_ = S.rows();
        setCbC_((int)S.rows(), (int)S.nVals(), &S.vecColPtr()[0], &S.vecRowIdx()[0], reinterpret_cast< const double * >(&S.vecVals()[0]));
        nSetMatrix++;
    }
    virtual void solve(const CVector & b, CVector & x){
        if (!solveCbC_) throwError("oracle driver: complex solve needs the complex callbacks");
        if (x.size() != nC_) x.resize(nC_);
        solveCbC_((int)nC_, reinterpret_cast< const double * >(&b[0]), reinterpret_cast< double * >(&x[0]));
        nSolve++;
    }
    virtual void solve(const RVector & b, RVector & x){
        Stopwatch sw(true);
        Index n = S_->rows();
        if (x.size() != n) x.resize(n);
        if (solveCb_){
            solveCb_((int)n, &b[0], &x[0]);
        } else {
            pcg_(b, x);
        }
        nSolve++; tSolve += sw.duration();
    }
    void pcg_(const RVector & b, RVector & x){
        const RSparseMatrix & A = *S_;
        Index n = A.rows();
        x *= 0.0;
        double b2 = std::sqrt(dot(b, b));
        if (b2 == 0.0) return;
        RVector r(b), z(r / diag_), p(z);
        double rz = dot(r, z);
        for (int it = 0; it < 200000; it++){
            RVector Ap(A * p);
            double a = rz / dot(p, Ap);
            x += p * a; r -= Ap * a;
            iters_++;
            if (std::sqrt(dot(r, r)) <= tol_ * b2) break;
            z = r / diag_; double rzn = dot(r, z);
            p = z + p * (rzn / rz); rz = rzn;
        }
    }
    const RSparseMatrix * S_;
    RVector diag_;
    set_matrix_cb setCb_;
    solve_cb solveCb_;
    set_matrix_c_cb setCbC_ = 0;
    solve_c_cb solveCbC_ = 0;
    Index nC_ = 0;
    long nSetMatrix, nSolve;
    double tSetMatrix, tSolve;
    double tol_;
    long iters_;
};

struct RefHandle {
    Mesh * mesh;
    DataContainerERT * data;
    DCMultiElectrodeModelling * fop;
    InjectedSolver * solver;
    RMatrix * subPots;
    bool sr;
    RMatrix * prim;      // numeric primary potentials handed to the SR fop (topography), owned here
};

} // namespace

extern "C" {

// Build the reference Mesh + DataContainerERT + fop from flat arrays.
//   xyz[nNodes*3], nodeMarker[nNodes]; cells[nCells*nloc] (nloc 3/6/4/10, reference local order),
//   bounds[nBounds*nlocb] with their markers (only marked outer faces need to be listed),
//   sensors[nSensors*3], abmn[nData*4] (-1 = unused electrode).
void * ref_create(int dim, int nNodes, const double * xyz, const int * nodeMarker,
                  int nCells, int nloc, const int * cells, const int * cellMarker,
                  int nBounds, int nlocb, const int * bounds, const int * boundMarker,
                  int nSensors, const double * sensors,
                  int nData, const int * abmn, int sr, int verbose){
    RefHandle * h = new RefHandle();
    h->mesh = new Mesh(dim);
    Mesh & mesh = *h->mesh;
    for (int i = 0; i < nNodes; i++) mesh.createNode(xyz[3*i], xyz[3*i+1], xyz[3*i+2], nodeMarker[i]);
    for (int c = 0; c < nCells; c++){
        IndexArray ids(nloc);
        for (int j = 0; j < nloc; j++) ids[j] = cells[c*nloc + j];
        mesh.createCell(ids, cellMarker[c]);
    }
    for (int b = 0; b < nBounds; b++){
        IndexArray ids(nlocb);
        for (int j = 0; j < nlocb; j++) ids[j] = bounds[b*nlocb + j];
        mesh.createBoundary(ids, boundMarker[b], false);
    }
    mesh.createNeighborInfos();

    h->data = new DataContainerERT();
    for (int s = 0; s < nSensors; s++)
        h->data->createSensor(RVector3(sensors[3*s], sensors[3*s+1], sensors[3*s+2]));
    h->data->resize(nData);
    for (int d = 0; d < nData; d++)
        h->data->createFourPointData(d, abmn[4*d], abmn[4*d+1], abmn[4*d+2], abmn[4*d+3]);

    h->sr = sr != 0;
    if (h->sr) h->fop = new DCSRMultiElectrodeModelling(mesh, *h->data, verbose > 0);
    else       h->fop = new DCMultiElectrodeModelling(mesh, *h->data, verbose > 0);
    h->solver = new InjectedSolver();
    h->fop->setSolver(h->solver);
    h->fop->setThreadCount(1);
    h->subPots = new RMatrix();
    if (verbose >= 0) h->fop->collectSubPotentials(*h->subPots);   // verbose < 0: complex handle, see ref_create_complex
    return h;
}

// Complex-resistivity handle (total field only: DCSRMultiElectrodeModelling::calculateK throws for complex,
// dcfemmodelling.cpp:2156).  The k-resolved potentials cannot be collected (collectSubPotentials takes an RMatrix, the
// complex run needs a CMatrix the fop creates itself, :1672-1676); solution() holds the k-summed [re; im] rows.
void * ref_create_complex(int dim, int nNodes, const double * xyz, const int * nodeMarker,
                          int nCells, int nloc, const int * cells, const int * cellMarker,
                          int nBounds, int nlocb, const int * bounds, const int * boundMarker,
                          int nSensors, const double * sensors, int nData, const int * abmn){
    RefHandle * h = (RefHandle *)ref_create(dim, nNodes, xyz, nodeMarker, nCells, nloc, cells, cellMarker, nBounds, nlocb, bounds,
                                            boundMarker, nSensors, sensors, nData, abmn, 0, -1);
    h->fop->setVerbose(false);
    h->fop->setComplex(true);
    return h;
}
void ref_set_complex_callbacks(void * vh, set_matrix_c_cb a, solve_c_cb b){
    RefHandle * h = (RefHandle *)vh; h->solver->setCbC_ = a; h->solver->solveCbC_ = b;
}
// complex response (dcfemmodelling.cpp:1103-1118): model = [re(rho); im(rho)], out = [re(rhoa); im(rhoa)]
void ref_response_complex(void * vh, int nModel2, const double * model, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel2); for (int i = 0; i < nModel2; i++) m[i] = model[i];
    RVector r(h->fop->response(m));
    for (Index i = 0; i < r.size(); i++) out[i] = r[i];
}
// complex Jacobian (dcfemmodelling.cpp:1447-1461, 1410-1444): out = interleaved (re, im) row-major [rows x cols]
void ref_create_jacobian_complex(void * vh, int nModel2, const double * model, int * rowsCols, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel2); for (int i = 0; i < nModel2; i++) m[i] = model[i];
    if (!out) { h->fop->createJacobian(m); }
    CMatrix * J = dynamic_cast< CMatrix * >(h->fop->jacobian());
    rowsCols[0] = (int)J->rows(); rowsCols[1] = (int)J->cols();
    if (out) for (Index i = 0; i < J->rows(); i++)
        std::memcpy(out + 2*i*J->cols(), reinterpret_cast< const double * >(&(*J)[i][0]), 2*J->cols()*sizeof(double));
}
int ref_solution_rows(void * vh){ return (int)((RefHandle *)vh)->fop->solution().rows(); }

void ref_destroy(void * vh){
    RefHandle * h = (RefHandle *)vh;
    delete h->fop; delete h->solver; delete h->subPots; delete h->data; delete h->mesh; if (h->prim) delete h->prim; delete h;
}

void ref_set_solver_callbacks(void * vh, set_matrix_cb a, solve_cb b){
    RefHandle * h = (RefHandle *)vh; h->solver->setCb_ = a; h->solver->solveCb_ = b;
}
void ref_solver_stats(void * vh, double * out4){
    RefHandle * h = (RefHandle *)vh;
    out4[0] = (double)h->solver->nSetMatrix; out4[1] = h->solver->tSetMatrix;
    out4[2] = (double)h->solver->nSolve;     out4[3] = h->solver->tSolve;
}
void ref_set_pcg_tol(void * vh, double tol){ ((RefHandle *)vh)->solver->tol_ = tol; }
// Time the linear-solve stage of calculateK (dcfemmodelling.cpp:2152-2296) for the first nSrc current
// patterns and all wavenumbers: out = {seconds in setMatrix, seconds in solve, seconds total calculateK, iterations}
void ref_time_partial_solve(void * vh, int nModel, const double * model, int nSrc, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    h->fop->mapERTModel(m, -9e99);
    std::vector < ElectrodeShape * > eA, eB;
    h->fop->createCurrentPattern(eA, eB, true);
    eA.resize(nSrc); eB.resize(nSrc);
    h->fop->preCalculate(eA, eB);
    RMatrix sol(nSrc * h->fop->kValues().size(), h->fop->mesh()->nodeCount());
    double t0s = h->solver->tSetMatrix, t0v = h->solver->tSolve; long it0 = h->solver->iters_;
    Stopwatch sw(true);
    for (Index k = 0; k < h->fop->kValues().size(); k++) h->fop->calculateK(eA, eB, sol, k);
    out[2] = sw.duration();
    out[0] = h->solver->tSetMatrix - t0s; out[1] = h->solver->tSolve - t0v; out[3] = (double)(h->solver->iters_ - it0);
}
// same bounded solve, but the k-resolved total potentials of the first nSrc sources are handed back
// (row i + k * nSrc, node-major as the reference stores them): bench.py compares them with the GPU's potentials
void ref_partial_solve_pots(void * vh, int nModel, const double * model, int nSrc, double * out, double * solOut){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    h->fop->mapERTModel(m, -9e99);
    std::vector < ElectrodeShape * > eA, eB;
    h->fop->createCurrentPattern(eA, eB, true);
    eA.resize(nSrc); eB.resize(nSrc);
    h->fop->preCalculate(eA, eB);
    const Index nK = h->fop->kValues().size(), N = h->fop->mesh()->nodeCount();
    RMatrix sol(nSrc * nK, N);
    double t0s = h->solver->tSetMatrix, t0v = h->solver->tSolve; long it0 = h->solver->iters_;
    Stopwatch sw(true);
    for (Index k = 0; k < nK; k++) h->fop->calculateK(eA, eB, sol, k);
    out[2] = sw.duration();
    out[0] = h->solver->tSetMatrix - t0s; out[1] = h->solver->tSolve - t0v; out[3] = (double)(h->solver->iters_ - it0);
    if (solOut) for (Index r = 0; r < sol.rows(); r++) std::memcpy(solOut + (size_t)r * N, &sol[r][0], N * sizeof(double));
}
void ref_set_threads(void * vh, int n){ ((RefHandle *)vh)->fop->setThreadCount(n); }

int ref_n_k(void * vh){ return (int)((RefHandle *)vh)->fop->kValues().size(); }
void ref_get_kw(void * vh, double * k, double * w){
    RefHandle * h = (RefHandle *)vh;
    for (Index i = 0; i < h->fop->kValues().size(); i++){ k[i] = h->fop->kValues()[i]; w[i] = h->fop->weights()[i]; }
}
void ref_set_kw(void * vh, int n, const double * k, const double * w){
    RefHandle * h = (RefHandle *)vh;
    RVector kv(n), wv(n);
    for (int i = 0; i < n; i++){ kv[i] = k[i]; wv[i] = w[i]; }
    h->fop->setkValues(kv); h->fop->setWeights(wv);
}
int ref_topography(void * vh){ return ((RefHandle *)vh)->fop->topography() ? 1 : 0; }
int ref_n_electrodes(void * vh){ return (int)((RefHandle *)vh)->fop->electrodes().size(); }
// node id of every electrode (mID; -1 for free electrodes)
void ref_electrode_nodes(void * vh, int * out){
    RefHandle * h = (RefHandle *)vh;
    for (Index i = 0; i < h->fop->electrodes().size(); i++) out[i] = h->fop->electrodes()[i]->mID();
}

// analytic geometric factors (bertMisc.cpp:131) stored into the data container as token "k"
void ref_geometric_factors(void * vh, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector k(h->fop->calcGeometricFactor(*h->data));
    h->data->set("k", k);
    for (Index i = 0; i < k.size(); i++) out[i] = k[i];
}
void ref_set_k(void * vh, const double * kin){
    RefHandle * h = (RefHandle *)vh;
    RVector k(h->data->size());
    for (Index i = 0; i < k.size(); i++) k[i] = kin[i];
    h->data->set("k", k);
}

void ref_response(void * vh, int nModel, const double * model, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    RVector r(h->fop->response(m));
    for (Index i = 0; i < r.size(); i++) out[i] = r[i];
}
// cell resistivities after mapERTModel (dcfemmodelling.cpp:1211)
void ref_mapped_model(void * vh, int nModel, const double * model, double * outCells){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    h->fop->mapERTModel(m, -9e99);
    const RVector & a = h->fop->mesh()->cellAttributes();
    for (Index i = 0; i < a.size(); i++) outCells[i] = a[i];
}
// drop cached potentials so that the next createJacobian recomputes them (dcfemmodelling.cpp:1262)
void ref_clear_potentials(void * vh){ ((RefHandle *)vh)->subPots->clear(); }

void ref_create_jacobian(void * vh, int nModel, const double * model, int * rowsCols){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    h->fop->createJacobian(m);
    RMatrix & J = h->fop->jacobianRef();
    rowsCols[0] = (int)J.rows(); rowsCols[1] = (int)J.cols();
}
void ref_get_jacobian(void * vh, double * out){
    RMatrix & J = ((RefHandle *)vh)->fop->jacobianRef();
    for (Index i = 0; i < J.rows(); i++) std::memcpy(out + i*J.cols(), &J[i][0], J.cols()*sizeof(double));
}
// raw createSensitivityCol output (no k/rho^2 scaling) for given potentials (bertJacobian.cpp:267)
void ref_sensitivity_only(void * vh, int nThreads, int nRowsPots, const double * pots, double * outJ, int * rowsCols){
    RefHandle * h = (RefHandle *)vh;
    Index N = h->fop->mesh()->nodeCount();
    RMatrix P(nRowsPots, N);
    for (int i = 0; i < nRowsPots; i++) std::memcpy(&P[i][0], pots + (size_t)i*N, N*sizeof(double));
    RMatrix J;
    std::vector < std::pair < Index, Index > > cl;
    createSensitivityCol(J, *h->fop->mesh(), *h->data, P, h->fop->weights(), h->fop->kValues(), cl, nThreads, false);
    rowsCols[0] = (int)J.rows(); rowsCols[1] = (int)J.cols();
    if (outJ) for (Index i = 0; i < J.rows(); i++) std::memcpy(outJ + i*J.cols(), &J[i][0], J.cols()*sizeof(double));
}

// coverageDCtrans (bertJacobian.cpp:569-598) on a dense row-major matrix handed in by the caller
void ref_coverage_trans(int rows, int cols, const double * J, const double * dd, const double * mm, double * out){
    RMatrix S(rows, cols);
    for (int i = 0; i < rows; i++) std::memcpy(&S[i][0], J + (size_t)i*cols, cols*sizeof(double));
    RVector d(rows), m(cols);
    for (int i = 0; i < rows; i++) d[i] = dd[i];
    for (int j = 0; j < cols; j++) m[j] = mm[j];
    RVector cov(coverageDCtrans(S, d, m));
    for (int j = 0; j < cols; j++) out[j] = cov[j];
}
// createCoverage (bertJacobian.cpp:600-628) with the parameter mesh given as flat arrays; out has nCells entries
void ref_create_coverage(int rows, int cols, const double * J, int dim, int nNodes, const double * xyz,
                         int nCells, int nloc, const int * cells, const int * cellMarker,
                         const double * response, const double * model, double * out){
    RMatrix S(rows, cols);
    for (int i = 0; i < rows; i++) std::memcpy(&S[i][0], J + (size_t)i*cols, cols*sizeof(double));
    Mesh mesh(dim);
    for (int i = 0; i < nNodes; i++) mesh.createNode(xyz[3*i], xyz[3*i+1], xyz[3*i+2], 0);
    for (int c = 0; c < nCells; c++){
        IndexArray ids(nloc);
        for (int j = 0; j < nloc; j++) ids[j] = cells[c*nloc + j];
        mesh.createCell(ids, cellMarker[c]);
    }
    RVector cov;
    if (response && model){
        RVector r(rows), m(cols);
        for (int i = 0; i < rows; i++) r[i] = response[i];
        for (int j = 0; j < cols; j++) m[j] = model[j];
        cov = createCoverage(S, mesh, r, m);
    } else {
        cov = createCoverage(S, mesh);
    }
    for (int c = 0; c < nCells; c++) out[c] = cov[c];
}

int ref_subpot_rows(void * vh){ return (int)((RefHandle *)vh)->subPots->rows(); }
// k-resolved potentials, row = electrode + nE*kIdx (dcfemmodelling.cpp:1681)
void ref_get_subpotentials(void * vh, double * out){
    RMatrix & U = *((RefHandle *)vh)->subPots;
    for (Index i = 0; i < U.rows(); i++) std::memcpy(out + i*U.cols(), &U[i][0], U.cols()*sizeof(double));
}
// k-summed potentials solutions_ (dcfemmodelling.cpp:1707)
void ref_get_solutions(void * vh, double * out){
    const RMatrix & U = ((RefHandle *)vh)->fop->solution();
    for (Index i = 0; i < U.rows(); i++) std::memcpy(out + i*U.cols(), &U[i][0], U.cols()*sizeof(double));
}
// analytic primary potentials (SR only), same row layout
int ref_get_primary(void * vh, double * out){
    RefHandle * h = (RefHandle *)vh;
    if (!h->sr) return 0;
    RMatrix & U = dynamic_cast< DCSRMultiElectrodeModelling * >(h->fop)->primaryPotential();
    if (out) for (Index i = 0; i < U.rows(); i++) std::memcpy(out + i*U.cols(), &U[i][0], U.cols()*sizeof(double));
    return (int)U.rows();
}

// Numeric primary potentials the way checkPrimpotentials_ builds them with topography (dcfemmodelling.cpp:2009-2056):
// k-resolved potentials of a total-field run (rho = 1) on the P2 mesh of handle vp -- collected by its response() --
// interpolated to the node positions of the SR handle vs (interpolate.h:75) and handed over with setPrimaryPotential.
// (The reference's own temporary fop inside checkPrimpotentials_ cannot take the injected solver, and no CHOLMOD is
// installed here; this performs the same three steps from the outside.)
int ref_set_primary_from(void * vs, void * vp){
    RefHandle * s = (RefHandle *)vs; RefHandle * p = (RefHandle *)vp;
    if (!s->sr || p->subPots->rows() == 0) return 0;
    if (s->prim) delete s->prim;
    s->prim = new RMatrix();
    interpolate(*p->mesh, *p->subPots, s->mesh->positions(), *s->prim, false);
    dynamic_cast< DCSRMultiElectrodeModelling * >(s->fop)->setPrimaryPotential(*s->prim);
    return (int)s->prim->rows();
}
// SparseMatrix::fillStiffnessMatrix (kind 0) / fillMassMatrix (kind 1) with per-cell coefficients (sparsematrix.h:1034-1065)
int ref_fill_matrix(void * vh, int kind, const double * coef, double * vals){
    RefHandle * h = (RefHandle *)vh;
    Mesh & mesh = *h->mesh;
    RVector a(mesh.cellCount());
    for (Index i = 0; i < mesh.cellCount(); i++) a[i] = coef[i];
    RSparseMatrix S;
    if (kind == 0) S.fillStiffnessMatrix(mesh, a); else S.fillMassMatrix(mesh, a);
    if (vals) for (Index i = 0; i < S.nVals(); i++) vals[i] = S.vals()[i];
    return (int)S.nVals();
}

// CSR pattern exactly as SparseMatrix::buildSparsityPattern (sparsematrix.h:966)
int ref_pattern(void * vh, int * rowptr, int * colidx){
    RefHandle * h = (RefHandle *)vh;
    RSparseMatrix S; S.buildSparsityPattern(*h->fop->mesh());
    if (rowptr){
        std::memcpy(rowptr, &S.vecColPtr()[0], (S.rows()+1)*sizeof(int));
        std::memcpy(colidx, &S.vecRowIdx()[0], S.nVals()*sizeof(int));
    }
    return (int)S.nVals();
}
// System matrix values for wavenumber k and per-cell rho, assembled the way calculateK does
// (dcfemmodelling.cpp:2180-2182): domain + mixed boundary (+ optional calibration rows).
//   what: 0 = domain only, 1 = domain + boundary
void ref_assemble(void * vh, double k, const double * rhoCells, int what, double * vals, double * seconds){
    RefHandle * h = (RefHandle *)vh;
    Mesh & mesh = *h->fop->mesh();
    RVector rho(mesh.cellCount()); for (Index i = 0; i < rho.size(); i++) rho[i] = rhoCells[i];
    mesh.setCellAttributes(rho);
    RVector3 src(0.0, 0.0, 0.0); int n = 0;
    for (Index i = 0; i < h->fop->electrodes().size(); i++) { src += h->fop->electrodes()[i]->pos(); n++; }
    src /= (double)n;
    Stopwatch sw(true);
    RSparseMatrix S; S.buildSparsityPattern(mesh);
    double tPattern = sw.duration(true);
    dcfemDomainAssembleStiffnessMatrix(S, mesh, k, true);
    if (what > 0) dcfemBoundaryAssembleStiffnessMatrix(S, mesh, src, k);
    double tAsm = sw.duration(true);
    if (vals) std::memcpy(vals, &S.vecVals()[0], S.nVals()*sizeof(double));
    if (seconds){ seconds[0] = tPattern; seconds[1] = tAsm; }
}

// Export the (possibly refined) forward mesh as flat arrays; call with NULLs to query sizes.
//   sizes: nNodes, nCells, nloc, nBoundsMarked, nlocb
void ref_mesh_sizes(void * vh, int * sizes){
    Mesh & mesh = *((RefHandle *)vh)->fop->mesh();
    sizes[0] = (int)mesh.nodeCount(); sizes[1] = (int)mesh.cellCount();
    sizes[2] = (int)mesh.cell(0).nodeCount();
    int nb = 0, nlb = 0;
    for (Index i = 0; i < mesh.boundaryCount(); i++) if (mesh.boundary(i).marker() != 0){ nb++; nlb = (int)mesh.boundary(i).nodeCount(); }
    sizes[3] = nb; sizes[4] = nlb;
}
void ref_mesh_export(void * vh, double * xyz, int * nodeMarker, int * cells, int * cellMarker,
                     int * bounds, int * boundMarker){
    Mesh & mesh = *((RefHandle *)vh)->fop->mesh();
    for (Index i = 0; i < mesh.nodeCount(); i++){
        for (int d = 0; d < 3; d++) xyz[3*i+d] = mesh.node(i).pos()[d];
        nodeMarker[i] = mesh.node(i).marker();
    }
    Index nloc = mesh.cell(0).nodeCount();
    for (Index c = 0; c < mesh.cellCount(); c++){
        for (Index j = 0; j < nloc; j++) cells[c*nloc+j] = (int)mesh.cell(c).node(j).id();
        cellMarker[c] = mesh.cell(c).marker();
    }
    Index b = 0;
    for (Index i = 0; i < mesh.boundaryCount(); i++){
        Boundary & bd = mesh.boundary(i);
        if (bd.marker() == 0) continue;
        for (Index j = 0; j < bd.nodeCount(); j++) bounds[b*bd.nodeCount()+j] = (int)bd.node(j).id();
        boundMarker[b] = bd.marker(); b++;
    }
}

// Refine a flat P1 mesh with the reference's own createH2()/createP2() (mesh.cpp:1278-1315)
// and hand the result back through a plain handle usable with ref_mesh_sizes/ref_mesh_export.
// kind: 1 = H2, 2 = P2.  (Used to pin the package's own refinement conventions.)
struct MeshOnly { Mesh * m; };
void * ref_refine(int dim, int nNodes, const double * xyz, const int * nodeMarker,
                  int nCells, int nloc, const int * cells, const int * cellMarker,
                  int nBounds, int nlocb, const int * bounds, const int * boundMarker, int kind){
    Mesh mesh(dim);
    for (int i = 0; i < nNodes; i++) mesh.createNode(xyz[3*i], xyz[3*i+1], xyz[3*i+2], nodeMarker[i]);
    for (int c = 0; c < nCells; c++){
        IndexArray ids(nloc); for (int j = 0; j < nloc; j++) ids[j] = cells[c*nloc + j];
        mesh.createCell(ids, cellMarker[c]);
    }
    for (int b = 0; b < nBounds; b++){
        IndexArray ids(nlocb); for (int j = 0; j < nlocb; j++) ids[j] = bounds[b*nlocb + j];
        mesh.createBoundary(ids, boundMarker[b], false);
    }
    mesh.createNeighborInfos();
    RefHandle * h = new RefHandle();
    h->mesh = new Mesh(dim);
    if (kind == 1) *h->mesh = mesh.createH2(); else *h->mesh = mesh.createP2();
    h->data = 0; h->solver = 0; h->subPots = 0; h->sr = false;
    h->fop = new DCMultiElectrodeModelling(false);
    h->fop->setMesh(*h->mesh);
    return h;
}

// Reference element matrices for one cell (elementmatrix.cpp:798 ux2uy2uz2, :683 u2)
void ref_element_matrices(void * vh, int cellId, double * K, double * M){
    Mesh & mesh = *((RefHandle *)vh)->fop->mesh();
    Cell & c = mesh.cell(cellId);
    ElementMatrix < double > Se;
    Index n = c.nodeCount();
    Se.ux2uy2uz2(c);
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) K[i*n+j] = Se.getVal(i, j);
    Se.u2(c);
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) M[i*n+j] = Se.getVal(i, j);
}

// Bessel functions as the reference evaluates them (numericbase.h:80-180)
void ref_bessel(int n, const double * x, double * k0, double * k1){
    for (int i = 0; i < n; i++){ k0[i] = besselK0(x[i]); k1[i] = besselK1(x[i]); }
}
// initKWaveList(rMin, rMax, nLeg, nLag) (bertMisc.cpp:87)
void ref_kwave_list(double rMin, double rMax, int nLeg, int nLag, double * k, double * w){
    RVector kv, wv; initKWaveList(rMin, rMax, nLeg, nLag, kv, wv);
    for (Index i = 0; i < kv.size(); i++){ k[i] = kv[i]; w[i] = wv[i]; }
}

// Time the reference's stages on the current model (bench.py --impl reference):
// out = {t_response, t_createJacobian}
void ref_time_forward_jacobian(void * vh, int nModel, const double * model, int nThreads, double * out){
    RefHandle * h = (RefHandle *)vh;
    RVector m(nModel); for (int i = 0; i < nModel; i++) m[i] = model[i];
    h->fop->setThreadCount(nThreads);
    Stopwatch sw(true);
    RVector r(h->fop->response(m));
    out[0] = sw.duration(true);
    h->fop->createJacobian(m);
    out[1] = sw.duration(true);
}

} // extern "C"
