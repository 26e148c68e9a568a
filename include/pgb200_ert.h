/* pgb200_ert.h -- C ABI of the B200-native ERT forward + Jacobian path.
 *
 * Drop-in target: the work pyGIMLi does inside
 *   DCSRMultiElectrodeModelling::response(model)      core/src/bert/dcfemmodelling.cpp:1085
 *   DCSRMultiElectrodeModelling::createJacobian(model) core/src/bert/dcfemmodelling.cpp:1446
 * reached from pygimli/physics/ert/ertModelling.py:213 (`self._core.response(mod)`) and :238
 * (`self._core.createJacobian(mod)`).  Plain pointers and sizes only; no torch types.
 *
 * Life cycle:  build a pgb200_plan on the host (geometry-only, once per mesh/scheme)
 *   -> pgb200_ert_create()  uploads it, assembles the rho=1 matrices and the analytic primary
 *      potentials (both geometry-only, cached like primPot_ in dcfemmodelling.cpp:1991-1999)
 *   -> pgb200_ert_response() / pgb200_ert_create_jacobian() per Gauss-Newton step
 *   -> pgb200_ert_destroy().
 * All functions return 0 on success, non-zero on failure; pgb200_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PGB200_ERT_H
#define PGB200_ERT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgb200_ert pgb200_ert; /* opaque handle */

/* Host-side plan.  All pointers are HOST pointers, copied during create.  Index arrays are
 * int32 (the reference keeps CSR indices in `int` "to be cholmod compatible",
 * core/src/sparsematrix.h:1112-1116). */
typedef struct pgb200_plan {
    int dim;            /* 2 or 3                                                           */
    int nloc;           /* nodes per cell: 3 Tri3, 6 Tri6, 4 Tet4, 10 Tet10                 */
    int n_nodes;        /* N                                                                */
    int n_cells;        /* C                                                                */
    int nnz;            /* CSR entries                                                      */
    int n_elec;         /* nE electrodes = current patterns (pole sources, :1518-1523)      */
    int n_k;            /* wavenumbers (1 in 3-D)                                           */
    int n_model;        /* M = max cell marker + 1 (bertJacobian.cpp:280)                   */
    int n_data;         /* D                                                                */
    int sr;             /* 1: singularity removal (DCSRMultiElectrodeModelling), 0: total field */
    int fullspace;      /* 1: no surface found -> full-space analytic potentials            */
    double surface_z;   /* mirror plane of the analytic primary potential (bertMisc.cpp:196)*/

    const double *pos;          /* [N*3] node coordinates                                   */
    const int *cells;           /* [C*nloc] node ids, original cell order                   */
    const int *cell_marker;     /* [C]                                                      */
    const int *rowptr;          /* [N+1]   CSR pattern == SparseMatrix::buildSparsityPattern */
    const int *colidx;          /* [nnz]   (sparsematrix.h:966-1032), columns ascending      */
    const int *diag_pos;        /* [N]     CSR slot of the diagonal                          */

    int n_colors;               /* conflict-free cell colours for the atomic-free scatter    */
    const int *color_ptr;       /* [n_colors+1] ranges into the colour-ordered cell list     */
    const int *color_order;     /* [C]     original cell id of colour-ordered slot           */
    const int *cells_col;       /* [nloc*C]      SoA node ids in colour order                */
    const int *pos_col;         /* [nloc*nloc*C] SoA CSR slot of local entry (i,j)           */

    const double *k_values;     /* [n_k]  (bertMisc.cpp:87-105 or setkValues)                */
    const double *k_weights;    /* [n_k]                                                     */

    int n_bc_slots;             /* mixed-BC faces (marker -2, dcfemmodelling.cpp:243-299)    */
    int n_bc_entries;
    const int *bc_slot;         /* [n_bc_slots]   CSR slot                                   */
    const int *bc_ptr;          /* [n_bc_slots+1] range into entries                         */
    const int *bc_owner;        /* [n_bc_entries] owner cell (1/rho of that cell)            */
    const double *bc_coef;      /* [n_k*n_bc_entries] beta_b(k) * |face| * Uhat_ij           */

    int n_dir_zero;             /* homogeneous Dirichlet rows/cols (marker -3, :141-161)     */
    int n_dir_nodes;
    const int *dir_zero_slots;  /* [n_dir_zero]  CSR slots forced to 0                       */
    const int *dir_diag_slots;  /* [n_dir_nodes] CSR slots forced to 1                       */
    const int *dir_nodes;       /* [n_dir_nodes] rows whose RHS is forced to 0               */

    const double *el_pos;       /* [nE*3] electrode positions                                */
    const int *sing_node;       /* [nE]   node whose analytic value is patched, -1 none      */
    const double *sing_val;     /* [n_k*nE] patched value (electrode.cpp:154-189)            */
    const int *pick_ptr;        /* [nE+1] potential pick-up / delta-RHS stencil              */
    const int *pick_idx;        /*        node ids                                           */
    const double *pick_w;       /*        shape-function weights (1 for node electrodes)     */
    const int *src_cell_ptr;    /* [nE+1] cells around the electrode (rho at the source,     */
    const int *src_cells;       /*        geometric mean, electrode.cpp:102-120)             */

    int n_pro_levels;           /* background prolongation (mesh.cpp:2247-2316)              */
    int pro_nf;                 /* faces per cell                                            */
    const int *pro_level_ptr;   /* [n_pro_levels+1] ranges into pro_cells                    */
    const int *pro_cells;       /* cells filled at each level                                */
    const int *pro_nb;          /* [n*pro_nf] neighbour cells                                */
    const double *pro_w;        /* [n*pro_nf] normalised weights (0 = unused)                */

    int n_jac_cells;            /* cells with marker >= 0, sorted by marker (:298-299)       */
    const int *jac_cells;       /* [n_jac_cells]                                             */
    const int *jac_col_ptr;     /* [M+1] ranges into jac_cells                               */

    const int *abmn;            /* [D*4] electrode indices, -1 = unused                      */
    const double *k_fac;        /* [D] geometric factors                                     */

    int topography;             /* 1: non-flat surface or pure-Neumann domain (dcfemmodelling.cpp:718-754): no analytic
                                 * primary potentials / analytic branches; with sr = 1 the primary potentials must be
                                 * supplied through pgb200_ert_set_primary_dev before the first solve (:2009-2056)  */
    int ref_node;               /* reference-electrode node (marker -999, bert/bert.h:31) in the internal numbering or -1:
                                 * every current pattern is the dipole (electrode, reference), the right-hand side gets -1
                                 * there (dcfemmodelling.cpp:1009-1015, 1517-1523, 1868-1870)                     */
    int ref_last;               /* 1: pure-Neumann domain without a -999 node -- the LAST electrode is the current reference
                                 * (:1054-1064): its own pattern does not exist (zero potentials) and createJacobian fails
                                 * with the reference's length error (bertJacobian.cpp:283-291)                  */
} pgb200_plan;

/* One coarse level of the aggregation hierarchy of the multilevel preconditioner (host pointers, copied).
 * Geometry only: built once per mesh by pygimli_b200/amg_setup.py.                                         */
typedef struct pgb200_amg_level {
    int n;                  /* nodes of this level                                                       */
    int nnz;                /* CSR entries of this level                                                 */
    const int *rowptr;      /* [n+1]                                                                     */
    const int *colidx;      /* [nnz]                                                                     */
    const int *diag_pos;    /* [n]                                                                       */
    const int *gal_ptr;     /* [nnz+1] entry s of this level = sum of finer entries gal_idx[gal_ptr[s]..) */
    const int *gal_idx;     /* [nnz of the finer level]                                                  */
    const int *agg;         /* [n of the finer level] finer node -> node of this level                   */
    const int *mem_ptr;     /* [n+1] members of every aggregate ...                                      */
    const int *mem_idx;     /* [n of the finer level] ... as finer-level node ids                        */
} pgb200_amg_level;

/* ---- compiled plan builder: what the reference takes (mesh + data container) -> plan ---------------------------
 * ModellingBase::setMesh(mesh) / setData(dataContainer) (core/src/modellingbase.h:68-99) hand the reference a pointer-graph
 * mesh and a DataContainerERT; the flat-array equivalents below carry the same information (original numbering, the
 * reference's marker conventions).  pgb200_plan_build does all geometry-only set-up on the host (csrc/plan_builder.cpp:
 * pattern by sort + unique, colouring, electrode matching, wavenumbers, mixed-BC table, prolongation levels, Jacobian
 * columns), pgb200_ert_open adds the device side and the aggregation hierarchy -- one call from a reference-side binding. */
typedef struct pgb200_mesh_in {
    int dim;                    /* 2 or 3                                                                      */
    int nloc;                   /* nodes per cell: 3 / 6 / 4 / 10                                              */
    int n_nodes, n_cells;
    int n_bounds, nlb;          /* marked boundary faces and nodes per face (2 / 3 edges, 3 / 6 triangles)     */
    const double *pos;          /* [N*3]                                                                       */
    const int *node_marker;     /* [N]  -99 electrode, -999 reference, -1000 calibration (bert/bert.h:30-32)   */
    const int *cells;           /* [C*nloc]                                                                    */
    const int *cell_marker;     /* [C]  >= 0 model index, < 0 background                                       */
    const int *bounds;          /* [B*nlb]                                                                     */
    const int *bound_marker;    /* [B]  -1 Neumann (surface), -2 mixed, -3 Dirichlet (gimli.h:234-240)         */
} pgb200_mesh_in;
typedef struct pgb200_scheme_in {
    int n_elec, n_data;
    const double *sensors;      /* [nE*3] sensor positions (DataContainer::sensorPositions)                    */
    const int *abmn;            /* [D*4]  tokens a b m n, -1 = unused                                          */
    const double *k_fac;        /* [D] token k or NULL: analytic factors on a flat earth, numeric with topography */
} pgb200_scheme_in;
typedef struct pgb200_built_plan pgb200_built_plan;     /* owns every array of the plan */
/* n_k_user > 0: wavenumbers / weights set by the caller (setkValues / setWeights, dcfemmodelling.h:239-243) */
int pgb200_plan_build(const pgb200_mesh_in *mesh, const pgb200_scheme_in *scheme, int sr, int n_k_user, const double *k_user,
                      const double *w_user, pgb200_built_plan **out);
int pgb200_plan_free(pgb200_built_plan *plan);
const char *pgb200_plan_error(void);
const pgb200_plan *pgb200_plan_view(const pgb200_built_plan *plan);
/* named arrays / scalars of a built plan for host-side consumers and tests; type: 0 int32, 1 float64, 2 int64.  Names:
 * node_perm node_inv pos cells cell_marker rowptr colidx diag_pos ref_rowptr ref_colidx ref_slot color_ptr color_order
 * cells_col pos_col k w bc_slot bc_ptr bc_owner bc_coef dir_zero_slots dir_diag_slots dir_nodes el_pos sing_val pick_w
 * min_radius el_node el_node_ref el_cell sing_node pick_ptr pick_idx src_cell_ptr src_cells pro_level_ptr pro_cells pro_nb
 * pro_w jac_cells jac_col_ptr abmn k_fac, level<l>.{rowptr,colidx,diag_pos,gal_ptr,gal_idx,agg,mem_ptr,mem_idx}      */
int pgb200_plan_array(const pgb200_built_plan *plan, const char *name, const void **ptr, long long *count, int *type);
int pgb200_plan_scalar(const pgb200_built_plan *plan, const char *name, double *out);
/* aggregation hierarchy from the rho = 1 values of the first wavenumber (host); returns the number of levels, < 0 on error */
int pgb200_plan_build_hierarchy(pgb200_built_plan *plan, const double *vals1, double theta, int passes, int min_size, int max_levels);
const pgb200_amg_level *pgb200_plan_levels(const pgb200_built_plan *plan);

/* ---- host-only helpers (no GPU needed) -------------------------------------------- */
const char *pgb200_last_error(void);
int pgb200_version(void);
/* Greedy conflict colouring: cells sharing a node get different colours.
 * Returns the number of colours (<= 0 on failure); color[C] receives the colour per cell. */
int pgb200_color_cells(int n_cells, int nloc, const int *cells, int n_nodes, int *color);
/* Streamed row panels of the SpMM kernel (csrc/stream_panels.h) for a CSR pattern; the library builds them itself inside
 * pgb200_ert_create / pgb200_ert_set_hierarchy, this entry point exists for the host-side tests.  Two calls: with
 * panel_row_ptr == NULL only counts[10] = {panels, chunks, halo entries, crp_stride, max rows, max chunk halo, max chunk
 * entries, nnz, runs, max runs per chunk} is filled; then with arrays of those sizes (runs[3 * runs] = {first halo entry
 * in its chunk, column, length}).  Returns 0 on success.                                                       */
int pgb200_build_stream_panels(int n_rows, const int *rowptr, const int *colidx, int rmax, int hc, int max_chunks, int *counts,
                               int *panel_row_ptr, int *panel_chunk_ptr, int *chunk_halo_ptr, int *halo_cols, int *chunk_ent_ptr,
                               int *ent_src, unsigned *ent_idx, int *crp, int *chunk_run_ptr, int *runs);

/* The same panels in the 8-row-group form of the FP64 tensor-core SpMM (k_spmm_mma): rows of a panel in groups of 8, per
 * (chunk, group) k-steps of 4 union columns.  counts[8] = {panels, chunks, halo entries, k-steps, meta words, meta group
 * stride, max k-steps per chunk, max meta words per chunk}; a_src[32 * k-steps] = CSR slot of A-fragment element
 * (4 * row-in-group + column-in-step) or -1; meta per chunk = [k-step range start per group | one word per k-step with the
 * four staged-row indices, one byte each].  Host-side tests only.  Returns 0 on success.                        */
int pgb200_build_mma_panels(int n_rows, const int *rowptr, const int *colidx, int groups, int hc, int max_chunks, int rowb_hint,
                            int *counts, int *panel_row_ptr, int *panel_chunk_ptr, int *chunk_halo_ptr, int *halo_cols,
                            int *chunk_ks_ptr, int *a_src, int *chunk_meta_ptr, unsigned *meta);

/* Greedy pairwise aggregation along the strongest negative coupling (multilevel preconditioner set-up); a pair is
 * formed only if the coupling is at least theta times the strongest coupling of BOTH nodes, left-over nodes join a
 * neighbouring aggregate only across such a strong coupling (theta = 0: unconditional matching).  agg[n] receives the
 * aggregate id per node; returns the number of aggregates, -1 on a null argument.  With group != NULL only nodes of the
 * same group are matched (aggregates stay inside SpMM row panels).                               */
int pgb200_pairwise_aggregate(int n, const int *rowptr, const int *colidx, const double *vals, const int *group, double theta,
                              int *agg);

/* ---- life cycle ------------------------------------------------------------------- */
int pgb200_ert_create(const pgb200_plan *plan, int device, pgb200_ert **out);
/* mesh + scheme -> ready handle: pgb200_plan_build + pgb200_ert_create + aggregation hierarchy (multilevel = 1) in one call.
 * The handle owns the built plan (pgb200_ert_plan); n_k_user / k_user / w_user as in pgb200_plan_build.           */
int pgb200_ert_open(const pgb200_mesh_in *mesh, const pgb200_scheme_in *scheme, int sr, int n_k_user, const double *k_user,
                    const double *w_user, int multilevel, int device, pgb200_ert **out);
/* the same for a plan the caller built and keeps alive (takes no ownership): device set-up + hierarchy */
int pgb200_ert_open_plan(pgb200_built_plan *plan, int multilevel, int device, pgb200_ert **out);
const pgb200_built_plan *pgb200_ert_plan(const pgb200_ert *h);
int pgb200_ert_destroy(pgb200_ert *h);
/* Install (n_levels > 0) or remove the aggregation hierarchy of the multilevel preconditioner.  */
int pgb200_ert_set_hierarchy(pgb200_ert *h, int n_levels, const pgb200_amg_level *levels);
/* multilevel = 0: Jacobi-PCG, 1: V(1,1) aggregation multigrid preconditioner (default when a hierarchy is
 * installed); coarse_sweeps: damped-Jacobi sweeps on the coarsest level (default 8).              */
int pgb200_ert_set_preconditioner(pgb200_ert *h, int multilevel, int coarse_sweeps);
/* replay blocks of 6 multilevel-PCG iterations as one CUDA graph (default on; off while profiling) */
int pgb200_ert_set_graph(pgb200_ert *h, int on);
/* on != 0: the block-PCG starts from the potentials of the previous solve on this handle instead of zero (Gauss-Newton /
 * time-lapse loops, where consecutive models differ little).  Off by default; the convergence criterion is unchanged
 * (||r|| <= tol ||b|| per source column).                                                      */
int pgb200_ert_set_warm_start(pgb200_ert *h, int on);
/* CUDA stream (cudaStream_t) all work is enqueued on; 0/NULL = the handle's own blocking stream,
 * which is implicitly ordered with the legacy default stream.                                */
int pgb200_ert_set_stream(pgb200_ert *h, void *stream);
/* Block-PCG controls: relative residual tolerance ||r||/||b|| per source column,
 * iteration cap, and how many iterations run between convergence checks.               */
int pgb200_ert_set_solver(pgb200_ert *h, double rel_tol, int max_iter, int check_every);
/* Restrict this handle to a shard: current sources [src_begin, src_end) are solved here,
 * data rows [row_begin, row_end) are written by the Jacobian (multi-GPU).               */
int pgb200_ert_set_shard(pgb200_ert *h, int src_begin, int src_end, int row_begin, int row_end);
/* Replace geometric factors (token "k" of the data container).                          */
int pgb200_ert_set_kfac(pgb200_ert *h, const double *k_fac_host);

/* ---- the path: HOST buffers (what a reference-side binding calls) ------------------- */
/* response: model_host[n_model_in] (n_model_in == M: per marker, == C: per cell,
 * dcfemmodelling.cpp:1211-1218)  ->  rhoa_host[D] = sqrt(|resp * respRez|) (:1196).     */
int pgb200_ert_response(pgb200_ert *h, const double *model_host, int n_model_in, double *rhoa_host);
/* createJacobian: uses the potentials of the last response() if present (:1262), else
 * solves (analytic branch for homogeneous models, :1272-1301).  J stays in HBM.          */
int pgb200_ert_create_jacobian(pgb200_ert *h, const double *model_host, int n_model_in);
/* mapERTModel (dcfemmodelling.cpp:1211-1218): model vector -> cell resistivities rho_cells_host[C]      */
int pgb200_ert_map_model(pgb200_ert *h, const double *model_host, int n_model_in, double *rho_cells_host);
/* copy J to the host, row-major [D_local x M].                                           */
int pgb200_ert_jacobian_copy(pgb200_ert *h, double *j_host);
/* y = J x  and  y = J^T x  with host vectors (jacobian().mult / transMult).              */
int pgb200_ert_jacobian_mult(pgb200_ert *h, const double *x_host, double *y_host);
int pgb200_ert_jacobian_tmult(pgb200_ert *h, const double *x_host, double *y_host);
/* Error-/transform-weighted Jacobian of the inversion (MultLeftRightMatrix, pygimli/frameworks/inversion.py:705-708):
 *   mult_lr   y[D_local] = left .* (J (right .* x)),    tmult_lr   y[M] = right .* (J^T (left .* x)).
 * left has D_local entries, right has M; either may be NULL (= ones).  J never leaves HBM.   */
int pgb200_ert_jacobian_mult_lr(pgb200_ert *h, const double *left_host, const double *right_host, const double *x_host, double *y_host);
int pgb200_ert_jacobian_tmult_lr(pgb200_ert *h, const double *left_host, const double *right_host, const double *x_host, double *y_host);
/* coverageDCtrans (core/src/bert/bertJacobian.cpp:569-598, RMatrix branch): cov[j] = sum_i |J_ij dd_i| / |mm_j|.
 * dd has D_local entries, mm has M; mm == NULL returns the undivided column sums (the partial result of a row shard,
 * to be summed over the ranks and divided afterwards).                                        */
int pgb200_ert_coverage_trans(pgb200_ert *h, const double *dd_host, const double *mm_host, double *cov_host);

/* Generic FEM matrices on the path's element kernels (SparseMatrix::fillStiffnessMatrix / fillMassMatrix,
 * core/src/sparsematrix.h:1034-1065, as used by pygimli/solver/solver.py:1953, 2033):
 *   vals = sum_c a[c] * int grad N_i . grad N_j  +  b[c] * int N_i N_j   over the handle's mesh, no boundary terms.
 * a_cells_host / b_cells_host: [C] per-cell coefficients in the plan's cell order, either may be NULL (term absent);
 * vals_host: [nnz] in the plan's CSR order.                                                  */
int pgb200_ert_fill_matrix(pgb200_ert *h, const double *a_cells_host, const double *b_cells_host, double *vals_host);

/* Numeric primary potentials for the singularity-removal path with topography (checkPrimpotentials_,
 * dcfemmodelling.cpp:2009-2056: total-field solve for rho = 1 on the P2-refined mesh, taken at the nodes of this mesh).
 * src_dev: node-major potentials of the primary handle ([n_src_nodes][src_ld], column = electrode + nE * k, see
 * pgb200_ert_potentials_info); row_map_host[i] = row of src_dev holding node i of THIS handle (plan numbering).  */
int pgb200_ert_set_primary_dev(pgb200_ert *h, const double *src_dev, long long src_ld, const int *row_map_host);

/* ---- the path: DEVICE buffers (inputs already resident in HBM) ---------------------- */
int pgb200_ert_response_dev(pgb200_ert *h, const double *model_dev, int n_model_in, double *rhoa_dev);
int pgb200_ert_create_jacobian_dev(pgb200_ert *h, const double *model_dev, int n_model_in);
/* J in HBM is stored column-major: element (d, j) at ptr[j * ld + d] (each model column is a
 * contiguous run of data rows, written with coalesced 128-bit stores).                    */
int pgb200_ert_jacobian_info(pgb200_ert *h, void **dev_ptr, int *rows, int *cols, long long *ld);
/* drop cached potentials (mesh/data change semantics, dcfemmodelling.cpp:712, :777)       */
int pgb200_ert_clear_potentials(pgb200_ert *h);
/* device pointer + leading dimension of the k-resolved potentials U[node][ld], column
 * s = electrode + nE * kIdx (subSolutions_ transposed, :1681); used for the NCCL all-gather */
int pgb200_ert_potentials_info(pgb200_ert *h, void **dev_ptr, int *n_nodes, int *n_src, long long *ld);
/* after the caller's all-gather filled the other shards' columns; fails if this handle's own shard was never solved */
int pgb200_ert_mark_potentials_valid(pgb200_ert *h);
/* bit 0: this handle's source shard holds solved potentials; bit 1: all nS columns are valid (the Jacobian may use
 * them).  < 0 on a null handle.                                                              */
int pgb200_ert_potentials_state(pgb200_ert *h);

/* ---- multi-GPU staging (one handle per GPU; the exchange itself is NCCL in the caller) - */
/* solve this shard's sources; leaves U[:, shard] and the PARTIAL electrode matrix in HBM    */
int pgb200_ert_forward_dev(pgb200_ert *h, const double *model_dev, int n_model_in);
/* device pointer of the nE x nE electrode-potential matrix (all-reduce SUM across shards)   */
int pgb200_ert_pm_info(pgb200_ert *h, void **dev_ptr, int *n);
/* ABMN + reciprocity combine from the (reduced) electrode matrix -> rhoa_dev[D]             */
int pgb200_ert_finish_response_dev(pgb200_ert *h, double *rhoa_dev);
/* pack (unpack=0) / unpack (unpack=1) the potential columns [c0,c1) to/from a contiguous
 * [N x (c1-c0)] device buffer, the unit of the all-gather of potentials                     */
int pgb200_ert_pack_potentials(pgb200_ert *h, int c0, int c1, double *buf_dev, int unpack);

/* ---- introspection for parity tests / bench ----------------------------------------- */
/* what: "vals" [nK*nnz], "vals1" [nK*nnz] (rho=1), "rho" [C], "rho_src" [nE], "prim" [nS*N]
 * (row = e + nE*k), "pots" [nS*N], "solutions" [nE*N], "pm" [nE*nE], "resp" [D], "resp_rez" [D],
 * "rel_res" [nS].  Returns the number of doubles written (or needed when out == NULL), < 0 on error. */
long long pgb200_ert_get(pgb200_ert *h, const char *what, double *out_host, long long capacity);
/* stats[0]=PCG iterations of the last solve, [1]=max relative residual, [2]=kernel launches
 * since create/reset, [3..8] = accumulated ms since reset: map, assemble, rhs, solve, epilogue,
 * jacobian, [9]=SpMM launches timed, [10]=their total ms, [11]=Jacobian kernel ms (total),
 * [12]=Jacobian passes timed, [13]=PCG iterations since reset, [14]=solves since reset, [15]=algorithmic bytes of
 * the timed SpMM launches (each with its own active column window), [16]=warm-started solves since reset   */
int pgb200_ert_stats(pgb200_ert *h, double *stats, int n);
int pgb200_ert_reset_stats(pgb200_ert *h);
/* which code paths the last solve / Jacobian plan took (parity tests assert that the kernels the benchmark times are
 * the ones compared with the reference): [0] column pairs per lane of the streamed SpMM (0 = plain gather kernel),
 * [1] column tiles, [2] unused, [3] CUDA-graph launches of the last solve, [4] Jacobian chunks, [5] register tiles per
 * thread of the widest chunk, [6] 1 if every chunk uses pre-resolved Gram offsets, [7] coarse levels of the multilevel
 * preconditioner, [8] shared-memory slots of the streamed kernel's ring, [9] coarse levels on the streamed kernel  */
int pgb200_ert_path_info(pgb200_ert *h, int *out, int n);
/* ---- complex resistivity (induced polarisation): DCMultiElectrodeModelling with setComplex(true) ------------------------
 * Replaces CSparseMatrix assembly (core/src/bert/dcfemmodelling.cpp:235-242), the complex total-field solves (:1755-1925,
 * cholmodWrapper.cpp:167-222) and createJacobian_(CVector, CMatrix) (:1410-1444).  The handle must be opened as a TOTAL-FIELD
 * problem (sr = 0) on a scheme that lists every electrode twice (sensors = [s_0..s_{n-1}, s_0..s_{n-1}]) and whose rows are the
 * four real blocks of the complex sensitivity -- (a,b,m,n), (a+n,b+n,m+n,n+n), (a,b,m+n,n+n), (a+n,b+n,m,n) for every datum --,
 * with the wavenumbers of the original electrode list; pygimli_b200.CoreB200.setComplex(True) builds exactly that.  Inside a
 * wavenumber group the first n source columns carry the real parts of the potentials, the next n the imaginary parts.
 * S = S_r + i S_i is complex symmetric: conjugate-orthogonal CG preconditioned by the real multilevel cycle of S_r.      */
int pgb200_ert_set_complex(pgb200_ert *h, int on);
/* model_host = [Re rho (n_in) | Im rho (n_in)].  After the call pgb200_ert_pm_info holds the doubled electrode matrix
 * (row i < n: real part of the potentials of source i at the electrodes, row n + i: imaginary part).              */
int pgb200_ert_complex_forward(pgb200_ert *h, const double *model_host, int n_in);
/* j_host: D x M complex values (D = scheme rows / 4), row-major, interleaved (re, im); rows scaled by kfac_host[d] / m_j^2
 * when n_in == M.  Solves first if the handle has no potentials of a complex forward call.                        */
int pgb200_ert_complex_jacobian(pgb200_ert *h, const double *model_host, int n_in, const double *kfac_host, double *j_host);

/* Measurement aid: `reps` back-to-back launches of the fine-level streamed SpMM (role 0: SpMM + p.Ap, 1: post-smoothing +
 * r.z, 2: residual) on the assembled matrix and the PCG work vectors (overwritten); CUDA-event time per launch in ms. */
int pgb200_ert_bench_spmm(pgb200_ert *h, int role, int reps, double *ms_per_launch);
int pgb200_ert_set_profile(pgb200_ert *h, int on);
/* on == 2 additionally records one CUDA event per kernel launch (no CUDA graph); pgb200_ert_get_trace returns, for the
 * launches since then, (source line in csrc/pgb200_ert.cu) * 256 + 16 * role + multilevel level of each launch (role of a
 * streamed SpMM launch: 1 SpMM, 2 post-smoothing, 3 residual; 0 otherwise) and the time since the previous launch
 * finished [ms]: a warm per-kernel timeline of a step (profiles/summarize_trace.py).  Returns the entry count.  */
int pgb200_ert_get_trace(pgb200_ert *h, int *lines, float *ms, int cap);
/* 0: plain gather kernels (A/B evidence); non-zero (default): the streamed, warp-specialised row-panel kernel */
int pgb200_ert_set_spmm_variant(pgb200_ert *h, int panel_staged);

/* ---- single-kernel entry points (device pointers; unit tests and micro-benchmarks) --- */
/* Y[N x ld] = A X with per-wavenumber values: column s uses vals[(s / nE) * nnz + .]       */
int pgb200_spmm(const int *rowptr, const int *colidx, const double *vals, long long nnz,
                const double *X, double *Y, int n_rows, int n_elec, int n_k, long long ld, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PGB200_ERT_H */
