// pgb200_ert.cu -- host side of the C ABI declared in include/pgb200_ert.h.
// Owns the device memory of one (mesh, scheme) plan and sequences the kernels of
// ert_kernels.cuh on one CUDA stream.  No CPU fallback anywhere: every compute entry
// point needs a CUDA device and fails loudly otherwise.
#include "../../include/pgb200_ert.h"
#include "ert_kernels.cuh"
#include "stream_panels.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <initializer_list>
#include <string>
#include <vector>

using namespace pgb;

namespace {

thread_local std::string g_err;

#define PGB_FAIL(msg) do { g_err = std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; return 1; } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; return 1; } } while (0)
#define CKR(call) do { int r_ = (call); if (r_) return r_; } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    int alloc(size_t count) {
        release(); n = count;
        if (count == 0) return 0;
        CK(cudaMalloc((void **)&p, count * sizeof(T)));
        return 0;
    }
    int upload(const T *host, size_t count, cudaStream_t st) {
        CKR(alloc(count));
        if (count) CK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct JacChunk {
    int nP, nPp, nd, idx_in_smem, kfac_in_smem, out_identity, resolved, mt, threads;
    size_t plist_off, data_off;   // offsets into the concatenated device arrays
    size_t smem;
};

struct Jac2Chunk {
    int nd, nP, nPp, n_tiles, PL, mt, out_identity, out_base, kfac_in_smem, off_in_smem;
    size_t data_off, plist_off, tile_off;       // offsets into the concatenated device arrays (data rows, P basis, tiles)
    size_t smem; unsigned slot_bytes, rec_bytes, uq_off;
};

enum Phase { PH_MAP = 0, PH_ASM, PH_RHS, PH_SOLVE, PH_EPI, PH_JAC, PH_COUNT };

} // namespace

struct GraphKey {
    int c0, c1; double tol; int use_panels, levels, sweeps; void *stream; void *vals;
    bool operator==(const GraphKey &o) const {
        return c0 == o.c0 && c1 == o.c1 && tol == o.tol && use_panels == o.use_panels &&
               levels == o.levels && sweeps == o.sweeps && stream == o.stream && vals == o.vals;
    }
};

// device side of stream_panels.h for one matrix level, plus the packed entry arrays of its two value sets
struct StreamDev {
    bool ok = false;
    int n_panels = 0, n_chunks = 0, crp_stride = 0, max_chunk_halo = 0, max_chunk_ent = 0; size_t nnz = 0;
    DevBuf<int> panel_row_ptr, panel_chunk_ptr, chunk_halo_ptr, halo_cols, chunk_ent_ptr, crp, ent_src, chunk_run_ptr, runs;
    DevBuf<unsigned> ent_idx;
    DevBuf<PanelEntry> ent_a, ent_dw;       // [nK][nnz]: A (SpMM, post-smoothing) and A * diag(dw) (pre-smoothing residual)
    // 8-row-group form (k_spmm_mma): when mma is set the arrays above that describe single entries are not allocated
    bool mma = false; int max_chunk_ks = 0, max_chunk_meta = 0, hc_used = 0; size_t n_frag = 0;
    DevBuf<int> chunk_ks_ptr, chunk_meta_ptr, a_src, cdesc; DevBuf<unsigned> meta;
    DevBuf<double> av_a, av_dw;             // [nK][n_frag] packed A fragments of the two value sets
    // gather form (k_spmm_gather): built by pgb200_ert_set_shard for narrow source shards
    bool g_ok = false; int g_groups = 0, g_rows = 0; size_t g_frag = 0;
    DevBuf<int> g_ks_ptr, g_cols, g_a_src; DevBuf<double> g_av_a, g_av_dw;
    MmaLevel mlevel() const {
        MmaLevel L; L.panel_row_ptr = panel_row_ptr.p; L.panel_chunk_ptr = panel_chunk_ptr.p; L.chunk_halo_ptr = chunk_halo_ptr.p;
        L.halo_cols = halo_cols.p; L.chunk_ks_ptr = chunk_ks_ptr.p; L.chunk_meta_ptr = chunk_meta_ptr.p; L.chunk_run_ptr = chunk_run_ptr.p;
        L.runs = runs.p; L.meta = meta.p; L.n_panels = n_panels; L.cdesc = reinterpret_cast<const int4 *>(cdesc.p);
        return L;
    }
    StreamLevel level() const {
        StreamLevel L; L.panel_row_ptr = panel_row_ptr.p; L.panel_chunk_ptr = panel_chunk_ptr.p; L.chunk_halo_ptr = chunk_halo_ptr.p;
        L.halo_cols = halo_cols.p; L.chunk_ent_ptr = chunk_ent_ptr.p; L.crp = crp.p; L.n_panels = n_panels; L.crp_stride = crp_stride;
        L.chunk_run_ptr = chunk_run_ptr.p; L.runs = runs.p;
        return L;
    }
};

struct AmgLevel {      // one coarse level of the aggregation hierarchy (device arrays)
    int n = 0; size_t nnz = 0; int n_finer = 0;
    DevBuf<int> rowptr, colidx, diag_pos, gal_ptr, gal_idx, agg, mem_ptr, mem_idx;
    DevBuf<double> vals, vals_dw, dinvw, R, X, Z;
    StreamDev stream;      // streamed row panels (levels that are large enough)
};

struct pgb200_ert {
    // sizes
    int dim = 0, nloc = 0, elem = 0, N = 0, C = 0, nE = 0, nK = 0, nS = 0, M = 0, D = 0, sr = 1, fullspace = 0, topography = 0;
    int ref_node = -1, ref_last = 0;          // current reference of the dipole patterns (plan.ref_node / plan.ref_last)
    bool prim_set = false;
    size_t nnz = 0, ld = 0;
    double surface_z = 0.0;
    int device = 0;
    cudaStream_t st = 0;
    // solver controls
    double tol = 1e-12; int max_iter = 20000; int check_every = 25;
    // shard
    int c0 = 0, c1 = 0, row0 = 0, row1 = 0;
    // plan (device)
    DevBuf<double> pos, kvals, kw, bc_coef, el_pos, sing_val, pick_w, pro_w, kfac;
    DevBuf<int> cells, cell_marker, rowptr, colidx, diag_pos, color_order, cells_col, pos_col, bc_slot, bc_ptr, bc_owner,
        dir_zero, dir_diag, dir_nodes, sing_node, pick_ptr, pick_idx, src_cell_ptr, src_cells, pro_cells, pro_nb,
        jac_cells, jac_col_ptr, abmn;
    std::vector<int> color_ptr, pro_level_ptr;
    StreamDev stream; int use_panels = 1, use_panels_build = 1;   // streamed row panels of the fine level (use_panels 0: plain gather SpMM, A/B evidence)
    int stream_rmax = std::min(60, ST_CONSUMER_WARPS * ST_RPW), stream_hc = 104, stream_chunks = 2; // panel limits (stream_panels.h)
    int use_mma = 1, mma_hc = 104, mma_chunks = 10, mma_dbg = 0; DevBuf<long long> mma_dbg_buf;
    // complex resistivity (pgb200_ert_set_complex): plan built with every electrode listed twice, see ert_kernels.cuh
    int is_complex = 0; DevBuf<double> c_rho_r, c_y2, c_scal, c_model, c_out;      // FP64 tensor-core form of the streamed SpMM (k_spmm_mma); 0: k_spmm_stream
    DevBuf<double> dot_part; DevBuf<unsigned> dot_counter;    // deterministic column dots: per-CTA partial rows + tickets
    int dot_slots = 0; size_t smem_optin = 0;
    std::vector<double> h_kvals;
    int n_colors = 0, n_bc_slots = 0, n_bc_entries = 0, n_dir_zero = 0, n_dir_nodes = 0, pro_nf = 0, n_jac_cells = 0;
    std::vector<int> h_abmn; std::vector<double> h_kfac;
    // state (device)
    DevBuf<double> model, rho, rho_src, vals, vals1, dinv, prim, B, X, R, P, AP, U, scal, pM, resp, resp_rez, rhoa, Jt, tmp, xin, yout;
    DevBuf<int> flags;
    bool pots_valid = false, shard_solved = false, have_vals = false;   // pots_valid: ALL nS columns of U hold potentials of one model
    int model_len = 0;
    std::vector<double> h_model;
    // jacobian plan
    DevBuf<int> j_plist, j_qlist, j_out, j_cta_ptr; DevBuf<JacDatum> j_idx; std::vector<int> h_jac_col_ptr; int j_grid = 0;
    DevBuf<double> j_kfac;
    std::vector<JacChunk> chunks; int nQ = 0, nQp = 0;
    // second-generation Jacobian (k_jacobian2): element records, basis lists, tile lists, pre-resolved term offsets
    int jac_v2 = 1, recn = 0;
    DevBuf<double> erec, mu_tab, UDp, UDq, j2_kfac;
    DevBuf<int> j2_pa, j2_pb, j2_qa, j2_qb, j2_out;
    DevBuf<unsigned short> j2_tp, j2_tq, j2_off;
    std::vector<Jac2Chunk> chunks2; int j2_nQ = 0, j2_nQp = 0, j2_shared = 0, j2_terms = 0, j2_grid = 0; bool j2_ok = false;
    size_t ldJ = 0; int j_rows = 0; bool jac_valid = false;
    // multilevel preconditioner
    std::vector<AmgLevel *> amg; int use_amg = 1, coarse_sweeps = 8; DevBuf<double> Z0, X0, dinvw0, vals_dw0; DevBuf<unsigned long long> gmax;
    std::vector<double> col_relres; int shrink_window = 1;
    // CUDA graph of 6 PCG iterations
    int use_graph = 1; cudaGraphExec_t gexec = nullptr; GraphKey gkey{}; int glaunches = 0; long long launches_per_block = 0;
    cudaStream_t own_st = nullptr;
    // stats
    int last_iters = 0; double last_relres = 0.0; long long launches = 0;
    cudaEvent_t ev[PH_COUNT + 1]; bool ev_ok = false; float ph_ms[PH_COUNT] = {0};
    bool ph_rec[PH_COUNT + 1] = {false};
    std::vector<cudaEvent_t> tev; std::vector<int> tline; int n_tev = 0; int trace = 0; int cur_tag = 0;   // tag: multilevel level of the launch
    int cur_role = 0;   // trace only: epilogue role of a streamed SpMM launch (1 SpMM, 2 post-smoothing, 3 residual)
    int prof = 0; std::vector<cudaEvent_t> pev; int n_pev = 0; double spmm_ms = 0.0, spmm_bytes = 0.0; int spmm_timed = 0; double jac_ms = 0.0;
    cudaEvent_t jev[2]; bool jac_timed = false; int jac_launches = 0; long long total_iters = 0; int solves = 0;
    double *h_pinned = nullptr; size_t h_pinned_n = 0;
    int num_sms = 148;
    int use_subcycle = 1;      // 1: coarsest level fused (default), 2: all levels below STREAM_MIN_ROWS, 0: off
    size_t sub_smem[5] = {0, 0, 0, 0, 0};      // fused small multigrid levels (k_amg_subcycle)
    int warm_start = 0, warm_used = 0; bool x_warm_ok = false;   // initial guess = previous solution (Gauss-Newton loops)
    pgb200_built_plan *built = nullptr; bool owns_built = false;    // pgb200_ert_open: the plan the handle was built from
    // which code paths the last solve / Jacobian took (pgb200_ert_path_info)
    int pi_panel_nc = 0, pi_tiles = 0, pi_two_k = 0, pi_graph_launches = 0, pi_slots = 0, pi_stream_levels = 0;
};

inline void note_launch(pgb200_ert *h, int line) {
    h->launches++;
    if (h->trace && h->n_tev < (int)h->tev.size()) { cudaEventRecord(h->tev[h->n_tev], h->st); h->tline[h->n_tev++] = line * 256 + ((h->cur_role & 15) << 4) + (h->cur_tag & 15); }
}
// every kernel launch is counted; in trace mode (set_profile(h, 2)) an event is recorded after each one, tagged with
// the source line of the launch, so that consecutive events give a warm per-kernel timeline of one step
#define LAUNCH(h) note_launch((h), __LINE__)

namespace {

int ensure_pinned(pgb200_ert *h, size_t n) {
    if (h->h_pinned_n >= n) return 0;
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned = nullptr; h->h_pinned_n = 0;
    CK(cudaMallocHost((void **)&h->h_pinned, n * sizeof(double)));
    h->h_pinned_n = n;
    return 0;
}

template <int E>
int launch_assemble(pgb200_ert *h, const double *rho, double *vals) {
    CK(cudaMemsetAsync(vals, 0, sizeof(double) * h->nnz * h->nK, h->st));
    for (int c = 0; c < h->n_colors; c++) {
        const int first = h->color_ptr[c], count = h->color_ptr[c + 1] - first;
        if (count <= 0) continue;
        k_assemble<E><<<cdiv(count, 128), 128, 0, h->st>>>(h->pos.p, h->cells_col.p, h->pos_col.p, h->color_order.p, rho,
                                                          h->C, first, count, h->kvals.p, h->nK, h->nnz, vals);
        LAUNCH(h);
    }
    CK(cudaGetLastError());
    return 0;
}

template <int E>
int launch_assemble_generic(pgb200_ert *h, const double *a, const double *b, double *vals) {
    CK(cudaMemsetAsync(vals, 0, sizeof(double) * h->nnz, h->st));
    for (int c = 0; c < h->n_colors; c++) {
        const int first = h->color_ptr[c], count = h->color_ptr[c + 1] - first;
        if (count <= 0) continue;
        k_assemble_generic<E><<<cdiv(count, 128), 128, 0, h->st>>>(h->pos.p, h->cells_col.p, h->pos_col.p, h->color_order.p, a, b,
                                                                  h->C, first, count, vals);
        LAUNCH(h);
    }
    CK(cudaGetLastError());
    return 0;
}

int assemble(pgb200_ert *h, const double *rho, double *vals) {
    switch (h->elem) {
        case TRI3:  CKR(launch_assemble<TRI3>(h, rho, vals)); break;
        case TRI6:  CKR(launch_assemble<TRI6>(h, rho, vals)); break;
        case TET4:  CKR(launch_assemble<TET4>(h, rho, vals)); break;
        case TET10: CKR(launch_assemble<TET10>(h, rho, vals)); break;
        default: PGB_FAIL("unknown element type");
    }
    if (h->n_bc_slots) {
        k_boundary_add<<<cdiv(h->n_bc_slots, 128), 128, 0, h->st>>>(h->bc_slot.p, h->bc_ptr.p, h->bc_owner.p, h->bc_coef.p,
                                                                   h->n_bc_slots, h->n_bc_entries, rho, h->nK, h->nnz, vals);
        LAUNCH(h);
    }
    if (h->n_dir_zero) {
        k_dirichlet_zero<<<cdiv(h->n_dir_zero, 128), 128, 0, h->st>>>(h->dir_zero.p, h->n_dir_zero, h->nK, h->nnz, vals); LAUNCH(h);
        k_dirichlet_diag<<<cdiv(h->n_dir_nodes, 128), 128, 0, h->st>>>(h->dir_diag.p, h->n_dir_nodes, h->nK, h->nnz, vals); LAUNCH(h);
    }
    CK(cudaGetLastError());
    return 0;
}

template <int MODE, bool DOT>
int launch_spmm(pgb200_ert *h, const double *vals, const double *vals1, const double *rho_src, const double *X, double *Y,
                int c0, int c1, double *dots) {
    const int ncols = c1 - c0;
    if (ncols <= 0) return 0;
    // choose columns per thread to minimise idle lanes in the last column block
    int best = 1; double beste = 0.0;
    for (int cpt : {4, 2, 1}) {
        const int w = SPMM_TX * cpt; const double eff = (double)ncols / (double)(cdiv(ncols, w) * w);
        if (eff > beste + 0.05) { beste = eff; best = cpt; }
    }
    dim3 block(SPMM_TX, SPMM_TY), grid(cdiv(h->N, SPMM_ROWS), cdiv(ncols, SPMM_TX * best));
#define SPMM_GO(CPT) k_spmm<CPT, MODE, DOT><<<grid, block, 0, h->st>>>(h->rowptr.p, h->colidx.p, vals, vals1, rho_src, h->nnz, X, Y, h->N, h->nE, c0, c1, h->ld, dots)
    if (best == 4) SPMM_GO(4); else if (best == 2) SPMM_GO(2); else SPMM_GO(1);
#undef SPMM_GO
    LAUNCH(h);
    return 0;
}

// ---- streamed row-panel SpMM (the PCG hot kernel; ert_kernels.cuh k_spmm_stream) -------------------------------------
// host layout -> device, once per pattern
int stream_upload(pgb200_ert *h, StreamDev &D, int n, const int *rowptr_host, const int *colidx_host) {
    D.ok = false; D.mma = false;
    if (!h->use_panels_build) return 0;
    StreamPanelsHost S;
    std::string err = "off";
    if (h->use_mma) {
        // staged-row pitch of the most common launch: the whole source block for one wavenumber, else one electrode group
        const int rowb_hint = 8 * (h->nK == 1 ? (int)h->ld : ((h->nE + 1) & ~1));
        D.hc_used = std::max(h->mma_hc, (MM_ROWS + 1) / 2);
        err = build_stream_panels(n, rowptr_host, colidx_host, MM_ROWS, D.hc_used, h->mma_chunks, S, MM_CONSUMER_WARPS, rowb_hint);
        D.mma = err.empty() && S.max_chunk_halo <= 128;
    }
    if (!D.mma) err = build_stream_panels(n, rowptr_host, colidx_host, h->stream_rmax, h->stream_hc, h->stream_chunks, S);
    if (!err.empty()) return 0;                  // (a row wider than the halo limit) -> plain kernels for this level
    cudaStream_t st = h->st;
    CKR(D.panel_row_ptr.upload(S.panel_row_ptr.data(), S.panel_row_ptr.size(), st));
    CKR(D.panel_chunk_ptr.upload(S.panel_chunk_ptr.data(), S.panel_chunk_ptr.size(), st));
    CKR(D.chunk_halo_ptr.upload(S.chunk_halo_ptr.data(), S.chunk_halo_ptr.size(), st));
    CKR(D.halo_cols.upload(S.halo_cols.data(), S.halo_cols.size(), st));
    if (D.mma) {
        CKR(D.chunk_ks_ptr.upload(S.chunk_ks_ptr.data(), S.chunk_ks_ptr.size(), st));
        CKR(D.chunk_meta_ptr.upload(S.chunk_meta_ptr.data(), S.chunk_meta_ptr.size(), st));
        CKR(D.a_src.upload(S.a_src.data(), S.a_src.size(), st));
        CKR(D.meta.upload(S.meta.data(), S.meta.size(), st));
        std::vector<int> cd(8 * (size_t)S.n_chunks);
        for (int c = 0; c < S.n_chunks; c++) {
            int *d = cd.data() + 8 * (size_t)c;
            d[0] = S.chunk_halo_ptr[c]; d[1] = S.chunk_halo_ptr[c + 1] - S.chunk_halo_ptr[c];
            d[2] = S.chunk_ks_ptr[c]; d[3] = S.chunk_ks_ptr[c + 1] - S.chunk_ks_ptr[c];
            d[4] = S.chunk_meta_ptr[c]; d[5] = S.chunk_meta_ptr[c + 1] - S.chunk_meta_ptr[c];
            d[6] = S.chunk_run_ptr[c]; d[7] = S.chunk_run_ptr[c + 1] - S.chunk_run_ptr[c];
        }
        CKR(D.cdesc.upload(cd.data(), cd.size(), st));
    } else {
        CKR(D.chunk_ent_ptr.upload(S.chunk_ent_ptr.data(), S.chunk_ent_ptr.size(), st));
        CKR(D.crp.upload(S.crp.data(), S.crp.size(), st));
        CKR(D.ent_src.upload(S.ent_src.data(), S.ent_src.size(), st));
        CKR(D.ent_idx.upload(S.ent_idx.data(), S.ent_idx.size(), st));
    }
    CKR(D.chunk_run_ptr.upload(S.chunk_run_ptr.data(), S.chunk_run_ptr.size(), st));
    std::vector<int> runs(3 * S.run_start.size());
    for (size_t i = 0; i < S.run_start.size(); i++) { runs[3 * i] = S.run_start[i]; runs[3 * i + 1] = S.run_col[i]; runs[3 * i + 2] = S.run_len[i]; }
    CKR(D.runs.upload(runs.data(), runs.size(), st));
    if (S.max_chunk_halo > 128) return 0;         // the producer warp handles at most 4 x 32 copies per stage
    CK(cudaStreamSynchronize(st));               // the host vectors go out of scope
    D.n_panels = S.n_panels; D.n_chunks = S.n_chunks; D.crp_stride = S.crp_stride; D.max_chunk_halo = S.max_chunk_halo;
    D.max_chunk_ent = S.max_chunk_ent; D.nnz = (size_t)S.nnz;
    if (D.mma) {
        D.max_chunk_ks = S.max_chunk_ks; D.max_chunk_meta = S.max_chunk_meta; D.n_frag = (size_t)S.n_ks * 32;
        CKR(D.av_a.alloc(std::max<size_t>(1, D.n_frag * h->nK))); CKR(D.av_dw.alloc(std::max<size_t>(1, D.n_frag * h->nK)));
    } else {
        CKR(D.ent_a.alloc(D.nnz * h->nK)); CKR(D.ent_dw.alloc(D.nnz * h->nK));
    }
    D.ok = true;
    return 0;
}
// which: 0 = A (SpMM, post-smoothing), 1 = A * diag(dw) (pre-smoothing residual)
int stream_pack(pgb200_ert *h, StreamDev &D, const double *vals, int which) {
    if (!D.ok || D.nnz == 0) return 0;
    if (D.g_ok && D.g_frag) {
        k_pack_mma<<<cdiv((long long)D.g_frag, 256), 256, 0, h->st>>>(D.g_a_src.p, D.g_frag, D.nnz, h->nK, vals, which ? D.g_av_dw.p : D.g_av_a.p); LAUNCH(h);
    }
    if (D.mma) {
        if (D.n_frag == 0) return 0;
        k_pack_mma<<<cdiv((long long)D.n_frag, 256), 256, 0, h->st>>>(D.a_src.p, D.n_frag, D.nnz, h->nK, vals, which ? D.av_dw.p : D.av_a.p); LAUNCH(h);
        return 0;
    }
    k_pack_entries<<<cdiv((long long)D.nnz, 256), 256, 0, h->st>>>(D.ent_src.p, D.ent_idx.p, D.nnz, h->nK, vals, which ? D.ent_dw.p : D.ent_a.p); LAUNCH(h);
    return 0;
}
inline size_t up128(size_t x) { return (x + 127) / 128 * 128; }

template <int NCP, int EPI, bool DOT>
int stream_go(pgb200_ert *h, const StreamArgs &A, size_t smem) {
    k_spmm_stream<NCP, EPI, DOT><<<h->num_sms, ST_THREADS, smem, h->st>>>(A);
    h->cur_role = EPI + 1; LAUNCH(h); h->cur_role = 0;
    return 0;
}
template <int EPI>
int stream_configure_epi(size_t smem) {
    CK(cudaFuncSetAttribute(k_spmm_stream<1, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_spmm_stream<2, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_spmm_stream<1, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_spmm_stream<2, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}
// per handle (the attribute is per device/context): dynamic shared memory of the streamed kernels
int stream_configure(pgb200_ert *h) {
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_spmm_stream<2, EPI_POST, true>));
    h->smem_optin = (size_t)optin - fa.sharedSizeBytes - 256;          // static: mbarriers, dot scratch
    CKR(stream_configure_epi<EPI_SPMM>(h->smem_optin));
    CKR(stream_configure_epi<EPI_POST>(h->smem_optin));
    CKR(stream_configure_epi<EPI_RESIDUAL>(h->smem_optin));
    return 0;
}

template <int NT, int EPI, bool DOT>
int mma_go(pgb200_ert *h, const MmaArgs &A, size_t smem) {
    k_spmm_mma<NT, EPI, DOT><<<h->num_sms, MM_THREADS, smem, h->st>>>(A);
    h->cur_role = EPI + 1; LAUNCH(h); h->cur_role = 0;
    return 0;
}
template <int NT, int EPI>
int mma_configure_nt(size_t smem) {
    CK(cudaFuncSetAttribute(k_spmm_mma<NT, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_spmm_mma<NT, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}
template <int EPI>
int mma_configure_epi(size_t smem) {
    CKR((mma_configure_nt<2, EPI>(smem))); CKR((mma_configure_nt<4, EPI>(smem))); CKR((mma_configure_nt<7, EPI>(smem)));
    CKR((mma_configure_nt<10, EPI>(smem))); CKR((mma_configure_nt<13, EPI>(smem))); CKR((mma_configure_nt<16, EPI>(smem)));
    return 0;
}
int mma_configure(pgb200_ert *h) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_spmm_mma<16, EPI_POST, true>));
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    h->smem_optin = std::min(h->smem_optin, (size_t)optin - fa.sharedSizeBytes - 256);
    CKR(mma_configure_epi<EPI_SPMM>(h->smem_optin));
    CKR(mma_configure_epi<EPI_POST>(h->smem_optin));
    CKR(mma_configure_epi<EPI_RESIDUAL>(h->smem_optin));
    return 0;
}

// tile geometry shared by the two streamed kernels: tiles per wavenumber, tile width
struct TileGeo { int k_lo, n_k, tpk, pw, n_tiles; };
inline TileGeo tile_geo(int nE, int c0, int c1, int wmax) {
    TileGeo g;
    g.k_lo = c0 / nE;
    const int k_hi = (c1 - 1) / nE;
    // widest column span of one wavenumber group inside the window, from an even start
    int span;
    if (k_hi == g.k_lo) span = ((c1 + 1) & ~1) - (c0 & ~1);
    else span = nE + ((nE & 1) ? 1 : 0);
    g.n_k = k_hi - g.k_lo + 1;
    g.tpk = cdiv(span, wmax);
    g.pw = (cdiv(span, g.tpk) + 1) & ~1;
    g.n_tiles = g.n_k * g.tpk;
    return g;
}

template <int EPI>
int launch_mma(pgb200_ert *h, const StreamDev &D, int which, const double *X, double *Y, int c0, int c1, double *dots, const PanelExtra &ex) {
    MmaArgs A;
    A.L = D.mlevel(); A.aval = which ? D.av_dw.p : D.av_a.p; A.n_frag = D.n_frag; A.X = X; A.Y = Y; A.ld = h->ld; A.nE = h->nE; A.c0 = c0; A.c1 = c1; A.ex = ex;
    // slot = [X rows | A fragments | meta]; the B fragments of the last n-tile may read up to 56 bytes past a row: the X
    // region is followed by the A region of the same slot, so those reads stay inside the slot
    const size_t a_b = up128((size_t)std::max(1, D.max_chunk_ks) * 256), m_b = up128((size_t)std::max(4, D.max_chunk_meta) * 4);
    const size_t per_col = (size_t)D.max_chunk_halo * 8;
    // chunk 1 stays resident through chunk 0 when the panel's own rows span both: at least 3 slots then
    const int min_slots = D.hc_used < MM_ROWS ? 3 : 2;
    long long wfit = ((long long)(h->smem_optin / min_slots) - (long long)a_b - (long long)m_b - 128) / (long long)per_col;
    int wmax = (int)std::min<long long>(ST_MAX_TILE_W, wfit) & ~1;
    if (wmax < 2) PGB_FAIL("streamed SpMM: a halo chunk does not fit shared memory");
    const TileGeo g = tile_geo(h->nE, c0, c1, wmax);
    A.k_lo = g.k_lo; A.tpk = g.tpk; A.pw = g.pw; A.n_tiles = g.n_tiles;
    A.x_bytes = (uint32_t)up128((size_t)D.max_chunk_halo * A.pw * 8);
    A.a_bytes = (uint32_t)a_b;
    A.slot_bytes = (uint32_t)(A.x_bytes + a_b + m_b);
    A.slots = (int)std::min<size_t>(ST_MAX_SLOTS, h->smem_optin / A.slot_bytes);
    if (A.slots < min_slots) PGB_FAIL("streamed SpMM: internal slot sizing error");
    const size_t smem = (size_t)A.slots * A.slot_bytes;
    const int G = h->num_sms;
    A.cpt = A.n_tiles <= G ? std::max(1, G / A.n_tiles) : 1;
    A.fullrows = (A.n_tiles == 1 && (c0 & ~1) == 0 && A.pw == (int)h->ld) ? 1 : 0;
    // partial-width tiles (2.5-D wavenumber groups, wide multi-GPU shards): one bulk copy per halo row dealt over the 128 producer
    // lanes (measured 10 % faster than the cp.async row copies, which PGB200_PARTIAL_BULK=0 restores)
    { static const int pb = getenv("PGB200_PARTIAL_BULK") ? atoi(getenv("PGB200_PARTIAL_BULK")) : 1; if (!A.fullrows && pb) A.fullrows = 2; }
    A.dot_part = h->dot_part.p; A.dot_counter = h->dot_counter.p; A.dots = dots;
    A.dbg = h->mma_dbg; A.dbg_buf = h->mma_dbg_buf.p;
    if (dots && A.n_tiles > (int)h->dot_counter.n) PGB_FAIL("streamed SpMM: too many column tiles for the dot tickets");
    const int nt = cdiv(A.pw, 8);
    h->pi_panel_nc = 100 + nt; h->pi_tiles = std::max(h->pi_tiles, A.n_tiles); h->pi_slots = A.slots;
#define MMA_GO(NT) { if (dots) return mma_go<NT, EPI, true>(h, A, smem); return mma_go<NT, EPI, false>(h, A, smem); }
    if (nt <= 2) MMA_GO(2)
    if (nt <= 4) MMA_GO(4)
    if (nt <= 7) MMA_GO(7)
    if (nt <= 10) MMA_GO(10)
    if (nt <= 13) MMA_GO(13)
    MMA_GO(16)
#undef MMA_GO
}

// ---- narrow column windows: gather form (k_spmm_gather) ---------------------------------------------------------------
constexpr int GATHER_MAX_W = 32;
int gather_upload(pgb200_ert *h, StreamDev &D, int n, const int *rowptr_dev, const int *colidx_dev, size_t nnz) {
    D.g_ok = false;
    if (!D.ok || n <= 0) return 0;
    std::vector<int> rp((size_t)n + 1), ci(nnz);
    CK(cudaMemcpyAsync(rp.data(), rowptr_dev, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(ci.data(), colidx_dev, sizeof(int) * nnz, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    GatherGroupsHost G;
    build_gather_groups(n, rp.data(), ci.data(), G);
    CKR(D.g_ks_ptr.upload(G.ks_ptr.data(), G.ks_ptr.size(), h->st));
    CKR(D.g_cols.upload(G.cols.data(), G.cols.size(), h->st));
    CKR(D.g_a_src.upload(G.a_src.data(), G.a_src.size(), h->st));
    CK(cudaStreamSynchronize(h->st));
    D.g_groups = G.n_groups; D.g_rows = n; D.g_frag = (size_t)G.n_ks * 32;
    CKR(D.g_av_a.alloc(std::max<size_t>(1, D.g_frag * h->nK))); CKR(D.g_av_dw.alloc(std::max<size_t>(1, D.g_frag * h->nK)));
    D.g_ok = true;
    return 0;
}
template <int NT, int EPI, bool DOT>
int gather_go(pgb200_ert *h, const GatherArgs &A, int grid) {
    k_spmm_gather<NT, EPI, DOT><<<grid, GA_THREADS, 0, h->st>>>(A);
    h->cur_role = EPI + 1; LAUNCH(h); h->cur_role = 0;
    return 0;
}
template <int EPI>
int launch_gather(pgb200_ert *h, const StreamDev &D, int which, const double *X, double *Y, int c0, int c1, double *dots, const PanelExtra &ex) {
    GatherArgs A;
    A.ks_ptr = D.g_ks_ptr.p; A.cols = D.g_cols.p; A.aval = which ? D.g_av_dw.p : D.g_av_a.p; A.n_frag = D.g_frag;
    A.n_groups = D.g_groups; A.n_rows = D.g_rows; A.X = X; A.Y = Y; A.ld = h->ld;
    A.kk = c0 / h->nE; A.cs = c0 & ~1; A.wc = ((c1 + 1) & ~1) - A.cs; A.v0 = c0; A.v1 = c1;
    A.dot_part = h->dot_part.p; A.dot_counter = h->dot_counter.p; A.dots = dots; A.ex = ex; A.dbg = h->mma_dbg;
    static const int force_per_sm = getenv("PGB200_GATHER_CTAS") ? atoi(getenv("PGB200_GATHER_CTAS")) : 0;
    const int per_sm = force_per_sm ? force_per_sm : (A.wc <= 16 ? 4 : 3);
    const int grid = std::max(1, std::min(std::min(per_sm * h->num_sms, 2 * h->dot_slots), cdiv(D.g_groups, GA_THREADS / 32)));
    h->pi_panel_nc = 200 + cdiv(A.wc, 8); h->pi_tiles = std::max(h->pi_tiles, 1);
    if (A.wc <= 16) { if (dots) return gather_go<2, EPI, true>(h, A, grid); return gather_go<2, EPI, false>(h, A, grid); }
    if (dots) return gather_go<4, EPI, true>(h, A, grid);
    return gather_go<4, EPI, false>(h, A, grid);
}

// Y = op(A X) on the column window [c0, c1); dots != nullptr: deterministic per-column dot of the epilogue
// which: 0 = the level's matrix A, 1 = A * diag(dw)
template <int EPI>
int launch_stream(pgb200_ert *h, const StreamDev &D, int which, const double *X, double *Y, int c0, int c1, double *dots,
                  const PanelExtra &ex) {
    if (c1 <= c0) return 0;
    if (D.g_ok && c0 / h->nE == (c1 - 1) / h->nE && ((c1 + 1) & ~1) - (c0 & ~1) <= GATHER_MAX_W)
        return launch_gather<EPI>(h, D, which, X, Y, c0, c1, dots, ex);
    if (D.mma) return launch_mma<EPI>(h, D, which, X, Y, c0, c1, dots, ex);
    const PanelEntry *ent = which ? D.ent_dw.p : D.ent_a.p;
    const int nE = h->nE;
    StreamArgs A;
    A.L = D.level(); A.ent = ent; A.nnz = D.nnz; A.X = X; A.Y = Y; A.ld = h->ld; A.nE = nE; A.c0 = c0; A.c1 = c1; A.ex = ex;
    A.k_lo = c0 / nE;
    const int k_hi = (c1 - 1) / nE;
    // widest column span of one wavenumber group inside the window, from an even start
    int span;
    if (k_hi == A.k_lo) span = ((c1 + 1) & ~1) - (c0 & ~1);
    else span = nE + ((nE & 1) ? 1 : 0);
    // tile width: at least two slots must fit
    const size_t ent_b = up128((size_t)std::max(1, D.max_chunk_ent) * sizeof(PanelEntry)), crp_b = up128((size_t)D.crp_stride * 4);
    const size_t per_col = (size_t)D.max_chunk_halo * 8;
    long long wfit = ((long long)(h->smem_optin / 2) - (long long)ent_b - (long long)crp_b - 128) / (long long)per_col;
    int wmax = (int)std::min<long long>(ST_MAX_TILE_W, wfit) & ~1;
    if (wmax < 2) PGB_FAIL("streamed SpMM: a halo chunk does not fit shared memory");
    A.tpk = cdiv(span, wmax);
    A.pw = (cdiv(span, A.tpk) + 1) & ~1;
    A.n_tiles = (k_hi - A.k_lo + 1) * A.tpk;
    A.x_bytes = (uint32_t)up128((size_t)D.max_chunk_halo * A.pw * 8);
    A.ent_bytes = (uint32_t)ent_b;
    A.slot_bytes = (uint32_t)(A.x_bytes + ent_b + crp_b);
    A.slots = (int)std::min<size_t>(ST_MAX_SLOTS, h->smem_optin / A.slot_bytes);
    if (A.slots < 2) PGB_FAIL("streamed SpMM: internal slot sizing error");
    const size_t smem = (size_t)A.slots * A.slot_bytes;
    const int G = h->num_sms;
    A.cpt = A.n_tiles <= G ? std::max(1, G / A.n_tiles) : 1;
    A.fullrows = (A.n_tiles == 1 && (c0 & ~1) == 0 && A.pw == (int)h->ld) ? 1 : 0;
    A.dot_part = h->dot_part.p; A.dot_counter = h->dot_counter.p; A.dots = dots;
    if (dots && A.n_tiles > (int)h->dot_counter.n) PGB_FAIL("streamed SpMM: too many column tiles for the dot tickets");
    h->pi_panel_nc = A.pw > 64 ? 2 : 1; h->pi_tiles = std::max(h->pi_tiles, A.n_tiles); h->pi_slots = A.slots;
    if (A.pw > 64) { if (dots) return stream_go<2, EPI, true>(h, A, smem); return stream_go<2, EPI, false>(h, A, smem); }
    if (dots) return stream_go<1, EPI, true>(h, A, smem);
    return stream_go<1, EPI, false>(h, A, smem);
}
bool panel_path_ok(const pgb200_ert *h) { return h->use_panels && h->stream.ok; }

// launch geometry of the flat element-wise kernels (ert_kernels.cuh, flat_map): column chunks of at most FLAT_T columns,
// rows per CTA = rows per pass x passes
constexpr int FLAT_MAX_GX = 296;           // 2 CTAs of 512 threads per SM: bounds the partial rows of the deterministic dots
struct FlatCfg { int cw; dim3 grid; };
inline FlatCfg flat_cfg(int n_rows, int c0, int c1, int gx_cap = 8 * FLAT_MAX_GX) {
    FlatCfg f;
    const int w = std::max(1, c1 - c0);
    const int nchunk = cdiv(w, FLAT_T);
    f.cw = cdiv(w, nchunk);
    const int rpp = FLAT_T / f.cw;
    // up to 12 passes per CTA, fewer when that would leave less than ~4 CTAs per SM (small levels, narrow shards); the
    // rows are split evenly over the CTAs (contiguous ranges)
    const int passes = std::max(1, std::min(12, n_rows / (rpp * 600)));
    f.grid = dim3(std::max(1, std::min(cdiv(n_rows, rpp * passes), gx_cap)), nchunk);
    return f;
}
inline DotOut dot_out(pgb200_ert *h, double *out0, double *out1) {
    DotOut D; D.part = h->dot_part.p; D.plane = (size_t)h->dot_slots * h->ld; D.counter = h->dot_counter.p; D.out0 = out0; D.out1 = out1; D.ld = h->ld;
    return D;
}

// rows per CTA of the plain multilevel kernels: 32 on big levels, 8 on small ones so that they still fill the GPU
inline int amg_rows_per_cta(int n) { return n >= 60000 ? 32 : 8; }

constexpr int STREAM_MIN_ROWS = 6000;     // coarse levels with at least this many rows use the streamed row-panel kernel
constexpr int AMG_SPLIT_BELOW = 6000;      // levels smaller than this use the one-row-per-CTA kernels (latency-bound; measured: 11 k rows is already better off with the row-per-thread kernels)

template <int CPT>
int amg_post_cpt(pgb200_ert *h, const int *rowptr, const int *colidx, const double *vals, size_t nnz, const double *dinvw, int n,
                 const double *X, const double *R, double *Z, int c0, int c1, double *dots) {
    static_assert(CPT <= AMG_TY, "split kernels finish column group m in thread row m");
    if (!dots && n < AMG_SPLIT_BELOW) {
        dim3 block(AMG_TX, AMG_TY), grid(n, cdiv(c1 - c0, AMG_TX * CPT));
        k_amg_post_split<CPT><<<grid, block, 0, h->st>>>(rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, h->nE, c0, c1, h->ld);
        LAUNCH(h);
        return 0;
    }
    const int rows = amg_rows_per_cta(n);
    dim3 block(AMG_TX, AMG_TY), grid(cdiv(n, rows), cdiv(c1 - c0, AMG_TX * CPT));
    if (dots) k_amg_post<CPT, true><<<grid, block, 0, h->st>>>(rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, h->nE, c0, c1, h->ld, dots, rows);
    else k_amg_post<CPT, false><<<grid, block, 0, h->st>>>(rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, h->nE, c0, c1, h->ld, nullptr, rows);
    LAUNCH(h);
    return 0;
}
int pick_cpt(int ncols) {
    // wide column blocks amortise the index/value loads over 4 accumulators; only narrow shards use fewer
    if (ncols >= 48) return 4;
    if (ncols >= 24) return 2;
    return 1;
}
int amg_post(pgb200_ert *h, const int *rowptr, const int *colidx, const double *vals, size_t nnz, const double *dinvw, int n,
             const double *X, const double *R, double *Z, int c0, int c1, double *dots) {
    // (plain-kernel path only: the column dots are accumulated with atomics into a zeroed buffer)
    if (dots) CK(cudaMemsetAsync(dots + c0, 0, sizeof(double) * (c1 - c0), h->st));
    switch (pick_cpt(c1 - c0)) {
        case 4: return amg_post_cpt<4>(h, rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, c0, c1, dots);
        case 2: return amg_post_cpt<2>(h, rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, c0, c1, dots);
        default: return amg_post_cpt<1>(h, rowptr, colidx, vals, nnz, dinvw, n, X, R, Z, c0, c1, dots);
    }
}
int amg_restrict(pgb200_ert *h, const int *rowptr, const int *colidx, const double *vals_dw, size_t nnz, int n_f,
                 const AmgLevel *L, const double *R, int c0, int c1) {
    const int cpt = pick_cpt(c1 - c0);
    if (L->n < AMG_SPLIT_BELOW) {
        dim3 block(AMG_TX, AMG_TY), grid(L->n, cdiv(c1 - c0, AMG_TX * cpt));
#define RSGO(C) k_amg_restrict_split<C><<<grid, block, 0, h->st>>>(rowptr, colidx, vals_dw, nnz, L->mem_ptr.p, L->mem_idx.p, L->n, R, L->R.p, h->nE, c0, c1, h->ld)
        if (cpt == 4) RSGO(4); else if (cpt == 2) RSGO(2); else RSGO(1);
#undef RSGO
        LAUNCH(h);
        return 0;
    }
    const int rows = amg_rows_per_cta(L->n);
    dim3 block(AMG_TX, AMG_TY), grid(cdiv(L->n, rows), cdiv(c1 - c0, AMG_TX * cpt));
#define RGO(C) k_amg_restrict<C><<<grid, block, 0, h->st>>>(rowptr, colidx, vals_dw, nnz, n_f, L->mem_ptr.p, L->mem_idx.p, L->n, R, L->R.p, h->nE, c0, c1, h->ld, rows)
    if (cpt == 4) RGO(4); else if (cpt == 2) RGO(2); else RGO(1);
#undef RGO
    LAUNCH(h);
    return 0;
}

// recompute the coarse matrices and smoother weights for the current vals (once per assembled model)
int amg_setup_values(pgb200_ert *h) {
    if (h->amg.empty()) return 0;
    const int nK = h->nK;
    auto smoother = [&](const int *rowptr, const int *colidx, const int *diag_pos, int n, size_t nnz, const double *vals, double *dinvw,
                        double *vals_dw) -> int {
        CK(cudaMemsetAsync(h->gmax.p, 0, sizeof(unsigned long long) * nK, h->st));
        k_row_ratio<<<cdiv(n, 128), 128, 0, h->st>>>(rowptr, diag_pos, n, nK, nnz, vals, h->gmax.p); LAUNCH(h);
        k_inv_diag_w<<<cdiv(n, 128), 128, 0, h->st>>>(diag_pos, n, nK, nnz, vals, h->gmax.p, dinvw); LAUNCH(h);
        if (vals_dw) { k_scale_cols<<<cdiv((long long)nnz, 256), 256, 0, h->st>>>(colidx, nnz, n, nK, vals, dinvw, vals_dw); LAUNCH(h); }
        return 0;
    };
    CKR(smoother(h->rowptr.p, h->colidx.p, h->diag_pos.p, h->N, h->nnz, h->vals.p, h->dinvw0.p, h->vals_dw0.p));
    CKR(stream_pack(h, h->stream, h->vals_dw0.p, 1));
    const double *vf = h->vals.p; size_t nnz_f = h->nnz;
    for (AmgLevel *L : h->amg) {
        k_galerkin<<<cdiv((long long)L->nnz, 128), 128, 0, h->st>>>(L->gal_ptr.p, L->gal_idx.p, (int)L->nnz, nK, nnz_f, L->nnz, vf, L->vals.p); LAUNCH(h);
        CKR(smoother(L->rowptr.p, L->colidx.p, L->diag_pos.p, L->n, L->nnz, L->vals.p, L->dinvw.p, L->vals_dw.p));
        CKR(stream_pack(h, L->stream, L->vals.p, 0));
        CKR(stream_pack(h, L->stream, L->vals_dw.p, 1));
        vf = L->vals.p; nnz_f = L->nnz;
    }
    CK(cudaGetLastError());
    return 0;
}

// fused sub-cycle over the levels lv[g0..nl] (ert_kernels.cuh k_amg_subcycle); returns false if it does not apply
struct LvRef { const int *rowptr, *colidx; const double *vals, *vals_dw; size_t nnz; const double *dinvw; int n; const double *R; double *X, *Z; const StreamDev *st; };
int amg_subcycle(pgb200_ert *h, const std::vector<LvRef> &lv, int g0, int c0, int c1, bool &done) {
    done = false;
    const int nl = (int)lv.size() - 1, nsub = nl - g0 + 1;
    if (!h->use_subcycle || g0 < 1 || nsub < 1 || nsub > SUB_MAX_LEVELS || c1 <= c0) return 0;
    SubArgs A{};
    A.nl = nsub; A.sweeps = h->coarse_sweeps; A.ld = h->ld; A.nE = h->nE; A.c0 = c0; A.c1 = c1;
    int off = 0;
    for (int s2 = 0; s2 < nsub; s2++) {
        const LvRef &L = lv[g0 + s2];
        SubLevel &S = A.lv[s2];
        S.rowptr = L.rowptr; S.colidx = L.colidx; S.vals = L.vals; S.vals_dw = L.vals_dw; S.dinvw = L.dinvw; S.n = L.n; S.nnz = L.nnz; S.off = off;
        if (s2 + 1 < nsub) { AmgLevel *T = h->amg[g0 + s2]; S.mem_ptr = T->mem_ptr.p; S.mem_idx = T->mem_idx.p; S.agg = T->agg.p; }
        off += L.n;
    }
    A.rows_total = off;
    A.Rin = lv[g0].R; A.Zout = lv[g0].Z;
    const int ncols = c1 - c0;
    int cpc = ncols <= 160 ? 1 : (ncols <= 320 ? 2 : 4);
    while (cpc > 1 && 3 * (size_t)cpc * off * 8 > h->smem_optin) cpc >>= 1;
    size_t smem = 3 * (size_t)cpc * off * 8;
    if (smem > h->smem_optin) return 0;
    const LvRef &last = lv[nl];
    const size_t cache = last.nnz * 12 + ((size_t)last.n + 1) * 4 + 16;
    A.cache_last = (smem + cache <= h->smem_optin) ? 1 : 0;
    if (A.cache_last) smem += cache;
    const int spk = cdiv(h->nE, cpc), k_lo = c0 / h->nE, k_hi = (c1 - 1) / h->nE;
    (void)k_lo;
    const int grid = (k_hi + 1) * spk;                 // CTAs of wavenumber groups below the window exit at once
#define SUBGO(C) do { if (smem > h->sub_smem[C]) { CK(cudaFuncSetAttribute(k_amg_subcycle<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); h->sub_smem[C] = smem; } \
                      k_amg_subcycle<C><<<grid, SUB_THREADS, smem, h->st>>>(A); } while (0)
    if (cpc == 1) SUBGO(1); else if (cpc == 2) SUBGO(2); else SUBGO(4);
#undef SUBGO
    LAUNCH(h);
    done = true;
    return 0;
}

// one V(1,1) cycle: Z0 = M^-1 R (level 0 residual = h->R); dots != nullptr: dots[c] = R.Z per column (deterministic on the
// streamed path)
int amg_vcycle(pgb200_ert *h, int c0, int c1, double *dots) {
    const int nl = (int)h->amg.size();
    typedef LvRef Lv;
    std::vector<Lv> lv(nl + 1);
    lv[0] = {h->rowptr.p, h->colidx.p, h->vals.p, h->vals_dw0.p, h->nnz, h->dinvw0.p, h->N, h->R.p, h->X0.p, h->Z0.p, &h->stream};
    for (int l = 0; l < nl; l++) { AmgLevel *L = h->amg[l]; lv[l + 1] = {L->rowptr.p, L->colidx.p, L->vals.p, L->vals_dw.p, L->nnz, L->dinvw.p, L->n, L->R.p, L->X.p, L->Z.p, &L->stream}; }
    auto streamed = [&](int l) { return h->use_panels && lv[l].st->ok; };
    // the levels below STREAM_MIN_ROWS run as ONE fused launch (g0 = first of them); everything above goes level by level
    // (measured on the 1M-tet case: fusing all levels below 6000 rows into one launch is SLOWER than one launch per
    //  operator application -- 214 us against 140 us: with one CTA per column slice the dependent index -> value -> x
    //  chains of the few-thousand-row levels are L2-latency-bound, while the per-level kernels spread every level over
    //  all SMs.  The coarsest level alone, with its matrix staged in shared memory, is where fusion pays: 7 launches -> 1)
    int g0 = nl + 1;
    if (nl > 0 && h->use_panels) {
        if (h->use_subcycle == 2) { g0 = 1; while (g0 <= nl && lv[g0].st->ok) g0++; }
        else if (h->use_subcycle == 1) g0 = std::max(1, nl);
    }
    const int down_to = std::min(nl, g0);
    // downward: residual after one damped-Jacobi sweep from zero, restricted
    for (int l = 0; l < down_to; l++) {
        h->cur_tag = l;
        if (!streamed(l)) { CKR(amg_restrict(h, lv[l].rowptr, lv[l].colidx, lv[l].vals_dw, lv[l].nnz, lv[l].n, h->amg[l], lv[l].R, c0, c1)); continue; }
        if (streamed(l)) {
            // residual through the streamed kernel (into X of this level, free until the prolongation), then a
            // deterministic member sum
            PanelExtra ex{};
            CKR(launch_stream<EPI_RESIDUAL>(h, *lv[l].st, 1, lv[l].R, lv[l].X, c0, c1, nullptr, ex));
            AmgLevel *L = h->amg[l];
            const FlatCfg fc = flat_cfg(L->n, c0, c1);
            k_amg_sum_members<<<fc.grid, FLAT_T, 0, h->st>>>(L->mem_ptr.p, L->mem_idx.p, L->n, lv[l].X, L->R.p, c0, c1, h->ld, fc.cw); LAUNCH(h);
        } else {
            CKR(amg_restrict(h, lv[l].rowptr, lv[l].colidx, lv[l].vals_dw, lv[l].nnz, lv[l].n, h->amg[l], lv[l].R, c0, c1));
        }
    }
    // coarsest level: fixed number of Jacobi sweeps
    const double *E = nullptr;
    bool fused = false;
    if (g0 <= nl) { h->cur_tag = g0; CKR(amg_subcycle(h, lv, g0, c0, c1, fused)); }
    if (fused) E = lv[g0].Z;
    else {
        for (int l = down_to; l < nl; l++) {
            h->cur_tag = l;
            CKR(amg_restrict(h, lv[l].rowptr, lv[l].colidx, lv[l].vals_dw, lv[l].nnz, lv[l].n, h->amg[l], lv[l].R, c0, c1));
        }
        Lv &c = lv[nl];
        h->cur_tag = nl;
        const FlatCfg fc = flat_cfg(c.n, c0, c1);
        k_amg_prolong<<<fc.grid, FLAT_T, 0, h->st>>>(c.dinvw, c.n, nullptr, c.R, nullptr, c.X, h->nE, c0, c1, h->ld, fc.cw); LAUNCH(h);
        double *a = c.X, *b = c.Z;
        const int sweeps = (nl == 0) ? 1 : h->coarse_sweeps;
        for (int s = 0; s + 1 < sweeps; s++) {
            CKR(amg_post(h, c.rowptr, c.colidx, c.vals, c.nnz, c.dinvw, c.n, a, c.R, b, c0, c1, nullptr));
            std::swap(a, b);
        }
        if (nl == 0) { CKR(amg_post(h, c.rowptr, c.colidx, c.vals, c.nnz, c.dinvw, c.n, a, c.R, b, c0, c1, dots)); a = b; }
        E = a;
    }
    // upward: prolongate, post-smooth
    for (int l = (fused ? g0 : nl) - 1; l >= 0; l--) {
        Lv &f = lv[l];
        h->cur_tag = l;
        const FlatCfg fc = flat_cfg(f.n, c0, c1);
        k_amg_prolong<<<fc.grid, FLAT_T, 0, h->st>>>(f.dinvw, f.n, h->amg[l]->agg.p, f.R, E, f.X, h->nE, c0, c1, h->ld, fc.cw); LAUNCH(h);
        if (streamed(l)) {
            PanelExtra ex{}; ex.R = f.R; ex.dinvw = f.dinvw; ex.n = f.n;
            CKR(launch_stream<EPI_POST>(h, *f.st, 0, f.X, f.Z, c0, c1, l == 0 ? dots : nullptr, ex));
        } else {
            CKR(amg_post(h, f.rowptr, f.colidx, f.vals, f.nnz, f.dinvw, f.n, f.X, f.R, f.Z, c0, c1, l == 0 ? dots : nullptr));
        }
        E = f.Z;
    }
    h->cur_tag = 0;
    CK(cudaGetLastError());
    return 0;
}

// scal layout: [0] rz_a [1] rz_b [2] rz_c [3] pAp [4] rr_a [5] rr_b [6] bb   (each ld doubles)
int pcg_solve(pgb200_ert *h) {
    const int s0 = h->c0, s1 = h->c1, ncols = s1 - s0;      // this shard's source columns
    int c0 = s0, c1 = s1;                                    // ACTIVE window: shrinks as wavenumber groups converge
    h->last_iters = 0; h->last_relres = 0.0;
    h->pi_panel_nc = 0; h->pi_tiles = 0; h->pi_two_k = 0; h->pi_graph_launches = 0;
    if (ncols <= 0) return 0;
    h->col_relres.assign(h->ld, 0.0);
    const size_t ld = h->ld;
    double *S = h->scal.p;
    auto sc = [&](int i) { return S + (size_t)i * ld; };
    CK(cudaMemsetAsync(S, 0, sizeof(double) * 7 * ld, h->st));
    FlatCfg fc = flat_cfg(h->N, c0, c1), fd = flat_cfg(h->N, c0, c1, FLAT_MAX_GX);   // per-iteration vector kernels; fd: those with column dots
    auto regrid = [&]() { fc = flat_cfg(h->N, c0, c1); fd = flat_cfg(h->N, c0, c1, FLAT_MAX_GX); };
    const bool amg = h->use_amg && !h->amg.empty();
    // warm start (opt-in, pgb200_ert_set_warm_start): X still holds the secondary potentials of the previous model on this
    // shard; a Gauss-Newton step changes the model little, so r0 = b - A x is small compared with b
    const double *ax = nullptr;
    if (h->warm_start && h->x_warm_ok) {
        if (panel_path_ok(h)) { PanelExtra ex{}; CKR(launch_stream<EPI_SPMM>(h, h->stream, 0, h->X.p, h->AP.p, c0, c1, nullptr, ex)); }
        else CKR((launch_spmm<0, false>(h, h->vals.p, nullptr, nullptr, h->X.p, h->AP.p, c0, c1, nullptr)));
        ax = h->AP.p;
        h->warm_used++;
    }
    h->x_warm_ok = false;
    k_pcg_init<<<fd.grid, FLAT_T, 0, h->st>>>(h->B.p, h->dinv.p, h->X.p, h->R.p, h->P.p, ax, h->N, h->nE, c0, c1, ld, fd.cw,
                                              dot_out(h, amg ? nullptr : sc(0), sc(6))); LAUNCH(h);
    if (amg) {
        CKR(amg_vcycle(h, c0, c1, sc(0)));
        CK(cudaMemcpyAsync(h->P.p, h->Z0.p, sizeof(double) * (size_t)h->N * ld, cudaMemcpyDeviceToDevice, h->st));
    }
    CK(cudaGetLastError());
    CKR(ensure_pinned(h, 2 * ld));
    const double tol2 = h->tol * h->tol;
    int it = 0; bool converged = false;
    // one PCG iteration; the scalar buffers rotate with the iteration number (rz: period 3, rr: period 2)
    auto body = [&](int i, bool timed) -> int {
        const int rz_old = i % 3, rz_new = (i + 1) % 3;
        const int rr_cur = 4 + (i % 2);
        if (timed) CK(cudaEventRecord(h->pev[h->n_pev++], h->st));
        if (panel_path_ok(h)) {
            PanelExtra ex{};
            CKR(launch_stream<EPI_SPMM>(h, h->stream, 0, h->P.p, h->AP.p, c0, c1, sc(3), ex));
        } else {
            CK(cudaMemsetAsync(sc(3) + c0, 0, sizeof(double) * (c1 - c0), h->st));
            CKR((launch_spmm<0, true>(h, h->vals.p, nullptr, nullptr, h->P.p, h->AP.p, c0, c1, sc(3))));
        }
        if (timed) {
            CK(cudaEventRecord(h->pev[h->n_pev++], h->st));
            // algorithmic bytes of THIS launch: the CSR of the wavenumber groups inside the active window, X read
            // and Y written once for the window's columns
            const int nkw = (c1 - 1) / h->nE - c0 / h->nE + 1;
            h->spmm_bytes += 12.0 * (double)h->nnz * nkw + 4.0 * (h->N + 1) + 16.0 * (double)h->N * (c1 - c0);
        }
        if (amg) {
            k_pcg_update_xr<false><<<fd.grid, FLAT_T, 0, h->st>>>(h->P.p, h->AP.p, nullptr, h->X.p, h->R.p, h->N, h->nE, c0, c1, ld,
                                                                 sc(rz_old), sc(3), fd.cw, dot_out(h, nullptr, sc(rr_cur))); LAUNCH(h);
            CKR(amg_vcycle(h, c0, c1, sc(rz_new)));
            k_pcg_update_p<false><<<fc.grid, FLAT_T, 0, h->st>>>(h->Z0.p, nullptr, h->P.p, h->N, h->nE, c0, c1, ld, sc(rz_old), sc(rz_new),
                                                                sc(rr_cur), sc(6), tol2, fc.cw); LAUNCH(h);
        } else {
            k_pcg_update_xr<true><<<fd.grid, FLAT_T, 0, h->st>>>(h->P.p, h->AP.p, h->dinv.p, h->X.p, h->R.p, h->N, h->nE, c0, c1, ld,
                                                                sc(rz_old), sc(3), fd.cw, dot_out(h, sc(rz_new), sc(rr_cur))); LAUNCH(h);
            k_pcg_update_p<true><<<fc.grid, FLAT_T, 0, h->st>>>(h->R.p, h->dinv.p, h->P.p, h->N, h->nE, c0, c1, ld, sc(rz_old), sc(rz_new),
                                                               sc(rr_cur), sc(6), tol2, fc.cw); LAUNCH(h);
        }
        return 0;
    };
    // residual norms of iteration i_done - 1 -> host; true when every column meets the tolerance
    auto check = [&](int i_done) -> int {
        const int rr_cur = 4 + ((i_done - 1) % 2);
        CK(cudaMemcpyAsync(h->h_pinned, sc(rr_cur), sizeof(double) * ld, cudaMemcpyDeviceToHost, h->st));
        CK(cudaMemcpyAsync(h->h_pinned + ld, sc(6), sizeof(double) * ld, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        int lo = c1, hi = c0;                                  // unconverged columns of the active window
        for (int c = c0; c < c1; c++) {
            const double bb = h->h_pinned[ld + c], rr = h->h_pinned[c];
            double rel = 0.0;
            if (bb > 0.0) rel = std::sqrt(rr / bb);
            else if (rr > 0.0) rel = INFINITY;
            if (!(rr == rr)) rel = INFINITY;
            h->col_relres[c] = rel;
            if (!(rel <= h->tol)) { lo = std::min(lo, c); hi = std::max(hi, c + 1); }
        }
        double worst = 0.0;
        for (int c = s0; c < s1; c++) worst = std::max(worst, h->col_relres[c]);
        h->last_relres = worst;
        converged = lo >= hi;
        if (!converged && h->shrink_window) {
            // columns are ordered by wavenumber, and whole wavenumber groups converge together (large k first):
            // drop the converged ends of the window.  Frozen columns keep their final X; all kernels take [c0,c1).
            lo &= ~1;                                            // even start: 16-byte aligned column tiles
            if (lo > c0 || hi < c1) { c0 = lo; c1 = hi; regrid(); }
        }
        return 0;
    };
    // CUDA-graph mode: blocks of 6 iterations (one full period of the buffer rotation) replayed as one graph launch.
    // Removes the launch gaps of the ~27 kernels per multilevel iteration (matters most for narrow multi-GPU shards).
    const bool graph_mode = h->use_graph && amg && !h->prof && h->st != 0;
    if (graph_mode) {
        const int blocks_per_check = ax ? 1 : std::max(1, h->check_every / 6);      // warm starts can be done after a few iterations
        int blocks = 0;
        while (it < h->max_iter && !converged) {
            GraphKey key{c0, c1, h->tol, h->use_panels, (int)h->amg.size(), h->coarse_sweeps, (void *)h->st, (void *)h->vals.p};
            if (it == 0) {
                const long long before = h->launches;
                for (int j = 0; j < 6; j++) CKR(body(j, false));          // warm-up block (also sets kernel attributes)
                h->launches_per_block = h->launches - before;
            } else {
                if (!h->gexec || !(h->gkey == key)) {
                    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
                    const long long before = h->launches;
                    cudaGraph_t g = nullptr;
                    CK(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
                    int rc = 0;
                    for (int j = 0; j < 6 && !rc; j++) rc = body(j, false);
                    cudaError_t ce = cudaStreamEndCapture(h->st, &g);
                    h->launches = before;
                    if (rc || ce != cudaSuccess) { if (g) cudaGraphDestroy(g); if (!rc) g_err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return 1; }
                    ce = cudaGraphInstantiate(&h->gexec, g, 0);
                    cudaGraphDestroy(g);
                    if (ce != cudaSuccess) { h->gexec = nullptr; g_err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return 1; }
                    h->gkey = key;
                }
                CK(cudaGraphLaunch(h->gexec, h->st));
                h->pi_graph_launches++;
                h->launches += h->launches_per_block;
            }
            it += 6; blocks++;
            // close to the tolerance the check runs after every block, so that at most 5 iterations are wasted
            if (blocks % blocks_per_check == 0 || it >= h->max_iter || (h->last_relres > 0.0 && h->last_relres <= 50.0 * h->tol)) CKR(check(it));
        }
    } else {
        while (it < h->max_iter) {
            const bool timed = h->prof && h->n_pev + 2 <= (int)h->pev.size();
            CKR(body(it, timed));
            it++;
            if ((it % h->check_every == 0) || it >= h->max_iter) {
                CKR(check(it));
                if (converged) break;
            }
        }
    }
    CK(cudaGetLastError());
    h->last_iters = it; h->total_iters += it; h->solves++;
    h->x_warm_ok = converged;
    if (!converged) {
        char buf[256];
        snprintf(buf, sizeof buf, "block-PCG did not reach rel. residual %.1e in %d iterations (worst column %.3e)", h->tol, it, h->last_relres);
        PGB_FAIL(buf);
    }
    return 0;
}

void phase_begin(pgb200_ert *h, int ph) { cudaEventRecord(h->ev[ph], h->st); h->ph_rec[ph] = true; }

int map_model(pgb200_ert *h, const double *model_dev, int n_in) {
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(int) * 4, h->st));
    k_check_model<<<cdiv(n_in, 256), 256, 0, h->st>>>(model_dev, n_in, h->flags.p); LAUNCH(h);
    k_map_model<<<cdiv(h->C, 256), 256, 0, h->st>>>(model_dev, n_in, h->cell_marker.p, h->C, h->rho.p); LAUNCH(h);
    if (n_in != h->C) {
        const int nl = (int)h->pro_level_ptr.size() - 1;
        for (int l = 0; l < nl; l++) {
            const int b = h->pro_level_ptr[l], n = h->pro_level_ptr[l + 1] - b;
            if (n <= 0) continue;
            k_prolong_level<<<cdiv(n, 128), 128, 0, h->st>>>(h->pro_cells.p + b, h->pro_nb.p + (size_t)b * h->pro_nf,
                                                            h->pro_w.p + (size_t)b * h->pro_nf, n, h->pro_nf, h->rho.p); LAUNCH(h);
        }
    }
    CK(cudaGetLastError());
    return 0;
}

// assemble S(rho), right-hand sides, solve, total potentials  (calculateK, dcfemmodelling.cpp:2152-2296)
int forward_solve(pgb200_ert *h) {
    if (h->sr && !h->prim_set)
        PGB_FAIL("topography: singularity removal needs numeric primary potentials (pgb200_ert_set_primary_dev) before the first solve");
    phase_begin(h, PH_ASM);
    CKR(assemble(h, h->rho.p, h->vals.p));
    k_count_singular<<<cdiv(h->N, 256), 256, 0, h->st>>>(h->diag_pos.p, h->N, h->nK, h->nnz, h->vals.p, h->flags.p + 1); LAUNCH(h);
    k_inv_diag<<<cdiv(h->N, 256), 256, 0, h->st>>>(h->diag_pos.p, h->N, h->nK, h->nnz, h->vals.p, h->dinv.p); LAUNCH(h);
    h->have_vals = true;
    CKR(stream_pack(h, h->stream, h->vals.p, 0));
    if (h->use_amg) CKR(amg_setup_values(h));
    phase_begin(h, PH_RHS);
    const int c0 = h->c0, c1 = h->c1;
    if (h->sr) {
        k_rho_src<<<cdiv(h->nE, 64), 64, 0, h->st>>>(h->src_cell_ptr.p, h->src_cells.p, h->rho.p, h->nE, h->rho_src.p); LAUNCH(h);
        CKR((launch_spmm<1, false>(h, h->vals.p, h->vals1.p, h->rho_src.p, h->prim.p, h->B.p, c0, c1, nullptr)));
    } else {
        CK(cudaMemsetAsync(h->B.p, 0, sizeof(double) * h->N * h->ld, h->st));
        if (c1 > c0) { k_delta_rhs<<<cdiv(c1 - c0, 128), 128, 0, h->st>>>(h->pick_ptr.p, h->pick_idx.p, h->pick_w.p, h->nE, c0, c1, h->ld, h->B.p); LAUNCH(h); }
        if (c1 > c0 && (h->ref_node >= 0 || h->ref_last)) {
            k_ref_rhs<<<cdiv(c1 - c0, 128), 128, 0, h->st>>>(h->pick_ptr.p, h->pick_idx.p, h->pick_w.p, h->nE, c0, c1, h->ld, h->ref_node, h->ref_last, h->B.p); LAUNCH(h);
        }
    }
    // NB: the reference zeroes right-hand-side rows only for calibration nodes (:2281-2283), which never exist on
    // the non-Neumann domains handled here; rows of -3 Dirichlet faces keep S1*p/rho_s - S*p = p (1 - rho_s).
    CK(cudaGetLastError());
    // model / matrix sanity before iterating
    int hf[4];
    CK(cudaMemcpyAsync(hf, h->flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (hf[0]) PGB_FAIL("response for model with negative or zero resistivity is not defined");
    if (hf[1]) PGB_FAIL("stiffness matrix has rows with diagonal < 1e-12 (the reference would force them to homogeneous Dirichlet); unsupported model");
    phase_begin(h, PH_SOLVE);
    CKR(pcg_solve(h));
    phase_begin(h, PH_EPI);
    if (c1 > c0) {
        dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(c1 - c0, 32));
        k_finalize_pots<<<g, b, 0, h->st>>>(h->X.p, h->sr ? h->prim.p : nullptr, h->rho_src.p, 0.0, h->N, h->nE, c0, c1, h->ld, h->U.p); LAUNCH(h);
    }
    CK(cudaGetLastError());
    // a source shard fills only its own columns: the Jacobian may use U only after the caller's all-gather
    // (pgb200_ert_mark_potentials_valid)
    h->shard_solved = true;
    h->pots_valid = (c0 == 0 && c1 == h->nS);
    return 0;
}

int analytic_pots(pgb200_ert *h, double scale) {
    const int c0 = h->c0, c1 = h->c1;
    if (c1 > c0) {
        dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(c1 - c0, 32));
        k_finalize_pots<<<g, b, 0, h->st>>>(nullptr, h->prim.p, nullptr, scale, h->N, h->nE, c0, c1, h->ld, h->U.p); LAUNCH(h);
    }
    CK(cudaGetLastError());
    h->shard_solved = true;
    h->pots_valid = (c0 == 0 && c1 == h->nS);
    return 0;
}

// partial electrode-potential matrix of this shard's sources
int pickup(pgb200_ert *h) {
    CK(cudaMemsetAsync(h->pM.p, 0, sizeof(double) * h->nE * h->nE, h->st));
    // only this shard's source columns [c0,c1) are summed; other ranks add theirs by all-reduce
    k_pickup<<<dim3(cdiv(h->nE, 64), h->nE), 64, 0, h->st>>>(h->U.p, h->ld, h->kw.p, h->nK, h->nE, h->pick_ptr.p, h->pick_idx.p,
                                                              h->pick_w.p, h->c0, h->c1, h->pM.p); LAUNCH(h);
    CK(cudaGetLastError());
    return 0;
}
int finish_response(pgb200_ert *h, double *rhoa_dev) {
    k_response<<<cdiv(h->D, 128), 128, 0, h->st>>>(h->pM.p, h->nE, h->abmn.p, h->kfac.p, h->D, h->resp.p, h->resp_rez.p, rhoa_dev); LAUNCH(h);
    CK(cudaGetLastError());
    return 0;
}

int build_jac2_plan(pgb200_ert *h);

int jac_plane(int nPp, int nQp) {               // plane stride of the element-major Gram block: >= #tiles and = 4 (mod 16), so that
    const int nt = (nPp / 4) * (nQp / 4);       // 16 consecutive q of one p hit 16 different 8-byte banks: (q & 3) * 4 + (q >> 2)
    return (nt + 11) / 16 * 16 + 4;
}
size_t jac_smem(int NL, int nPp, int nQp) {
    // double-buffered gathers, V, K, M, corner coordinates (2 x 4 x 3), G (16 planes + zero slot)
    return sizeof(double) * (2 * (size_t)NL * nPp + 3 * (size_t)NL * nQp + 2 * (size_t)NL * NL + 24 + 16 * (size_t)jac_plane(nPp, nQp) + 1);
}

// Build the chunked electrode-pair plan of the Jacobian for data rows [row0,row1)
int build_jac_plan(pgb200_ert *h) {
    const int NL = h->nloc;
    const int r0 = h->row0, r1 = h->row1, nd = r1 - r0;
    h->chunks.clear(); h->j_rows = nd; h->jac_valid = false;
    if (nd <= 0) return 0;
    std::vector<int> qmap(h->nE, -1), qlist;
    for (int d = r0; d < r1; d++) for (int t = 2; t < 4; t++) { int e = h->h_abmn[4 * d + t]; if (e >= 0 && qmap[e] < 0) qmap[e] = 1; }
    for (int e = 0; e < h->nE; e++) if (qmap[e] > 0) { qmap[e] = (int)qlist.size(); qlist.push_back(e); }
    h->nQ = (int)qlist.size(); h->nQp = std::max(4, (h->nQ + 3) / 4 * 4);
    std::vector<int> order(nd);
    for (int i = 0; i < nd; i++) order[i] = r0 + i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const int ax = h->h_abmn[4 * x], ay = h->h_abmn[4 * y];
        if (ax != ay) return ax < ay;
        return h->h_abmn[4 * x + 1] < h->h_abmn[4 * y + 1];
    });
    int dev_smem = 0;
    CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const size_t smem_cap = std::min<size_t>((size_t)dev_smem, 220 * 1024) - 256;
    const int max_tiles = 1024;
    std::vector<int> plist_all, outr; std::vector<JacDatum> idx; std::vector<double> kf;
    size_t i = 0;
    while (i < (size_t)nd) {
        std::vector<int> pmap(h->nE, -1), plist;
        JacChunk ch; ch.plist_off = plist_all.size(); ch.data_off = idx.size(); ch.nd = 0; ch.idx_in_smem = 0; ch.kfac_in_smem = 0;
        while (i < (size_t)nd) {
            const int d = order[i];
            const int a = h->h_abmn[4 * d], b = h->h_abmn[4 * d + 1];
            int add = 0;
            if (a >= 0 && pmap[a] < 0) add++;
            if (b >= 0 && b != a && pmap[b] < 0) add++;
            const int nPp_new = std::max(4, ((int)plist.size() + add + 3) / 4 * 4);
            const bool fits = jac_smem(NL, nPp_new, h->nQp) <= smem_cap && (nPp_new / 4) * (h->nQp / 4) <= max_tiles;
            if (!fits) {
                if (plist.empty()) PGB_FAIL("Jacobian tile does not fit shared memory (too many potential electrodes)");
                break;
            }
            if (a >= 0 && pmap[a] < 0) { pmap[a] = (int)plist.size(); plist.push_back(a); }
            if (b >= 0 && pmap[b] < 0) { pmap[b] = (int)plist.size(); plist.push_back(b); }
            const int m = h->h_abmn[4 * d + 2], n = h->h_abmn[4 * d + 3];
            JacDatum jd;
            jd.a = (unsigned short)(a >= 0 ? pmap[a] : 0xFFFF); jd.b = (unsigned short)(b >= 0 ? pmap[b] : 0xFFFF);
            jd.m = (unsigned short)(m >= 0 ? qmap[m] : 0xFFFF); jd.n = (unsigned short)(n >= 0 ? qmap[n] : 0xFFFF);
            idx.push_back(jd);
            outr.push_back(d - r0); kf.push_back(h->h_kfac[d]);
            ch.nd++; i++;
        }
        ch.nP = (int)plist.size(); ch.nPp = std::max(4, (ch.nP + 3) / 4 * 4);
        ch.smem = jac_smem(NL, ch.nPp, h->nQp);
        // resolved form: the record holds the offsets of G[a][m], G[a][n], G[b][m], G[b][n] inside the Gram block
        // (unused electrode -> the zero slot at offset nQp, the padding column of row 0)
        const int PL = jac_plane(ch.nPp, h->nQp), tilesQ = h->nQp / 4;
        ch.resolved = (16 * (size_t)PL + 1 <= 65535) ? 1 : 0;
        if (ch.resolved) {
            for (int q = 0; q < ch.nd; q++) {
                JacDatum &jd = idx[ch.data_off + q];
                const int a = jd.a, b = jd.b, m = jd.m, n = jd.n;
                auto off = [&](int p, int qq) { return (unsigned short)((p == 0xFFFF || qq == 0xFFFF) ? 16 * PL : ((p & 3) * 4 + (qq & 3)) * PL + (p >> 2) * tilesQ + (qq >> 2)); };
                // v = (G[a][m] - G[a][n]) - (G[b][m] - G[b][n])  ->  record order {am, an, bm, bn}
                jd.a = off(a, m); jd.b = off(a, n); jd.m = off(b, m); jd.n = off(b, n);
            }
        }
        const size_t isz = sizeof(JacDatum);
        if (ch.smem + (size_t)ch.nd * isz <= smem_cap) { ch.idx_in_smem = 1; ch.smem += (size_t)ch.nd * isz; }
        if (ch.smem + (size_t)ch.nd * sizeof(double) <= smem_cap) { ch.kfac_in_smem = 1; ch.smem += (size_t)ch.nd * sizeof(double); }
        ch.out_identity = 1;
        for (int q = 0; q < ch.nd; q++) if (outr[ch.data_off + q] != (int)ch.data_off + q) { ch.out_identity = 0; break; }
        const int ntiles = (ch.nPp / 4) * (h->nQp / 4);
        ch.mt = ntiles <= JAC_MAX_THREADS ? 1 : 2;
        ch.threads = std::max(128, ((ntiles + ch.mt - 1) / ch.mt + 31) / 32 * 32);
        plist_all.insert(plist_all.end(), plist.begin(), plist.end());
        h->chunks.push_back(ch);
    }
    CKR(h->j_plist.upload(plist_all.data(), plist_all.size(), h->st));
    CKR(h->j_qlist.upload(qlist.data(), qlist.size(), h->st));
    CKR(h->j_idx.upload(idx.data(), idx.size(), h->st));
    CKR(h->j_out.upload(outr.data(), outr.size(), h->st));
    CKR(h->j_kfac.upload(kf.data(), kf.size(), h->st));
    CK(cudaStreamSynchronize(h->st));
    h->ldJ = ((size_t)nd + 1) / 2 * 2;
    return 0;
}

template <int E>
int launch_jacobian(pgb200_ert *h, const double *rho_col) {
    size_t maxsm = 0;
    for (auto &c : h->chunks) maxsm = std::max(maxsm, c.smem);
    CK(cudaFuncSetAttribute(k_jacobian<E, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxsm));
    CK(cudaFuncSetAttribute(k_jacobian<E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxsm));
    for (auto &c : h->chunks) {
        JacArgs A;
        A.pos = h->pos.p; A.cells = h->cells.p; A.nloc = h->nloc;
        A.jac_cells = h->jac_cells.p; A.jac_col_ptr = h->jac_col_ptr.p; A.col_begin = 0; A.col_end = h->M;
        A.U = h->U.p; A.ld = h->ld; A.nE = h->nE; A.nK = h->nK; A.kvals = h->kvals.p; A.kw = h->kw.p;
        A.plist = h->j_plist.p + c.plist_off; A.nP = c.nP; A.nPp = c.nPp;
        A.qlist = h->j_qlist.p; A.nQ = h->nQ; A.nQp = h->nQp; A.gplane = jac_plane(c.nPp, h->nQp);
        A.idx = h->j_idx.p + c.data_off; A.resolved = c.resolved;
        A.idx_in_smem = c.idx_in_smem; A.kfac_in_smem = c.kfac_in_smem; A.out_identity = c.out_identity; A.out_base = (int)c.data_off;
        A.out_row = h->j_out.p + c.data_off; A.kfac = h->j_kfac.p + c.data_off; A.nd = c.nd;
        A.rho_col = rho_col; A.Jt = h->Jt.p; A.ldJ = h->ldJ;
        int occ = 1;
        if (c.mt == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_jacobian<E, 1>, c.threads, c.smem));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_jacobian<E, 2>, c.threads, c.smem));
        const int grid = std::max(1, std::min(h->M, h->num_sms * std::max(1, occ)));
        if (grid != h->j_grid) {
            // contiguous column ranges per CTA, balanced by the number of cells (+1 per column for the epilogue)
            std::vector<int> ptr(grid + 1, h->M);
            const long long total = (long long)h->h_jac_col_ptr[h->M] + h->M;
            int col = 0; ptr[0] = 0;
            for (int b = 1; b < grid; b++) {
                const long long target = total * b / grid;
                while (col < h->M && (long long)h->h_jac_col_ptr[col] + col < target) col++;
                ptr[b] = col;
            }
            CKR(h->j_cta_ptr.upload(ptr.data(), ptr.size(), h->st));
            h->j_grid = grid;
        }
        A.cta_col_ptr = h->j_cta_ptr.p;
        if (c.mt == 1) k_jacobian<E, 1><<<grid, c.threads, c.smem, h->st>>>(A);
        else k_jacobian<E, 2><<<grid, c.threads, c.smem, h->st>>>(A);
        LAUNCH(h);
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- second-generation Jacobian: host plan ---------------------------------------------------------------------------
// one side of the measurement in a chosen basis: list of (a, b) basis entries (b = -1: a single electrode) and, per datum,
// up to two signed indices into the list (dipole basis: one index; electrode basis: + index of a / m, - index of b / n)
struct JacSide {
    bool dipole = false;
    std::vector<std::pair<int, int>> list;
    std::vector<int> i0, i1;       // per datum; i1 = -1: no second term
};
static JacSide jac_side(const std::vector<int> &abmn, int r0, int r1, int t0) {
    JacSide S;
    const int nd = r1 - r0;
    std::vector<std::pair<int, int>> dip; std::vector<int> el;
    for (int d = r0; d < r1; d++) {
        const int x = abmn[4 * d + t0], y = abmn[4 * d + t0 + 1];
        dip.push_back({x, y});
        if (x >= 0) el.push_back(x);
        if (y >= 0) el.push_back(y);
    }
    std::vector<std::pair<int, int>> ud(dip); std::sort(ud.begin(), ud.end()); ud.erase(std::unique(ud.begin(), ud.end()), ud.end());
    std::sort(el.begin(), el.end()); el.erase(std::unique(el.begin(), el.end()), el.end());
    // dipoles whose first electrode is missing cannot be a basis entry (u_a - u_b needs a); fall back to electrodes
    bool ok = true;
    for (auto &p : ud) if (p.first < 0) ok = false;
    S.dipole = ok && ud.size() <= el.size() + el.size() / 4 + 4;
    S.i0.assign(nd, -1); S.i1.assign(nd, -1);
    if (S.dipole) {
        S.list = ud;
        for (int d = 0; d < nd; d++) S.i0[d] = (int)(std::lower_bound(ud.begin(), ud.end(), dip[d]) - ud.begin());
    } else {
        for (int e : el) S.list.push_back({e, -1});
        for (int d = 0; d < nd; d++) {
            if (dip[d].first >= 0) S.i0[d] = (int)(std::lower_bound(el.begin(), el.end(), dip[d].first) - el.begin());
            if (dip[d].second >= 0) S.i1[d] = (int)(std::lower_bound(el.begin(), el.end(), dip[d].second) - el.begin());
        }
    }
    return S;
}

int build_jac2_plan(pgb200_ert *h) {
    h->j2_ok = false; h->chunks2.clear();
    if (!h->jac_v2) return 0;
    const int NL = h->nloc, r0 = h->row0, r1 = h->row1, nd = r1 - r0;
    if (nd <= 0) { h->j2_ok = true; return 0; }
    JacSide P = jac_side(h->h_abmn, r0, r1, 0), Q = jac_side(h->h_abmn, r0, r1, 2);
    // one shared list (symmetric Gram block) when both sides use the same kind of basis and their union is not much larger
    bool shared = false;
    if (P.dipole == Q.dipole) {
        std::vector<std::pair<int, int>> uni(P.list); uni.insert(uni.end(), Q.list.begin(), Q.list.end());
        std::sort(uni.begin(), uni.end()); uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
        if (uni.size() <= std::max(P.list.size(), Q.list.size()) * 6 / 5 + 2) {
            auto remap = [&](JacSide &S) {
                std::vector<int> m(S.list.size());
                for (size_t i = 0; i < S.list.size(); i++) m[i] = (int)(std::lower_bound(uni.begin(), uni.end(), S.list[i]) - uni.begin());
                for (auto &v : S.i0) if (v >= 0) v = m[v];
                for (auto &v : S.i1) if (v >= 0) v = m[v];
                S.list = uni;
            };
            remap(P); remap(Q); shared = true;
        }
    }
    const int terms = (P.dipole ? 1 : 2) * (Q.dipole ? 1 : 2);
    h->j2_terms = terms; h->j2_shared = shared ? 1 : 0;
    h->j2_nQ = (int)Q.list.size(); h->j2_nQp = std::max(4, (h->j2_nQ + 3) / 4 * 4);
    const int nQp = h->j2_nQp, tilesQ = nQp / 4;
    int dev_smem = 0;
    CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const size_t cap = (size_t)dev_smem - 1024;
    const unsigned rec_bytes = (unsigned)h->recn * 8u;
    // device-side concatenations
    std::vector<int> pa, pb, outr; std::vector<unsigned short> ttp, ttq, toff; std::vector<double> kf;
    // a chunk over `rows` (indices relative to r0) with P basis entries `plist` (indices into P.list; identity when shared)
    auto make_chunk = [&](const std::vector<int> &rows, bool sh, Jac2Chunk &c, bool commit) -> bool {
        // P sub-list of this chunk
        std::vector<int> pmap(P.list.size(), -1), plist;
        if (sh) { plist.resize(P.list.size()); for (size_t i = 0; i < plist.size(); i++) { plist[i] = (int)i; pmap[i] = (int)i; } }
        else {
            std::vector<char> used(P.list.size(), 0);
            for (int d : rows) { if (P.i0[d] >= 0) used[P.i0[d]] = 1; if (P.i1[d] >= 0) used[P.i1[d]] = 1; }
            for (size_t i = 0; i < used.size(); i++) if (used[i]) { pmap[i] = (int)plist.size(); plist.push_back((int)i); }
        }
        c.nP = (int)plist.size(); c.nPp = std::max(4, (c.nP + 3) / 4 * 4);
        const int tilesP = c.nPp / 4;
        // tiles that some term needs
        std::vector<int> tslot((size_t)tilesP * tilesQ, -1);
        auto canon = [&](int p, int q, int &tp, int &tq, int &e) {
            tp = p >> 2; tq = q >> 2; e = (p & 3) * 4 + (q & 3);
            if (sh && tp > tq) { std::swap(tp, tq); e = (q & 3) * 4 + (p & 3); }
        };
        std::vector<std::pair<int, int>> tl;
        auto need = [&](int p, int q) { if (p < 0 || q < 0) return; int tp, tq, e; canon(pmap[p], q, tp, tq, e); int &sl = tslot[(size_t)tp * tilesQ + tq]; if (sl < 0) { sl = (int)tl.size(); tl.push_back({tp, tq}); } };
        for (int d : rows) { need(P.i0[d], Q.i0[d]); need(P.i0[d], Q.i1[d]); need(P.i1[d], Q.i0[d]); need(P.i1[d], Q.i1[d]); }
        c.n_tiles = (int)tl.size();
        if (c.n_tiles > J2_MAX_MT * J2_GT) return false;
        c.mt = c.n_tiles <= J2_GT ? 1 : 2;
        c.PL = std::max(1, c.n_tiles) | 1;
        if (16 * (size_t)c.PL + 2 > 65535) return false;
        c.nd = (int)rows.size();
        c.rec_bytes = rec_bytes;
        c.uq_off = rec_bytes + (unsigned)(NL * c.nPp * 8);
        c.slot_bytes = (unsigned)up128(rec_bytes + (size_t)NL * c.nPp * 8 + (sh ? 0 : (size_t)NL * nQp * 8));
        size_t base = (size_t)J2_SLOTS * c.slot_bytes + 2 * (size_t)NL * nQp * 8 + 2 * (16 * (size_t)c.PL + 2) * 8;
        if (base > cap) return false;
        c.kfac_in_smem = 0; c.off_in_smem = 0;
        const size_t kfb = (size_t)((c.nd + 1) & ~1) * 8, ofb = (size_t)c.nd * terms * 2 + 16;
        if (base + kfb + ofb <= cap) { c.kfac_in_smem = 1; c.off_in_smem = 1; base += kfb + ofb; }
        else if (base + ofb <= cap) { c.off_in_smem = 1; base += ofb; }
        c.smem = base;
        if (!commit) return true;
        c.data_off = kf.size(); c.plist_off = pa.size(); c.tile_off = ttp.size();
        for (int i : plist) { pa.push_back(P.list[i].first); pb.push_back(P.list[i].second); }
        for (auto &t : tl) { ttp.push_back((unsigned short)t.first); ttq.push_back((unsigned short)t.second); }
        const unsigned short zero = (unsigned short)(16 * c.PL);
        auto off = [&](int p, int q) -> unsigned short {
            if (p < 0 || q < 0) return zero;
            int tp, tq, e; canon(pmap[p], q, tp, tq, e);
            return (unsigned short)(e * c.PL + tslot[(size_t)tp * tilesQ + tq]);
        };
        c.out_identity = 1; c.out_base = (int)c.data_off;
        for (size_t x = 0; x < rows.size(); x++) {
            const int d = rows[x];
            // v = t0 [- t1] [- (t2 - t3)]
            if (terms == 1) toff.push_back(off(P.i0[d], Q.i0[d]));
            else if (terms == 2 && P.dipole) { toff.push_back(off(P.i0[d], Q.i0[d])); toff.push_back(off(P.i0[d], Q.i1[d])); }
            else if (terms == 2) { toff.push_back(off(P.i0[d], Q.i0[d])); toff.push_back(off(P.i1[d], Q.i0[d])); }
            else { toff.push_back(off(P.i0[d], Q.i0[d])); toff.push_back(off(P.i0[d], Q.i1[d])); toff.push_back(off(P.i1[d], Q.i0[d])); toff.push_back(off(P.i1[d], Q.i1[d])); }
            outr.push_back(d); kf.push_back(h->h_kfac[r0 + d]);
            if (d != (int)c.data_off + (int)x) c.out_identity = 0;
        }
        return true;
    };
    std::vector<int> all(nd);
    for (int i = 0; i < nd; i++) all[i] = i;
    Jac2Chunk c{};
    if (make_chunk(all, shared, c, false)) { make_chunk(all, shared, c, true); h->chunks2.push_back(c); }
    else {
        // rows ordered by their first P index; greedy chunks that fit (no symmetry inside a chunk: its P list is a subset)
        h->j2_shared = 0;
        std::vector<int> order(all);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return std::max(P.i0[x], P.i1[x]) < std::max(P.i0[y], P.i1[y]); });
        size_t i = 0;
        while (i < order.size()) {
            // largest prefix [i, j) that fits, by doubling + bisection on the row count
            size_t lo = i + 1, hi = order.size();
            auto fits = [&](size_t j) { std::vector<int> rows(order.begin() + i, order.begin() + j); Jac2Chunk t{}; return make_chunk(rows, false, t, false); };
            if (!fits(lo)) PGB_FAIL("Jacobian: a single data row does not fit the Gram buffers");
            if (!fits(hi)) { while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (fits(mid)) lo = mid; else hi = mid; } } else lo = hi;
            std::vector<int> rows(order.begin() + i, order.begin() + lo);
            Jac2Chunk cc{};
            make_chunk(rows, false, cc, true);
            h->chunks2.push_back(cc);
            i = lo;
        }
    }
    std::vector<int> qa, qb;
    for (auto &e : Q.list) { qa.push_back(e.first); qb.push_back(e.second); }
    CKR(h->j2_pa.upload(pa.data(), pa.size(), h->st)); CKR(h->j2_pb.upload(pb.data(), pb.size(), h->st));
    CKR(h->j2_qa.upload(qa.data(), qa.size(), h->st)); CKR(h->j2_qb.upload(qb.data(), qb.size(), h->st));
    CKR(h->j2_tp.upload(ttp.data(), ttp.size(), h->st)); CKR(h->j2_tq.upload(ttq.data(), ttq.size(), h->st));
    CKR(h->j2_off.upload(toff.data(), toff.size(), h->st)); CKR(h->j2_out.upload(outr.data(), outr.size(), h->st));
    CKR(h->j2_kfac.upload(kf.data(), kf.size(), h->st));
    int maxPp = 4;
    for (auto &cc : h->chunks2) maxPp = std::max(maxPp, cc.nPp);
    CKR(h->UDp.alloc((size_t)h->N * h->nK * maxPp));
    if (!h->j2_shared) CKR(h->UDq.alloc((size_t)h->N * h->nK * nQp));
    CK(cudaStreamSynchronize(h->st));
    h->j2_ok = true;
    return 0;
}

template <int E, int MT, int EW>
int jac2_go_ew(pgb200_ert *h, const Jac2Args &A, int terms, int grid, size_t smem) {
#define J2GO(T) do { CK(cudaFuncSetAttribute(k_jacobian2<E, MT, T, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                     k_jacobian2<E, MT, T, EW><<<grid, j2_threads(EW), smem, h->st>>>(A); } while (0)
    if (terms == 1) J2GO(1); else if (terms == 2) J2GO(2); else J2GO(4);
#undef J2GO
    LAUNCH(h);
    return 0;
}
// epilogue warps: 8 when the store of J bounds the kernel, 4 when the Gram blocks do -- bytes of J written per (tile, cell) of
// Gram work; c3: 9700 rows x 8 B / (325 tiles x 4 cells) = 60, c4: 3320 x 8 / (384 x 6) = 11.5
template <int E, int MT>
int jac2_go(pgb200_ert *h, const Jac2Args &A, int terms, int grid, size_t smem) {
    const double cells_per_col = std::max(1.0, (double)h->n_jac_cells / std::max(1, h->M));
    const double bytes_per_work = 8.0 * A.nd / (std::max(1, A.n_tiles) * cells_per_col * h->nK);
    static const int force = getenv("PGB200_J2_EPI") ? atoi(getenv("PGB200_J2_EPI")) : 0;
    const bool wide = force ? force >= 8 : bytes_per_work > 45.0;     // c5 (30) is faster with 4 warps: 4.77 vs 5.08 ms
    if (wide) return jac2_go_ew<E, MT, 8>(h, A, terms, grid, smem);
    return jac2_go_ew<E, MT, 4>(h, A, terms, grid, smem);
}

template <int E>
int launch_jacobian2(pgb200_ert *h, const double *rho_col) {
    const int nQp = h->j2_nQp;
    auto basis = [&](const int *la, const int *lb, int nL, int nLp, double *UD) -> int {
        dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(h->nK * nLp, 32));
        k_basis_pots<<<g, b, 0, h->st>>>(h->U.p, h->ld, h->N, h->nE, h->nK, la, lb, nL, nLp, UD, (size_t)h->nK * nLp); LAUNCH(h);
        return 0;
    };
    if (!h->j2_shared) CKR(basis(h->j2_qa.p, h->j2_qb.p, h->j2_nQ, nQp, h->UDq.p));
    const int grid = std::max(1, std::min(h->M, h->num_sms));
    if (grid != h->j2_grid) {
        // contiguous column ranges per CTA, balanced by the number of cells (+1 per column for the epilogue)
        std::vector<int> ptr(grid + 1, h->M);
        const long long total = (long long)h->h_jac_col_ptr[h->M] + h->M;
        int col = 0; ptr[0] = 0;
        for (int b = 1; b < grid; b++) {
            const long long target = total * b / grid;
            while (col < h->M && (long long)h->h_jac_col_ptr[col] + col < target) col++;
            ptr[b] = col;
        }
        CKR(h->j_cta_ptr.upload(ptr.data(), ptr.size(), h->st));
        h->j2_grid = grid; h->j_grid = 0;
    }
    for (auto &c : h->chunks2) {
        CKR(basis(h->j2_pa.p + c.plist_off, h->j2_pb.p + c.plist_off, c.nP, c.nPp, h->UDp.p));
        Jac2Args A;
        A.cells = h->cells.p; A.C = h->C; A.jac_cells = h->jac_cells.p; A.jac_col_ptr = h->jac_col_ptr.p; A.cta_col_ptr = h->j_cta_ptr.p;
        A.erec = h->erec.p; A.recn = h->recn; A.mu = h->mu_tab.p;
        A.UDp = h->UDp.p; A.ldUDp = (size_t)h->nK * c.nPp; A.UDq = h->j2_shared ? h->UDp.p : h->UDq.p; A.ldUDq = h->j2_shared ? A.ldUDp : (size_t)h->nK * nQp;
        A.nPp = c.nPp; A.nQp = h->j2_shared ? c.nPp : nQp; A.shared = h->j2_shared;
        A.nK = h->nK; A.kvals = h->kvals.p; A.kw = h->kw.p;
        A.tile_tp = h->j2_tp.p + c.tile_off; A.tile_tq = h->j2_tq.p + c.tile_off; A.n_tiles = c.n_tiles; A.PL = c.PL; A.mt = c.mt;
        A.toff = h->j2_off.p + c.data_off * h->j2_terms; A.terms = h->j2_terms;
        A.kfac = h->j2_kfac.p + c.data_off; A.out_row = h->j2_out.p + c.data_off; A.nd = c.nd; A.out_identity = c.out_identity; A.out_base = c.out_base;
        A.off_in_smem = c.off_in_smem; A.kfac_in_smem = c.kfac_in_smem;
        A.rho_col = rho_col; A.Jt = h->Jt.p; A.ldJ = h->ldJ;
        A.slot_bytes = c.slot_bytes; A.uq_off = c.uq_off; A.rec_bytes = c.rec_bytes;
        if (c.mt == 1) CKR((jac2_go<E, 1>(h, A, h->j2_terms, grid, c.smem)));
        else CKR((jac2_go<E, 2>(h, A, h->j2_terms, grid, c.smem)));
    }
    CK(cudaGetLastError());
    return 0;
}

int jacobian(pgb200_ert *h, const double *rho_col) {
    if (h->j_rows <= 0) { h->jac_valid = true; return 0; }
    if (h->Jt.n < (size_t)h->M * h->ldJ) CKR(h->Jt.alloc((size_t)h->M * h->ldJ));
    phase_begin(h, PH_JAC);
    if (h->prof) CK(cudaEventRecord(h->jev[0], h->st));
    if (h->jac_v2 && h->j2_ok) {
        switch (h->elem) {
            case TRI3:  CKR(launch_jacobian2<TRI3>(h, rho_col)); break;
            case TRI6:  CKR(launch_jacobian2<TRI6>(h, rho_col)); break;
            case TET4:  CKR(launch_jacobian2<TET4>(h, rho_col)); break;
            case TET10: CKR(launch_jacobian2<TET10>(h, rho_col)); break;
            default: PGB_FAIL("unknown element type");
        }
    } else {
        switch (h->elem) {
            case TRI3:  CKR(launch_jacobian<TRI3>(h, rho_col)); break;
            case TRI6:  CKR(launch_jacobian<TRI6>(h, rho_col)); break;
            case TET4:  CKR(launch_jacobian<TET4>(h, rho_col)); break;
            case TET10: CKR(launch_jacobian<TET10>(h, rho_col)); break;
            default: PGB_FAIL("unknown element type");
        }
    }
    if (h->prof) { CK(cudaEventRecord(h->jev[1], h->st)); h->jac_timed = true; }
    h->jac_valid = true;
    return 0;
}

int finish_timing(pgb200_ert *h) {
    CK(cudaEventRecord(h->ev[PH_COUNT], h->st));
    CK(cudaStreamSynchronize(h->st));
    int order[PH_COUNT + 1], n = 0;
    for (int p = 0; p < PH_COUNT; p++) { if (h->ph_rec[p]) order[n++] = p; }
    order[n] = PH_COUNT;
    for (int i = 0; i < n; i++) { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev[order[i]], h->ev[order[i + 1]]); h->ph_ms[order[i]] += ms; }
    for (int p = 0; p <= PH_COUNT; p++) h->ph_rec[p] = false;
    if (h->prof) {
        for (int i = 0; i + 1 < h->n_pev; i += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, h->pev[i], h->pev[i + 1]); h->spmm_ms += ms; h->spmm_timed++; }
        h->n_pev = 0;
        if (h->jac_timed) { float ms = 0.f; if (cudaEventElapsedTime(&ms, h->jev[0], h->jev[1]) == cudaSuccess) { h->jac_ms += ms; h->jac_launches++; } h->jac_timed = false; }
    }
    return 0;
}

double host_stddev(const std::vector<double> &v) {
    if (v.size() < 2) return 0.0;
    double mean = 0.0; for (double x : v) mean += x; mean /= (double)v.size();
    double s = 0.0; for (double x : v) s += (x - mean) * (x - mean);
    return std::sqrt(s / (double)(v.size() - 1));
}

} // namespace

// =====================================================================================
extern "C" {

const char *pgb200_last_error(void) { return g_err.c_str(); }
int pgb200_version(void) { return 100; }

int pgb200_color_cells(int n_cells, int nloc, const int *cells, int n_nodes, int *color) {
    // sequential greedy with a 256-bit "colours used around this node" mask per node
    struct Mask { unsigned long long w[4]; };
    std::vector<Mask> used((size_t)n_nodes, Mask{{0, 0, 0, 0}});
    int ncol = 0;
    for (int c = 0; c < n_cells; c++) {
        Mask f{{0, 0, 0, 0}};
        for (int j = 0; j < nloc; j++) { const Mask &u = used[cells[(size_t)c * nloc + j]]; for (int w = 0; w < 4; w++) f.w[w] |= u.w[w]; }
        int col = -1;
        for (int w = 0; w < 4 && col < 0; w++) if (~f.w[w]) col = w * 64 + __builtin_ctzll(~f.w[w]);
        if (col < 0) { g_err = "colouring needs more than 256 colours"; return -1; }
        color[c] = col; ncol = std::max(ncol, col + 1);
        for (int j = 0; j < nloc; j++) used[cells[(size_t)c * nloc + j]].w[col >> 6] |= 1ull << (col & 63);
    }
    return ncol;
}

// Streamed row panels (stream_panels.h) of a CSR pattern -- exported for the host-side tests, which replay the kernel's
// traversal on the CPU.  Two calls: with out == NULL the sizes are returned in counts[10] = {n_panels, n_chunks, halo
// entries, crp_stride, max_rows, max_chunk_halo, max_chunk_ent, nnz, runs, max runs per chunk}; with the arrays allocated
// they are filled (panel_row_ptr[n_panels+1], panel_chunk_ptr[n_panels+1], chunk_halo_ptr[n_chunks+1], halo_cols[halo
// entries], chunk_ent_ptr[n_chunks+1], ent_src[nnz], ent_idx[nnz], crp[n_chunks*crp_stride], chunk_run_ptr[n_chunks+1],
// runs[3*runs] = {first halo entry in the chunk, column, length}).
int pgb200_build_stream_panels(int n_rows, const int *rowptr, const int *colidx, int rmax, int hc, int max_chunks, int *counts,
                               int *panel_row_ptr, int *panel_chunk_ptr, int *chunk_halo_ptr, int *halo_cols, int *chunk_ent_ptr,
                               int *ent_src, unsigned *ent_idx, int *crp, int *chunk_run_ptr, int *runs) {
    if (!rowptr || !colidx || !counts) { g_err = "null argument"; return 1; }
    StreamPanelsHost S;
    const std::string err = build_stream_panels(n_rows, rowptr, colidx, rmax, hc, max_chunks, S);
    if (!err.empty()) { g_err = err; return 1; }
    const int c[10] = {S.n_panels, S.n_chunks, (int)S.halo_cols.size(), S.crp_stride, S.max_rows, S.max_chunk_halo, S.max_chunk_ent, (int)S.nnz,
                       (int)S.run_start.size(), S.max_chunk_runs};
    for (int i = 0; i < 10; i++) counts[i] = c[i];
    if (!panel_row_ptr) return 0;
    std::copy(S.panel_row_ptr.begin(), S.panel_row_ptr.end(), panel_row_ptr);
    std::copy(S.panel_chunk_ptr.begin(), S.panel_chunk_ptr.end(), panel_chunk_ptr);
    std::copy(S.chunk_halo_ptr.begin(), S.chunk_halo_ptr.end(), chunk_halo_ptr);
    std::copy(S.halo_cols.begin(), S.halo_cols.end(), halo_cols);
    std::copy(S.chunk_ent_ptr.begin(), S.chunk_ent_ptr.end(), chunk_ent_ptr);
    std::copy(S.ent_src.begin(), S.ent_src.end(), ent_src);
    std::copy(S.ent_idx.begin(), S.ent_idx.end(), ent_idx);
    std::copy(S.crp.begin(), S.crp.end(), crp);
    std::copy(S.chunk_run_ptr.begin(), S.chunk_run_ptr.end(), chunk_run_ptr);
    for (size_t i = 0; i < S.run_start.size(); i++) { runs[3 * i] = S.run_start[i]; runs[3 * i + 1] = S.run_col[i]; runs[3 * i + 2] = S.run_len[i]; }
    return 0;
}

// 8-row-group form of the streamed panels (k_spmm_mma) -- exported for the host-side tests.  Two calls like above:
// counts[8] = {n_panels, n_chunks, halo entries, k-steps, meta words, meta_gstride, max k-steps per chunk, max meta words per
// chunk}; arrays: panel_row_ptr[n_panels+1], panel_chunk_ptr[n_panels+1], chunk_halo_ptr[n_chunks+1], halo_cols[halo entries],
// chunk_ks_ptr[n_chunks+1], a_src[32 * k-steps], chunk_meta_ptr[n_chunks+1], meta[meta words].
int pgb200_build_mma_panels(int n_rows, const int *rowptr, const int *colidx, int groups, int hc, int max_chunks, int rowb_hint,
                            int *counts, int *panel_row_ptr, int *panel_chunk_ptr, int *chunk_halo_ptr, int *halo_cols,
                            int *chunk_ks_ptr, int *a_src, int *chunk_meta_ptr, unsigned *meta) {
    if (!rowptr || !colidx || !counts) { g_err = "null argument"; return 1; }
    StreamPanelsHost S;
    const std::string err = build_stream_panels(n_rows, rowptr, colidx, 8 * groups, hc, max_chunks, S, groups, rowb_hint);
    if (!err.empty()) { g_err = err; return 1; }
    const int c[8] = {S.n_panels, S.n_chunks, (int)S.halo_cols.size(), (int)S.n_ks, (int)S.meta.size(), S.meta_gstride, S.max_chunk_ks, S.max_chunk_meta};
    for (int i = 0; i < 8; i++) counts[i] = c[i];
    if (!panel_row_ptr) return 0;
    std::copy(S.panel_row_ptr.begin(), S.panel_row_ptr.end(), panel_row_ptr);
    std::copy(S.panel_chunk_ptr.begin(), S.panel_chunk_ptr.end(), panel_chunk_ptr);
    std::copy(S.chunk_halo_ptr.begin(), S.chunk_halo_ptr.end(), chunk_halo_ptr);
    std::copy(S.halo_cols.begin(), S.halo_cols.end(), halo_cols);
    std::copy(S.chunk_ks_ptr.begin(), S.chunk_ks_ptr.end(), chunk_ks_ptr);
    std::copy(S.a_src.begin(), S.a_src.end(), a_src);
    std::copy(S.chunk_meta_ptr.begin(), S.chunk_meta_ptr.end(), chunk_meta_ptr);
    std::copy(S.meta.begin(), S.meta.end(), meta);
    return 0;
}

// Pairwise aggregation for the multilevel preconditioner: nodes are visited in order; an unaggregated node is matched
// with the unaggregated neighbour it is most strongly coupled to (most negative off-diagonal), provided the coupling is
// STRONG for both of them: -a_ij >= theta * max_k(-a_ik) and >= theta * max_k(-a_jk).  Nodes left alone join the
// aggregate of their strongest neighbour if that coupling is strong for them, and stay singletons otherwise.  The
// threshold keeps the coarsening out of the weak directions of stretched cells (graded meshes), where the point
// smoother does not smooth: measured 188 -> 118 PCG iterations on the graded 3-D mesh, 504 -> 241 in 2.5-D, 729 -> 337 on
// P2 triangles, at the same operator complexity (theta = 0.25; theta = 0 is the unconditional matching).
int pgb200_pairwise_aggregate(int n, const int *rowptr, const int *colidx, const double *vals, const int *group, double theta,
                              int *agg) {
    if (n < 0 || !rowptr || !colidx || !vals || !agg) { g_err = "null argument"; return -1; }
    std::vector<double> rowmax((size_t)n, 0.0);
    for (int i = 0; i < n; i++) {
        double m = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++)
            if (colidx[p] != i && -vals[p] > m) m = -vals[p];
        rowmax[i] = m;
    }
    for (int i = 0; i < n; i++) agg[i] = -1;
    int na = 0;
    for (int i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        int best = -1; double bs = theta * rowmax[i];
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const int j = colidx[p];
            if (j == i || agg[j] >= 0) continue;
            if (group && group[j] != group[i]) continue;
            const double sgn = -vals[p];
            if (sgn > bs && sgn >= theta * rowmax[j]) { bs = sgn; best = j; }
        }
        if (best >= 0) { agg[i] = na; agg[best] = na; na++; }
    }
    for (int i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        int best = -1; double bs = theta * rowmax[i];
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const int j = colidx[p];
            if (j == i || agg[j] < 0) continue;
            if (group && group[j] != group[i]) continue;
            const double sgn = -vals[p];
            if (sgn > bs) { bs = sgn; best = j; }
        }
        agg[i] = (best >= 0) ? agg[best] : na++;
    }
    return na;
}

int pgb200_ert_create(const pgb200_plan *p, int device, pgb200_ert **out) {
    if (!p || !out) PGB_FAIL("null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) PGB_FAIL("no CUDA device: the B200 ERT path has no CPU fallback");
    if (device < 0 || device >= ndev) PGB_FAIL("invalid CUDA device index");
    CK(cudaSetDevice(device));
    pgb200_ert *h = new pgb200_ert();
    *out = h;
    h->device = device;
    CK(cudaStreamCreate(&h->own_st));          // blocking stream: implicitly ordered with the legacy default stream
    h->st = h->own_st;
    CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    h->dim = p->dim; h->nloc = p->nloc; h->N = p->n_nodes; h->C = p->n_cells; h->nnz = (size_t)p->nnz;
    h->nE = p->n_elec; h->nK = p->n_k; h->nS = h->nE * h->nK; h->M = p->n_model; h->D = p->n_data; h->sr = p->sr;
    h->fullspace = p->fullspace; h->surface_z = p->surface_z; h->topography = p->topography;
    h->ref_node = p->ref_node; h->ref_last = p->ref_last;
    if (h->ref_node >= p->n_nodes) PGB_FAIL("plan.ref_node out of range");
    if (p->sr && (h->ref_node >= 0 || h->ref_last))
        PGB_FAIL("singularity removal with a current reference electrode (reference-electrode node -999 or pure-Neumann domain: dipole patterns, "
                 "dcfemmodelling.cpp:1054-1064, 1517-1523) is not on the B200 path; use the total-field operator (sr = 0)");
    if (p->dim == 2 && p->nloc == 3) h->elem = TRI3; else if (p->dim == 2 && p->nloc == 6) h->elem = TRI6;
    else if (p->dim == 3 && p->nloc == 4) h->elem = TET4; else if (p->dim == 3 && p->nloc == 10) h->elem = TET10;
    else PGB_FAIL("unsupported cell type (need Tri3/Tri6/Tet4/Tet10)");
    h->ld = ((size_t)h->nS + 3) / 4 * 4;
    h->c0 = 0; h->c1 = h->nS; h->row0 = 0; h->row1 = h->D;
    cudaStream_t st = h->st;
    const int N = h->N, C = h->C, nE = h->nE, nK = h->nK, NL = h->nloc;
    CKR(h->pos.upload(p->pos, (size_t)N * 3, st));
    CKR(h->cells.upload(p->cells, (size_t)C * NL, st));
    CKR(h->cell_marker.upload(p->cell_marker, C, st));
    CKR(h->rowptr.upload(p->rowptr, (size_t)N + 1, st));
    CKR(h->colidx.upload(p->colidx, h->nnz, st));
    CKR(h->diag_pos.upload(p->diag_pos, N, st));
    h->n_colors = p->n_colors; h->color_ptr.assign(p->color_ptr, p->color_ptr + p->n_colors + 1);
    CKR(h->color_order.upload(p->color_order, C, st));
    CKR(h->cells_col.upload(p->cells_col, (size_t)C * NL, st));
    CKR(h->pos_col.upload(p->pos_col, (size_t)C * NL * NL, st));
    CKR(h->kvals.upload(p->k_values, nK, st)); CKR(h->kw.upload(p->k_weights, nK, st));
    h->h_kvals.assign(p->k_values, p->k_values + nK);
    h->n_bc_slots = p->n_bc_slots; h->n_bc_entries = p->n_bc_entries;
    CKR(h->bc_slot.upload(p->bc_slot, p->n_bc_slots, st)); CKR(h->bc_ptr.upload(p->bc_ptr, (size_t)p->n_bc_slots + 1, st));
    CKR(h->bc_owner.upload(p->bc_owner, p->n_bc_entries, st)); CKR(h->bc_coef.upload(p->bc_coef, (size_t)nK * p->n_bc_entries, st));
    h->n_dir_zero = p->n_dir_zero; h->n_dir_nodes = p->n_dir_nodes;
    CKR(h->dir_zero.upload(p->dir_zero_slots, p->n_dir_zero, st)); CKR(h->dir_diag.upload(p->dir_diag_slots, p->n_dir_nodes, st));
    CKR(h->dir_nodes.upload(p->dir_nodes, p->n_dir_nodes, st));
    CKR(h->el_pos.upload(p->el_pos, (size_t)nE * 3, st)); CKR(h->sing_node.upload(p->sing_node, nE, st));
    CKR(h->sing_val.upload(p->sing_val, (size_t)nK * nE, st));
    CKR(h->pick_ptr.upload(p->pick_ptr, (size_t)nE + 1, st));
    CKR(h->pick_idx.upload(p->pick_idx, p->pick_ptr[nE], st)); CKR(h->pick_w.upload(p->pick_w, p->pick_ptr[nE], st));
    CKR(h->src_cell_ptr.upload(p->src_cell_ptr, (size_t)nE + 1, st)); CKR(h->src_cells.upload(p->src_cells, p->src_cell_ptr[nE], st));
    h->pro_nf = p->pro_nf; h->pro_level_ptr.assign(p->pro_level_ptr, p->pro_level_ptr + p->n_pro_levels + 1);
    const size_t npro = p->n_pro_levels ? (size_t)p->pro_level_ptr[p->n_pro_levels] : 0;
    CKR(h->pro_cells.upload(p->pro_cells, npro, st)); CKR(h->pro_nb.upload(p->pro_nb, npro * p->pro_nf, st));
    CKR(h->pro_w.upload(p->pro_w, npro * p->pro_nf, st));
    // streamed row panels of the fine matrix (internal layout, built here from the pattern), their kernels' shared-memory
    // opt-in, and the partial rows / tickets of the deterministic column dots
    if (const char *e = getenv("PGB200_SUBCYCLE")) h->use_subcycle = atoi(e);
    if (const char *e = getenv("PGB200_STREAM_HC")) h->stream_hc = std::max(h->stream_rmax, atoi(e));
    if (const char *e = getenv("PGB200_STREAM_CHUNKS")) h->stream_chunks = std::max(1, atoi(e));
    if (const char *e = getenv("PGB200_STREAM_ROWS")) h->stream_rmax = std::min(ST_CONSUMER_WARPS * ST_RPW, std::max(1, atoi(e)));
    h->stream_hc = std::max(h->stream_hc, h->stream_rmax);
    if (const char *e = getenv("PGB200_SPMM_MMA")) h->use_mma = atoi(e);
    if (const char *e = getenv("PGB200_MMA_HC")) h->mma_hc = std::min(128, std::max((MM_ROWS + 1) / 2, atoi(e)));
    CKR(stream_configure(h));
    CKR(mma_configure(h));
    CKR(stream_upload(h, h->stream, N, p->rowptr, p->colidx));
    h->dot_slots = std::max(FLAT_MAX_GX, h->num_sms);
    CKR(h->dot_part.alloc(2 * (size_t)h->dot_slots * h->ld));
    CKR(h->dot_counter.alloc((size_t)std::max(64, h->nS + 8)));
    CK(cudaMemsetAsync(h->dot_counter.p, 0, sizeof(unsigned) * h->dot_counter.n, st));
    h->n_jac_cells = p->n_jac_cells;
    CKR(h->jac_cells.upload(p->jac_cells, p->n_jac_cells, st)); CKR(h->jac_col_ptr.upload(p->jac_col_ptr, (size_t)h->M + 1, st));
    h->h_jac_col_ptr.assign(p->jac_col_ptr, p->jac_col_ptr + h->M + 1);
    CKR(h->abmn.upload(p->abmn, (size_t)h->D * 4, st)); CKR(h->kfac.upload(p->k_fac, h->D, st));
    h->h_abmn.assign(p->abmn, p->abmn + (size_t)h->D * 4); h->h_kfac.assign(p->k_fac, p->k_fac + h->D);
    for (int d = 0; d < h->D * 4; d++) if (h->h_abmn[d] >= nE || h->h_abmn[d] < -1) PGB_FAIL("Collect matrix too small: electrode index out of range in the data (datamap.cpp:184-194)");

    // state
    const size_t blk = (size_t)N * h->ld;
    CKR(h->model.alloc(std::max(h->M, C))); CKR(h->rho.alloc(C)); CKR(h->rho_src.alloc(nE));
    CKR(h->vals.alloc(h->nnz * nK)); CKR(h->dinv.alloc((size_t)N * nK));
    CKR(h->B.alloc(blk)); CKR(h->X.alloc(blk)); CKR(h->R.alloc(blk)); CKR(h->P.alloc(blk)); CKR(h->AP.alloc(blk)); CKR(h->U.alloc(blk));
    CK(cudaMemsetAsync(h->U.p, 0, blk * sizeof(double), st));
    CK(cudaMemsetAsync(h->P.p, 0, blk * sizeof(double), st)); CK(cudaMemsetAsync(h->AP.p, 0, blk * sizeof(double), st));
    CKR(h->scal.alloc(7 * h->ld)); CKR(h->pM.alloc((size_t)nE * nE)); CKR(h->resp.alloc(h->D)); CKR(h->resp_rez.alloc(h->D)); CKR(h->rhoa.alloc(h->D));
    CKR(h->flags.alloc(4)); CKR(h->xin.alloc(3 * (size_t)std::max(h->M, h->D))); CKR(h->yout.alloc(std::max(h->M, h->D)));
    for (int i = 0; i <= PH_COUNT; i++) CK(cudaEventCreate(&h->ev[i]));
    CK(cudaEventCreate(&h->jev[0])); CK(cudaEventCreate(&h->jev[1]));
    h->ev_ok = true;
    // geometry-only device work: primary potentials and the rho = 1 matrices
    CKR(h->prim.alloc(blk));
    CK(cudaMemsetAsync(h->prim.p, 0, blk * sizeof(double), st));
    if (!h->topography) {
        // flat earth: analytic primary potentials (exactDCSolution); with topography they are numeric and arrive
        // through pgb200_ert_set_primary_dev
        dim3 b(32, 8), g(cdiv(N, 8), cdiv(h->nS, 32));
        k_primary<<<g, b, 0, st>>>(h->pos.p, N, h->el_pos.p, nE, h->sing_node.p, h->sing_val.p, h->kvals.p, nK, h->surface_z,
                                  h->fullspace, h->prim.p, h->ld); LAUNCH(h);
        CK(cudaGetLastError());
        h->prim_set = true;
    }
    // geometry-only element records of the Jacobian kernel: per cell the stiffness matrix and the size
    if (const char *e = getenv("PGB200_JACOBIAN_V1")) h->jac_v2 = atoi(e) ? 0 : 1;
    if (h->jac_v2) {
        h->recn = (NL * NL + 1 + 1) / 2 * 2;                 // [K | size | pad]: a multiple of 16 bytes (one bulk copy)
        CKR(h->erec.alloc((size_t)C * h->recn)); CKR(h->mu_tab.alloc((size_t)NL * NL));
        switch (h->elem) {
            case TRI3:  k_element_records<TRI3><<<cdiv(C, 128), 128, 0, st>>>(h->pos.p, h->cells.p, C, h->recn, h->erec.p); k_mass_unit_table<TRI3><<<1, 128, 0, st>>>(h->mu_tab.p); break;
            case TRI6:  k_element_records<TRI6><<<cdiv(C, 128), 128, 0, st>>>(h->pos.p, h->cells.p, C, h->recn, h->erec.p); k_mass_unit_table<TRI6><<<1, 128, 0, st>>>(h->mu_tab.p); break;
            case TET4:  k_element_records<TET4><<<cdiv(C, 128), 128, 0, st>>>(h->pos.p, h->cells.p, C, h->recn, h->erec.p); k_mass_unit_table<TET4><<<1, 128, 0, st>>>(h->mu_tab.p); break;
            default:    k_element_records<TET10><<<cdiv(C, 128), 128, 0, st>>>(h->pos.p, h->cells.p, C, h->recn, h->erec.p); k_mass_unit_table<TET10><<<1, 128, 0, st>>>(h->mu_tab.p); break;
        }
        LAUNCH(h); LAUNCH(h);
        CK(cudaGetLastError());
    }
    // rho = 1 matrices: the S1 of the singularity-removal right-hand side, and the (geometry-only) strength
    // information the multilevel hierarchy is built from
    CKR(h->vals1.alloc(h->nnz * nK)); CKR(assemble(h, nullptr, h->vals1.p));
    CKR(build_jac_plan(h));
    CKR(build_jac2_plan(h));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int pgb200_ert_open_plan(pgb200_built_plan *plan, int multilevel, int device, pgb200_ert **out) {
    if (!plan || !out) PGB_FAIL("null argument");
    const pgb200_plan *v = pgb200_plan_view(plan);
    CKR(pgb200_ert_create(v, device, out));
    pgb200_ert *h = *out;
    h->built = plan;
    if (multilevel) {
        // aggregation hierarchy from the rho = 1 matrix of the smallest wavenumber (geometry only)
        std::vector<double> v1((size_t)h->nnz);
        CK(cudaMemcpy(v1.data(), h->vals1.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost));
        const int nl = pgb200_plan_build_hierarchy(plan, v1.data(), 0.25, 2, 256, 12);
        if (nl < 0) PGB_FAIL(std::string("hierarchy: ") + pgb200_plan_error());
        CKR(pgb200_ert_set_hierarchy(h, nl, pgb200_plan_levels(plan)));
        CKR(pgb200_ert_set_preconditioner(h, nl > 0 ? 1 : 0, 8));
    } else CKR(pgb200_ert_set_preconditioner(h, 0, 8));
    return 0;
}
int pgb200_ert_open(const pgb200_mesh_in *mesh, const pgb200_scheme_in *scheme, int sr, int n_k_user, const double *k_user,
                    const double *w_user, int multilevel, int device, pgb200_ert **out) {
    if (!out) PGB_FAIL("null argument");
    pgb200_built_plan *plan = nullptr;
    if (pgb200_plan_build(mesh, scheme, sr, n_k_user, k_user, w_user, &plan)) PGB_FAIL(std::string(pgb200_plan_error()));
    const int rc = pgb200_ert_open_plan(plan, multilevel, device, out);
    if (rc) { if (*out) { (*out)->built = nullptr; pgb200_ert_destroy(*out); *out = nullptr; } pgb200_plan_free(plan); return rc; }
    (*out)->owns_built = true;
    return 0;
}
const pgb200_built_plan *pgb200_ert_plan(const pgb200_ert *h) { return h ? h->built : nullptr; }

int pgb200_ert_destroy(pgb200_ert *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->ev_ok) { for (int i = 0; i <= PH_COUNT; i++) cudaEventDestroy(h->ev[i]); cudaEventDestroy(h->jev[0]); cudaEventDestroy(h->jev[1]); }
    for (auto e : h->pev) cudaEventDestroy(e);
    for (auto e : h->tev) cudaEventDestroy(e);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    for (AmgLevel *L : h->amg) delete L;
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->own_st) cudaStreamDestroy(h->own_st);
    if (h->built && h->owns_built) pgb200_plan_free(h->built);
    delete h;
    return 0;
}

int pgb200_ert_set_hierarchy(pgb200_ert *h, int n_levels, const pgb200_amg_level *lv) {
    if (!h) PGB_FAIL("null handle");
    CK(cudaSetDevice(h->device));
    for (AmgLevel *L : h->amg) delete L;
    h->amg.clear();
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    if (n_levels <= 0) return 0;
    if (!lv) PGB_FAIL("null level array");
    cudaStream_t st = h->st;
    int n_finer = h->N; size_t nnz_finer = h->nnz;
    for (int l = 0; l < n_levels; l++) {
        const pgb200_amg_level &s = lv[l];
        AmgLevel *L = new AmgLevel();
        h->amg.push_back(L);
        L->n = s.n; L->nnz = (size_t)s.nnz; L->n_finer = n_finer;
        CKR(L->rowptr.upload(s.rowptr, (size_t)s.n + 1, st)); CKR(L->colidx.upload(s.colidx, L->nnz, st));
        CKR(L->diag_pos.upload(s.diag_pos, s.n, st));
        CKR(L->gal_ptr.upload(s.gal_ptr, L->nnz + 1, st)); CKR(L->gal_idx.upload(s.gal_idx, nnz_finer, st));
        CKR(L->agg.upload(s.agg, n_finer, st)); CKR(L->mem_ptr.upload(s.mem_ptr, (size_t)s.n + 1, st)); CKR(L->mem_idx.upload(s.mem_idx, n_finer, st));
        CKR(L->vals.alloc(L->nnz * h->nK)); CKR(L->vals_dw.alloc(L->nnz * h->nK)); CKR(L->dinvw.alloc((size_t)s.n * h->nK));
        const size_t blk = (size_t)s.n * h->ld;
        CKR(L->R.alloc(blk)); CKR(L->X.alloc(blk)); CKR(L->Z.alloc(blk));
        if (s.n >= STREAM_MIN_ROWS) CKR(stream_upload(h, L->stream, s.n, s.rowptr, s.colidx));   // large coarse levels run on the streamed kernel too
        CK(cudaMemsetAsync(L->R.p, 0, blk * sizeof(double), st)); CK(cudaMemsetAsync(L->X.p, 0, blk * sizeof(double), st));
        CK(cudaMemsetAsync(L->Z.p, 0, blk * sizeof(double), st));
        n_finer = s.n; nnz_finer = L->nnz;
    }
    const size_t blk0 = (size_t)h->N * h->ld;
    if (!h->Z0.p) { CKR(h->Z0.alloc(blk0)); CKR(h->X0.alloc(blk0)); CKR(h->dinvw0.alloc((size_t)h->N * h->nK)); CKR(h->vals_dw0.alloc(h->nnz * h->nK)); CKR(h->gmax.alloc(h->nK));
        CK(cudaMemsetAsync(h->Z0.p, 0, blk0 * sizeof(double), st)); CK(cudaMemsetAsync(h->X0.p, 0, blk0 * sizeof(double), st)); }
    CK(cudaStreamSynchronize(st));
    h->have_vals = false;
    return 0;
}
int pgb200_ert_set_graph(pgb200_ert *h, int on) { if (!h) PGB_FAIL("null handle"); h->use_graph = on != 0; return 0; }
/* 0: Jacobi-PCG; 1: multilevel V-cycle preconditioner (needs a hierarchy); sweeps: Jacobi sweeps on the coarsest level */
int pgb200_ert_set_preconditioner(pgb200_ert *h, int multilevel, int coarse_sweeps) {
    if (!h) PGB_FAIL("null handle");
    h->use_amg = multilevel != 0; if (coarse_sweeps > 0) h->coarse_sweeps = coarse_sweeps;
    return 0;
}

int pgb200_ert_set_warm_start(pgb200_ert *h, int on) { if (!h) PGB_FAIL("null handle"); h->warm_start = on != 0; if (!on) h->x_warm_ok = false; return 0; }

int pgb200_ert_set_stream(pgb200_ert *h, void *stream) {
    if (!h) PGB_FAIL("null handle");
    // NULL / legacy default stream: keep the handle's own blocking stream (implicitly ordered with the legacy stream;
    // CUDA graphs cannot be captured on the legacy stream itself)
    h->st = stream ? (cudaStream_t)stream : h->own_st;
    return 0;
}

int pgb200_ert_set_solver(pgb200_ert *h, double rel_tol, int max_iter, int check_every) {
    if (!h) PGB_FAIL("null handle");
    if (!(rel_tol > 0.0) || max_iter <= 0 || check_every <= 0) PGB_FAIL("invalid solver parameters");
    h->tol = rel_tol; h->max_iter = max_iter; h->check_every = check_every;
    return 0;
}

int pgb200_ert_set_shard(pgb200_ert *h, int src_begin, int src_end, int row_begin, int row_end) {
    if (!h) PGB_FAIL("null handle");
    if (src_begin < 0 || src_end > h->nS || src_begin > src_end || row_begin < 0 || row_end > h->D || row_begin > row_end) PGB_FAIL("invalid shard");
    CK(cudaSetDevice(h->device));
    h->c0 = src_begin; h->c1 = src_end; h->pots_valid = false; h->shard_solved = false; h->x_warm_ok = false;
    {
        // narrow source shard inside one wavenumber group: the levels that run on the streamed kernels get the gather form
        const bool narrow = src_end > src_begin && src_begin / h->nE == (src_end - 1) / h->nE &&
                            ((src_end + 1) & ~1) - (src_begin & ~1) <= GATHER_MAX_W && !getenv("PGB200_NO_GATHER");
        if (narrow && h->use_panels && !h->stream.g_ok) {
            CKR(gather_upload(h, h->stream, h->N, h->rowptr.p, h->colidx.p, h->nnz));
            for (AmgLevel *L : h->amg) if (L->stream.ok) CKR(gather_upload(h, L->stream, L->n, L->rowptr.p, L->colidx.p, L->nnz));
            h->have_vals = false;                  // the packed fragments are filled by the next assembly
        }
    }
    if (row_begin != h->row0 || row_end != h->row1) { h->row0 = row_begin; h->row1 = row_end; CKR(build_jac_plan(h)); CKR(build_jac2_plan(h)); }
    return 0;
}

int pgb200_ert_set_kfac(pgb200_ert *h, const double *k) {
    if (!h || !k) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    h->h_kfac.assign(k, k + h->D);
    CK(cudaMemcpyAsync(h->kfac.p, k, sizeof(double) * h->D, cudaMemcpyHostToDevice, h->st));
    CKR(build_jac_plan(h));
    CKR(build_jac2_plan(h));
    return 0;
}

int pgb200_ert_clear_potentials(pgb200_ert *h) { if (!h) PGB_FAIL("null handle"); h->pots_valid = false; h->shard_solved = false; return 0; }
int pgb200_ert_mark_potentials_valid(pgb200_ert *h) {
    if (!h) PGB_FAIL("null handle");
    if (!h->shard_solved) PGB_FAIL("potentials cannot be marked valid: this handle's source shard has not been solved (run the forward solve first)");
    h->pots_valid = true;
    return 0;
}
int pgb200_ert_potentials_state(pgb200_ert *h) { if (!h) return -1; return (h->shard_solved ? 1 : 0) | (h->pots_valid ? 2 : 0); }

// ---- forward --------------------------------------------------------------------------
static int response_common(pgb200_ert *h, int n_in, double *rhoa_dev) {
    phase_begin(h, PH_MAP);
    CKR(map_model(h, h->model.p, n_in));
    CKR(forward_solve(h));
    CKR(pickup(h));
    CKR(finish_response(h, rhoa_dev));
    return 0;
}

int pgb200_ert_response_dev(pgb200_ert *h, const double *model_dev, int n_in, double *rhoa_dev) {
    if (!h || !model_dev || !rhoa_dev) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    CK(cudaMemcpyAsync(h->model.p, model_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, h->st));
    h->model_len = n_in;
    CKR(response_common(h, n_in, rhoa_dev));
    CKR(finish_timing(h));
    return 0;
}

int pgb200_ert_response(pgb200_ert *h, const double *model_host, int n_in, double *rhoa_host) {
    if (!h || !model_host || !rhoa_host) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    h->h_model.assign(model_host, model_host + n_in);
    CK(cudaMemcpyAsync(h->model.p, model_host, sizeof(double) * n_in, cudaMemcpyHostToDevice, h->st));
    h->model_len = n_in;
    CKR(response_common(h, n_in, h->rhoa.p));
    CK(cudaMemcpyAsync(rhoa_host, h->rhoa.p, sizeof(double) * h->D, cudaMemcpyDeviceToHost, h->st));
    CKR(finish_timing(h));
    return 0;
}

// mapERTModel (dcfemmodelling.cpp:1211-1218): cell resistivities for a model vector (per marker or per cell),
// including the prolongation into background cells; rho_cells_host[C]
int pgb200_ert_map_model(pgb200_ert *h, const double *model_host, int n_in, double *rho_cells_host) {
    if (!h || !model_host || !rho_cells_host) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    CK(cudaMemcpyAsync(h->model.p, model_host, sizeof(double) * n_in, cudaMemcpyHostToDevice, h->st));
    CKR(map_model(h, h->model.p, n_in));
    CK(cudaMemcpyAsync(rho_cells_host, h->rho.p, sizeof(double) * h->C, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->pots_valid = h->pots_valid;   // potentials of the last response stay valid, as in the reference
    return 0;
}

// multi-GPU staging: solve this shard's sources and leave the partial electrode matrix in HBM
int pgb200_ert_forward_dev(pgb200_ert *h, const double *model_dev, int n_in) {
    if (!h || !model_dev) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    CK(cudaMemcpyAsync(h->model.p, model_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, h->st));
    h->model_len = n_in;
    phase_begin(h, PH_MAP);
    CKR(map_model(h, h->model.p, n_in));
    CKR(forward_solve(h));
    CKR(pickup(h));
    CKR(finish_timing(h));
    return 0;
}
int pgb200_ert_pm_info(pgb200_ert *h, void **dev_ptr, int *n) { if (!h) PGB_FAIL("null handle"); *dev_ptr = h->pM.p; *n = h->nE * h->nE; return 0; }
int pgb200_ert_finish_response_dev(pgb200_ert *h, double *rhoa_dev) {
    if (!h || !rhoa_dev) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    CKR(finish_response(h, rhoa_dev));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
// contiguous [N x (c1-c0)] copy of this shard's potential columns and its inverse (NCCL all-gather)
__global__ void k_pack_cols(const double *__restrict__ U, size_t ld, int N, int c0, int c1, double *__restrict__ buf, int unpack, double *__restrict__ Uw) {
    const int w = c1 - c0;
    const int c = blockIdx.y * blockDim.x + threadIdx.x, row = blockIdx.x * blockDim.y + threadIdx.y;
    if (c >= w || row >= N) return;
    if (unpack) Uw[(size_t)row * ld + c0 + c] = buf[(size_t)row * w + c];
    else buf[(size_t)row * w + c] = U[(size_t)row * ld + c0 + c];
}
int pgb200_ert_pack_potentials(pgb200_ert *h, int c0, int c1, double *buf_dev, int unpack) {
    if (!h || !buf_dev) PGB_FAIL("null argument");
    if (c0 < 0 || c1 > h->nS || c0 >= c1) return 0;
    CK(cudaSetDevice(h->device));
    dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(c1 - c0, 32));
    k_pack_cols<<<g, b, 0, h->st>>>(h->U.p, h->ld, h->N, c0, c1, buf_dev, unpack, h->U.p); LAUNCH(h);
    CK(cudaGetLastError());
    return 0;
}

int pgb200_ert_fill_matrix(pgb200_ert *h, const double *a_cells_host, const double *b_cells_host, double *vals_host) {
    if (!h || !vals_host) PGB_FAIL("null argument");
    if (!a_cells_host && !b_cells_host) PGB_FAIL("fill_matrix: neither a stiffness nor a mass coefficient given");
    CK(cudaSetDevice(h->device));
    DevBuf<double> a, b, out;
    if (a_cells_host) CKR(a.upload(a_cells_host, (size_t)h->C, h->st));
    if (b_cells_host) CKR(b.upload(b_cells_host, (size_t)h->C, h->st));
    CKR(out.alloc(h->nnz));
    switch (h->elem) {
        case TRI3:  CKR(launch_assemble_generic<TRI3>(h, a.p, b.p, out.p)); break;
        case TRI6:  CKR(launch_assemble_generic<TRI6>(h, a.p, b.p, out.p)); break;
        case TET4:  CKR(launch_assemble_generic<TET4>(h, a.p, b.p, out.p)); break;
        case TET10: CKR(launch_assemble_generic<TET10>(h, a.p, b.p, out.p)); break;
        default: PGB_FAIL("unknown element type");
    }
    CK(cudaMemcpyAsync(vals_host, out.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

int pgb200_ert_set_primary_dev(pgb200_ert *h, const double *src_dev, long long src_ld, const int *row_map_host) {
    if (!h || !src_dev || !row_map_host) PGB_FAIL("null argument");
    if (src_ld < h->nS) PGB_FAIL("primary potentials: leading dimension smaller than the number of sources");
    CK(cudaSetDevice(h->device));
    DevBuf<int> map;
    CKR(map.upload(row_map_host, (size_t)h->N, h->st));
    dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(h->nS, 32));
    k_gather_rows<<<g, b, 0, h->st>>>(src_dev, (size_t)src_ld, map.p, h->N, h->nS, h->ld, h->prim.p); LAUNCH(h);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->st));
    h->prim_set = true; h->pots_valid = false; h->shard_solved = false; h->jac_valid = false; h->x_warm_ok = false;
    return 0;
}

// ---- Jacobian -------------------------------------------------------------------------
static int create_jacobian_common(pgb200_ert *h, int n_in) {
    if (h->ref_last) {
        // the reference fails here too: nE - 1 current patterns, createSensitivityCol wants nE rows (bertJacobian.cpp:283-291)
        char buf[200];
        snprintf(buf, sizeof buf, "potential matrix rowsize to small. %d < %d (the last electrode is the current reference of this pure-Neumann "
                                  "domain; add a reference-electrode node, marker -999)", h->nS - h->nK, h->nS);
        PGB_FAIL(buf);
    }
    const double *rho_col = (n_in == h->M) ? h->model.p : nullptr;      // scaling only if len(model) == J.cols (:1377)
    if (!h->pots_valid) {
        // prepareJacobianT_ (:1246-1309): no potentials yet -> solve, analytically for a homogeneous model
        const bool hetero = h->topography || host_stddev(h->h_model) > 1e-12 * 1e5;     // setAnalytical(!(topography || het)) :1272
        phase_begin(h, PH_MAP);
        CKR(map_model(h, h->model.p, n_in));
        if (!hetero) { CKR(analytic_pots(h, h->h_model[0])); }
        else { CKR(forward_solve(h)); }
    }
    CKR(jacobian(h, rho_col));
    return 0;
}

int pgb200_ert_create_jacobian(pgb200_ert *h, const double *model_host, int n_in) {
    if (!h || !model_host) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    h->h_model.assign(model_host, model_host + n_in);
    CK(cudaMemcpyAsync(h->model.p, model_host, sizeof(double) * n_in, cudaMemcpyHostToDevice, h->st));
    h->model_len = n_in;
    CKR(create_jacobian_common(h, n_in));
    CKR(finish_timing(h));
    return 0;
}

int pgb200_ert_create_jacobian_dev(pgb200_ert *h, const double *model_dev, int n_in) {
    if (!h || !model_dev) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    CK(cudaMemcpyAsync(h->model.p, model_dev, sizeof(double) * n_in, cudaMemcpyDeviceToDevice, h->st));
    h->model_len = n_in;
    if (!h->pots_valid) {
        h->h_model.resize(n_in);
        CK(cudaMemcpyAsync(h->h_model.data(), model_dev, sizeof(double) * n_in, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
    }
    CKR(create_jacobian_common(h, n_in));
    CKR(finish_timing(h));
    return 0;
}

int pgb200_ert_jacobian_info(pgb200_ert *h, void **dev_ptr, int *rows, int *cols, long long *ld) {
    if (!h) PGB_FAIL("null handle");
    if (!h->jac_valid) PGB_FAIL("no Jacobian: call createJacobian first");
    if (dev_ptr) *dev_ptr = h->Jt.p;
    if (rows) *rows = h->j_rows;
    if (cols) *cols = h->M;
    if (ld) *ld = (long long)h->ldJ;
    return 0;
}

int pgb200_ert_jacobian_copy(pgb200_ert *h, double *j_host) {
    if (!h || !j_host) PGB_FAIL("null argument");
    if (!h->jac_valid) PGB_FAIL("no Jacobian: call createJacobian first");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->j_rows * h->M;
    if (n == 0) return 0;
    if (h->tmp.n < n) CKR(h->tmp.alloc(n));
    dim3 b(32, 8), g(cdiv(h->j_rows, 32), cdiv(h->M, 32));
    k_jac_to_rowmajor<<<g, b, 0, h->st>>>(h->Jt.p, h->ldJ, h->j_rows, h->M, h->tmp.p); LAUNCH(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(j_host, h->tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

// stage optional host scaling vectors next to x: xin = [x | left | right] (each max(M, D) doubles)
static int jac_stage(pgb200_ert *h, const double *x, int nx, const double *left, const double *right, const double **dl, const double **dr) {
    const size_t blk = (size_t)std::max(h->M, h->D);
    if (nx > 0) CK(cudaMemcpyAsync(h->xin.p, x, sizeof(double) * nx, cudaMemcpyHostToDevice, h->st));
    *dl = nullptr; *dr = nullptr;
    if (left && h->j_rows) { CK(cudaMemcpyAsync(h->xin.p + blk, left, sizeof(double) * h->j_rows, cudaMemcpyHostToDevice, h->st)); *dl = h->xin.p + blk; }
    if (right) { CK(cudaMemcpyAsync(h->xin.p + 2 * blk, right, sizeof(double) * h->M, cudaMemcpyHostToDevice, h->st)); *dr = h->xin.p + 2 * blk; }
    return 0;
}

int pgb200_ert_jacobian_mult_lr(pgb200_ert *h, const double *left_host, const double *right_host, const double *x_host, double *y_host) {
    if (!h || !x_host || !y_host) PGB_FAIL("null argument");
    if (!h->jac_valid) PGB_FAIL("no Jacobian: call createJacobian first");
    CK(cudaSetDevice(h->device));
    const double *dl, *dr;
    CKR(jac_stage(h, x_host, h->M, left_host, right_host, &dl, &dr));
    CK(cudaMemsetAsync(h->yout.p, 0, sizeof(double) * std::max(1, h->j_rows), h->st));
    if (h->j_rows) { k_jac_mult<<<dim3(cdiv(h->j_rows, 128), cdiv(h->M, 256)), 128, 0, h->st>>>(h->Jt.p, h->ldJ, h->j_rows, h->M, h->xin.p, dl, dr, h->yout.p); LAUNCH(h); }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, h->yout.p, sizeof(double) * h->j_rows, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
int pgb200_ert_jacobian_mult(pgb200_ert *h, const double *x_host, double *y_host) {
    return pgb200_ert_jacobian_mult_lr(h, nullptr, nullptr, x_host, y_host);
}

int pgb200_ert_jacobian_tmult_lr(pgb200_ert *h, const double *left_host, const double *right_host, const double *x_host, double *y_host) {
    if (!h || !x_host || !y_host) PGB_FAIL("null argument");
    if (!h->jac_valid) PGB_FAIL("no Jacobian: call createJacobian first");
    CK(cudaSetDevice(h->device));
    const double *dl, *dr;
    CKR(jac_stage(h, x_host, h->j_rows, left_host, right_host, &dl, &dr));
    k_jac_tmult<0><<<cdiv(h->M, 8), 256, 0, h->st>>>(h->Jt.p, h->ldJ, h->j_rows, h->M, h->xin.p, dl, dr, h->yout.p); LAUNCH(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, h->yout.p, sizeof(double) * h->M, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
int pgb200_ert_jacobian_tmult(pgb200_ert *h, const double *x_host, double *y_host) {
    return pgb200_ert_jacobian_tmult_lr(h, nullptr, nullptr, x_host, y_host);
}

int pgb200_ert_coverage_trans(pgb200_ert *h, const double *dd_host, const double *mm_host, double *cov_host) {
    if (!h || !dd_host || !cov_host) PGB_FAIL("null argument");
    if (!h->jac_valid) PGB_FAIL("no Jacobian: call createJacobian first");
    CK(cudaSetDevice(h->device));
    const double *dl, *dr;
    CKR(jac_stage(h, dd_host, h->j_rows, nullptr, mm_host, &dl, &dr));
    if (mm_host) k_jac_tmult<1><<<cdiv(h->M, 8), 256, 0, h->st>>>(h->Jt.p, h->ldJ, h->j_rows, h->M, h->xin.p, nullptr, dr, h->yout.p);
    else k_jac_tmult<2><<<cdiv(h->M, 8), 256, 0, h->st>>>(h->Jt.p, h->ldJ, h->j_rows, h->M, h->xin.p, nullptr, nullptr, h->yout.p);
    LAUNCH(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cov_host, h->yout.p, sizeof(double) * h->M, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

int pgb200_ert_potentials_info(pgb200_ert *h, void **dev_ptr, int *n_nodes, int *n_src, long long *ld) {
    if (!h) PGB_FAIL("null handle");
    if (dev_ptr) *dev_ptr = h->U.p;
    if (n_nodes) *n_nodes = h->N;
    if (n_src) *n_src = h->nS;
    if (ld) *ld = (long long)h->ld;
    return 0;
}

// ---- introspection ----------------------------------------------------------------------
long long pgb200_ert_get(pgb200_ert *h, const char *what, double *out, long long cap) {
    if (!h || !what) { g_err = "null argument"; return -1; }
    if (cudaSetDevice(h->device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return -1; }
    const std::string w(what);
    const double *src = nullptr; long long n = 0; bool transposed = false;
    if (w == "vals") { src = h->vals.p; n = (long long)h->nnz * h->nK; }
    else if (w == "vals1") { src = h->vals1.p; n = h->vals1.p ? (long long)h->nnz * h->nK : 0; }
    else if (w == "rho") { src = h->rho.p; n = h->C; }
    else if (w == "rho_src") { src = h->rho_src.p; n = h->nE; }
    else if (w == "pm") { src = h->pM.p; n = (long long)h->nE * h->nE; }
    else if (w == "resp") { src = h->resp.p; n = h->D; }
    else if (w == "resp_rez") { src = h->resp_rez.p; n = h->D; }
    else if (w == "rel_res") { n = h->nS; }
    else if (w == "prim" || w == "pots" || w == "rhs" || w == "sec") { n = (long long)h->nS * h->N; transposed = true; }
    else if (w == "solutions") { n = (long long)h->nE * h->N; }
    else { g_err = "unknown quantity: " + w; return -1; }
    if (!out) return n;
    if (cap < n) { g_err = "output buffer too small"; return -1; }
    if (n == 0) return 0;
    cudaError_t e = cudaSuccess;
    if (w == "rel_res") {
        // per-column relative residuals recorded by the convergence checks of the last solve
        for (int c = 0; c < h->nS; c++) out[c] = c < (int)h->col_relres.size() ? h->col_relres[c] : 0.0;
    } else if (transposed || w == "solutions") {
        if (h->tmp.n < (size_t)n) { if (h->tmp.alloc((size_t)n)) return -1; }
        if (w == "solutions") {
            dim3 b(32, 8), g(cdiv(h->N, 8), cdiv(h->nE, 32));
            k_ksum<<<g, b, 0, h->st>>>(h->U.p, h->ld, h->kw.p, h->nK, h->nE, h->N, h->tmp.p);
        } else {
            const double *Usrc = w == "prim" ? h->prim.p : w == "pots" ? h->U.p : w == "rhs" ? h->B.p : h->X.p;
            dim3 b(32, 8), g(cdiv(h->nS, 32), cdiv(h->N, 32));
            k_transpose_out<<<g, b, 0, h->st>>>(Usrc, h->ld, h->N, h->nS, h->tmp.p);
        }
        cudaStreamSynchronize(h->st);
        e = cudaMemcpy(out, h->tmp.p, sizeof(double) * n, cudaMemcpyDeviceToHost);
    } else {
        cudaStreamSynchronize(h->st);
        e = cudaMemcpy(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) { g_err = std::string("copy failed: ") + cudaGetErrorString(e); return -1; }
    return n;
}

int pgb200_ert_stats(pgb200_ert *h, double *s, int n) {
    if (!h || !s) PGB_FAIL("null argument");
    double v[17] = {(double)h->last_iters, h->last_relres, (double)h->launches, h->ph_ms[PH_MAP], h->ph_ms[PH_ASM], h->ph_ms[PH_RHS],
                    h->ph_ms[PH_SOLVE], h->ph_ms[PH_EPI], h->ph_ms[PH_JAC], (double)h->spmm_timed, h->spmm_ms, h->jac_ms,
                    (double)h->jac_launches, (double)h->total_iters, (double)h->solves, h->spmm_bytes, (double)h->warm_used};
    for (int i = 0; i < n && i < 17; i++) s[i] = v[i];
    return 0;
}
int pgb200_ert_path_info(pgb200_ert *h, int *out, int n) {
    if (!h || !out) PGB_FAIL("null argument");
    int mt = 0, res = 1;
    for (auto &c : h->chunks) { mt = std::max(mt, c.mt); res = res && c.resolved; }
    int nch = (int)h->chunks.size();
    if (h->jac_v2 && h->j2_ok) { mt = 0; nch = (int)h->chunks2.size(); for (auto &c : h->chunks2) mt = std::max(mt, c.mt); res = 10 * h->j2_terms + h->j2_shared; }
    int sl = 0;
    for (AmgLevel *L : h->amg) sl += (h->use_panels && L->stream.ok) ? 1 : 0;
    const int v[10] = {h->pi_panel_nc, h->pi_tiles, h->pi_two_k, h->pi_graph_launches, nch, mt, res, (int)h->amg.size(), h->pi_slots, sl};
    for (int i = 0; i < n && i < 10; i++) out[i] = v[i];
    return 0;
}
int pgb200_ert_reset_stats(pgb200_ert *h) {
    if (!h) PGB_FAIL("null handle");
    h->launches = 0; h->spmm_ms = 0.0; h->spmm_bytes = 0.0; h->warm_used = 0; h->spmm_timed = 0; h->jac_ms = 0.0; h->n_pev = 0; h->jac_launches = 0; h->total_iters = 0; h->solves = 0;
    for (int p = 0; p < PH_COUNT; p++) h->ph_ms[p] = 0.f;
    return 0;
}
// Measurement aid: `reps` back-to-back launches of the fine-level streamed SpMM in epilogue role `role` (0 SpMM + p.Ap,
// 1 post-smoothing + r.z, 2 residual) on the current matrix and the PCG work vectors (their contents are overwritten);
// ms_per_launch = CUDA-event time / reps.  The matrix must have been assembled (any response() call).
int pgb200_ert_bench_spmm(pgb200_ert *h, int role, int reps, double *ms_per_launch) {
    if (!h || !ms_per_launch || reps < 1) { g_err = "null argument"; return 1; }
    if (!h->have_vals || !panel_path_ok(h)) PGB_FAIL("bench_spmm: no assembled matrix / streamed path off");
    CK(cudaSetDevice(h->device));
    const int c0 = h->c0, c1 = h->c1;
    double *S = h->scal.p;
    if (const char *e = getenv("PGB200_MMA_DBG_LATE")) h->mma_dbg = atoi(e);     // ablation runs (ab/bench_spmm.py): results are wrong
    if (h->mma_dbg & 8) { CKR(h->mma_dbg_buf.alloc(8 * 4096)); CK(cudaMemsetAsync(h->mma_dbg_buf.p, 0, 8 * 4096 * sizeof(long long), h->st)); }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = -2; i < reps; i++) {
        if (i == 0) CK(cudaEventRecord(e0, h->st));
        PanelExtra ex{};
        if (role == 1) { ex.R = h->R.p; ex.dinvw = h->dinvw0.p ? h->dinvw0.p : h->dinv.p; ex.n = h->N; CKR(launch_stream<EPI_POST>(h, h->stream, 0, h->P.p, h->AP.p, c0, c1, S + 3 * h->ld, ex)); }
        else if (role == 2) CKR(launch_stream<EPI_RESIDUAL>(h, h->stream, 0, h->P.p, h->AP.p, c0, c1, nullptr, ex));
        else CKR(launch_stream<EPI_SPMM>(h, h->stream, 0, h->P.p, h->AP.p, c0, c1, S + 3 * h->ld, ex));
    }
    CK(cudaEventRecord(e1, h->st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (h->mma_dbg & 8) {
        std::vector<long long> v(8 * 4096);
        CK(cudaMemcpy(v.data(), h->mma_dbg_buf.p, v.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const long long t0 = v[0];
        for (int st = 0; st < 4096 && v[8 * st + 0]; st++) {
            fprintf(stderr, "stage %3d  prod: wait %6lld issued %6lld (+%lld)   cons: wait %6lld got %6lld loop %6lld done %6lld\n", st, v[8 * st] - t0, v[8 * st + 1] - t0,
                    v[8 * st + 2] - v[8 * st + 1], v[8 * st + 3] - t0, v[8 * st + 4] - t0, v[8 * st + 5] - t0, v[8 * st + 6] - t0);
        }
    }
    h->mma_dbg = 0;
    *ms_per_launch = (double)ms / reps;
    return 0;
}

// ---- complex resistivity -------------------------------------------------------------------------------
// The handle must come from a plan that lists every electrode twice (sensors = [sensors | sensors], total-field scheme, sr = 0):
// within a wavenumber group the first nE/2 source columns carry real parts, the others imaginary parts.
int pgb200_ert_set_complex(pgb200_ert *h, int on) {
    if (!h) PGB_FAIL("null handle");
    if (on && (h->sr || (h->nE & 1))) PGB_FAIL("complex resistivity needs a total-field handle (sr = 0) built with every electrode listed twice");
    h->is_complex = on != 0;
    h->pots_valid = false; h->shard_solved = false; h->jac_valid = false;
    return 0;
}

namespace {
// z = D^-1 r on all columns (levels without a hierarchy: Jacobi)
__global__ void k_cplx_jacobi(const double *__restrict__ R, const double *__restrict__ dinv, double *__restrict__ Z, int N, int nE, int nS, size_t ld) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * nS) return;
    const int row = (int)(i / nS), c = (int)(i % nS);
    Z[(size_t)row * ld + c] = dinv[(size_t)(c / nE) * N + row] * R[(size_t)row * ld + c];
}

// model_dev: [re (n_in) | im (n_in)];  assemble S_r, S_i, solve all sources by COCG, potentials -> U
int complex_forward(pgb200_ert *h, const double *model_dev, int n_in) {
    if (!h->is_complex) PGB_FAIL("pgb200_ert_set_complex(h, 1) first");
    if (h->c0 != 0 || h->c1 != h->nS) PGB_FAIL("complex resistivity: source shards are not supported");
    const int N = h->N, nE = h->nE, nEc = nE / 2, nS = h->nS, nSc = nS / 2;
    const size_t ld = h->ld;
    cudaStream_t st = h->st;
    if (!h->c_rho_r.p) {
        CKR(h->c_rho_r.alloc((size_t)h->C)); CKR(h->c_y2.alloc((size_t)N * ld)); CKR(h->c_scal.alloc(8 * ld));
        if (!h->vals1.p) CKR(h->vals1.alloc(h->nnz * h->nK));
    }
    // mapERTModel(CVector) (:1199-1208): real and imaginary parts are mapped and prolongated separately
    phase_begin(h, PH_MAP);
    CKR(map_model(h, model_dev, n_in));
    CK(cudaMemcpyAsync(h->c_rho_r.p, h->rho.p, sizeof(double) * h->C, cudaMemcpyDeviceToDevice, st));
    CKR(map_model(h, model_dev + n_in, n_in));
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(int) * 4, st));              // the positivity check of the real path does not apply
    k_cplx_sigma<<<cdiv(h->C, 256), 256, 0, st>>>(h->c_rho_r.p, h->rho.p, h->C, h->c_rho_r.p, h->rho.p, h->flags.p); LAUNCH(h);
    // S_r and S_i (:235-242): two real assembly passes with 1/sigma_r and 1/sigma_i as "resistivities"
    phase_begin(h, PH_ASM);
    CKR(assemble(h, h->c_rho_r.p, h->vals.p));
    CKR(assemble(h, h->rho.p, h->vals1.p));
    if (h->n_dir_nodes) { k_set_slots<<<cdiv(h->n_dir_nodes, 128), 128, 0, st>>>(h->dir_diag.p, h->n_dir_nodes, h->nK, h->nnz, 0.0, h->vals1.p); LAUNCH(h); }
    k_count_singular<<<cdiv(N, 256), 256, 0, st>>>(h->diag_pos.p, N, h->nK, h->nnz, h->vals.p, h->flags.p + 1); LAUNCH(h);
    k_inv_diag<<<cdiv(N, 256), 256, 0, st>>>(h->diag_pos.p, N, h->nK, h->nnz, h->vals.p, h->dinv.p); LAUNCH(h);
    h->have_vals = true;
    CKR(stream_pack(h, h->stream, h->vals.p, 0));
    if (h->use_amg) CKR(amg_setup_values(h));
    phase_begin(h, PH_RHS);
    CK(cudaMemsetAsync(h->B.p, 0, sizeof(double) * N * ld, st));
    k_delta_rhs<<<cdiv(nS, 128), 128, 0, st>>>(h->pick_ptr.p, h->pick_idx.p, h->pick_w.p, nE, 0, nS, ld, h->B.p); LAUNCH(h);
    if (h->ref_node >= 0) { k_ref_rhs<<<cdiv(nS, 128), 128, 0, st>>>(h->pick_ptr.p, h->pick_idx.p, h->pick_w.p, nE, 0, nS, ld, h->ref_node, 0, h->B.p); LAUNCH(h); }
    if (h->ref_last) PGB_FAIL("complex resistivity on a pure-Neumann domain needs a reference-electrode node (marker -999)");
    k_cplx_zero_imag_cols<<<cdiv((long long)N * nS, 256), 256, 0, st>>>(h->B.p, N, nEc, nS, ld); LAUNCH(h);
    int hf[4];
    CK(cudaMemcpyAsync(hf, h->flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hf[0]) PGB_FAIL("complex response for abs model with negative or zero resistivity is not defined");
    if (hf[1]) PGB_FAIL("stiffness matrix has rows with diagonal < 1e-12 (the reference would force them to homogeneous Dirichlet); unsupported model");
    // ---- COCG ----
    phase_begin(h, PH_SOLVE);
    const bool amg = h->use_amg && !h->amg.empty();
    double *S = h->c_scal.p;
    auto row = [&](int i) { return S + (size_t)i * ld; };      // 0,1 rho | 2,3 rho' | 4,5 p^T q | 6 |r|^2 | 7 |b|^2
    const FlatCfg fd = flat_cfg(N, 0, nSc, FLAT_MAX_GX), fc = flat_cfg(N, 0, nSc);
    CK(cudaMemsetAsync(S, 0, sizeof(double) * 8 * ld, st));
    CK(cudaMemsetAsync(h->X.p, 0, sizeof(double) * N * ld, st));
    CK(cudaMemcpyAsync(h->R.p, h->B.p, sizeof(double) * N * ld, cudaMemcpyDeviceToDevice, st));
    auto precond = [&]() -> int {
        if (amg) return amg_vcycle(h, 0, nS, nullptr);
        k_cplx_jacobi<<<cdiv((long long)N * nS, 256), 256, 0, st>>>(h->R.p, h->dinv.p, h->Z0.p, N, nE, nS, ld); LAUNCH(h);
        return 0;
    };
    if (!h->Z0.p) CKR(h->Z0.alloc((size_t)N * ld));
    k_cplx_dot<true><<<fd.grid, FLAT_T, 0, st>>>(h->B.p, nullptr, N, nEc, nSc, ld, fd.cw, dot_out(h, row(7), nullptr)); LAUNCH(h);
    CKR(precond());
    CK(cudaMemcpyAsync(h->P.p, h->Z0.p, sizeof(double) * N * ld, cudaMemcpyDeviceToDevice, st));
    k_cplx_dot<false><<<fd.grid, FLAT_T, 0, st>>>(h->R.p, h->Z0.p, N, nEc, nSc, ld, fd.cw, dot_out(h, row(0), row(1))); LAUNCH(h);
    CKR(ensure_pinned(h, 2 * ld));
    const double tol2 = h->tol * h->tol;
    int it = 0; bool converged = false; int cur = 0;
    while (it < h->max_iter && !converged) {
        double *rho = row(2 * cur), *rho_new = row(2 * (1 - cur));
        // q = (S_r + i S_i) p
        if (panel_path_ok(h)) { PanelExtra ex{}; CKR(launch_stream<EPI_SPMM>(h, h->stream, 0, h->P.p, h->AP.p, 0, nS, nullptr, ex)); }
        else CKR((launch_spmm<0, false>(h, h->vals.p, nullptr, nullptr, h->P.p, h->AP.p, 0, nS, nullptr)));
        CKR((launch_spmm<0, false>(h, h->vals1.p, nullptr, nullptr, h->P.p, h->c_y2.p, 0, nS, nullptr)));
        k_cplx_combine<<<fd.grid, FLAT_T, 0, st>>>(h->AP.p, h->c_y2.p, h->P.p, N, nEc, nSc, ld, fd.cw, dot_out(h, row(4), row(5))); LAUNCH(h);
        k_cplx_update_xr<<<fd.grid, FLAT_T, 0, st>>>(h->P.p, h->AP.p, h->X.p, h->R.p, N, nEc, nSc, ld, rho, row(4), ld, fd.cw, dot_out(h, nullptr, row(6))); LAUNCH(h);
        CKR(precond());
        k_cplx_dot<false><<<fd.grid, FLAT_T, 0, st>>>(h->R.p, h->Z0.p, N, nEc, nSc, ld, fd.cw, dot_out(h, rho_new, rho_new + ld)); LAUNCH(h);
        k_cplx_update_p<<<fc.grid, FLAT_T, 0, st>>>(h->Z0.p, h->P.p, N, nEc, nSc, ld, rho, rho_new, ld, row(6), row(7), tol2, fc.cw); LAUNCH(h);
        cur = 1 - cur;
        it++;
        if (it % 6 == 0 || it >= h->max_iter) {
            CK(cudaMemcpyAsync(h->h_pinned, row(6), sizeof(double) * 2 * ld, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            double worst = 0.0;
            for (int j = 0; j < nSc; j++) {
                const double rr = h->h_pinned[j], bb = h->h_pinned[ld + j];
                double rel = bb > 0.0 ? std::sqrt(rr / bb) : (rr > 0.0 ? INFINITY : 0.0);
                if (!(rr == rr)) rel = INFINITY;
                worst = std::max(worst, rel);
            }
            h->last_relres = worst;
            converged = worst <= h->tol;
        }
    }
    CK(cudaGetLastError());
    h->last_iters = it; h->total_iters += it; h->solves++;
    h->x_warm_ok = false;
    if (!converged) {
        char buf[256];
        snprintf(buf, sizeof buf, "complex block-COCG did not reach rel. residual %.1e in %d iterations (worst column %.3e)", h->tol, it, h->last_relres);
        PGB_FAIL(buf);
    }
    phase_begin(h, PH_EPI);
    {
        dim3 b(32, 8), g(cdiv(N, 8), cdiv(nS, 32));
        k_finalize_pots<<<g, b, 0, st>>>(h->X.p, nullptr, nullptr, 0.0, N, nE, 0, nS, ld, h->U.p); LAUNCH(h);
    }
    h->shard_solved = true; h->pots_valid = true;
    CKR(pickup(h));              // electrode matrix of the doubled layout: pM[i][j], i < nE/2 real part, i >= nE/2 imaginary part
    return 0;
}
}  // namespace

// model_host = [re (n_in) | im (n_in)] (n_in = model cells or cells).  Solves; the electrode matrix (pgb200_ert_pm_info) and the
// potentials (pgb200_ert_potentials_info) are then those of the doubled layout.
int pgb200_ert_complex_forward(pgb200_ert *h, const double *model_host, int n_in) {
    if (!h || !model_host) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    if (h->c_model.n < (size_t)2 * n_in) CKR(h->c_model.alloc((size_t)2 * std::max(h->M, h->C)));
    CK(cudaMemcpyAsync(h->c_model.p, model_host, sizeof(double) * 2 * n_in, cudaMemcpyHostToDevice, h->st));
    h->model_len = n_in;
    CKR(complex_forward(h, h->c_model.p, n_in));
    CKR(finish_timing(h));
    return 0;
}

// complex Jacobian: the handle's scheme holds the four real blocks (4 D rows, see k_cplx_jacobian); j_host receives D x M
// complex values, row-major, interleaved (re, im); kfac_host[D]; scaling by k_d / m_j^2 when n_in == M (:1377, :1420-1428)
int pgb200_ert_complex_jacobian(pgb200_ert *h, const double *model_host, int n_in, const double *kfac_host, double *j_host) {
    if (!h || !model_host || !kfac_host || !j_host) PGB_FAIL("null argument");
    CK(cudaSetDevice(h->device));
    if (!h->is_complex) PGB_FAIL("pgb200_ert_set_complex(h, 1) first");
    if (h->D % 4) PGB_FAIL("complex Jacobian: the scheme must hold four blocks of rows");
    if (n_in != h->M && n_in != h->C) PGB_FAIL("model length must equal the number of model cells (max marker + 1) or the cell count");
    const int D = h->D / 4, M = h->M;
    if (h->c_model.n < (size_t)2 * n_in) CKR(h->c_model.alloc((size_t)2 * std::max(h->M, h->C)));
    CK(cudaMemcpyAsync(h->c_model.p, model_host, sizeof(double) * 2 * n_in, cudaMemcpyHostToDevice, h->st));
    if (!h->pots_valid) CKR(complex_forward(h, h->c_model.p, n_in));       // prepareJacobianT_: numeric potentials of this model
    CKR(jacobian(h, nullptr));                                            // four real blocks, unscaled
    if (h->c_out.n < (size_t)2 * D * M + D) CKR(h->c_out.alloc((size_t)2 * D * M + D));
    double *kf = h->c_out.p + (size_t)2 * D * M;
    CK(cudaMemcpyAsync(kf, kfac_host, sizeof(double) * D, cudaMemcpyHostToDevice, h->st));
    dim3 b(32, 8), g(cdiv(M, 32), cdiv(D, 32));
    k_cplx_jacobian<<<g, b, 0, h->st>>>(h->Jt.p, h->ldJ, D, M, kf, h->c_model.p, h->c_model.p + n_in, n_in == M ? 1 : 0, h->c_out.p); LAUNCH(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(j_host, h->c_out.p, sizeof(double) * 2 * D * M, cudaMemcpyDeviceToHost, h->st));
    CKR(finish_timing(h));
    return 0;
}

int pgb200_ert_set_spmm_variant(pgb200_ert *h, int panel_staged) {
    if (!h) PGB_FAIL("null handle");
    h->use_panels = panel_staged != 0;
    return 0;
}
int pgb200_ert_set_profile(pgb200_ert *h, int on) {
    if (!h) PGB_FAIL("null handle");
    CK(cudaSetDevice(h->device));
    h->prof = on;
    if (on && h->pev.empty()) { h->pev.resize(4096); for (auto &e : h->pev) CK(cudaEventCreate(&e)); }
    h->trace = (on == 2);
    if (h->trace && h->tev.empty()) { h->tev.resize(60000); h->tline.assign(60000, 0); for (auto &e : h->tev) CK(cudaEventCreate(&e)); }
    if (h->trace) { h->n_tev = 0; note_launch(h, 0); h->launches--; }      // start marker
    return 0;
}

// trace of the launches since set_profile(h, 2): source line of each launch (of csrc/pgb200_ert.cu) and the time from
// the previous launch's completion to its own, in ms.  Returns the number of entries (at most cap).
int pgb200_ert_get_trace(pgb200_ert *h, int *lines, float *ms, int cap) {
    if (!h) { g_err = "null handle"; return -1; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    int n = 0;
    for (int i = 1; i < h->n_tev && n < cap; i++, n++) {
        float t = 0.f;
        cudaEventElapsedTime(&t, h->tev[i - 1], h->tev[i]);
        if (lines) lines[n] = h->tline[i];          // source line * 256 + role * 16 + multilevel level
        if (ms) ms[n] = t;
    }
    return n;
}

int pgb200_spmm(const int *rowptr, const int *colidx, const double *vals, long long nnz, const double *X, double *Y,
                int n_rows, int n_elec, int n_k, long long ld, void *stream) {
    const int ncols = n_elec * n_k;
    dim3 block(SPMM_TX, SPMM_TY), grid(cdiv(n_rows, SPMM_ROWS), cdiv(ncols, SPMM_TX * 4));
    k_spmm<4, 0, false><<<grid, block, 0, (cudaStream_t)stream>>>(rowptr, colidx, vals, nullptr, nullptr, (size_t)nnz, X, Y, n_rows,
                                                                  n_elec, 0, ncols, (size_t)ld, nullptr);
    CK(cudaGetLastError());
    return 0;
}

} // extern "C"
