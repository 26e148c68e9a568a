// CUDA kernels of the B200 ERT forward + Jacobian path (FP64, sm_100a, no tensor cores:
// per-cell 3x3..10x10 contractions and CSR work, HBM/L2-bound -- see DESIGN.md).
//
// Layout conventions
//   * block vectors (potentials, PCG state) are node-major  X[node * ld + s],
//     s = electrode + nE * kIdx  (the reference's subSolutions_ row index,
//     dcfemmodelling.cpp:1681) -- one node's values for all sources are contiguous, so the
//     SpMM gathers whole rows and the Jacobian gathers a cell's nodes with coalesced loads.
//   * CSR values are stored per wavenumber: vals[kIdx * nnz + slot].
//   * J is written column-major Jt[col * ldJ + d].
#pragma once
#include "ert_device.cuh"
#include <stdint.h>

namespace pgb {

// ---------------------------------------------------------------------------------
// model mapping: rho_cell = model[marker]   (modellingbase.cpp:424-440) or cell-wise copy
// ---------------------------------------------------------------------------------
__global__ void k_map_model(const double *__restrict__ model, int n_model_in, const int *__restrict__ marker,
                            int C, double *__restrict__ rho) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (n_model_in == C) { rho[c] = model[c]; return; }
    int m = marker[c];
    rho[c] = (m >= 0 && m < n_model_in) ? model[m] : 0.0;
}
// one prolongation level (mesh.cpp:2276-2306): weighted mean of already filled neighbours
__global__ void k_prolong_level(const int *__restrict__ cells, const int *__restrict__ nb, const double *__restrict__ w,
                                int n, int nf, double *__restrict__ rho) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double acc = 0.0;
    for (int f = 0; f < nf; f++) {
        double wf = w[(size_t)t * nf + f];
        if (wf != 0.0) acc += wf * rho[nb[(size_t)t * nf + f]];
    }
    rho[cells[t]] = acc;
}
__global__ void k_check_model(const double *__restrict__ v, int n, int *flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(v[i] >= 1e-12)) atomicOr(flag, 1);   // min(model) < TOLERANCE -> error (:1133)
}
// std-dev test for the analytic branch of createJacobian (:1272-1274) on the host-visible scalars
__global__ void k_mean_var(const double *__restrict__ v, int n, double *out /* sum, sumsq */) {
    __shared__ double s0[256], s1[256];
    double a = 0.0, b = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a += v[i]; b += v[i] * v[i]; }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { atomicAdd(out, s0[0]); atomicAdd(out + 1, s1[0]); }
}

// ---------------------------------------------------------------------------------
// K1: coloured, atomic-free stiffness assembly (dcfemmodelling.cpp:163-228)
//   one thread per cell of ONE colour; cells of a colour share no node, so no two threads of a
//   launch touch the same CSR slot.  Node ids, scatter slots and rho are SoA in colour order
//   -> fully coalesced; the nK wavenumber matrices are produced in the same pass.
//   S(k) += (1/rho_c) (K_c + k^2 M_c);   rho == nullptr -> rho = 1 (the S1 matrix of the SR scheme)
// ---------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(128)
k_assemble(const double *__restrict__ pos, const int *__restrict__ cells_col, const int *__restrict__ pos_col,
           const int *__restrict__ color_order, const double *__restrict__ rho, int C, int first, int count,
           const double *__restrict__ kvals, int nK, size_t nnz, double *__restrict__ vals) {
    constexpr int NV = ElemTraits<E>::NV, NL = ElemTraits<E>::NL, DIM = ElemTraits<E>::DIM;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    double r = 1.0;
    if (rho) r = rho[color_order[slot]];
    if (fabs(r) <= 1e-12) return;                       // :185 skip |rho| <= TOLERANCE
    const double inv_rho = 1.0 / r;
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int n = cells_col[(size_t)v * C + slot];
        X[v][0] = pos[3 * (size_t)n]; X[v][1] = pos[3 * (size_t)n + 1]; X[v][2] = pos[3 * (size_t)n + 2];
    }
    double size, G[NV][NV];
    simplex_gram<DIM>(X, size, G);
#pragma unroll
    for (int i = 0; i < NL; i++) {
#pragma unroll
        for (int j = 0; j < NL; j++) {
            const double kij = stiff_entry<E>(i, j, size, G);
            const double mij = size * mass_unit<E>(i, j);
            const int p = pos_col[(size_t)(i * NL + j) * C + slot];
            for (int kk = 0; kk < nK; kk++) {
                const double k = kvals[kk];
                const double v = (k > 0.0) ? (mij * (k * k) + kij) : kij;    // :186-196
                vals[(size_t)kk * nnz + p] += v * inv_rho;
            }
        }
    }
}

// generic FEM matrices on the same element kernels (SparseMatrix::fillStiffnessMatrix / fillMassMatrix,
// core/src/sparsematrix.h:1034-1065):  vals += a_c * K_c + b_c * M_c  with per-cell coefficients in the ORIGINAL cell
// order (a == nullptr / b == nullptr: that term is absent)
template <int E>
__global__ void __launch_bounds__(128)
k_assemble_generic(const double *__restrict__ pos, const int *__restrict__ cells_col, const int *__restrict__ pos_col,
                   const int *__restrict__ color_order, const double *__restrict__ a, const double *__restrict__ b, int C, int first,
                   int count, double *__restrict__ vals) {
    constexpr int NV = ElemTraits<E>::NV, NL = ElemTraits<E>::NL, DIM = ElemTraits<E>::DIM;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int slot = first + t;
    const int cell = color_order[slot];
    const double wa = a ? a[cell] : 0.0, wb = b ? b[cell] : 0.0;
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int n = cells_col[(size_t)v * C + slot];
        X[v][0] = pos[3 * (size_t)n]; X[v][1] = pos[3 * (size_t)n + 1]; X[v][2] = pos[3 * (size_t)n + 2];
    }
    double size, G[NV][NV];
    simplex_gram<DIM>(X, size, G);
#pragma unroll
    for (int i = 0; i < NL; i++) {
#pragma unroll
        for (int j = 0; j < NL; j++) {
            double v = 0.0;
            if (a) v = wa * stiff_entry<E>(i, j, size, G);
            if (b) v = fma(wb, size * mass_unit<E>(i, j), v);
            vals[pos_col[(size_t)(i * NL + j) * C + slot]] += v;
        }
    }
}

// mixed boundary faces (:243-299): vals[k][slot] += sum_e coef[k][e] / rho[owner[e]]
__global__ void k_boundary_add(const int *__restrict__ slot, const int *__restrict__ ptr, const int *__restrict__ owner,
                               const double *__restrict__ coef, int n_slots, int n_entries, const double *__restrict__ rho,
                               int nK, size_t nnz, double *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots) return;
    const int p = slot[t];
    for (int kk = 0; kk < nK; kk++) {
        double acc = 0.0;
        for (int e = ptr[t]; e < ptr[t + 1]; e++) {
            const double r = rho ? rho[owner[e]] : 1.0;
            acc += coef[(size_t)kk * n_entries + e] / r;
        }
        vals[(size_t)kk * nnz + p] += acc;
    }
}
// homogeneous Dirichlet rows/cols (:141-161)
__global__ void k_dirichlet_zero(const int *__restrict__ zero_slots, int nz, int nK, size_t nnz, double *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nz) for (int kk = 0; kk < nK; kk++) vals[(size_t)kk * nnz + zero_slots[t]] = 0.0;
}
__global__ void k_dirichlet_diag(const int *__restrict__ diag_slots, int nd, int nK, size_t nnz, double *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nd) for (int kk = 0; kk < nK; kk++) vals[(size_t)kk * nnz + diag_slots[t]] = 1.0;
}
// :209-218 rows whose diagonal is < TOLERANCE would be forced to identity by the reference;
// we count them (the host refuses such a model rather than silently diverging)
__global__ void k_count_singular(const int *__restrict__ diag_pos, int N, int nK, size_t nnz, const double *__restrict__ vals, int *count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int kk = 0; kk < nK; kk++) if (vals[(size_t)kk * nnz + diag_pos[i]] < 1e-12) atomicAdd(count, 1);
}
__global__ void k_inv_diag(const int *__restrict__ diag_pos, int N, int nK, size_t nnz, const double *__restrict__ vals,
                           double *__restrict__ dinv /* [nK*N] */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int kk = 0; kk < nK; kk++) dinv[(size_t)kk * N + i] = 1.0 / vals[(size_t)kk * nnz + diag_pos[i]];
}

// ---------------------------------------------------------------------------------
// analytic primary potentials (bertMisc.cpp:186-260 + singular patch electrode.cpp:154-189)
//   prim[node * ld + (e + nE*kIdx)]
// ---------------------------------------------------------------------------------
__global__ void k_primary(const double *__restrict__ pos, int N, const double *__restrict__ el_pos, int nE,
                          const int *__restrict__ sing_node, const double *__restrict__ sing_val,
                          const double *__restrict__ kvals, int nK, double surface_z, int fullspace,
                          double *__restrict__ prim, size_t ld) {
    const int s = blockIdx.y * blockDim.x + threadIdx.x;     // columns fastest -> coalesced
    const int node = blockIdx.x * blockDim.y + threadIdx.y;
    if (s >= nE * nK || node >= N) return;
    const int kk = s / nE, e = s - kk * nE;
    const double k = kvals[kk];
    const int md = (k > 0.0) ? 1 : 2;                         // mirrored coordinate (:241)
    const double px = pos[3 * (size_t)node], py = pos[3 * (size_t)node + 1], pz = pos[3 * (size_t)node + 2];
    const double sx = el_pos[3 * e], sy = el_pos[3 * e + 1], sz = el_pos[3 * e + 2];
    double val;
    if (sing_node[e] == node) {
        val = sing_val[kk * nE + e];
    } else {
        const double dx = px - sx, dy = py - sy, dz = pz - sz;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        if (r < 1e-12) {
            val = 0.0;                                        // fallback (:235)
        } else {
            double mx = sx, my = sy, mz = sz;
            if (md == 1) my = 2.0 * surface_z - sy; else mz = 2.0 * surface_z - sz;
            const double ex = px - mx, ey = py - my, ez = pz - mz;
            const double rm = sqrt(ex * ex + ey * ey + ez * ez);
            const double PI_ = 3.14159265358979323846;
            if (fullspace) {
                val = (k == 0.0) ? 1.0 / (4.0 * PI_ * r) : as_bessel_k0(r * k) / (2.0 * PI_);
            } else if (k == 0.0) {
                val = (1.0 / r + 1.0 / rm) / (4.0 * PI_);
            } else {
                const double d2 = (sx - mx) * (sx - mx) + (sy - my) * (sy - my) + (sz - mz) * (sz - mz);
                if (d2 < 1e-12) val = as_bessel_k0(r * k) / PI_;                        // source == mirror (:220)
                else val = (as_bessel_k0(r * k) + as_bessel_k0(rm * k)) / (2.0 * PI_);
            }
        }
    }
    prim[(size_t)node * ld + s] = val;
}

// rho at the source: geometric mean over the cells around the electrode (electrode.cpp:102-120)
__global__ void k_rho_src(const int *__restrict__ ptr, const int *__restrict__ cells, const double *__restrict__ rho,
                          int nE, double *__restrict__ rho_src) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    double acc = 0.0; int n = ptr[e + 1] - ptr[e];
    for (int i = ptr[e]; i < ptr[e + 1]; i++) acc += log(rho[cells[i]]);
    rho_src[e] = exp(acc / (double)n);
}

// ---------------------------------------------------------------------------------
// K2: CSR x dense-block SpMM  Y = A X  (block of all sources; per-wavenumber values)
//   CTA = ROWS rows x (16*CPT) columns; a half-warp reads 16 consecutive doubles (128 B) of a
//   gathered X row, every thread carries CPT independent accumulators.  Optionally fuses
//   the per-column dot  sum_i X[i][s] * Y[i][s]  (p.Ap of PCG): serial over the thread's rows,
//   shared-memory tree over the row lanes, one atomicAdd per column and CTA.
//   MODE 0: Y = A X            MODE 1: Y = A1 X - rho_src[e] * A X  (the SR right-hand side,
//   dcfemmodelling.cpp:2252-2254, both products in one pass over the pattern)
// ---------------------------------------------------------------------------------
constexpr int SPMM_TX = 16, SPMM_TY = 8, SPMM_ROWS = 32;

template <int CPT, int MODE, bool DOT>
__global__ void __launch_bounds__(SPMM_TX * SPMM_TY)
k_spmm(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
       const double *__restrict__ vals1, const double *__restrict__ rho_src, size_t nnz,
       const double *__restrict__ X, double *__restrict__ Y, int N, int nE, int c0, int c1, size_t ld,
       double *__restrict__ dots) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int cbase = c0 + blockIdx.y * (SPMM_TX * CPT) + tx;
    int col[CPT]; size_t voff[CPT]; double rs[CPT]; bool ok[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) {
        col[m] = cbase + m * SPMM_TX;
        ok[m] = col[m] < c1;
        const int cc = ok[m] ? col[m] : c0;
        const int kk = cc / nE;
        voff[m] = (size_t)kk * nnz;
        rs[m] = (MODE == 1) ? rho_src[cc - kk * nE] : 0.0;
        col[m] = cc;
    }
    double part[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) part[m] = 0.0;
    const int row0 = blockIdx.x * SPMM_ROWS;
    for (int r = ty; r < SPMM_ROWS; r += SPMM_TY) {
        const int row = row0 + r;
        if (row >= N) break;
        const int pb = rowptr[row], pe = rowptr[row + 1];
        double acc[CPT];
#pragma unroll
        for (int m = 0; m < CPT; m++) acc[m] = 0.0;
        for (int p = pb; p < pe; p++) {
            const size_t xo = (size_t)colidx[p] * ld;
#pragma unroll
            for (int m = 0; m < CPT; m++) {
                double a = __ldg(vals + voff[m] + p);
                if (MODE == 1) a = __ldg(vals1 + voff[m] + p) - rs[m] * a;
                acc[m] = fma(a, __ldg(X + xo + col[m]), acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < CPT; m++) {
            if (ok[m]) {
                Y[(size_t)row * ld + col[m]] = acc[m];
                if (DOT) part[m] = fma(acc[m], __ldg(X + (size_t)row * ld + col[m]), part[m]);
            }
        }
    }
    if (DOT) {
        __shared__ double red[SPMM_TY][SPMM_TX * CPT];
#pragma unroll
        for (int m = 0; m < CPT; m++) red[ty][m * SPMM_TX + tx] = part[m];
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int m = 0; m < CPT; m++) {
                double s = 0.0;
#pragma unroll
                for (int y = 0; y < SPMM_TY; y++) s += red[y][m * SPMM_TX + tx];
                if (ok[m]) atomicAdd(dots + col[m], s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// K2': streamed row-panel SpMM  Y = A X  (+ fused epilogues), the PCG hot kernel.   Layout: stream_panels.h.
//
//   Persistent, warp-specialised: one CTA per SM = 1 PRODUCER warp + 15 CONSUMER warps, a ring of S shared-memory slots
//   guarded by full/empty mbarriers.
//     producer  for every stage (panel, chunk) of this CTA's work list: wait for the slot to be free, arm its mbarrier
//               with the byte count, then issue the TMA bulk copies (cp.async.bulk.shared.global, SASS UBLKCP): one per
//               halo row of the chunk (the tile's columns of X), one for the chunk's packed entries {value, row-in-chunk},
//               one for its per-row entry ranges.  The copies of stage n+1 run while the consumers work on stage n.
//     consumers every warp owns up to 4 rows of the panel and keeps their sums for all columns of the tile in registers
//               across the chunks (lane l: columns 2l, 2l+1 and 64+2l, 64+2l+1 -> two conflict-free 128-bit LDS per entry
//               and row: 7 shared-memory wavefronts per entry for 100 columns instead of 8 + a second pass over the
//               entries).  Inner loop: one 128-bit broadcast LDS of the entry, the X loads, 2-4 DFMA.
//               After chunk 0 (the panel's own X rows, processed last) the epilogue writes Y with 128-bit stores.
//   Column tiles never straddle a wavenumber group (one CSR value set per tile); 2.5-D windows are tile lists.
//   Per-column dot products (p.Ap, r.z) are DETERMINISTIC: lane partials -> fixed-order sum over the 15 warps -> one partial
//   row per CTA in global memory -> the CTA that takes the last ticket adds the rows in index order (no float atomics).
// ---------------------------------------------------------------------------------
// sum of base[sl * ld] over sl = first, first + step, ... < n, added in that order; the loads of 8 rows are issued
// together (the last CTA runs alone on the GPU: without this it would pay one L2 round trip per row)
__device__ __forceinline__ double ordered_row_sum(const double *base, size_t ld, int first, int step, int n) {
    double acc = 0.0;
    int sl = first;
    for (; sl + 7 * step < n; sl += 8 * step) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; u++) t[u] = __ldcg(base + (size_t)(sl + u * step) * ld);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += t[u];
    }
    for (; sl < n; sl += step) acc += __ldcg(base + (size_t)sl * ld);
    return acc;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of a global range (no shared-memory destination): the ring of k_spmm_mma asks for the NEXT stage's data
// while the current one is being copied, so that the copy into the freed slot later is served by L2 and not by DRAM
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct __align__(16) PanelEntry { double a; uint32_t idx; uint32_t pad; };   // value, row of its column in the staged chunk

// packed entries of one matrix in the streamed order: ent[k][p'] = {vals[k][ent_src[p']], ent_idx[p']}
__global__ void k_pack_entries(const int *__restrict__ ent_src, const unsigned *__restrict__ ent_idx, size_t nnz, int nK,
                               const double *__restrict__ vals, PanelEntry *__restrict__ ent) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int src = ent_src[p];
    const uint32_t idx = ent_idx[p];
    for (int kk = 0; kk < nK; kk++) {
        PanelEntry e; e.a = vals[(size_t)kk * nnz + src]; e.idx = idx; e.pad = 0;
        ent[(size_t)kk * nnz + p] = e;
    }
}

// what a consumer warp does with the row sums  acc = (A X)[row]:
//   EPI_SPMM     Y = acc                         (+ optional dot  X_row . Y_row : p.Ap of PCG)
//   EPI_POST     Y = X_row + dw_row (R_row - acc) (+ dot R_row . Y_row): damped-Jacobi post-smoothing, r.z of PCG
//   EPI_RESIDUAL Y = X_row - acc: with values pre-scaled by dw this is the residual after the pre-smoothing sweep
//                from a zero guess, R - A (dw .* R); k_amg_sum_members then restricts it (piecewise-constant P)
enum PanelEpi : int { EPI_SPMM = 0, EPI_POST = 1, EPI_RESIDUAL = 2 };

struct PanelExtra {
    const double *R;          // EPI_POST: residual block
    const double *dinvw;      // EPI_POST: [nK][n] damped inverse diagonal
    int n;                    // rows of the level (stride of dinvw per wavenumber)
};

#ifndef PGB_ST_WARPS
#define PGB_ST_WARPS 15
#define PGB_ST_RPW 4
#define PGB_ST_UNROLL 4
#endif
constexpr int ST_CONSUMER_WARPS = PGB_ST_WARPS;    // + 1 producer warp (15 + 1 = 512 threads: 128 registers per thread available)
constexpr int ST_CONSUMERS = ST_CONSUMER_WARPS * 32;
constexpr int ST_THREADS = ST_CONSUMERS + 32;      // + the producer warp
constexpr int ST_UNROLL = PGB_ST_UNROLL;
constexpr int ST_RPW = PGB_ST_RPW;                 // rows per consumer warp: panels have at most ST_CONSUMER_WARPS * ST_RPW rows
constexpr int ST_MAX_SLOTS = 4;
constexpr int ST_MAX_TILE_W = 128;                 // columns per tile (2 column pairs per lane)

struct StreamLevel {       // device arrays of stream_panels.h
    const int *panel_row_ptr, *panel_chunk_ptr, *chunk_halo_ptr, *halo_cols, *chunk_ent_ptr, *crp, *chunk_run_ptr, *runs;
    int n_panels, crp_stride;
};

struct StreamArgs {
    StreamLevel L;
    const PanelEntry *ent; size_t nnz;            // packed entries [nK][nnz]
    const double *X; double *Y; size_t ld;
    int nE, c0, c1;                               // active column window [c0, c1)
    int k_lo, tpk, pw, n_tiles, cpt;              // tile geometry: first wavenumber, tiles per wavenumber, tile width, tiles, CTAs per tile
    int slots; uint32_t slot_bytes, x_bytes, ent_bytes;    // slot = [X rows | entries | row ranges]
    int fullrows;                                 // every tile spans whole rows of X (one wavenumber, whole window): run-merged bulk copies
    double *dot_part; unsigned *dot_counter; double *dots;  // deterministic dots (DOT kernels)
    PanelExtra ex;
};

// tile t -> wavenumber kk, copied columns [cs, cs + wc), valid columns [v0, v1); false: nothing to do
__device__ __forceinline__ bool stream_tile(const StreamArgs &A, int t, int &kk, int &cs, int &wc, int &v0, int &v1) {
    const int q = t / A.tpk, j = t - q * A.tpk;
    kk = A.k_lo + q;
    const int gb = max(kk * A.nE, A.c0), ge = min((kk + 1) * A.nE, A.c1);    // this group's part of the window
    cs = (gb & ~1) + j * A.pw;                                                // even start: 16-byte aligned bulk copies
    const int ce = min(cs + A.pw, (ge + 1) & ~1);
    v0 = max(cs, gb); v1 = min(ce, ge);
    if (v1 <= v0) return false;
    wc = ce - cs;
    return true;
}

template <int NCP, int EPI, bool DOT>
__global__ void __launch_bounds__(ST_THREADS, 1)
k_spmm_stream(const StreamArgs A) {
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ __align__(8) uint64_t full[ST_MAX_SLOTS], empty[ST_MAX_SLOTS];
    __shared__ double sdot[ST_MAX_TILE_W];
    __shared__ int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = A.slots;
    if (threadIdx.x == 0) {
        // full: one arrival from the producer's expect_tx (+ one per producer lane when the X rows come by cp.async)
        for (int s = 0; s < S; s++) { mbar_init(&full[s], A.fullrows ? 1 : 33); mbar_init(&empty[s], ST_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // work list of this CTA: tiles t_first, t_first + t_step, ...; panels j, j + cpt, ...
    int t_first, t_step, j;
    if (A.n_tiles <= (int)gridDim.x) { t_first = blockIdx.x / A.cpt; t_step = A.n_tiles; j = blockIdx.x - t_first * A.cpt; }
    else { t_first = blockIdx.x; t_step = gridDim.x; j = 0; }
    const int pstep = A.cpt;
    const StreamLevel &L = A.L;

    if (warp == ST_CONSUMER_WARPS) {
        // ------------------------------- producer -------------------------------
        uint32_t n = 0;
        for (int t = t_first; t < A.n_tiles; t += t_step) {
            int kk, cs, wc, v0, v1;
            if (!stream_tile(A, t, kk, cs, wc, v0, v1)) continue;
            const uint32_t rowb = (uint32_t)wc * 8u;
            const bool fullrows = A.fullrows != 0;
            const PanelEntry *entk = A.ent + (size_t)kk * A.nnz;
            for (int p = j; p < L.n_panels; p += pstep) {
                const int ch0 = L.panel_chunk_ptr[p], ch1 = L.panel_chunk_ptr[p + 1];
                for (int ch = ch1 - 1; ch >= ch0; ch--, n++) {
                    const uint32_t slot = n % (uint32_t)S, use = n / (uint32_t)S;
                    unsigned char *sb = st_smem + (size_t)slot * A.slot_bytes;
                    const int h0 = L.chunk_halo_ptr[ch], hn = L.chunk_halo_ptr[ch + 1] - h0;
                    const int e0 = L.chunk_ent_ptr[ch], ne = L.chunk_ent_ptr[ch + 1] - e0;
                    // what to copy is read BEFORE waiting for the slot: the index loads overlap the consumers' work
                    int src_row[4], dst_row[4], len[4], ncopy = 0;
                    if (fullrows) {
                        // the tile spans whole rows of X: one bulk copy per run of consecutive halo rows
                        const int q0 = L.chunk_run_ptr[ch], q1 = L.chunk_run_ptr[ch + 1];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = q0 + lane + 32 * u;
                            if (i < q1) { dst_row[u] = __ldg(L.runs + 3 * i); src_row[u] = __ldg(L.runs + 3 * i + 1); len[u] = __ldg(L.runs + 3 * i + 2); ncopy = u + 1; }
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = lane + 32 * u;
                            if (i < hn) { dst_row[u] = i; src_row[u] = __ldg(L.halo_cols + h0 + i); len[u] = 1; ncopy = u + 1; }
                        }
                    }
                    if (use > 0) mbar_wait(&empty[slot], (use & 1u) ^ 1u);
                    if (lane == 0) {
                        mbar_expect_tx(&full[slot], (fullrows ? (uint32_t)hn * rowb : 0u) + (uint32_t)ne * 16u + (uint32_t)L.crp_stride * 4u);
                        if (ne > 0) tma_bulk_g2s(sb + A.x_bytes, entk + e0, (uint32_t)ne * 16u, &full[slot]);
                        tma_bulk_g2s(sb + A.x_bytes + A.ent_bytes, L.crp + (size_t)ch * L.crp_stride, (uint32_t)L.crp_stride * 4u, &full[slot]);
                    }
                    __syncwarp();
                    if (fullrows) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (u < ncopy)
                                tma_bulk_g2s(sb + (size_t)dst_row[u] * rowb, A.X + (size_t)src_row[u] * A.ld + cs, (uint32_t)len[u] * rowb, &full[slot]);
                    } else {
                        // partial-width tiles (2.5-D wavenumber groups, multi-GPU column shards): the rows are short and
                        // strided, a bulk copy per row would be bound by the copy engine's per-request cost.  One
                        // 16-byte cp.async (LDGSTS) per lane moves a whole row per warp instruction; completion is
                        // tracked by the same mbarrier (cp.async.mbarrier.arrive.noinc, one arrival per lane).
                        // (kept lean on purpose: one warp issues every row of the stage, so each instruction in this loop
                        //  costs ~100 issue slots per stage -- 32-bit shared addresses, one wide multiply-add per row)
                        const bool l0 = 16u * lane < rowb, l1 = NCP == 2 && 512u + 16u * lane < rowb;
                        const uint32_t ldb = (uint32_t)A.ld * 8u;
                        const unsigned char *gl = reinterpret_cast<const unsigned char *>(A.X + cs) + 16 * lane;
                        uint32_t d = smem_u32(sb) + 16u * lane;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int nrow = min(32, hn - 32 * u);
#pragma unroll 4
                            for (int jj = 0; jj < nrow; jj++) {
                                const uint32_t src = (uint32_t)__shfl_sync(0xffffffffu, src_row[u], jj);
                                const unsigned char *g = gl + (size_t)src * ldb;          // one IMAD.WIDE
                                if (l0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                                if (l1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512u), "l"(g + 512) : "memory");
                                d += rowb;
                            }
                        }
                        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
                    }
                }
            }
        }
        return;
    }

    // ------------------------------- consumers -------------------------------
    uint32_t n = 0;
    for (int t = t_first; t < A.n_tiles; t += t_step) {
        int kk, cs, wc, v0, v1;
        if (!stream_tile(A, t, kk, cs, wc, v0, v1)) continue;
        const uint32_t rowb = (uint32_t)wc * 8u;
        // lane's columns: pair q covers cs + 64 q + 2 lane, +1
        int col[NCP]; bool has[NCP], ok_lo[NCP], ok_hi[NCP];
#pragma unroll
        for (int q = 0; q < NCP; q++) {
            col[q] = cs + 64 * q + 2 * lane;
            has[q] = (64 * q + 2 * lane) < wc;
            ok_lo[q] = has[q] && col[q] >= v0 && col[q] < v1;
            ok_hi[q] = has[q] && col[q] + 1 >= v0 && col[q] + 1 < v1;
        }
        double part[2 * NCP];
#pragma unroll
        for (int q = 0; q < 2 * NCP; q++) part[q] = 0.0;
        const double *dwk = (EPI == EPI_POST) ? A.ex.dinvw + (size_t)kk * A.ex.n : nullptr;

        for (int p = j; p < L.n_panels; p += pstep) {
            const int r0 = L.panel_row_ptr[p], nrows = L.panel_row_ptr[p + 1] - r0;
            const int nch = L.panel_chunk_ptr[p + 1] - L.panel_chunk_ptr[p];
            double acc[ST_RPW][2 * NCP];
#pragma unroll
            for (int rr = 0; rr < ST_RPW; rr++)
#pragma unroll
                for (int q = 0; q < 2 * NCP; q++) acc[rr][q] = 0.0;
            for (int c = nch - 1; c >= 0; c--, n++) {
                const uint32_t slot = n % (uint32_t)S, use = n / (uint32_t)S;
                mbar_wait(&full[slot], use & 1u);
                const unsigned char *sb = st_smem + (size_t)slot * A.slot_bytes;
                const PanelEntry *sE = reinterpret_cast<const PanelEntry *>(sb + A.x_bytes);
                const int *sR = reinterpret_cast<const int *>(sb + A.x_bytes + A.ent_bytes);
                const unsigned char *xl = sb + 16 * lane;
#pragma unroll
                for (int rr = 0; rr < ST_RPW; rr++) {
                    const int r = warp + ST_CONSUMER_WARPS * rr;
                    if (r < nrows) {
                        const int pb = sR[r], pe = sR[r + 1];
#pragma unroll ST_UNROLL
                        for (int e = pb; e < pe; e++) {
                            const int4 raw = *reinterpret_cast<const int4 *>(sE + e);      // one 128-bit broadcast load: {value, row}
                            const double a = __hiloint2double(raw.y, raw.x);
                            const unsigned char *xr = xl + (uint32_t)raw.z * rowb;
#pragma unroll
                            for (int q = 0; q < NCP; q++) {
                                if (has[q]) {
                                    const double2 x = *reinterpret_cast<const double2 *>(xr + 512 * q);
                                    acc[rr][2 * q] = fma(a, x.x, acc[rr][2 * q]);
                                    acc[rr][2 * q + 1] = fma(a, x.y, acc[rr][2 * q + 1]);
                                }
                            }
                        }
                    }
                }
                if (c == 0) {
                    // epilogue: chunk 0 starts with the panel's own X rows (halo index == local row)
                    double rres[EPI == EPI_POST ? ST_RPW : 1][2 * NCP], dwr[ST_RPW];
                    if (EPI == EPI_POST) {
                        // all global loads of the epilogue are issued before the first store (Y may alias nothing, but the
                        // compiler cannot know): one L2 round trip per panel instead of one per row
#pragma unroll
                        for (int rr = 0; rr < ST_RPW; rr++) {
                            const int r = warp + ST_CONSUMER_WARPS * rr;
                            dwr[rr] = 0.0;
#pragma unroll
                            for (int q = 0; q < 2 * NCP; q++) rres[rr][q] = 0.0;
                            if (r < nrows) {
                                const int row = r0 + r;
                                dwr[rr] = __ldg(dwk + row);
#pragma unroll
                                for (int q = 0; q < NCP; q++) {
                                    const size_t o = (size_t)row * A.ld + col[q];
                                    if (ok_lo[q] && ok_hi[q]) { const double2 rv = __ldg(reinterpret_cast<const double2 *>(A.ex.R + o)); rres[rr][2 * q] = rv.x; rres[rr][2 * q + 1] = rv.y; }
                                    else { if (ok_lo[q]) rres[rr][2 * q] = __ldg(A.ex.R + o); if (ok_hi[q]) rres[rr][2 * q + 1] = __ldg(A.ex.R + o + 1); }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < ST_RPW; rr++) {
                        const int r = warp + ST_CONSUMER_WARPS * rr;
                        if (r < nrows) {
                            const int row = r0 + r;
                            const unsigned char *xs = xl + (uint32_t)r * rowb;
                            const double dw = (EPI == EPI_POST) ? dwr[rr] : 0.0;
#pragma unroll
                            for (int q = 0; q < NCP; q++) {
                                if (has[q]) {
                                    const double2 x = *reinterpret_cast<const double2 *>(xs + 512 * q);
                                    const size_t o = (size_t)row * A.ld + col[q];
                                    double y0, y1;
                                    if (EPI == EPI_POST) {
                                        const double r_lo = rres[rr][2 * q], r_hi = rres[rr][2 * q + 1];
                                        y0 = fma(dw, r_lo - acc[rr][2 * q], x.x);
                                        y1 = fma(dw, r_hi - acc[rr][2 * q + 1], x.y);
                                        if (DOT) { if (ok_lo[q]) part[2 * q] = fma(r_lo, y0, part[2 * q]); if (ok_hi[q]) part[2 * q + 1] = fma(r_hi, y1, part[2 * q + 1]); }
                                    } else if (EPI == EPI_RESIDUAL) {
                                        y0 = x.x - acc[rr][2 * q]; y1 = x.y - acc[rr][2 * q + 1];
                                    } else {
                                        y0 = acc[rr][2 * q]; y1 = acc[rr][2 * q + 1];
                                        if (DOT) { if (ok_lo[q]) part[2 * q] = fma(y0, x.x, part[2 * q]); if (ok_hi[q]) part[2 * q + 1] = fma(y1, x.y, part[2 * q + 1]); }
                                    }
                                    if (ok_lo[q] && ok_hi[q]) *reinterpret_cast<double2 *>(A.Y + o) = make_double2(y0, y1);
                                    else { if (ok_lo[q]) A.Y[o] = y0; if (ok_hi[q]) A.Y[o + 1] = y1; }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
        }
        if (DOT) {
            // fixed-order sum over the warps, one partial row per CTA, the last CTA of the tile adds the rows in order
            const int nslot = (A.n_tiles <= (int)gridDim.x) ? A.cpt : 1;
            for (int w = 0; w < ST_CONSUMER_WARPS; w++) {
                if (warp == w) {
#pragma unroll
                    for (int q = 0; q < NCP; q++) {
                        if (has[q]) {
                            const int i = 64 * q + 2 * lane;
                            sdot[i] = (w == 0 ? 0.0 : sdot[i]) + part[2 * q];
                            sdot[i + 1] = (w == 0 ? 0.0 : sdot[i + 1]) + part[2 * q + 1];
                        }
                    }
                }
                named_bar_sync(1, ST_CONSUMERS);
            }
            double *mine = A.dot_part + (size_t)j * A.ld;
            for (int i = threadIdx.x; i < wc; i += ST_CONSUMERS) if (cs + i >= v0 && cs + i < v1) mine[cs + i] = sdot[i];
            __threadfence();
            named_bar_sync(1, ST_CONSUMERS);
            if (threadIdx.x == 0) s_ticket = (int)atomicAdd(A.dot_counter + t, 1u);
            named_bar_sync(1, ST_CONSUMERS);
            if (s_ticket == nslot - 1) {
                __threadfence();
                // 4 threads per column add every 4th partial row, then the four sums are added in order
                const int i = threadIdx.x >> 2, part4 = threadIdx.x & 3;
                for (int i0 = 0; i0 < wc; i0 += ST_CONSUMERS / 4) {
                    const int cc = cs + i0 + i;
                    const bool okc = (i0 + i) < wc && cc >= v0 && cc < v1;
                    double s4 = okc ? ordered_row_sum(A.dot_part + cc, A.ld, part4, 4, nslot) : 0.0;
                    const double s1 = __shfl_down_sync(0xffffffffu, s4, 1), s2 = __shfl_down_sync(0xffffffffu, s4, 2), s3 = __shfl_down_sync(0xffffffffu, s4, 3);
                    if (okc && part4 == 0) A.dots[cc] = ((s4 + s1) + s2) + s3;
                }
                if (threadIdx.x == 0) A.dot_counter[t] = 0u;
            }
            named_bar_sync(1, ST_CONSUMERS);
        }
    }
}

// ---------------------------------------------------------------------------------
// K2'': the same streamed row-panel SpMM with the products on the FP64 tensor-core pipe (DMMA m8n8k4).
//
//   Why: k_spmm_stream needs 8 distinct bytes of X from shared memory per DFMA lane (9 wavefronts per CSR entry for 100
//   columns) and is bound by the shared-memory pipe at ~30 % of the HBM roofline.  A DMMA holds its operands in register
//   FRAGMENTS that are reused across the 8 x 8 outputs: one k-step of the layout below (8 rows x 4 columns of A against 4
//   staged rows of X, all source columns) costs 2 wavefronts for A and 2 per 8 source columns for X = 28 wavefronts for 100
//   columns and ~10 CSR entries' worth of work, i.e. ~3 per entry.  The zeros of the 8 x 4 blocks are multiplied too
//   (~3x the flops of the CSR form), which the otherwise idle FP64 tensor pipe (64 FMA/clk/SM, measured) absorbs.
//
//   Layout (stream_panels.h, "8-row-group form"): a panel has up to 8 * MM_CONSUMER_WARPS rows; consumer warp w owns rows
//   [8w, 8w+8) of the panel and keeps their sums for all columns of the tile in C fragments (lane: row lane/4, columns
//   8t + 2(lane%4) + {0,1} of n-tile t) across the chunks.  Per (chunk, group) the host lists the k-steps: 32 packed A
//   values in fragment order (lane = 4 * row + column) and one word with the 4 staged-row indices.
//   Producer warp, ring of slots, tiles, deterministic dots: as k_spmm_stream.
// ---------------------------------------------------------------------------------
#ifndef PGB_MM_WARPS
#define PGB_MM_WARPS 12      // 3 consumer warpgroups + 1 producer warpgroup = 16 warps, on every SM sub-partition 3 consumers (the DMMA
#define PGB_MM_PRODUCERS 4   // pipe is per sub-partition) and 1 producer.  The kernel starts with 128 registers per thread; the
#endif                       // producers shrink to 56 and the consumers grow to 152 (setmaxnreg).  Issuing a bulk copy costs the
                             // issuing warp ~65 cycles (measured, ab/dmma/tma_fill2.cu): the copies of a stage are split over the producers
constexpr int MM_CONSUMER_WARPS = PGB_MM_WARPS;
constexpr int MM_PRODUCER_WARPS = PGB_MM_PRODUCERS;
constexpr int MM_CONSUMERS = MM_CONSUMER_WARPS * 32;
constexpr int MM_THREADS = MM_CONSUMERS + 32 * MM_PRODUCER_WARPS;
#ifndef PGB_MM_CREGS
#define PGB_MM_CREGS 152
#endif
constexpr int MM_PRODUCER_REGS = 56, MM_CONSUMER_REGS = PGB_MM_CREGS;       // 12 * 32 * 152 + 4 * 32 * 56 = 65536
static_assert(MM_CONSUMER_WARPS % 4 == 0 && MM_PRODUCER_WARPS == 4, "setmaxnreg works on warpgroups of 4 warps");
constexpr int MM_ROWS = 8 * MM_CONSUMER_WARPS;         // rows per panel
constexpr int MM_GSTRIDE = (MM_CONSUMER_WARPS + 1 + 3) / 4 * 4;

// packed A fragments of one matrix: aval[k][i] = vals[k][a_src[i]] (0 where a_src[i] < 0)
__global__ void k_pack_mma(const int *__restrict__ a_src, size_t n_frag, size_t nnz, int nK, const double *__restrict__ vals,
                           double *__restrict__ aval) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frag) return;
    const int src = a_src[i];
    for (int kk = 0; kk < nK; kk++) aval[(size_t)kk * n_frag + i] = src >= 0 ? vals[(size_t)kk * nnz + src] : 0.0;
}

struct MmaLevel {
    const int *panel_row_ptr, *panel_chunk_ptr, *chunk_halo_ptr, *halo_cols, *chunk_ks_ptr, *chunk_meta_ptr, *chunk_run_ptr, *runs;
    const int4 *cdesc;            // per chunk two int4: {first halo entry, halo entries, first k-step, k-steps}, {first meta word, meta words, first run, runs}
    const unsigned *meta;
    int n_panels;
};
struct MmaArgs {
    MmaLevel L;
    const double *aval; size_t n_frag;            // packed A fragments [nK][n_frag], n_frag = 32 * k-steps
    const double *X; double *Y; size_t ld;
    int nE, c0, c1;
    int k_lo, tpk, pw, n_tiles, cpt;
    int slots; uint32_t slot_bytes, x_bytes, a_bytes;      // slot = [X rows | A fragments | meta]
    int fullrows;
    int dbg;                                      // measurement only: 1 = skip the k-step loop, 2 = skip the X copies, 4 = skip the epilogue,
    long long *dbg_buf;                           //                   8 = per-stage clock stamps of CTA 0 into dbg_buf[stage * 8 + i]
    double *dot_part; unsigned *dot_counter; double *dots;
    PanelExtra ex;
};
__device__ __forceinline__ bool mma_tile(const MmaArgs &A, int t, int &kk, int &cs, int &wc, int &v0, int &v1) {
    const int q = t / A.tpk, j = t - q * A.tpk;
    kk = A.k_lo + q;
    const int gb = max(kk * A.nE, A.c0), ge = min((kk + 1) * A.nE, A.c1);
    cs = (gb & ~1) + j * A.pw;
    const int ce = min(cs + A.pw, (ge + 1) & ~1);
    v0 = max(cs, gb); v1 = min(ce, ge);
    if (v1 <= v0) return false;
    wc = ce - cs;
    return true;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// wait for the phase of a full-barrier; the returned token orders the stage's shared-memory reads behind the wait
__device__ __forceinline__ uint32_t mbar_wait_tok(uint64_t *bar, uint32_t phase) {
    // test_wait in a spin loop: try_wait may suspend the warp for an implementation-defined time, which adds its wake-up
    // latency to every stage of the ring
    uint32_t tok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "mov.u32 %0, %2;\n"
        "}\n" : "=r"(tok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return tok;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr, uint32_t tok) {
    double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr), "r"(tok)); return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr, uint32_t tok) {
    double2 v; asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "r"(tok)); return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr, uint32_t tok) {
    uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr), "r"(tok)); return v;
}

template <int NT, int EPI, bool DOT>
__global__ void __launch_bounds__(MM_THREADS, 1)
k_spmm_mma(const MmaArgs A) {
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ __align__(8) uint64_t full[ST_MAX_SLOTS], empty[ST_MAX_SLOTS];
    __shared__ double sdot[ST_MAX_TILE_W];
    __shared__ int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = A.slots;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], MM_PRODUCER_WARPS * (A.fullrows ? 1 : 33)); mbar_init(&empty[s], MM_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int t_first, t_step, j;
    if (A.n_tiles <= (int)gridDim.x) { t_first = blockIdx.x / A.cpt; t_step = A.n_tiles; j = blockIdx.x - t_first * A.cpt; }
    else { t_first = blockIdx.x; t_step = gridDim.x; j = 0; }
    const int pstep = A.cpt;
    const MmaLevel &L = A.L;

    if (warp >= MM_CONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MM_PRODUCER_REGS));
        // ------------------------------- producers -------------------------------
        // producer pw issues every MM_PRODUCER_WARPS-th copy of a stage (the first one also the A fragments and the meta
        // block) and arms the stage's barrier with the bytes of its own copies.  What to copy is fetched AHEAD of the ring:
        // the chunk record two stages ahead, the copy list one stage ahead -- when the ring is the bottleneck (the usual
        // case) no index load sits between a slot becoming free and its copies being issued.
        const int pw = warp - MM_CONSUMER_WARPS;
        constexpr int P = MM_PRODUCER_WARPS;
        constexpr int U = (4 + P - 1) / P;          // up to 128 copies per stage over P warps
        uint32_t slot = 0, use = 0;
        int dbg_stage = 0;
        for (int t = t_first; t < A.n_tiles; t += t_step) {
            int kk, cs, wc, v0, v1;
            if (!mma_tile(A, t, kk, cs, wc, v0, v1)) continue;
            if (j >= L.n_panels) continue;
            const uint32_t rowb = (uint32_t)wc * 8u;
            const bool fullrows = A.fullrows != 0;
            const double *avk = A.aval + (size_t)kk * A.n_frag;
            // position of the record fetch (two stages ahead): panel pp, its chunk range, chunk pc counting down
            int pp = j, pc0 = L.panel_chunk_ptr[pp], pc = L.panel_chunk_ptr[pp + 1] - 1;
            int pn0 = 0, pn1 = 0;
            if (pp + pstep < L.n_panels) { pn0 = L.panel_chunk_ptr[pp + pstep]; pn1 = L.panel_chunk_ptr[pp + pstep + 1]; }
            bool more = true;                              // the position (pp, pc) is a stage
            auto fetch_rec = [&](int4 &ra, int4 &rb, bool &ok) {
                ok = more;
                if (!more) return;
                ra = __ldg(L.cdesc + 2 * pc); rb = __ldg(L.cdesc + 2 * pc + 1);
                if (pc > pc0) pc--;
                else {
                    pp += pstep;
                    if (pp < L.n_panels) {
                        pc0 = pn0; pc = pn1 - 1;
                        if (pp + pstep < L.n_panels) { pn0 = L.panel_chunk_ptr[pp + pstep]; pn1 = L.panel_chunk_ptr[pp + pstep + 1]; }
                    } else more = false;
                }
            };
            struct Copies { int src[U], dst[U], len[U], n, rows; };
            auto fetch_copies = [&](const int4 &ra, const int4 &rb, Copies &C) {
                C.n = 0; C.rows = 0;
                if (A.fullrows == 2) {
                    // partial-width tile, one bulk copy per halo row (the rows are not contiguous in memory)
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = pw + P * (lane + 32 * u);
                        if (i < ra.y && !(A.dbg & 2)) { C.dst[u] = i; C.src[u] = __ldg(L.halo_cols + ra.x + i); C.len[u] = 1; C.n = u + 1; C.rows += 1; }
                    }
                } else if (fullrows) {
                    const int q0 = rb.z, q1 = (A.dbg & 2) ? q0 : q0 + rb.w;
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = q0 + pw + P * (lane + 32 * u);
                        if (i < q1) { C.dst[u] = __ldg(L.runs + 3 * i); C.src[u] = __ldg(L.runs + 3 * i + 1); C.len[u] = __ldg(L.runs + 3 * i + 2); C.n = u + 1; C.rows += C.len[u]; }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = pw + P * (lane + 32 * u);
                        if (i < ra.y) { C.dst[u] = i; C.src[u] = __ldg(L.halo_cols + ra.x + i); C.len[u] = 1; C.n = u + 1; }
                    }
                }
            };
            int4 a0, b0, a1, b1, a2, b2; bool ok0, ok1, ok2;
            Copies C0, C1;
            fetch_rec(a0, b0, ok0);
            fetch_rec(a1, b1, ok1);
            fetch_copies(a0, b0, C0);
            while (ok0) {
                // ahead of the ring: record of stage + 2, copy list of stage + 1
                fetch_rec(a2, b2, ok2);
                if (ok1) fetch_copies(a1, b1, C1);
                // this stage
                unsigned char *sb = st_smem + (size_t)slot * A.slot_bytes;
                const int hn = a0.y, k0 = a0.z, nks = a0.w, m0 = b0.x, nm = b0.y;
                int rows_mine = C0.rows;
                if (fullrows) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) rows_mine += __shfl_xor_sync(0xffffffffu, rows_mine, o);
                }
                const bool stamp = (A.dbg & 8) && blockIdx.x == 0 && pw == 0 && lane == 0;
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 0] = clock64();
                if (use > 0) mbar_wait_spin(&empty[slot], (use & 1u) ^ 1u);
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 1] = clock64();
                if (lane == 0) {
                    mbar_expect_tx(&full[slot], (uint32_t)rows_mine * rowb + (pw == 0 ? (uint32_t)nks * 256u + (uint32_t)nm * 4u : 0u));
                    if (pw == 0) {
                        if (nks > 0) tma_bulk_g2s(sb + A.x_bytes, avk + (size_t)k0 * 32, (uint32_t)nks * 256u, &full[slot]);
                        tma_bulk_g2s(sb + A.x_bytes + A.a_bytes, L.meta + m0, (uint32_t)nm * 4u, &full[slot]);
                    }
                }
                __syncwarp();
                if (fullrows) {
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (u < C0.n)
                            tma_bulk_g2s(sb + (size_t)C0.dst[u] * rowb, A.X + (size_t)C0.src[u] * A.ld + cs, (uint32_t)C0.len[u] * rowb, &full[slot]);
                } else {
                    // partial-width tiles: one 16-byte cp.async per lane moves a row per warp instruction (see k_spmm_stream);
                    // this warp's rows are pw, pw + P, ...
                    const bool l0 = 16u * lane < rowb, l1 = 512u + 16u * lane < rowb;
                    const uint32_t ldb = (uint32_t)A.ld * 8u;
                    const unsigned char *gl = reinterpret_cast<const unsigned char *>(A.X + cs) + 16 * lane;
                    uint32_t d = smem_u32(sb) + 16u * lane + (uint32_t)pw * rowb;
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int nrow = min(32, (hn - pw + P - 1) / P - 32 * u);
#pragma unroll 4
                        for (int jj = 0; jj < nrow; jj++) {
                            const uint32_t src = (uint32_t)__shfl_sync(0xffffffffu, C0.src[u], jj);
                            const unsigned char *g = gl + (size_t)src * ldb;
                            if (l0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                            if (l1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512u), "l"(g + 512) : "memory");
                            d += P * rowb;
                        }
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
                }
                // the next stage goes to L2 now (its slot is still in use): DRAM latency leaves the ring's critical path
                if (ok1 && !(A.dbg & 32)) {
                    if (lane == 0 && pw == 0) {
                        if (a1.w > 0) tma_prefetch_l2(avk + (size_t)a1.z * 32, (uint32_t)a1.w * 256u);
                        tma_prefetch_l2(L.meta + b1.x, (uint32_t)b1.y * 4u);
                    }
                    if (A.fullrows == 1) {
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (u < C1.n) tma_prefetch_l2(A.X + (size_t)C1.src[u] * A.ld + cs, (uint32_t)C1.len[u] * rowb);
                    }
                }
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 2] = clock64();
                dbg_stage++;
                if (++slot == (uint32_t)S) { slot = 0; use++; }
                a0 = a1; b0 = b1; ok0 = ok1; C0 = C1;
                a1 = a2; b1 = b2; ok1 = ok2;
            }
        }
        return;
    }

    // ------------------------------- consumers -------------------------------
    // All shared-memory reads of the inner loop go through 32-bit shared addresses and non-volatile asm (the compiler may
    // interleave them with the DMMAs as it likes); every read carries the token the full-barrier wait of its stage
    // returned, so it can neither move above that wait nor be merged with a read of the same address in another stage.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MM_CONSUMER_REGS));
    const int lr = lane >> 2, lk = lane & 3;      // fragment coordinates: row / B column n = lr, k index / C column pair = lk
    const uint32_t smem0 = smem_u32(st_smem);
    uint32_t slot = 0, phase = 0;
    int dbg_stage = 0;
    for (int t = t_first; t < A.n_tiles; t += t_step) {
        int kk, cs, wc, v0, v1;
        if (!mma_tile(A, t, kk, cs, wc, v0, v1)) continue;
        const uint32_t rowb = (uint32_t)wc * 8u;
        const int nt = (wc + 7) >> 3;             // n-tiles of 8 source columns in use (<= NT; the others compute on whatever
                                                  // follows the row in the slot and are never stored)
        double part[DOT ? NT : 1][2];
#pragma unroll
        for (int q = 0; q < (DOT ? NT : 1); q++) { part[q][0] = 0.0; part[q][1] = 0.0; }
        const double *dwk = (EPI == EPI_POST) ? A.ex.dinvw + (size_t)kk * A.ex.n : nullptr;

        int p = j;
        int r0 = 0, r1 = 0, ch0 = 0, ch1 = 0, hc0 = 0;
        if (p < L.n_panels) {
            r0 = L.panel_row_ptr[p]; r1 = L.panel_row_ptr[p + 1]; ch0 = L.panel_chunk_ptr[p]; ch1 = L.panel_chunk_ptr[p + 1];
            hc0 = L.chunk_halo_ptr[ch0 + 1] - L.chunk_halo_ptr[ch0];
        }
        while (p < L.n_panels) {
            const int nrows = r1 - r0, nch = ch1 - ch0, row_base = r0;
            const int n0 = min(nrows, hc0);           // own rows staged with chunk 0; the others are the first rows of chunk 1
            // the next panel's extents are fetched now: their latency hides behind this panel's work
            const int pn = p + pstep;
            if (pn < L.n_panels) {
                r0 = L.panel_row_ptr[pn]; r1 = L.panel_row_ptr[pn + 1]; ch0 = L.panel_chunk_ptr[pn]; ch1 = L.panel_chunk_ptr[pn + 1];
                hc0 = L.chunk_halo_ptr[ch0 + 1] - L.chunk_halo_ptr[ch0];
            }
            uint32_t sb1 = 0, slot1 = 0;              // chunk 1 stays resident until the epilogue when it holds own rows
            double acc[NT][2];
#pragma unroll
            for (int q = 0; q < NT; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
            for (int c = nch - 1; c >= 0; c--) {
                const bool stamp = (A.dbg & 8) && blockIdx.x == 0 && threadIdx.x == 0;
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 3] = clock64();
                const uint32_t tok = mbar_wait_tok(&full[slot], phase);
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 4] = clock64();
                const uint32_t sb = smem0 + slot * A.slot_bytes;
                const uint32_t sM = sb + A.x_bytes + A.a_bytes;
                const int ks0 = (int)lds_u32(sM + 4u * warp, tok), ks1 = (A.dbg & 1) ? ks0 : (int)lds_u32(sM + 4u * warp + 4u, tok);
                uint32_t aA = sb + A.x_bytes + 8u * lane + 256u * (uint32_t)ks0;
                uint32_t aK = sM + 4u * MM_GSTRIDE + 4u * (uint32_t)ks0;
                const uint32_t xb = sb + 8u * lr;
#pragma unroll 2
                for (int ks = ks0; ks < ks1; ks++, aA += 256u, aK += 4u) {
                    const double a = lds_f64(aA, tok);
                    const uint32_t idx = __byte_perm(lds_u32(aK, tok), 0u, 0x4440u + lk);     // byte lk of the word
                    const uint32_t xr = xb + idx * rowb;
                    double b[NT];
#pragma unroll
                    for (int q = 0; q < NT; q++) b[q] = lds_f64(xr + 64u * q, tok);
#pragma unroll
                    for (int q = 0; q < NT; q++) dmma884(acc[q][0], acc[q][1], a, b[q]);
                }
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 5] = clock64();
                // epilogue after the panel's last chunk (c == 0): the results are formed IN PLACE in the accumulators while the
                // slot is still held (own X rows from shared memory), then the slot is released, then Y is stored -- the burst
                // of stores (a whole panel of Y) no longer sits between two stages of the ring
                const int r = 8 * warp + lr;
                const bool epi = c == 0 && !(A.dbg & 4) && r < nrows;
                const int row = row_base + r;
                if (epi) {
                    const uint32_t xs = (r < n0 ? sb + (uint32_t)r * rowb : sb1 + (uint32_t)(r - n0) * rowb) + 16u * lk;
                    if (EPI == EPI_POST) {
                        // R in two batches (registers: a batch of R + accumulators + dot partials), each one L2 round trip
                        const double dw = __ldg(dwk + row);
                        constexpr int H = (NT + 1) / 2;
#pragma unroll
                        for (int half = 0; half < 2; half++) {
                            double rres[H][2];
#pragma unroll
                            for (int qq = 0; qq < H; qq++) {
                                const int q = half * H + qq;
                                rres[qq][0] = 0.0; rres[qq][1] = 0.0;
                                if (q < NT && q < nt) {
                                    const int lc = 8 * q + 2 * lk, col = cs + lc;
                                    const bool ok0 = lc < wc && col >= v0 && col < v1, ok1 = lc + 1 < wc && col + 1 >= v0 && col + 1 < v1;
                                    const size_t o = (size_t)row * A.ld + col;
                                    if (ok0 && ok1) { const double2 rv = __ldg(reinterpret_cast<const double2 *>(A.ex.R + o)); rres[qq][0] = rv.x; rres[qq][1] = rv.y; }
                                    else { if (ok0) rres[qq][0] = __ldg(A.ex.R + o); if (ok1) rres[qq][1] = __ldg(A.ex.R + o + 1); }
                                }
                            }
#pragma unroll
                            for (int qq = 0; qq < H; qq++) {
                                const int q = half * H + qq;
                                if (q < NT && q < nt) {
                                    const int lc = 8 * q + 2 * lk, col = cs + lc;
                                    const bool ok0 = lc < wc && col >= v0 && col < v1, ok1 = lc + 1 < wc && col + 1 >= v0 && col + 1 < v1;
                                    const double2 x = lds_f64x2(xs + 64u * q, tok);
                                    acc[q][0] = fma(dw, rres[qq][0] - acc[q][0], x.x);
                                    acc[q][1] = fma(dw, rres[qq][1] - acc[q][1], x.y);
                                    if (DOT) { if (ok0) part[q][0] = fma(rres[qq][0], acc[q][0], part[q][0]); if (ok1) part[q][1] = fma(rres[qq][1], acc[q][1], part[q][1]); }
                                }
                            }
                        }
                    } else if (EPI == EPI_RESIDUAL || DOT) {
#pragma unroll
                        for (int q = 0; q < NT; q++) {
                            if (q < nt) {
                                const int lc = 8 * q + 2 * lk, col = cs + lc;
                                const bool ok0 = lc < wc && col >= v0 && col < v1, ok1 = lc + 1 < wc && col + 1 >= v0 && col + 1 < v1;
                                const double2 x = lds_f64x2(xs + 64u * q, tok);
                                if (EPI == EPI_RESIDUAL) { acc[q][0] = x.x - acc[q][0]; acc[q][1] = x.y - acc[q][1]; }
                                else { if (ok0) part[q][0] = fma(acc[q][0], x.x, part[q][0]); if (ok1) part[q][1] = fma(acc[q][1], x.y, part[q][1]); }
                            }
                        }
                    }
                }
                __syncwarp();
                if (c == 1 && n0 < nrows) { sb1 = sb; slot1 = slot; }        // released after the epilogue
                else if (lane == 0) {
                    mbar_arrive(&empty[slot]);
                    if (c == 0 && n0 < nrows) mbar_arrive(&empty[slot1]);
                }
                if (epi && !(A.dbg & 16)) {
#pragma unroll
                    for (int q = 0; q < NT; q++) {
                        if (q < nt) {
                            const int lc = 8 * q + 2 * lk, col = cs + lc;
                            const bool ok0 = lc < wc && col >= v0 && col < v1, ok1 = lc + 1 < wc && col + 1 >= v0 && col + 1 < v1;
                            const size_t o = (size_t)row * A.ld + col;
                            if (ok0 && ok1) *reinterpret_cast<double2 *>(A.Y + o) = make_double2(acc[q][0], acc[q][1]);
                            else { if (ok0) A.Y[o] = acc[q][0]; if (ok1) A.Y[o + 1] = acc[q][1]; }
                        }
                    }
                }
                if (stamp) A.dbg_buf[(size_t)dbg_stage * 8 + 6] = clock64();
                dbg_stage++;
                if (++slot == (uint32_t)S) { slot = 0; phase ^= 1u; }
            }
            p = pn;
        }
        if (DOT && !(A.dbg & 64)) {
            // rows of the fragment first (lanes with equal lk, fixed shuffle order), then a fixed-order sum over the warps,
            // one partial row per CTA, the last CTA of the tile adds the rows in index order
#pragma unroll
            for (int q = 0; q < NT; q++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    double v = part[q][h];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    part[q][h] = v;
                }
            }
            const int nslot = (A.n_tiles <= (int)gridDim.x) ? A.cpt : 1;
            for (int w = 0; w < MM_CONSUMER_WARPS; w++) {
                if (warp == w && lr == 0) {
#pragma unroll
                    for (int q = 0; q < NT; q++) {
                        if (q < nt) {
                            const int i = 8 * q + 2 * lk;
                            sdot[i] = (w == 0 ? 0.0 : sdot[i]) + part[q][0];
                            sdot[i + 1] = (w == 0 ? 0.0 : sdot[i + 1]) + part[q][1];
                        }
                    }
                }
                named_bar_sync(1, MM_CONSUMERS);
            }
            double *mine = A.dot_part + (size_t)j * A.ld;
            for (int i = threadIdx.x; i < wc; i += MM_CONSUMERS) if (cs + i >= v0 && cs + i < v1) mine[cs + i] = sdot[i];
            __threadfence();
            named_bar_sync(1, MM_CONSUMERS);
            if (threadIdx.x == 0) s_ticket = (int)atomicAdd(A.dot_counter + t, 1u);
            named_bar_sync(1, MM_CONSUMERS);
            if (s_ticket == nslot - 1) {
                __threadfence();
                const int i = threadIdx.x >> 2, part4 = threadIdx.x & 3;
                for (int i0 = 0; i0 < wc; i0 += MM_CONSUMERS / 4) {
                    const int cc = cs + i0 + i;
                    const bool okc = (i0 + i) < wc && cc >= v0 && cc < v1;
                    double s4 = okc ? ordered_row_sum(A.dot_part + cc, A.ld, part4, 4, nslot) : 0.0;
                    const double s1 = __shfl_down_sync(0xffffffffu, s4, 1), s2 = __shfl_down_sync(0xffffffffu, s4, 2), s3 = __shfl_down_sync(0xffffffffu, s4, 3);
                    if (okc && part4 == 0) A.dots[cc] = ((s4 + s1) + s2) + s3;
                }
                if (threadIdx.x == 0) A.dot_counter[t] = 0u;
            }
            named_bar_sync(1, MM_CONSUMERS);
        }
    }
}

// ---------------------------------------------------------------------------------
// K2-narrow: the SpMM of NARROW column windows (the source shards of a multi-GPU run: 13 - 32 of 100 columns).  A narrow block
// vector fits L2 many times over and a stage of the streamed kernels would carry a few KB, so their per-stage latencies
// dominate (measured: 86 us for 13 columns against 124 us for 100).  Here nothing is staged: a warp takes an 8-row group,
// walks its k-steps (stream_panels.h, gather form) -- A fragment from global memory, the four rows of X through the read-only
// path -- and feeds the same DMMA m8n8k4 tiles; 32 warps per SM hide the L2 latency.  Same three epilogues, same
// deterministic dots (per-CTA partial rows, last ticket adds them in order).
// ---------------------------------------------------------------------------------
struct GatherArgs {
    const int *ks_ptr, *cols;            // gather form of the level
    const double *aval; size_t n_frag;   // packed A fragments [nK][n_frag]
    int n_groups, n_rows;
    const double *X; double *Y; size_t ld;
    int kk, cs, wc, v0, v1;              // wavenumber group; copied columns [cs, cs + wc) with cs even; valid columns [v0, v1)
    int dbg;                             // measurement only (profiles/bench_spmm.py): 64 = no dot tail, 256 = no final sum of the last CTA
    double *dot_part; unsigned *dot_counter; double *dots;
    PanelExtra ex;
};
constexpr int GA_THREADS = 256;
template <int NT, int EPI, bool DOT>
__global__ void __launch_bounds__(GA_THREADS, NT <= 2 ? 4 : 3)       // latency-bound: as many warps per SM as the registers allow
k_spmm_gather(const GatherArgs A) {
    __shared__ double sdot[8 * NT];
    __shared__ int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
    const int nwarps = GA_THREADS / 32;
    const double *av = A.aval + (size_t)A.kk * A.n_frag + lane;
    const int ldm1 = (int)A.ld - 1;
    // column of the lane's B element in n-tile q (clamped: the values of columns outside the window are never stored)
    int bcol[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) bcol[q] = min(A.cs + 8 * q + lr, ldm1);
    double part[DOT ? NT : 1][2];
#pragma unroll
    for (int q = 0; q < (DOT ? NT : 1); q++) { part[q][0] = 0.0; part[q][1] = 0.0; }
    const double *dwk = (EPI == EPI_POST) ? A.ex.dinvw + (size_t)A.kk * A.ex.n : nullptr;
    const int gstep = gridDim.x * nwarps;
    int g = blockIdx.x * nwarps + warp;
    int k0 = 0, k1 = 0;
    if (g < A.n_groups) { k0 = __ldg(A.ks_ptr + g); k1 = __ldg(A.ks_ptr + g + 1); }
    for (; g < A.n_groups; g += gstep) {
        // the NEXT group of this warp: its k-step range is fetched now, and once it is known its A fragments and column
        // lists are prefetched into L2 (they stream from DRAM exactly once; the demand loads then see L2 latency)
        int k0n = 0, k1n = 0;
        if (g + gstep < A.n_groups) { k0n = __ldg(A.ks_ptr + g + gstep); k1n = __ldg(A.ks_ptr + g + gstep + 1); }
        double acc[NT][2];
#pragma unroll
        for (int q = 0; q < NT; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
        constexpr int KU = NT <= 2 ? 4 : 2;          // k-steps in flight per warp
        for (int ks = k0; ks < k1; ks += KU) {
            int col[KU]; double a[KU]; double b[KU][NT];
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const bool ok = ks + u < k1;
                col[u] = ok ? __ldg(A.cols + 4 * (size_t)(ks + u) + lk) : 0;
                a[u] = ok ? __ldcs(av + 32 * (size_t)(ks + u)) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const double *xr = A.X + (size_t)col[u] * A.ld;
#pragma unroll
                for (int q = 0; q < NT; q++) b[u][q] = __ldg(xr + bcol[q]);
            }
#pragma unroll
            for (int u = 0; u < KU; u++)
#pragma unroll
                for (int q = 0; q < NT; q++) dmma884(acc[q][0], acc[q][1], a[u], b[u][q]);
        }
        {
            // 128-byte lines of the next group's A fragments (2 per k-step) and column list (1 per 8 k-steps)
            const int nl = 2 * (k1n - k0n);
            for (int i = lane; i < nl; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.aval + (size_t)A.kk * A.n_frag + 32 * (size_t)k0n + 16 * (size_t)i));
            if (lane < (k1n - k0n + 7) / 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.cols + 4 * (size_t)k0n + 32 * (size_t)lane));
        }
        const int row = 8 * g + lr;
        if (row < A.n_rows) {
            const double dw = (EPI == EPI_POST) ? __ldg(dwk + row) : 0.0;
#pragma unroll
            for (int q = 0; q < NT; q++) {
                const int lc = 8 * q + 2 * lk, colq = A.cs + lc;
                const bool ok0 = lc < A.wc && colq >= A.v0 && colq < A.v1, ok1 = lc + 1 < A.wc && colq + 1 >= A.v0 && colq + 1 < A.v1;
                if (!(ok0 || ok1)) continue;
                const size_t o = (size_t)row * A.ld + colq;
                double x0 = 0.0, x1 = 0.0, r0 = 0.0, r1 = 0.0;
                if (EPI != EPI_SPMM || DOT) {
                    if (ok0 && ok1) { const double2 v = __ldg(reinterpret_cast<const double2 *>(A.X + o)); x0 = v.x; x1 = v.y; }
                    else { if (ok0) x0 = __ldg(A.X + o); if (ok1) x1 = __ldg(A.X + o + 1); }
                }
                if (EPI == EPI_POST) {
                    if (ok0 && ok1) { const double2 v = __ldg(reinterpret_cast<const double2 *>(A.ex.R + o)); r0 = v.x; r1 = v.y; }
                    else { if (ok0) r0 = __ldg(A.ex.R + o); if (ok1) r1 = __ldg(A.ex.R + o + 1); }
                }
                double y0, y1;
                if (EPI == EPI_POST) {
                    y0 = fma(dw, r0 - acc[q][0], x0); y1 = fma(dw, r1 - acc[q][1], x1);
                    if (DOT) { if (ok0) part[q][0] = fma(r0, y0, part[q][0]); if (ok1) part[q][1] = fma(r1, y1, part[q][1]); }
                } else if (EPI == EPI_RESIDUAL) {
                    y0 = x0 - acc[q][0]; y1 = x1 - acc[q][1];
                } else {
                    y0 = acc[q][0]; y1 = acc[q][1];
                    if (DOT) { if (ok0) part[q][0] = fma(y0, x0, part[q][0]); if (ok1) part[q][1] = fma(y1, x1, part[q][1]); }
                }
                if (ok0 && ok1) *reinterpret_cast<double2 *>(A.Y + o) = make_double2(y0, y1);
                else { if (ok0) A.Y[o] = y0; if (ok1) A.Y[o + 1] = y1; }
            }
        }
        k0 = k0n; k1 = k1n;
    }
    if (DOT && !(A.dbg & 64)) {
#pragma unroll
        for (int q = 0; q < NT; q++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                double v = part[q][hh];
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                part[q][hh] = v;
            }
        for (int w = 0; w < nwarps; w++) {
            if (warp == w && lr == 0) {
#pragma unroll
                for (int q = 0; q < NT; q++) {
                    const int i = 8 * q + 2 * lk;
                    sdot[i] = (w == 0 ? 0.0 : sdot[i]) + part[q][0];
                    sdot[i + 1] = (w == 0 ? 0.0 : sdot[i + 1]) + part[q][1];
                }
            }
            __syncthreads();
        }
        double *mine = A.dot_part + (size_t)blockIdx.x * A.ld;
        for (int i = threadIdx.x; i < A.wc; i += GA_THREADS) if (A.cs + i >= A.v0 && A.cs + i < A.v1) mine[A.cs + i] = sdot[i];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = (int)atomicAdd(A.dot_counter, 1u);
        __syncthreads();
        if (s_ticket == (int)gridDim.x - 1 && !(A.dbg & 256)) {
            // last CTA: GA_THREADS / (8 NT) threads per column add interleaved slices of the partial rows (fixed order), then
            // one thread per column adds the slice sums in order
            __threadfence();
            __shared__ double sfin[GA_THREADS];
            constexpr int WP = 8 * NT, SL = GA_THREADS / WP;
            const int c = threadIdx.x % WP, sl = threadIdx.x / WP, cc = A.cs + c;
            const bool okc = c < A.wc && cc >= A.v0 && cc < A.v1;
            sfin[threadIdx.x] = okc ? ordered_row_sum(A.dot_part + cc, A.ld, sl, SL, (int)gridDim.x) : 0.0;
            __syncthreads();
            if (sl == 0 && okc) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < SL; j++) v += sfin[j * WP + c];
                A.dots[cc] = v;
            }
            if (threadIdx.x == 0) *A.dot_counter = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------
// block-PCG vector kernels (one independent CG per source column, all columns advance in the same launches).
// Flat mapping: the FLAT_T threads of a CTA tile the column window [c0,c1) (chunks of at most cw columns along
// blockIdx.y) as rpp = FLAT_T / w consecutive rows of w columns, so that a warp touches 32 consecutive elements of the
// row-major block whatever the window width (100 columns -> 5 rows per pass, 14 columns of an 8-GPU shard -> 36), and
// every thread keeps ONE column (per-column scalars and dot partials stay in registers).  The rows are split evenly
// over the CTAs of a column chunk (contiguous ranges, flat_rows).
// Per-column dot products are DETERMINISTIC (no floating-point atomics): thread partials meet in shared memory in a
// fixed order, every CTA writes one partial row, the CTA that draws the last ticket adds the rows in index order.
// ---------------------------------------------------------------------------------
constexpr int FLAT_T = 512;
struct FlatMap { int col, roff, rpp, w; bool active; };
__device__ __forceinline__ FlatMap flat_map(int c0, int c1, int cw) {
    FlatMap f;
    const int cs = c0 + blockIdx.y * cw;
    f.w = min(cw, c1 - cs);
    f.rpp = FLAT_T / f.w;
    f.roff = threadIdx.x / f.w;
    f.col = cs + threadIdx.x - f.roff * f.w;
    f.active = f.roff < f.rpp;
    return f;
}
// contiguous, balanced row range of this CTA
__device__ __forceinline__ void flat_rows(int n, int &lo, int &hi) {
    lo = (int)(((long long)n * blockIdx.x) / gridDim.x);
    hi = (int)(((long long)n * (blockIdx.x + 1)) / gridDim.x);
}
struct DotOut {
    double *part;          // [2][slots][ld] partial rows (plane 0 / 1 for the two sums a kernel may produce)
    size_t plane;          // slots * ld
    unsigned *counter;     // [column chunks] tickets, zero between launches
    double *out0, *out1;   // final per-column sums (nullptr: that sum is not wanted)
    size_t ld;
};
// v0, v1: this thread's partial sums for its column.  All threads of the CTA must call.
__device__ __forceinline__ void flat_col_finalize(double v0, double v1, const FlatMap &f, const DotOut &D) {
    __shared__ double red0[FLAT_T], red1[FLAT_T];
    __shared__ int ticket;
    red0[threadIdx.x] = f.active ? v0 : 0.0; red1[threadIdx.x] = f.active ? v1 : 0.0;
    __syncthreads();
    if ((int)threadIdx.x < f.w) {
        double a = 0.0, b = 0.0;
        for (int j = 0; j < f.rpp; j++) { a += red0[threadIdx.x + j * f.w]; b += red1[threadIdx.x + j * f.w]; }
        double *row = D.part + (size_t)blockIdx.x * D.ld + f.col;
        if (D.out0) row[0] = a;
        if (D.out1) row[D.plane] = b;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = (int)atomicAdd(D.counter + blockIdx.y, 1u);
    __syncthreads();
    if (ticket != (int)gridDim.x - 1) return;
    __threadfence();
    // last CTA of this column chunk: thread (roff, col) adds rows roff, roff + rpp, ...; then the roff partials in order
    double a = 0.0, b = 0.0;
    if (f.active) {
        if (D.out0) a = ordered_row_sum(D.part + f.col, D.ld, f.roff, f.rpp, (int)gridDim.x);
        if (D.out1) b = ordered_row_sum(D.part + D.plane + f.col, D.ld, f.roff, f.rpp, (int)gridDim.x);
    }
    __syncthreads();
    red0[threadIdx.x] = a; red1[threadIdx.x] = b;
    __syncthreads();
    if ((int)threadIdx.x < f.w) {
        double sa = 0.0, sb = 0.0;
        for (int j = 0; j < f.rpp; j++) { sa += red0[threadIdx.x + j * f.w]; sb += red1[threadIdx.x + j * f.w]; }
        if (D.out0) D.out0[f.col] = sa;
        if (D.out1) D.out1[f.col] = sb;
    }
    if (threadIdx.x == 0) D.counter[blockIdx.y] = 0u;
}

// x = 0; r = b; p = z = Dinv r (Jacobi; with the multilevel preconditioner p is set by the first cycle); rz = r.z; bb = b.b
// AX != nullptr: warm start from the x the buffer already holds (the potentials of the previous model): r = b - A x
__global__ void __launch_bounds__(FLAT_T)
k_pcg_init(const double *__restrict__ B, const double *__restrict__ dinv, double *__restrict__ Xv, double *__restrict__ R,
           double *__restrict__ P, const double *__restrict__ AX, int N, int nE, int c0, int c1, size_t ld, int cw, const DotOut D) {
    const FlatMap f = flat_map(c0, c1, cw);
    double s0 = 0.0, s1 = 0.0;
    if (f.active) {
        const double *dk = dinv + (size_t)(f.col / nE) * N;
        int lo, hi; flat_rows(N, lo, hi);
        for (int row = lo + f.roff; row < hi; row += f.rpp) {
            const size_t o = (size_t)row * ld + f.col;
            const double b = B[o];
            double r = b;
            if (AX) r = b - AX[o]; else Xv[o] = 0.0;
            const double z = dk[row] * r;
            R[o] = r; P[o] = z;
            s0 = fma(r, z, s0); s1 = fma(b, b, s1);
        }
    }
    flat_col_finalize(s0, s1, f, D);
}

// alpha = rz/pAp;  x += alpha p;  r -= alpha Ap;  rz_new = r.Dinv r (Jacobi only);  rr = r.r
template <bool JACOBI>
__global__ void __launch_bounds__(FLAT_T)
k_pcg_update_xr(const double *__restrict__ P, const double *__restrict__ AP, const double *__restrict__ dinv,
                double *__restrict__ Xv, double *__restrict__ R, int N, int nE, int c0, int c1, size_t ld,
                const double *__restrict__ rz, const double *__restrict__ pAp, int cw, const DotOut D) {
    const FlatMap f = flat_map(c0, c1, cw);
    double s0 = 0.0, s1 = 0.0;
    if (f.active) {
        const double den = pAp[f.col], num = rz[f.col];
        const double alpha = (den > 0.0 && num > 0.0) ? num / den : 0.0;
        const double *dk = JACOBI ? dinv + (size_t)(f.col / nE) * N : nullptr;
        int lo, hi; flat_rows(N, lo, hi);
#pragma unroll 4
        for (int row = lo + f.roff; row < hi; row += f.rpp) {
            const size_t o = (size_t)row * ld + f.col;
            const double rn = fma(-alpha, AP[o], R[o]);
            Xv[o] = fma(alpha, P[o], Xv[o]);
            R[o] = rn;
            if (JACOBI) s0 = fma(rn * dk[row], rn, s0);
            s1 = fma(rn, rn, s1);
        }
    }
    flat_col_finalize(s0, s1, f, D);     // D.out0 = rz_new (Jacobi) or nullptr, D.out1 = rr
}
// beta = rz_new/rz;  p = z + beta p   (p = 0 once the column has converged: it freezes)
// JACOBI: z = Dinv r on the fly; otherwise R points at the preconditioned residual Z and dinv is unused
template <bool JACOBI>
__global__ void __launch_bounds__(FLAT_T)
k_pcg_update_p(const double *__restrict__ R, const double *__restrict__ dinv, double *__restrict__ P, int N, int nE,
               int c0, int c1, size_t ld, const double *__restrict__ rz, const double *__restrict__ rz_new,
               const double *__restrict__ rr, const double *__restrict__ bb, double tol2, int cw) {
    const FlatMap f = flat_map(c0, c1, cw);
    if (!f.active) return;
    const int col = f.col;
    const bool done = !(rr[col] > tol2 * bb[col]);
    const double den = rz[col];
    const double beta = (den > 0.0) ? rz_new[col] / den : 0.0;
    const double *dk = JACOBI ? dinv + (size_t)(col / nE) * N : nullptr;
    int lo, hi; flat_rows(N, lo, hi);
#pragma unroll 4
    for (int row = lo + f.roff; row < hi; row += f.rpp) {
        const size_t o = (size_t)row * ld + col;
        P[o] = done ? 0.0 : fma(beta, P[o], JACOBI ? dk[row] * R[o] : R[o]);
    }
}

// ---------------------------------------------------------------------------------
// complex resistivity (reference: DCMultiElectrodeModelling with setComplex(true),
// dcfemmodelling.cpp:235-242 CSparseMatrix assembly, :1755-1925 complex total-field solves, :1410-1444 complex Jacobian).
// S = S_r + i S_i is complex SYMMETRIC (sigma = 1/rho complex per cell, real element matrices): solved by COCG
// (conjugate orthogonal CG: the CG recurrence with the unconjugated product x^T y) preconditioned by the real multilevel
// cycle of S_r applied to real and imaginary parts alike.  Layout: the handle is built with every electrode listed twice
// (plan with 2 nE "electrodes"): inside wavenumber group k, columns [0, nE) hold the REAL parts and [nE, 2 nE) the IMAGINARY
// parts of the nE complex sources, so every real kernel of the path (SpMM, multilevel cycle, pick-up, Jacobian) runs
// unchanged on the 2 nS real columns.  The kernels below work per COMPLEX source j (flat mapping over [0, nS/2)).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int cplx_col(int j, int nEc) { return (j / nEc) * 2 * nEc + (j % nEc); }   // real part; imaginary: + nEc

// 1 / sigma_r and 1 / sigma_i per cell for the two assembly passes (sigma = 1 / (rho_r + i rho_i)); flag: |rho| <= 1e-12
__global__ void k_cplx_sigma(const double *__restrict__ rr, const double *__restrict__ ri, int C, double *__restrict__ inv_sr,
                             double *__restrict__ inv_si, int *flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double a = rr[c], b = ri[c], d = a * a + b * b;
    if (!(d > 1e-24)) { atomicOr(flag, 1); inv_sr[c] = 0.0; inv_si[c] = 0.0; return; }     // both passes skip the cell
    inv_sr[c] = d / a;            // sigma_r =  a / d
    inv_si[c] = -d / b;           // sigma_i = -b / d   (b == 0: -inf -> contributes 0)
}
__global__ void k_set_slots(const int *__restrict__ slots, int n, int nK, size_t nnz, double value, double *__restrict__ vals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) for (int kk = 0; kk < nK; kk++) vals[(size_t)kk * nnz + slots[t]] = value;
}
// right-hand sides: the imaginary-part columns of the doubled layout are zero (real unit currents)
__global__ void k_cplx_zero_imag_cols(double *__restrict__ B, int N, int nEc, int nS2, size_t ld) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * nS2) return;
    const int row = (int)(i / nS2), c = (int)(i % nS2);
    if ((c % (2 * nEc)) >= nEc) B[(size_t)row * ld + c] = 0.0;
}
// out0[j] + i out1[j] = sum_rows a b (unconjugated);  NORM: out0[j] = sum |a|^2 (b unused)
template <bool NORM>
__global__ void __launch_bounds__(FLAT_T)
k_cplx_dot(const double *__restrict__ Av, const double *__restrict__ Bv, int N, int nEc, int nSc, size_t ld, int cw, const DotOut D) {
    const FlatMap f = flat_map(0, nSc, cw);
    double s0 = 0.0, s1 = 0.0;
    if (f.active) {
        const int cr = cplx_col(f.col, nEc), ci = cr + nEc;
        int lo, hi; flat_rows(N, lo, hi);
        for (int row = lo + f.roff; row < hi; row += f.rpp) {
            const size_t o = (size_t)row * ld;
            const double ar = Av[o + cr], ai = Av[o + ci];
            if (NORM) { s0 = fma(ar, ar, s0); s0 = fma(ai, ai, s0); }
            else { const double br = Bv[o + cr], bi = Bv[o + ci]; s0 += ar * br - ai * bi; s1 += ar * bi + ai * br; }
        }
    }
    flat_col_finalize(s0, s1, f, D);
}
// q = (S_r + i S_i) p from Y1 = S_r [p_r | p_i] and Y2 = S_i [p_r | p_i]:  q_r = Y1_r - Y2_i,  q_i = Y2_r + Y1_i  (written
// over Y1);  out = p^T q
__global__ void __launch_bounds__(FLAT_T)
k_cplx_combine(double *__restrict__ Y1, const double *__restrict__ Y2, const double *__restrict__ P, int N, int nEc, int nSc, size_t ld,
               int cw, const DotOut D) {
    const FlatMap f = flat_map(0, nSc, cw);
    double s0 = 0.0, s1 = 0.0;
    if (f.active) {
        const int cr = cplx_col(f.col, nEc), ci = cr + nEc;
        int lo, hi; flat_rows(N, lo, hi);
        for (int row = lo + f.roff; row < hi; row += f.rpp) {
            const size_t o = (size_t)row * ld;
            const double qr = Y1[o + cr] - Y2[o + ci], qi = Y2[o + cr] + Y1[o + ci];
            Y1[o + cr] = qr; Y1[o + ci] = qi;
            const double pr = P[o + cr], pi = P[o + ci];
            s0 += pr * qr - pi * qi; s1 += pr * qi + pi * qr;
        }
    }
    flat_col_finalize(s0, s1, f, D);
}
// alpha = rho / (p^T q);  x += alpha p;  r -= alpha q;  out1 = |r|^2     (scalars: [re | im] at j and ldS + j)
__global__ void __launch_bounds__(FLAT_T)
k_cplx_update_xr(const double *__restrict__ P, const double *__restrict__ Q, double *__restrict__ Xv, double *__restrict__ R, int N, int nEc,
                 int nSc, size_t ld, const double *__restrict__ rho, const double *__restrict__ pq, size_t ldS, int cw, const DotOut D) {
    const FlatMap f = flat_map(0, nSc, cw);
    double s1 = 0.0;
    if (f.active) {
        const int cr = cplx_col(f.col, nEc), ci = cr + nEc;
        const double nr = rho[f.col], ni = rho[ldS + f.col], dr = pq[f.col], di = pq[ldS + f.col];
        const double dd = dr * dr + di * di;
        const double ar = dd > 0.0 ? (nr * dr + ni * di) / dd : 0.0, ai = dd > 0.0 ? (ni * dr - nr * di) / dd : 0.0;
        int lo, hi; flat_rows(N, lo, hi);
        for (int row = lo + f.roff; row < hi; row += f.rpp) {
            const size_t o = (size_t)row * ld;
            const double pr = P[o + cr], pi = P[o + ci], qr = Q[o + cr], qi = Q[o + ci];
            Xv[o + cr] += ar * pr - ai * pi; Xv[o + ci] += ar * pi + ai * pr;
            const double rr = R[o + cr] - (ar * qr - ai * qi), ri = R[o + ci] - (ar * qi + ai * qr);
            R[o + cr] = rr; R[o + ci] = ri;
            s1 = fma(rr, rr, s1); s1 = fma(ri, ri, s1);
        }
    }
    flat_col_finalize(0.0, s1, f, D);
}
// beta = rho_new / rho;  p = z + beta p   (p = 0 once the column has converged)
__global__ void __launch_bounds__(FLAT_T)
k_cplx_update_p(const double *__restrict__ Z, double *__restrict__ P, int N, int nEc, int nSc, size_t ld, const double *__restrict__ rho,
                const double *__restrict__ rho_new, size_t ldS, const double *__restrict__ rr, const double *__restrict__ bb, double tol2, int cw) {
    const FlatMap f = flat_map(0, nSc, cw);
    if (!f.active) return;
    const int cr = cplx_col(f.col, nEc), ci = cr + nEc;
    const bool done = !(rr[f.col] > tol2 * bb[f.col]);
    const double nr = rho_new[f.col], ni = rho_new[ldS + f.col], dr = rho[f.col], di = rho[ldS + f.col];
    const double dd = dr * dr + di * di;
    const double br = dd > 0.0 ? (nr * dr + ni * di) / dd : 0.0, bi = dd > 0.0 ? (ni * dr - nr * di) / dd : 0.0;
    int lo, hi; flat_rows(N, lo, hi);
    for (int row = lo + f.roff; row < hi; row += f.rpp) {
        const size_t o = (size_t)row * ld;
        const double pr = P[o + cr], pi = P[o + ci];
        P[o + cr] = done ? 0.0 : Z[o + cr] + (br * pr - bi * pi);
        P[o + ci] = done ? 0.0 : Z[o + ci] + (br * pi + bi * pr);
    }
}
// complex Jacobian from the four real blocks of the doubled scheme (rows [0,D): ab_r.mn_r, [D,2D): ab_i.mn_i, [2D,3D):
// ab_r.mn_i, [3D,4D): ab_i.mn_r):  J = ((J0 - J1) + i (J2 + J3)) * k_d / m_col^2   (:1410-1444, no conjugation);
// out: row-major [D][M] interleaved (re, im).  scale == 0: unscaled (model length != columns)
__global__ void k_cplx_jacobian(const double *__restrict__ Jt, size_t ldJ, int D, int M, const double *__restrict__ kfac,
                                const double *__restrict__ m_re, const double *__restrict__ m_im, int scale, double *__restrict__ out) {
    __shared__ double t0[32][33], t1[32][33];
    const int c0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int col = c0 + jj, d = d0 + threadIdx.x;
        double re = 0.0, im = 0.0;
        if (col < M && d < D) {
            const double *cj = Jt + (size_t)col * ldJ;
            re = cj[d] - cj[D + d]; im = cj[2 * D + d] + cj[3 * D + d];
            if (scale) {
                const double a = m_re[col], b = m_im[col];
                const double sr = a * a - b * b, si = 2.0 * a * b, dd = sr * sr + si * si;      // m^2
                const double k = kfac[d];
                const double xr = (re * sr + im * si) / dd * k, xi = (im * sr - re * si) / dd * k;
                re = xr; im = xi;
            }
        }
        t0[jj][threadIdx.x] = re; t1[jj][threadIdx.x] = im;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int d = d0 + jj, col = c0 + threadIdx.x;
        if (d < D && col < M) { out[2 * ((size_t)d * M + col)] = t0[threadIdx.x][jj]; out[2 * ((size_t)d * M + col) + 1] = t1[threadIdx.x][jj]; }
    }
}

// ---------------------------------------------------------------------------------
// multilevel preconditioner (unsmoothed aggregation, V(1,1), damped Jacobi) -- see amg_setup.py
// All kernels work on node-major block vectors of one level and the column window [c0,c1).
// ---------------------------------------------------------------------------------
// coarse matrix values: plain sums of finer-level entries (Galerkin product with piecewise-constant P)
__global__ void k_galerkin(const int *__restrict__ gal_ptr, const int *__restrict__ gal_idx, int n_slots, int nK,
                           size_t nnz_f, size_t nnz_c, const double *__restrict__ vals_f, double *__restrict__ vals_c) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int b = gal_ptr[s], e = gal_ptr[s + 1];
    for (int kk = 0; kk < nK; kk++) {
        double acc = 0.0;
        for (int p = b; p < e; p++) acc += vals_f[(size_t)kk * nnz_f + gal_idx[p]];
        vals_c[(size_t)kk * nnz_c + s] = acc;
    }
}
// Gershgorin bound of lambda_max(D^-1 A) per wavenumber: max_i sum_j |a_ij| / a_ii  (positive doubles order like integers)
__global__ void k_row_ratio(const int *__restrict__ rowptr, const int *__restrict__ diag_pos, int n, int nK, size_t nnz,
                            const double *__restrict__ vals, unsigned long long *gmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int kk = 0; kk < nK; kk++) {
        const double *v = vals + (size_t)kk * nnz;
        double s = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) s += fabs(v[p]);
        const double g = s / v[diag_pos[i]];
        if (g > 0.0) atomicMax(gmax + kk, (unsigned long long)__double_as_longlong(g));
    }
}
// dinvw = omega / a_ii with omega = 1.6 / max(2, Gershgorin bound): damped Jacobi that stays convergent
__global__ void k_inv_diag_w(const int *__restrict__ diag_pos, int n, int nK, size_t nnz, const double *__restrict__ vals,
                             const unsigned long long *__restrict__ gmax, double *__restrict__ dinvw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int kk = 0; kk < nK; kk++) {
        const double g = __longlong_as_double((long long)gmax[kk]);
        const double omega = 1.6 / fmax(2.0, g);
        dinvw[(size_t)kk * n + i] = omega / vals[(size_t)kk * nnz + diag_pos[i]];
    }
}

// vals_dw[k][p] = vals[k][p] * dinvw[k][col(p)]
__global__ void k_scale_cols(const int *__restrict__ colidx, size_t nnz, int n, int nK, const double *__restrict__ vals,
                             const double *__restrict__ dinvw, double *__restrict__ vals_dw) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int j = colidx[p];
    for (int kk = 0; kk < nK; kk++) vals_dw[(size_t)kk * nnz + p] = vals[(size_t)kk * nnz + p] * dinvw[(size_t)kk * n + j];
}

constexpr int AMG_TX = 16, AMG_TY = 8;   // rows per CTA are a launch parameter: small (coarse) levels use fewer rows for more CTAs

// Z = X + dw .* (R - A X)   (post-smoothing / Jacobi sweep); optional fused dot  sum_i R_i Z_i
template <int CPT, bool DOT>
__global__ void __launch_bounds__(AMG_TX * AMG_TY)
k_amg_post(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals, size_t nnz,
           const double *__restrict__ dinvw, int n, const double *__restrict__ X, const double *__restrict__ R,
           double *__restrict__ Z, int nE, int c0, int c1, size_t ld, double *__restrict__ dots, int AMG_ROWS) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int cbase = c0 + blockIdx.y * (AMG_TX * CPT) + tx;
    int col[CPT]; size_t voff[CPT]; const double *dw[CPT]; bool ok[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) {
        const int c = cbase + m * AMG_TX;
        ok[m] = c < c1;
        col[m] = ok[m] ? c : c0;
        const int kk = col[m] / nE;
        voff[m] = (size_t)kk * nnz; dw[m] = dinvw + (size_t)kk * n;
    }
    double part[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) part[m] = 0.0;
    const int row0 = blockIdx.x * AMG_ROWS;
    for (int r = ty; r < AMG_ROWS; r += AMG_TY) {
        const int row = row0 + r;
        if (row >= n) break;
        double acc[CPT];
#pragma unroll
        for (int m = 0; m < CPT; m++) acc[m] = 0.0;
        // the entries of a row are independent: batches of 4 keep 4 index loads, then 4*CPT gathers, in flight (the
        // coarse levels are latency-bound, not bandwidth-bound: a few thousand rows, ~15 entries each)
        const int pb = rowptr[row], pe = rowptr[row + 1];
        int p = pb;
        for (; p + 4 <= pe; p += 4) {
            size_t xo[4];
#pragma unroll
            for (int u = 0; u < 4; u++) xo[u] = (size_t)__ldg(colidx + p + u) * ld;
            double a[4][CPT], x[4][CPT];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int m = 0; m < CPT; m++) { a[u][m] = __ldg(vals + voff[m] + p + u); x[u][m] = __ldg(X + xo[u] + col[m]); }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int m = 0; m < CPT; m++) acc[m] = fma(a[u][m], x[u][m], acc[m]);
        }
        for (; p < pe; p++) {
            const size_t xo = (size_t)colidx[p] * ld;
#pragma unroll
            for (int m = 0; m < CPT; m++) acc[m] = fma(__ldg(vals + voff[m] + p), __ldg(X + xo + col[m]), acc[m]);
        }
#pragma unroll
        for (int m = 0; m < CPT; m++) {
            if (ok[m]) {
                const size_t o = (size_t)row * ld + col[m];
                const double rr = R[o];
                const double z = fma(dw[m][row], rr - acc[m], X[o]);
                Z[o] = z;
                if (DOT) part[m] = fma(rr, z, part[m]);
            }
        }
    }
    if (DOT) {
        __shared__ double red[AMG_TY][AMG_TX * CPT];
#pragma unroll
        for (int m = 0; m < CPT; m++) red[ty][m * AMG_TX + tx] = part[m];
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int m = 0; m < CPT; m++) {
                double sacc = 0.0;
#pragma unroll
                for (int y = 0; y < AMG_TY; y++) sacc += red[y][m * AMG_TX + tx];
                if (ok[m]) atomicAdd(dots + col[m], sacc);
            }
        }
    }
}

// pre-smoothing from a zero guess + residual + restriction, fused:
//   RC[I] = sum_{i in aggregate I} ( R_i - sum_j a_ij * dw_j * R_j )      (vals_dw[p] = a_ij * dw_j, see k_scale_cols)
template <int CPT>
__global__ void __launch_bounds__(AMG_TX * AMG_TY)
k_amg_restrict(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals_dw, size_t nnz,
               int n_f, const int *__restrict__ mem_ptr, const int *__restrict__ mem_idx, int n_c,
               const double *__restrict__ R, double *__restrict__ RC, int nE, int c0, int c1, size_t ld, int AMG_ROWS) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int cbase = c0 + blockIdx.y * (AMG_TX * CPT) + tx;
    int col[CPT]; size_t voff[CPT]; bool ok[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) {
        const int c = cbase + m * AMG_TX;
        ok[m] = c < c1;
        col[m] = ok[m] ? c : c0;
        voff[m] = (size_t)(col[m] / nE) * nnz;
    }
    (void)n_f;
    const int row0 = blockIdx.x * AMG_ROWS;
    for (int r = ty; r < AMG_ROWS; r += AMG_TY) {
        const int I = row0 + r;
        if (I >= n_c) break;
        double acc[CPT];
#pragma unroll
        for (int m = 0; m < CPT; m++) acc[m] = 0.0;
        for (int q = mem_ptr[I]; q < mem_ptr[I + 1]; q++) {
            const int i = mem_idx[q];
#pragma unroll
            for (int m = 0; m < CPT; m++) acc[m] += __ldg(R + (size_t)i * ld + col[m]);
            const int pb = rowptr[i], pe = rowptr[i + 1];
            int p = pb;
            for (; p + 4 <= pe; p += 4) {                      // batches of 4 independent entries (see k_amg_post)
                size_t xo[4];
#pragma unroll
                for (int u = 0; u < 4; u++) xo[u] = (size_t)__ldg(colidx + p + u) * ld;
                double a[4][CPT], x[4][CPT];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int m = 0; m < CPT; m++) { a[u][m] = __ldg(vals_dw + voff[m] + p + u); x[u][m] = __ldg(R + xo[u] + col[m]); }
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int m = 0; m < CPT; m++) acc[m] = fma(-a[u][m], x[u][m], acc[m]);
            }
            for (; p < pe; p++) {
                const size_t xo = (size_t)colidx[p] * ld;
#pragma unroll
                for (int m = 0; m < CPT; m++) acc[m] = fma(-__ldg(vals_dw + voff[m] + p), __ldg(R + xo + col[m]), acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < CPT; m++) if (ok[m]) RC[(size_t)I * ld + col[m]] = acc[m];
    }
}

// Small levels (a few thousand rows and fewer) are latency-bound: one thread walking the ~15 entries of its row --
// index load, gather, FMA, each dependent on the last -- takes longer than the data movement.  The "split" variants give
// ONE row to a CTA and spread its entries over the AMG_TY thread rows (partial sums meet in shared memory), which cuts
// the dependent chain by 8.
template <int CPT>
__global__ void __launch_bounds__(AMG_TX * AMG_TY)
k_amg_post_split(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals, size_t nnz,
                 const double *__restrict__ dinvw, int n, const double *__restrict__ X, const double *__restrict__ R,
                 double *__restrict__ Z, int nE, int c0, int c1, size_t ld) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.x;
    const int cbase = c0 + blockIdx.y * (AMG_TX * CPT) + tx;
    int col[CPT]; size_t voff[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) {
        const int c = cbase + m * AMG_TX;
        col[m] = c < c1 ? c : c0;
        voff[m] = (size_t)(col[m] / nE) * nnz;
    }
    double acc[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) acc[m] = 0.0;
    for (int p = rowptr[row] + ty; p < rowptr[row + 1]; p += AMG_TY) {
        const size_t xo = (size_t)__ldg(colidx + p) * ld;
#pragma unroll
        for (int m = 0; m < CPT; m++) acc[m] = fma(__ldg(vals + voff[m] + p), __ldg(X + xo + col[m]), acc[m]);
    }
    __shared__ double red[AMG_TY][AMG_TX * CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) red[ty][m * AMG_TX + tx] = acc[m];
    __syncthreads();
    if (ty < CPT) {                                             // thread row m finishes column group m
        const int c = cbase + ty * AMG_TX;
        if (c < c1) {
            double s = 0.0;
#pragma unroll
            for (int y = 0; y < AMG_TY; y++) s += red[y][ty * AMG_TX + tx];
            const size_t o = (size_t)row * ld + c;
            Z[o] = fma(dinvw[(size_t)(c / nE) * n + row], R[o] - s, X[o]);
        }
    }
}
template <int CPT>
__global__ void __launch_bounds__(AMG_TX * AMG_TY)
k_amg_restrict_split(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals_dw, size_t nnz,
                     const int *__restrict__ mem_ptr, const int *__restrict__ mem_idx, int n_c,
                     const double *__restrict__ R, double *__restrict__ RC, int nE, int c0, int c1, size_t ld) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int I = blockIdx.x;
    const int cbase = c0 + blockIdx.y * (AMG_TX * CPT) + tx;
    int col[CPT]; size_t voff[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) {
        const int c = cbase + m * AMG_TX;
        col[m] = c < c1 ? c : c0;
        voff[m] = (size_t)(col[m] / nE) * nnz;
    }
    double acc[CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) acc[m] = 0.0;
    int cnt = 0;                                               // entries of the aggregate's rows seen so far
    const int qb = mem_ptr[I], qe = mem_ptr[I + 1];
    for (int q = qb; q < qe; q++) {
        const int i = mem_idx[q];
        if (((q - qb) & (AMG_TY - 1)) == ty) {
#pragma unroll
            for (int m = 0; m < CPT; m++) acc[m] += __ldg(R + (size_t)i * ld + col[m]);
        }
        const int pb = rowptr[i], pe = rowptr[i + 1];
        for (int p = pb + ((ty - cnt) & (AMG_TY - 1)); p < pe; p += AMG_TY) {
            const size_t xo = (size_t)__ldg(colidx + p) * ld;
#pragma unroll
            for (int m = 0; m < CPT; m++) acc[m] = fma(-__ldg(vals_dw + voff[m] + p), __ldg(R + xo + col[m]), acc[m]);
        }
        cnt += pe - pb;
    }
    __shared__ double red[AMG_TY][AMG_TX * CPT];
#pragma unroll
    for (int m = 0; m < CPT; m++) red[ty][m * AMG_TX + tx] = acc[m];
    __syncthreads();
    if (ty < CPT) {
        const int c = cbase + ty * AMG_TX;
        if (c < c1) {
            double s = 0.0;
#pragma unroll
            for (int y = 0; y < AMG_TY; y++) s += red[y][ty * AMG_TX + tx];
            RC[(size_t)I * ld + c] = s;
        }
    }
    (void)n_c;
}

// RC[I] = sum of RES over the members of aggregate I (deterministic restriction of a fine residual block); flat mapping
__global__ void __launch_bounds__(FLAT_T)
k_amg_sum_members(const int *__restrict__ mem_ptr, const int *__restrict__ mem_idx, int n_c, const double *__restrict__ RES,
                  double *__restrict__ RC, int c0, int c1, size_t ld, int cw) {
    const FlatMap f = flat_map(c0, c1, cw);
    if (!f.active) return;
    int lo, hi; flat_rows(n_c, lo, hi);
#pragma unroll 2
    for (int I = lo + f.roff; I < hi; I += f.rpp) {
        double acc = 0.0;
        for (int q = mem_ptr[I]; q < mem_ptr[I + 1]; q++) acc += RES[(size_t)mem_idx[q] * ld + f.col];
        RC[(size_t)I * ld + f.col] = acc;
    }
}

// X = dw .* R + EC[agg]   (pre-smoothed iterate plus prolongated coarse correction); EC == nullptr -> X = dw .* R
__global__ void __launch_bounds__(FLAT_T)
k_amg_prolong(const double *__restrict__ dinvw, int n, const int *__restrict__ agg, const double *__restrict__ R,
              const double *__restrict__ EC, double *__restrict__ X, int nE, int c0, int c1, size_t ld, int cw) {
    const FlatMap f = flat_map(c0, c1, cw);
    if (!f.active) return;
    const double *dk = dinvw + (size_t)(f.col / nE) * n;
    int lo, hi; flat_rows(n, lo, hi);
#pragma unroll 4
    for (int row = lo + f.roff; row < hi; row += f.rpp) {
        const size_t o = (size_t)row * ld + f.col;
        double x = dk[row] * R[o];
        if (EC) x += EC[(size_t)agg[row] * ld + f.col];
        X[o] = x;
    }
}

// ---------------------------------------------------------------------------------
// Small levels of the multilevel cycle, fused: ONE launch runs the whole sub-cycle of the levels below STREAM_MIN_ROWS
// (restrict ... coarsest sweeps ... prolongate + post-smooth) for a slice of CPC source columns per CTA.
//   The source columns are independent, so a CTA that owns a column slice needs no grid-wide synchronisation: every
//   vector of every small level (R, X, Z of its columns) lives in shared memory for the whole sub-cycle, the level
//   matrices stream from L2 (coalesced: half a warp per matrix row, lanes = consecutive CSR entries, shuffle reduction).
//   Replaces ~20 latency-bound launches per PCG iteration (8-16 us each on levels of a few thousand rows and fewer).
// ---------------------------------------------------------------------------------
constexpr int SUB_MAX_LEVELS = 8;
constexpr int SUB_THREADS = 512;
struct SubLevel {
    const int *rowptr, *colidx, *mem_ptr, *mem_idx, *agg;      // mem/agg: transfer to the NEXT (coarser) sub-level
    const double *vals, *vals_dw, *dinvw;                       // [nK][nnz], [nK][nnz], [nK][n]
    int n; size_t nnz;
    int off;                                                    // offset (in rows) of this level inside a column's shared vectors
};
struct SubArgs {
    SubLevel lv[SUB_MAX_LEVELS]; int nl;                        // nl sub-levels, the last one is the coarsest
    int rows_total;                                             // sum of n over the sub-levels
    const double *Rin; double *Zout; size_t ld;                 // residual of the first sub-level in, its correction out
    int nE, c0, c1, sweeps;
    int cache_last;                                             // stage the coarsest level's matrix in shared memory (it is applied `sweeps` times)
};

template <int CPC>
__device__ __forceinline__ void sub_half_reduce(double (&v)[CPC]) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < CPC; c++) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
}

template <int CPC>
__global__ void __launch_bounds__(SUB_THREADS)
k_amg_subcycle(const SubArgs A) {
    extern __shared__ __align__(16) double sub_sm[];
    // column slice of this CTA: slices never straddle a wavenumber group
    const int spk = (A.nE + CPC - 1) / CPC;                      // slices per wavenumber group
    const int kk = blockIdx.x / spk, cs = kk * A.nE + (blockIdx.x - kk * spk) * CPC;
    const int ce = min(cs + CPC, (kk + 1) * A.nE);
    const int v0 = max(cs, A.c0), v1 = min(ce, A.c1);
    if (v1 <= v0) return;
    const int RT = A.rows_total;
    double *sR = sub_sm, *sX = sub_sm + (size_t)CPC * RT, *sZ = sub_sm + 2 * (size_t)CPC * RT;     // [CPC][rows_total] each
    const int tid = threadIdx.x, half = tid >> 4, hl = tid & 15, nhalf = SUB_THREADS / 16;
    // 1. residual of the first sub-level
    for (int x = tid; x < A.lv[0].n * CPC; x += SUB_THREADS) {
        const int row = x / CPC, c = x - row * CPC;
        sR[(size_t)c * RT + row] = (cs + c >= v0 && cs + c < v1) ? A.Rin[(size_t)row * A.ld + cs + c] : 0.0;
    }
    __syncthreads();
    // 2. downward: residual after one damped-Jacobi sweep from zero, restricted (k_amg_restrict)
    for (int l = 0; l + 1 < A.nl; l++) {
        const SubLevel &L = A.lv[l], &Lc = A.lv[l + 1];
        const double *vdw = L.vals_dw + (size_t)kk * L.nnz;
        for (int base = 0; base < Lc.n; base += nhalf) {           // uniform trip count: the half-warp reductions need whole warps
            const int I = base + half;
            const bool act = I < Lc.n;
            double acc[CPC];
#pragma unroll
            for (int c = 0; c < CPC; c++) acc[c] = 0.0;
            for (int q = act ? L.mem_ptr[I] : 0, qe = act ? L.mem_ptr[I + 1] : 0; q < qe; q++) {
                const int i = L.mem_idx[q];
                if (hl == 0) {
#pragma unroll
                    for (int c = 0; c < CPC; c++) acc[c] += sR[(size_t)c * RT + L.off + i];
                }
                for (int p = L.rowptr[i] + hl; p < L.rowptr[i + 1]; p += 16) {
                    const double a = vdw[p]; const int j = L.colidx[p];
#pragma unroll
                    for (int c = 0; c < CPC; c++) acc[c] = fma(-a, sR[(size_t)c * RT + L.off + j], acc[c]);
                }
            }
            sub_half_reduce<CPC>(acc);
            if (act && hl == 0) {
#pragma unroll
                for (int c = 0; c < CPC; c++) sR[(size_t)c * RT + Lc.off + I] = acc[c];
            }
        }
        __syncthreads();
    }
    // 3. coarsest level: x = dw r, then sweeps - 1 damped-Jacobi sweeps
    {
        const SubLevel &L = A.lv[A.nl - 1];
        const double *dw = L.dinvw + (size_t)kk * L.n, *va = L.vals + (size_t)kk * L.nnz;
        const int *rp = L.rowptr, *cj = L.colidx;
        if (A.cache_last) {
            // the matrix of the coarsest level is applied sweeps - 1 times: from shared memory the dependent
            // index -> value -> x chain costs tens of cycles instead of L2 round trips
            double *cv = sub_sm + 3 * (size_t)CPC * RT;
            int *cc = reinterpret_cast<int *>(cv + L.nnz), *cr = cc + L.nnz;
            for (size_t x = tid; x < L.nnz; x += SUB_THREADS) { cv[x] = va[x]; cc[x] = L.colidx[x]; }
            for (int x = tid; x <= L.n; x += SUB_THREADS) cr[x] = L.rowptr[x];
            va = cv; cj = cc; rp = cr;
        }
        for (int x = tid; x < L.n * CPC; x += SUB_THREADS) { const int row = x % L.n, c = x / L.n; sX[(size_t)c * RT + L.off + row] = dw[row] * sR[(size_t)c * RT + L.off + row]; }
        __syncthreads();
        double *a = sX, *b = sZ;
        const int sw = A.sweeps;
        for (int s = 0; s + 1 < sw; s++) {
            for (int base = 0; base < L.n; base += nhalf) {
                const int row = base + half;
                const bool act = row < L.n;
                double acc[CPC];
#pragma unroll
                for (int c = 0; c < CPC; c++) acc[c] = 0.0;
                for (int p = act ? rp[row] + hl : 0, pe = act ? rp[row + 1] : 0; p < pe; p += 16) {
                    const double e = va[p]; const int j = cj[p];
#pragma unroll
                    for (int c = 0; c < CPC; c++) acc[c] = fma(e, a[(size_t)c * RT + L.off + j], acc[c]);
                }
                sub_half_reduce<CPC>(acc);
                if (act && hl == 0) {
#pragma unroll
                    for (int c = 0; c < CPC; c++) { const size_t o = (size_t)c * RT + L.off + row; b[o] = fma(dw[row], sR[o] - acc[c], a[o]); }
                }
            }
            __syncthreads();
            double *t = a; a = b; b = t;
        }
        // the coarsest result must sit in sZ for the upward pass (E of the next finer level)
        if (a != sZ) { for (int x = tid; x < L.n * CPC; x += SUB_THREADS) { const int row = x % L.n, c = x / L.n; sZ[(size_t)c * RT + L.off + row] = a[(size_t)c * RT + L.off + row]; } __syncthreads(); }
    }
    // 4. upward: x = dw r + E[agg];  z = x + dw (r - A x)
    for (int l = A.nl - 2; l >= 0; l--) {
        const SubLevel &L = A.lv[l], &Lc = A.lv[l + 1];
        const double *dw = L.dinvw + (size_t)kk * L.n, *va = L.vals + (size_t)kk * L.nnz;
        for (int x = tid; x < L.n * CPC; x += SUB_THREADS) {
            const int row = x % L.n, c = x / L.n;
            sX[(size_t)c * RT + L.off + row] = fma(dw[row], sR[(size_t)c * RT + L.off + row], sZ[(size_t)c * RT + Lc.off + L.agg[row]]);
        }
        __syncthreads();
        for (int base = 0; base < L.n; base += nhalf) {
            const int row = base + half;
            const bool act = row < L.n;
            double acc[CPC];
#pragma unroll
            for (int c = 0; c < CPC; c++) acc[c] = 0.0;
            for (int p = act ? L.rowptr[row] + hl : 0, pe = act ? L.rowptr[row + 1] : 0; p < pe; p += 16) {
                const double e = va[p]; const int j = L.colidx[p];
#pragma unroll
                for (int c = 0; c < CPC; c++) acc[c] = fma(e, sX[(size_t)c * RT + L.off + j], acc[c]);
            }
            sub_half_reduce<CPC>(acc);
            if (act && hl == 0) {
#pragma unroll
                for (int c = 0; c < CPC; c++) { const size_t o = (size_t)c * RT + L.off + row; sZ[o] = fma(dw[row], sR[o] - acc[c], sX[o]); }
            }
        }
        __syncthreads();
    }
    // 5. correction of the first sub-level back to HBM
    for (int x = tid; x < A.lv[0].n * CPC; x += SUB_THREADS) {
        const int row = x / CPC, c = x - row * CPC;
        if (cs + c >= v0 && cs + c < v1) A.Zout[(size_t)row * A.ld + cs + c] = sZ[(size_t)c * RT + row];
    }
}

// numeric primary potentials: prim[i][c] = SRC[map[i]][c]  (rows of the P2 primary solve at this mesh's nodes)
__global__ void k_gather_rows(const double *__restrict__ SRC, size_t ld_src, const int *__restrict__ map, int n, int ncols,
                              size_t ld, double *__restrict__ OUT) {
    const int col = blockIdx.y * blockDim.x + threadIdx.x;
    const int row = blockIdx.x * blockDim.y + threadIdx.y;
    if (col >= ncols || row >= n) return;
    OUT[(size_t)row * ld + col] = SRC[(size_t)map[row] * ld_src + col];
}

// total field  U = X + rho_src * prim   (:2287)  /  analytic branch  U = scale * prim (:1295-1300)
__global__ void k_finalize_pots(const double *__restrict__ Xv, const double *__restrict__ prim, const double *__restrict__ rho_src,
                                double scale, int N, int nE, int c0, int c1, size_t ld, double *__restrict__ U) {
    const int col = c0 + blockIdx.y * blockDim.x + threadIdx.x;
    const int row = blockIdx.x * blockDim.y + threadIdx.y;
    if (col >= c1 || row >= N) return;
    const size_t o = (size_t)row * ld + col;
    if (Xv) {
        const int e = col % nE;
        U[o] = prim ? Xv[o] + rho_src[e] * prim[o] : Xv[o];
    } else {
        U[o] = scale * prim[o];
    }
}
// zero the RHS rows of Dirichlet nodes (:2281-2283)
__global__ void k_zero_rows(const int *__restrict__ rows, int n, int c0, int c1, size_t ld, double *__restrict__ B) {
    const int col = c0 + blockIdx.y * blockDim.x + threadIdx.x;
    const int i = blockIdx.x;
    if (col < c1 && i < n) B[(size_t)rows[i] * ld + col] = 0.0;
}
// total-field right-hand side: delta at the electrode (electrode.cpp:133-151, :274-287)
__global__ void k_delta_rhs(const int *__restrict__ pick_ptr, const int *__restrict__ pick_idx, const double *__restrict__ pick_w,
                            int nE, int c0, int c1, size_t ld, double *__restrict__ B) {
    const int col = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= c1) return;
    const int e = col % nE;
    for (int t = pick_ptr[e]; t < pick_ptr[e + 1]; t++) B[(size_t)pick_idx[t] * ld + col] = pick_w[t];
}

// current reference of the dipole patterns (dcfemmodelling.cpp:1517-1523, 1868-1870): a reference-electrode node (-999)
// takes -1 in every pattern; without one on a pure-Neumann domain the LAST electrode is the reference (:1054-1064) and
// its own pattern does not exist (zero right-hand side, zero potentials)
__global__ void k_ref_rhs(const int *__restrict__ pick_ptr, const int *__restrict__ pick_idx, const double *__restrict__ pick_w,
                          int nE, int c0, int c1, size_t ld, int ref_node, int ref_last, double *__restrict__ B) {
    const int col = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= c1) return;
    const int e = col % nE;
    if (ref_node >= 0) B[(size_t)ref_node * ld + col] = -1.0;
    if (ref_last) {
        if (e == nE - 1) { for (int t = pick_ptr[e]; t < pick_ptr[e + 1]; t++) B[(size_t)pick_idx[t] * ld + col] = 0.0; }
        else for (int t = pick_ptr[nE - 1]; t < pick_ptr[nE]; t++) B[(size_t)pick_idx[t] * ld + col] = -pick_w[t];
    }
}

// ---------------------------------------------------------------------------------
// forward epilogue (dcfemmodelling.cpp:1707-1729, datamap.cpp:57-97, :159-231, :1140-1196)
// ---------------------------------------------------------------------------------
// pM[e][e'] = sum_k w_k * sum_t pick_w[t] U[pick_idx[t]][e + nE k]   (potential of source e at electrode e')
__global__ void k_pickup(const double *__restrict__ U, size_t ld, const double *__restrict__ w, int nK, int nE,
                         const int *__restrict__ pick_ptr, const int *__restrict__ pick_idx, const double *__restrict__ pick_w,
                         int c0, int c1, double *__restrict__ pM) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;      // source
    const int ep = blockIdx.y;                                 // receiving electrode
    if (e >= nE) return;
    double acc = 0.0;
    for (int kk = 0; kk < nK; kk++) {
        const int s = e + nE * kk;
        if (s < c0 || s >= c1) continue;                       // other shards add their part (all-reduce)
        double u = 0.0;
        for (int t = pick_ptr[ep]; t < pick_ptr[ep + 1]; t++) u += pick_w[t] * U[(size_t)pick_idx[t] * ld + e + nE * kk];
        acc += u * w[kk];
    }
    pM[(size_t)e * nE + ep] = acc;
}
__global__ void k_response(const double *__restrict__ pM, int nE, const int *__restrict__ abmn, const double *__restrict__ kfac, int D,
                           double *__restrict__ resp, double *__restrict__ resp_rez, double *__restrict__ rhoa) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int a = abmn[4 * d], b = abmn[4 * d + 1], m = abmn[4 * d + 2], n = abmn[4 * d + 3];
    auto P = [&](int s, int r) -> double { return (s >= 0 && r >= 0) ? pM[(size_t)s * nE + r] : 0.0; };
    const double u  = (P(a, m) - P(a, n)) - (P(b, m) - P(b, n));
    const double ur = (P(m, a) - P(m, b)) - (P(n, a) - P(n, b));           // reciprocity: a<->m, b<->n
    const double r1 = rint(u / 1e-10) * 1e-10 * kfac[d];                   // round(u, 1e-10) * k
    const double r2 = rint(ur / 1e-10) * 1e-10 * kfac[d];
    resp[d] = r1; resp_rez[d] = r2;
    rhoa[d] = sqrt(fabs(r1 * r2));
}
// solutions_[e][node] = sum_k w_k U[node][e + nE k]  (only for introspection)
__global__ void k_ksum(const double *__restrict__ U, size_t ld, const double *__restrict__ w, int nK, int nE, int N, double *__restrict__ out) {
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    const int node = blockIdx.x * blockDim.y + threadIdx.y;
    if (e >= nE || node >= N) return;
    double acc = 0.0;
    for (int kk = 0; kk < nK; kk++) acc += w[kk] * U[(size_t)node * ld + e + nE * kk];
    out[(size_t)e * N + node] = acc;
}
// transposed copy for introspection: out[s][node] = U[node][s]
__global__ void k_transpose_out(const double *__restrict__ U, size_t ld, int N, int nS, double *__restrict__ out) {
    __shared__ double tile[32][33];
    int s = blockIdx.x * 32 + threadIdx.x, node = blockIdx.y * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8) if (s < nS && node + j < N) tile[threadIdx.y + j][threadIdx.x] = U[(size_t)(node + j) * ld + s];
    __syncthreads();
    int node2 = blockIdx.y * 32 + threadIdx.x, s2 = blockIdx.x * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8) if (s2 + j < nS && node2 < N) out[(size_t)(s2 + j) * N + node2] = tile[threadIdx.x][threadIdx.y + j];
}

// ---------------------------------------------------------------------------------
// K3: Jacobian (bertJacobian.cpp:70-119, elementmatrix.h:280-293, dcfemmodelling.cpp:1377-1383)
//   J[d][col] = k_d / rho_col^2 * sum_{cells c of col} sum_k w_k (u_a-u_b)^T (K_c + k^2 M_c) (u_m-u_n)
//
//   One CTA per model column.  For every cell of the column and every wavenumber the CTA
//   gathers the cell's nodal potentials for the "current-side" electrode list P and the
//   "potential-side" list Q (coalesced: one node's values for all sources are contiguous),
//   forms V = E U_Q, and accumulates the electrode-pair Gram block
//       G[p][q] += w_k * sum_i U_P[i][p] * V[i][q]
//   in registers (4x4 micro-tiles per thread, FP64 FMA pipe).  G is then dropped to shared
//   memory once per column and every datum is the 4-term ABMN combination
//       G[a][m] - G[a][n] - G[b][m] + G[b][n],
//   scaled and stored to the column-major J with consecutive threads writing consecutive rows.
// ---------------------------------------------------------------------------------
constexpr int JAC_MAX_THREADS = 640;   // MT = 1: one 4x4 micro-tile per thread (<= 640 tiles); MT = 2: two (<= 1024 tiles, 512 threads)

struct __align__(8) JacDatum { unsigned short a, b, m, n; };   // indices into plist / qlist, 0xFFFF = unused electrode

struct JacArgs {
    const double *pos; const int *cells; int nloc;
    const int *jac_cells; const int *jac_col_ptr; int col_begin, col_end;
    const int *cta_col_ptr;                   // [gridDim.x + 1] contiguous column range of every CTA
    const double *U; size_t ld; int nE; int nK; const double *kvals; const double *kw;
    const int *plist; int nP, nPp;            // current-side electrodes, padded to a multiple of 4
    const int *qlist; int nQ, nQp;
    int gplane;                               // plane stride of the Gram block in shared memory
    const JacDatum *idx;                      // [nd] 16-bit indices (a, b, m, n) into plist / qlist ...
    int resolved;                             // ... or, if the Gram block has < 65536 entries, the four shared-memory
                                              // offsets of G[a][m], G[a][n], G[b][m], G[b][n] (unused -> a zero slot)
    const int *out_row; const double *kfac; int nd;
    int idx_in_smem;                          // stage the index records in shared memory (reused by every column)
    int kfac_in_smem;                         // stage the geometric factors too
    int out_identity;                         // out_row[d] == out_base + d: no indirection on the store
    int out_base;
    const double *rho_col;                    // [M] model value per column or nullptr (no scaling)
    double *Jt; size_t ldJ;
};

// cursor over the work items (column, cell, wavenumber) of one CTA.  Every CTA owns a contiguous range of
// columns (balanced by cell count on the host): the column pointers and cell lists are then read sequentially
// (L1 hits), neighbouring cells share nodes, and adjacent columns of J are adjacent in memory.
struct JacCursor {
    int col, end, ci, ce, kk, cell;
    __device__ __forceinline__ bool valid() const { return col < end; }
    __device__ __forceinline__ void seek(const JacArgs &A) {                   // first column >= col that has cells
        while (col < end) {
            ci = A.jac_col_ptr[col]; ce = A.jac_col_ptr[col + 1];
            if (ci < ce) { cell = A.jac_cells[ci]; kk = 0; return; }
            col++;
        }
    }
    __device__ __forceinline__ void advance(const JacArgs &A) {
        if (++kk < A.nK) return;
        kk = 0;
        if (++ci < ce) { cell = A.jac_cells[ci]; return; }
        col++;
        seek(A);
    }
};

__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// Software-pipelined over the work items: while item t is contracted (and its column written), the node ids of
// item t+2 and the coordinates + nodal potentials of item t+1 are already in flight (cp.async, one group per item).
template <int E, int JAC_MAX_TILES>
__global__ void __launch_bounds__(JAC_MAX_TILES == 1 ? JAC_MAX_THREADS : 512)
k_jacobian(const JacArgs A) {
    constexpr int NV = ElemTraits<E>::NV, NL = ElemTraits<E>::NL, DIM = ElemTraits<E>::DIM;
    const int JAC_THREADS = blockDim.x;
    extern __shared__ __align__(16) double sm[];
    const int szP = NL * A.nPp, szQ = NL * A.nQp;
    double *sUp = sm;                               // [2][NL][nPp]
    double *sUq = sUp + 2 * szP;                    // [2][NL][nQp]
    double *sV  = sUq + 2 * szQ;                    // [NL][nQp]
    double *sK  = sV + szQ;                         // [NL*NL] stiffness
    double *sM  = sK + NL * NL;                     // [NL*NL] mass
    double *sXYZ = sM + NL * NL;                    // [2][NV][3] corner coordinates
    double *sG  = sXYZ + 2 * NV * 3;                // Gram block, element-major: 16 planes [a][b] of one value per 4x4 tile
    const int PL = A.gplane;                        // plane stride (>= number of tiles, = 4 mod 16: conflict-free tile stores and row reads)
    double *sKf = sG + 16 * (size_t)PL + 1;         // [nd] geometric factors (optional); sG[16 PL] is the zero slot
    void *sIdxRaw = A.kfac_in_smem ? (void *)(sKf + A.nd) : (void *)sKf;
    __shared__ int snode[3][NL];
    const int tid = threadIdx.x;
    const int tilesQ = A.nQp / 4, tilesP = A.nPp / 4, ntiles = tilesP * tilesQ;
    const JacDatum *idx = A.idx;
    const double *kfp = A.kfac;
    if (A.kfac_in_smem) { for (int d = tid; d < A.nd; d += JAC_THREADS) sKf[d] = A.kfac[d]; kfp = sKf; }
    if (A.idx_in_smem) {
        JacDatum *s16 = reinterpret_cast<JacDatum *>(sIdxRaw);
        for (int d = tid; d < A.nd; d += JAC_THREADS) s16[d] = A.idx[d];
        idx = s16;
    }
    if (tid == 0) sG[16 * PL] = 0.0;                // the zero slot unused electrodes point at
    const int my_lo = A.cta_col_ptr[blockIdx.x], my_hi = A.cta_col_ptr[blockIdx.x + 1];
    const bool scaled = A.rho_col != nullptr;

    // zero-fill the padding of the gather buffers once (the async copies only touch the used part)
    for (int x = tid; x < 2 * szP; x += JAC_THREADS) sUp[x] = 0.0;
    for (int x = tid; x < 2 * szQ; x += JAC_THREADS) sUq[x] = 0.0;

    auto load_ids = [&](const JacCursor &c, int slot) {            // node ids of an item -> snode[slot]
        if (c.valid() && tid < NL) cp_async4(&snode[slot][tid], A.cells + (size_t)c.cell * NL + tid);
    };
    auto load_item = [&](const JacCursor &c, int slot, int buf) {  // coordinates + nodal potentials; ids must be visible
        if (!c.valid()) return;
        if (tid < NV * 3) { const int v = tid / 3, d = tid - 3 * v; cp_async8(sXYZ + buf * NV * 3 + tid, A.pos + 3 * (size_t)snode[slot][v] + d); }
        double *up = sUp + buf * szP, *uq = sUq + buf * szQ;
        for (int x = tid; x < szP; x += JAC_THREADS) {
            const int i = x / A.nPp, p = x - i * A.nPp;
            if (p < A.nP) cp_async8(up + x, A.U + (size_t)snode[slot][i] * A.ld + A.plist[p] + A.nE * c.kk);
        }
        for (int x = tid; x < szQ; x += JAC_THREADS) {
            const int i = x / A.nQp, q = x - i * A.nQp;
            if (q < A.nQ) cp_async8(uq + x, A.U + (size_t)snode[slot][i] * A.ld + A.qlist[q] + A.nE * c.kk);
        }
    };

    JacCursor cur; cur.col = my_lo; cur.end = my_hi; cur.ci = cur.ce = cur.kk = cur.cell = 0; cur.seek(A);
    JacCursor nxt = cur; if (nxt.valid()) nxt.advance(A);
    JacCursor nn = nxt;  if (nn.valid()) nn.advance(A);
    // columns before the first non-empty one are all-zero
    for (int cz = my_lo; cz < min(cur.col, my_hi); cz++)
        for (int d = tid; d < A.nd; d += JAC_THREADS) A.Jt[(size_t)cz * A.ldJ + (A.out_identity ? A.out_base + d : A.out_row[d])] = 0.0;

    // prologue: ids(0), ids(1) -> then coordinates + potentials of item 0
    load_ids(cur, 0); load_ids(nxt, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    load_item(cur, 0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    double acc[JAC_MAX_TILES][16];
#pragma unroll
    for (int t = 0; t < JAC_MAX_TILES; t++)
#pragma unroll
        for (int x = 0; x < 16; x++) acc[t][x] = 0.0;

    for (int t = 0; cur.valid(); t++) {
        const int buf = t & 1, slot1 = (t + 1) % 3, slot2 = (t + 2) % 3;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                             // item t landed; everybody is done with item t-1
        // next items in flight while this one is processed
        load_ids(nn, slot2);
        load_item(nxt, slot1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tid < NL * NL) {
            double X[NV][3];
#pragma unroll
            for (int v = 0; v < NV; v++) { X[v][0] = sXYZ[buf * NV * 3 + 3 * v]; X[v][1] = sXYZ[buf * NV * 3 + 3 * v + 1]; X[v][2] = sXYZ[buf * NV * 3 + 3 * v + 2]; }
            double size, G[NV][NV];
            simplex_gram<DIM>(X, size, G);
            const int i = tid / NL, j = tid - i * NL;
            sK[tid] = stiff_entry<E>(i, j, size, G);
            sM[tid] = size * mass_unit<E>(i, j);
        }
        __syncthreads();
        const double *up = sUp + buf * szP, *uq = sUq + buf * szQ;
        {
            const double k = A.kvals[cur.kk], wk = A.kw[cur.kk];
            const double k2 = k * k;
            for (int x = tid; x < szQ; x += JAC_THREADS) {     // V = w_k (K + k^2 M) U_Q
                const int i = x / A.nQp, q = x - i * A.nQp;
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NL; j++) v = fma(fma(k2, sM[i * NL + j], sK[i * NL + j]), uq[j * A.nQp + q], v);
                sV[x] = v * wk;
            }
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < JAC_MAX_TILES; tt++) {
            const int tile = tid + tt * JAC_THREADS;
            if (tile < ntiles) {
                const int tp = tile / tilesQ, tq = tile - tp * tilesQ;
#pragma unroll
                for (int i = 0; i < NL; i++) {
                    const double2 u01 = *reinterpret_cast<const double2 *>(up + i * A.nPp + 4 * tp);
                    const double2 u23 = *reinterpret_cast<const double2 *>(up + i * A.nPp + 4 * tp + 2);
                    const double2 v01 = *reinterpret_cast<const double2 *>(sV + i * A.nQp + 4 * tq);
                    const double2 v23 = *reinterpret_cast<const double2 *>(sV + i * A.nQp + 4 * tq + 2);
                    const double u[4] = {u01.x, u01.y, u23.x, u23.y};
                    const double v[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) acc[tt][a * 4 + b] = fma(u[a], v[b], acc[tt][a * 4 + b]);
                }
            }
        }
        const bool last_of_col = !nxt.valid() || nxt.col != cur.col;
        if (last_of_col) {
            const int col = cur.col;
            // drop G to shared memory (every thread passed the V barrier, so the previous epilogue is over)
#pragma unroll
            for (int tt = 0; tt < JAC_MAX_TILES; tt++) {
                const int tile = tid + tt * JAC_THREADS;
                if (tile < ntiles) {
                    const int tp = tile / tilesQ, tq = tile - tp * tilesQ;
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) { sG[(a * 4 + b) * PL + tile] = acc[tt][a * 4 + b]; acc[tt][a * 4 + b] = 0.0; }
                    (void)tp; (void)tq;
                }
            }
            __syncthreads();
            double scale = 1.0;
            if (scaled) { const double r = A.rho_col[col]; scale = 1.0 / (r * r); }
            double *out = A.Jt + (size_t)col * A.ldJ;
            if (A.resolved && A.out_identity && scaled) {
                // the common case, kept lean: pre-resolved offsets, k_i / rho_j^2 scaling, rows in place
                const JacDatum *ip = idx + tid;
                const double *kp = kfp + tid;
                double *op = out + A.out_base + tid;
                const int n_it = (A.nd - tid + JAC_THREADS - 1) / JAC_THREADS;
#pragma unroll 4
                for (int it = 0; it < n_it; it++) {
                    const JacDatum e = *ip;                           // four pre-resolved offsets into G
                    const double v = (sG[e.a] - sG[e.b]) - (sG[e.m] - sG[e.n]);
                    *op = v * (*kp * scale);
                    ip += JAC_THREADS; kp += JAC_THREADS; op += JAC_THREADS;
                }
            } else if (A.resolved) {
#pragma unroll 4
                for (int d = tid; d < A.nd; d += JAC_THREADS) {
                    const JacDatum e = idx[d];
                    const double v = (sG[e.a] - sG[e.b]) - (sG[e.m] - sG[e.n]);
                    const double kf = scaled ? kfp[d] * scale : 1.0;  // k_i / rho_j^2 only if len(model) == cols (:1377)
                    out[A.out_identity ? A.out_base + d : __ldg(A.out_row + d)] = v * kf;
                }
            } else {
#pragma unroll 2
                for (int d = tid; d < A.nd; d += JAC_THREADS) {
                    const JacDatum e = idx[d];
                    const int ea = e.a == 0xFFFF ? -1 : e.a, eb = e.b == 0xFFFF ? -1 : e.b, em = e.m == 0xFFFF ? -1 : e.m, en = e.n == 0xFFFF ? -1 : e.n;
                    const double kf = scaled ? kfp[d] * scale : 1.0;
                    auto goff = [&](int p, int q) { return ((p & 3) * 4 + (q & 3)) * PL + (p >> 2) * tilesQ + (q >> 2); };
                    double v = 0.0;
                    if (ea >= 0 && em >= 0) v += sG[goff(ea, em)];
                    if (ea >= 0 && en >= 0) v -= sG[goff(ea, en)];
                    if (eb >= 0 && em >= 0) v -= sG[goff(eb, em)];
                    if (eb >= 0 && en >= 0) v += sG[goff(eb, en)];
                    out[A.out_identity ? A.out_base + d : __ldg(A.out_row + d)] = v * kf;
                }
            }
            // columns of this CTA without any model cell between this one and the next item are all-zero
            const int stop = nxt.valid() ? nxt.col : my_hi;
            for (int cz = col + 1; cz < stop; cz++)
                for (int d = tid; d < A.nd; d += JAC_THREADS) A.Jt[(size_t)cz * A.ldJ + (A.out_identity ? A.out_base + d : A.out_row[d])] = 0.0;
        }
        cur = nxt; nxt = nn;
        if (nn.valid()) nn.advance(A);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// K3': Jacobian, second generation (bertJacobian.cpp:70-119, elementmatrix.h:280-293, dcfemmodelling.cpp:1377-1383)
//   J[d][col] = k_d / rho_col^2 * sum_{cells c of col} sum_k w_k (u_a-u_b)^T (K_c + k^2 M_c) (u_m-u_n)
//
//   * BASIS: the host picks, per side of the measurement, either the electrodes or the distinct dipoles as basis
//     (k_basis_pots writes  UD[node][k][i] = u_a - u_b  once per Jacobian).  In dipole space a datum of a dipole-dipole
//     scheme is ONE entry of the Gram block instead of four; when both sides share one basis list the block is
//     symmetric and only the tiles on and above the diagonal are computed.
//   * TILE LIST: only the 4x4 register tiles that some datum needs are computed (block-sparse schemes, e.g. crosshole).
//   * WARP SPECIALISATION inside one persistent CTA per SM:
//       producer warp   TMA bulk copies of the next items (cell, wavenumber): the cell's precomputed element record
//                       (stiffness K_c and size, geometry-only, built at create) and the basis potentials of its nodes
//                       -- one copy per node row -- into a ring of slots (full/empty mbarriers)
//       Gram warps      V = w_k (K + k^2 M) U_Q for the NEXT item, then the register-tile update of the CURRENT one
//                       (one named barrier per item); at the end of a model column the tiles go to one of two
//                       shared-memory Gram buffers
//       epilogue warps  turn the finished Gram buffer into the column of J: 1 / 2 / 4 signed look-ups per datum through
//                       pre-resolved 16-bit offsets, times k_d / rho^2, 128-bit stores -- while the Gram warps are
//                       already working on the next column.
// ---------------------------------------------------------------------------------
constexpr int J2_GRAM_WARPS = 11;                              // + EW epilogue warps + 1 producer warp
constexpr int J2_GT = J2_GRAM_WARPS * 32;
// EW = epilogue warps (template parameter of k_jacobian2): 4 where the Gram blocks bound the kernel (many tiles per written
// byte: crosshole c4), 8 where the store of J does (complete schemes on many rows: c3, measured 5.8 -> 4.7 ms)
__host__ __device__ constexpr int j2_threads(int ew) { return 32 * (J2_GRAM_WARPS + ew + 1); }
constexpr int J2_SLOTS = 3;
constexpr int J2_MAX_MT = 2;                                   // register tiles per Gram thread

// per-cell element record [NL*NL stiffness | size | pad]: geometry only
template <int E>
__global__ void k_element_records(const double *__restrict__ pos, const int *__restrict__ cells, int C, int recn, double *__restrict__ rec) {
    constexpr int NV = ElemTraits<E>::NV, NL = ElemTraits<E>::NL, DIM = ElemTraits<E>::DIM;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int n = cells[(size_t)c * NL + v];
        X[v][0] = pos[3 * (size_t)n]; X[v][1] = pos[3 * (size_t)n + 1]; X[v][2] = pos[3 * (size_t)n + 2];
    }
    double size, G[NV][NV];
    simplex_gram<DIM>(X, size, G);
    double *r = rec + (size_t)c * recn;
#pragma unroll
    for (int i = 0; i < NL; i++)
#pragma unroll
        for (int j = 0; j < NL; j++) r[i * NL + j] = stiff_entry<E>(i, j, size, G);
    r[NL * NL] = size;
    for (int x = NL * NL + 1; x < recn; x++) r[x] = 0.0;
}
template <int E>
__global__ void k_mass_unit_table(double *__restrict__ mu) {
    constexpr int NL = ElemTraits<E>::NL;
    const int t = threadIdx.x;
    if (t < NL * NL) mu[t] = mass_unit<E>(t / NL, t % NL);
}
// basis potentials: UD[node][k][i] = U[node][a_i + nE k] - U[node][b_i + nE k]  (b_i < 0: electrode basis / pole);
// columns nL..nLp-1 of every k-block are zero padding
__global__ void k_basis_pots(const double *__restrict__ U, size_t ld, int N, int nE, int nK, const int *__restrict__ la,
                             const int *__restrict__ lb, int nL, int nLp, double *__restrict__ UD, size_t ldUD) {
    const int x = blockIdx.y * blockDim.x + threadIdx.x;          // k * nLp + i
    const int node = blockIdx.x * blockDim.y + threadIdx.y;
    if (x >= nK * nLp || node >= N) return;
    const int kk = x / nLp, i = x - kk * nLp;
    double v = 0.0;
    if (i < nL) {
        const int a = la[i], b = lb[i];
        v = U[(size_t)node * ld + a + nE * kk];
        if (b >= 0) v -= U[(size_t)node * ld + b + nE * kk];
    }
    UD[(size_t)node * ldUD + x] = v;
}

struct Jac2Args {
    const int *cells; int C;
    const int *jac_cells; const int *jac_col_ptr; const int *cta_col_ptr;
    const double *erec; int recn; const double *mu;            // element records, unit mass matrix
    const double *UDp, *UDq; size_t ldUDp, ldUDq; int nPp, nQp, shared; // basis potentials [N][nK][nPp] / [N][nK][nQp]; shared: Q list == P list
    int nK; const double *kvals, *kw;
    const unsigned short *tile_tp, *tile_tq; int n_tiles, PL, mt;   // Gram tiles; PL: plane stride of a Gram buffer
    const unsigned short *toff; int terms;                      // [nd][terms] pre-resolved offsets into a Gram buffer
    const double *kfac; const int *out_row; int nd, out_identity, out_base, off_in_smem, kfac_in_smem;
    const double *rho_col; double *Jt; size_t ldJ;
    uint32_t slot_bytes, uq_off;                                // ring slot: [record | U_P rows | U_Q rows]; uq_off: byte offset of U_Q
    uint32_t rec_bytes;
};

template <int E, int MT, int TERMS, int EW>
__global__ void __launch_bounds__(j2_threads(EW), 1)
k_jacobian2(const Jac2Args A) {
    constexpr int J2_EPI_WARPS = EW, J2_THREADS = j2_threads(EW), J2_ET = 32 * EW;
    constexpr int NL = ElemTraits<E>::NL;
    extern __shared__ __align__(128) unsigned char j2_smem[];
    __shared__ __align__(8) uint64_t full[J2_SLOTS], empty[J2_SLOTS], gfull[2], gempty[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    // shared-memory map
    unsigned char *ring = j2_smem;
    double *sV = reinterpret_cast<double *>(ring + (size_t)J2_SLOTS * A.slot_bytes);          // [2][NL][nQp]
    double *sG = sV + 2 * NL * A.nQp;                                                         // [2][16 PL + 2]
    const int gsz = 16 * A.PL + 2;
    double *sKf = sG + 2 * gsz;                                                               // [nd] (optional)
    unsigned short *sOff = reinterpret_cast<unsigned short *>(sKf + (A.kfac_in_smem ? ((A.nd + 1) & ~1) : 0));   // [nd][TERMS] (optional)
    if (tid == 0) {
        for (int s = 0; s < J2_SLOTS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&gfull[b], 1); mbar_init(&gempty[b], J2_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sG[16 * A.PL] = 0.0; sG[gsz + 16 * A.PL] = 0.0;                                        // zero slots of both buffers
    }
    // tables shared by the epilogue warps (read-only afterwards)
    if (A.kfac_in_smem) for (int d = tid; d < A.nd; d += J2_THREADS) sKf[d] = A.kfac[d];
    if (A.off_in_smem) for (int x = tid; x < A.nd * TERMS; x += J2_THREADS) sOff[x] = A.toff[x];
    __syncthreads();
    const int my_lo = A.cta_col_ptr[blockIdx.x], my_hi = A.cta_col_ptr[blockIdx.x + 1];
    const int ci_lo = A.jac_col_ptr[my_lo], ci_hi = A.jac_col_ptr[my_hi];
    const int nK = A.nK;
    const uint32_t rowP = (uint32_t)A.nPp * 8u, rowQ = (uint32_t)A.nQp * 8u;

    if (warp == J2_GRAM_WARPS + J2_EPI_WARPS) {
        // ------------------------------- producer -------------------------------
        const uint32_t bytes = A.rec_bytes + NL * rowP + (A.shared ? 0u : NL * rowQ);
        uint32_t n = 0;
        for (int base = ci_lo; base < ci_hi; base += 32) {
            // this batch: lane l holds the cell and node ids of cell base + l
            const int ci = base + lane;
            int cell = 0, nodes[NL];
            if (ci < ci_hi) {
                cell = __ldg(A.jac_cells + ci);
#pragma unroll
                for (int i = 0; i < NL; i++) nodes[i] = __ldg(A.cells + (size_t)cell * NL + i);
            } else {
#pragma unroll
                for (int i = 0; i < NL; i++) nodes[i] = 0;
            }
            const int nb = min(32, ci_hi - base);
            for (int jj = 0; jj < nb; jj++) {
                const int cj = __shfl_sync(0xffffffffu, cell, jj);
                int mynode = 0;
#pragma unroll
                for (int i = 0; i < NL; i++) { const int v = __shfl_sync(0xffffffffu, nodes[i], jj); if (lane == i || lane == NL + i) mynode = v; }
                for (int kk = 0; kk < nK; kk++, n++) {
                    const uint32_t slot = n % J2_SLOTS, use = n / J2_SLOTS;
                    if (use > 0) mbar_wait(&empty[slot], (use & 1u) ^ 1u);
                    unsigned char *sb = ring + (size_t)slot * A.slot_bytes;
                    if (lane == 31) {
                        mbar_expect_tx(&full[slot], bytes);
                        tma_bulk_g2s(sb, A.erec + (size_t)cj * A.recn, A.rec_bytes, &full[slot]);
                    }
                    __syncwarp();
                    if (lane < NL) tma_bulk_g2s(sb + A.rec_bytes + lane * rowP, A.UDp + (size_t)mynode * A.ldUDp + (size_t)kk * A.nPp, rowP, &full[slot]);
                    else if (!A.shared && lane < 2 * NL)
                        tma_bulk_g2s(sb + A.uq_off + (lane - NL) * rowQ, A.UDq + (size_t)mynode * A.ldUDq + (size_t)kk * A.nQp, rowQ, &full[slot]);
                }
            }
        }
        return;
    }

    if (warp >= J2_GRAM_WARPS) {
        // ------------------------------- epilogue warps -------------------------------
        const int te = tid - J2_GT;
        const unsigned short *off = A.off_in_smem ? sOff : A.toff;
        const double *kfp = A.kfac_in_smem ? sKf : A.kfac;
        const bool scaled = A.rho_col != nullptr;
        uint32_t nc = 0;                                         // non-empty columns seen
        for (int col = my_lo; col < my_hi; col++) {
            double *out = A.Jt + (size_t)col * A.ldJ;
            if (A.jac_col_ptr[col] == A.jac_col_ptr[col + 1]) {      // no cells: the column of J is zero
                for (int d = te; d < A.nd; d += J2_ET) out[A.out_identity ? A.out_base + d : A.out_row[d]] = 0.0;
                continue;
            }
            const uint32_t buf = nc & 1u, use = nc >> 1;
            mbar_wait(&gfull[buf], use & 1u);
            const double *G = sG + buf * gsz;
            double scale = 1.0;
            if (scaled) { const double r = A.rho_col[col]; scale = 1.0 / (r * r); }
            if (TERMS == 1 && A.out_identity && scaled && !(A.out_base & 1)) {
                // the common dipole-dipole case: one look-up per datum, two data per thread and pass, 128-bit stores
                const uint32_t *off2 = reinterpret_cast<const uint32_t *>(off);
                const double2 *kf2 = reinterpret_cast<const double2 *>(kfp);
                double2 *out2 = reinterpret_cast<double2 *>(out + A.out_base);
                const int np = A.nd >> 1;
#pragma unroll 4
                for (int d2 = te; d2 < np; d2 += J2_ET) {
                    const uint32_t o = off2[d2];
                    const double2 kf = kf2[d2];
                    out2[d2] = make_double2(G[o & 0xffffu] * (kf.x * scale), G[o >> 16] * (kf.y * scale));
                }
                if ((A.nd & 1) && te == 0) out[A.out_base + A.nd - 1] = G[off[A.nd - 1]] * (kfp[A.nd - 1] * scale);
            } else {
#pragma unroll 2
                for (int d = te; d < A.nd; d += J2_ET) {
                    const unsigned short *o = off + (size_t)d * TERMS;
                    double v = G[o[0]];
                    if (TERMS >= 2) v -= G[o[1]];
                    if (TERMS == 4) v -= (G[o[2]] - G[o[3]]);
                    const double kf = scaled ? kfp[d] * scale : 1.0;       // k_i / rho_j^2 only if len(model) == cols (:1377)
                    out[A.out_identity ? A.out_base + d : A.out_row[d]] = v * kf;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&gempty[buf]);
            nc++;
        }
        return;
    }

    // ------------------------------- Gram warps -------------------------------
    int tp[MT], tq[MT]; bool have[MT];
#pragma unroll
    for (int m = 0; m < MT; m++) {
        const int t = tid + m * J2_GT;
        have[m] = t < A.n_tiles;
        tp[m] = have[m] ? A.tile_tp[t] : 0; tq[m] = have[m] ? A.tile_tq[t] : 0;
    }
    double acc[MT][16];
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int x = 0; x < 16; x++) acc[m][x] = 0.0;
    const int szQ = NL * A.nQp;
    // V(item) = w_k (K + k^2 size mu) U_Q  into sV[which]
    auto compute_v = [&](uint32_t n_item, int kk) {
        const unsigned char *sb = ring + (size_t)(n_item % J2_SLOTS) * A.slot_bytes;
        const double *rec = reinterpret_cast<const double *>(sb);
        const double *uq = reinterpret_cast<const double *>(sb + (A.shared ? A.rec_bytes : A.uq_off));
        double *v = sV + (n_item & 1u) * szQ;
        const double k = A.kvals[kk], wk = A.kw[kk];
        const double k2s = k * k * rec[NL * NL];
        for (int x = tid; x < szQ; x += J2_GT) {
            const int i = x / A.nQp, q = x - i * A.nQp;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NL; j++) s = fma(fma(k2s, A.mu[i * NL + j], rec[i * NL + j]), uq[j * A.nQp + q], s);
            v[x] = s * wk;
        }
    };
    uint32_t n = 0, nc = 0;
    const uint32_t n_items = (uint32_t)(ci_hi - ci_lo) * (uint32_t)nK;
    int col = my_lo;
    while (col < my_hi && A.jac_col_ptr[col] == A.jac_col_ptr[col + 1]) col++;
    if (n_items > 0) {
        mbar_wait(&full[0], 0u);
        compute_v(0u, 0);
        named_bar_sync(2, J2_GT);
    }
    int ci = ci_lo, kk = 0;
    for (; n < n_items; n++) {
        // V of the next item while its data are fresh; the register tiles of this one
        int ci_n = ci, kk_n = kk + 1;
        if (kk_n == nK) { kk_n = 0; ci_n++; }
        if (n + 1 < n_items) {
            const uint32_t sl = (n + 1) % J2_SLOTS, use = (n + 1) / J2_SLOTS;
            mbar_wait(&full[sl], use & 1u);
            compute_v(n + 1, kk_n);
        }
        {
            const unsigned char *sb = ring + (size_t)(n % J2_SLOTS) * A.slot_bytes;
            const double *up = reinterpret_cast<const double *>(sb + A.rec_bytes);
            const double *v = sV + (n & 1u) * szQ;
#pragma unroll
            for (int m = 0; m < MT; m++) {
                if (have[m]) {
#pragma unroll
                    for (int i = 0; i < NL; i++) {
                        const double2 u01 = *reinterpret_cast<const double2 *>(up + i * A.nPp + 4 * tp[m]);
                        const double2 u23 = *reinterpret_cast<const double2 *>(up + i * A.nPp + 4 * tp[m] + 2);
                        const double2 v01 = *reinterpret_cast<const double2 *>(v + i * A.nQp + 4 * tq[m]);
                        const double2 v23 = *reinterpret_cast<const double2 *>(v + i * A.nQp + 4 * tq[m] + 2);
                        const double u[4] = {u01.x, u01.y, u23.x, u23.y};
                        const double w[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int b = 0; b < 4; b++) acc[m][a * 4 + b] = fma(u[a], w[b], acc[m][a * 4 + b]);
                    }
                }
            }
        }
        named_bar_sync(2, J2_GT);                     // V(n+1) visible; everybody is done with item n
        if (tid == 0) mbar_arrive(&empty[n % J2_SLOTS]);
        // last item of the column?
        const bool col_done = (kk == nK - 1) && (ci + 1 == A.jac_col_ptr[col + 1]);
        if (col_done) {
            const uint32_t buf = nc & 1u, use = nc >> 1;
            if (use > 0) mbar_wait(&gempty[buf], (use & 1u) ^ 1u);
            double *G = sG + buf * gsz;
#pragma unroll
            for (int m = 0; m < MT; m++) {
                if (have[m]) {
                    const int t = tid + m * J2_GT;
#pragma unroll
                    for (int x = 0; x < 16; x++) { G[x * A.PL + t] = acc[m][x]; acc[m][x] = 0.0; }
                }
            }
            named_bar_sync(2, J2_GT);
            if (tid == 0) mbar_arrive(&gfull[buf]);
            nc++;
            col++;
            while (col < my_hi && A.jac_col_ptr[col] == A.jac_col_ptr[col + 1]) col++;
        }
        ci = ci_n; kk = kk_n;
    }
}

// y = l .* (J (r .* x))  (J column-major [cols][ld]):  y[d] = l[d] * sum_j Jt[j][d] r[j] x[j];  l, r optional
// (MultLeftRightMatrix of the inversion, pygimli/frameworks/inversion.py:705-708)
__global__ void k_jac_mult(const double *__restrict__ Jt, size_t ld, int rows, int cols, const double *__restrict__ x,
                           const double *__restrict__ left, const double *__restrict__ right, double *__restrict__ y) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= rows) return;
    const int j0 = blockIdx.y * 256, j1 = min(cols, j0 + 256);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = j0;
    for (; j + 4 <= j1; j += 4) {
        const double *c = Jt + (size_t)j * ld + d;
        const double v0 = c[0], v1 = c[ld], v2 = c[2 * ld], v3 = c[3 * ld];
        a0 = fma(v0, right ? x[j] * right[j] : x[j], a0);
        a1 = fma(v1, right ? x[j + 1] * right[j + 1] : x[j + 1], a1);
        a2 = fma(v2, right ? x[j + 2] * right[j + 2] : x[j + 2], a2);
        a3 = fma(v3, right ? x[j + 3] * right[j + 3] : x[j + 3], a3);
    }
    for (; j < j1; j++) a0 = fma(Jt[(size_t)j * ld + d], right ? x[j] * right[j] : x[j], a0);
    const double acc = (a0 + a1) + (a2 + a3);
    atomicAdd(y + d, left ? left[d] * acc : acc);
}
// one warp per column j of J^T:
//   MODE 0  y[j] = r[j] * sum_d Jt[j][d] l[d] x[d]            (transMult; l, r optional)
//   MODE 1  y[j] = sum_d |Jt[j][d] x[d]| / |r[j]|             (coverageDCtrans, bertJacobian.cpp:569-598; x = dd, r = mm)
//   MODE 2  y[j] = sum_d |Jt[j][d] x[d]|                      (partial coverage of a row shard; divided after the all-reduce)
template <int MODE>
__global__ void k_jac_tmult(const double *__restrict__ Jt, size_t ld, int rows, int cols, const double *__restrict__ x,
                            const double *__restrict__ left, const double *__restrict__ right, double *__restrict__ y) {
    const int j = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (j >= cols) return;
    const int lane = threadIdx.x & 31;
    const double *c = Jt + (size_t)j * ld;
    double a0 = 0.0, a1 = 0.0;
    int d = lane;
    for (; d + 32 < rows; d += 64) {
        const double v0 = c[d], v1 = c[d + 32];
        const double x0 = (MODE == 0 && left) ? left[d] * x[d] : x[d], x1 = (MODE == 0 && left) ? left[d + 32] * x[d + 32] : x[d + 32];
        if (MODE == 0) { a0 = fma(v0, x0, a0); a1 = fma(v1, x1, a1); }
        else { a0 += fabs(v0 * x0); a1 += fabs(v1 * x1); }
    }
    if (d < rows) {
        const double x0 = (MODE == 0 && left) ? left[d] * x[d] : x[d];
        if (MODE == 0) a0 = fma(c[d], x0, a0); else a0 += fabs(c[d] * x0);
    }
    double acc = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        if (MODE == 0) y[j] = right ? right[j] * acc : acc;
        else if (MODE == 1) y[j] = acc / fabs(right[j]);
        else y[j] = acc;
    }
}
// row-major copy of the column-major J: out[d][j] = Jt[j][d]
__global__ void k_jac_to_rowmajor(const double *__restrict__ Jt, size_t ld, int rows, int cols, double *__restrict__ out) {
    __shared__ double tile[32][33];
    int d = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 32 + threadIdx.y;
    for (int t = 0; t < 32; t += 8) if (d < rows && j + t < cols) tile[threadIdx.y + t][threadIdx.x] = Jt[(size_t)(j + t) * ld + d];
    __syncthreads();
    int j2 = blockIdx.y * 32 + threadIdx.x, d2 = blockIdx.x * 32 + threadIdx.y;
    for (int t = 0; t < 32; t += 8) if (d2 + t < rows && j2 < cols) out[(size_t)(d2 + t) * cols + j2] = tile[threadIdx.x][threadIdx.y + t];
}

} // namespace pgb
