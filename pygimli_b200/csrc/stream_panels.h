// stream_panels.h -- host-side layout of the streamed row-panel SpMM (k_spmm_stream, ert_kernels.cuh).
//
// A matrix level (the fine stiffness matrix or a coarse level of the multilevel preconditioner) is cut into ROW PANELS of
// consecutive rows (rows are numbered along a space-filling curve, so a panel is a compact patch of the mesh).  The
// distinct columns a panel touches form its HALO list, ordered  [the panel's own rows | other columns ascending].  The halo list is cut into CHUNKS of at most `hc` entries; chunk 0 always holds the panel's own rows.  The CSR
// entries of a panel are re-ordered chunk-major: (chunk, row, ascending column), so that
//     one pipeline stage of the kernel = one (panel, chunk):
//        * the chunk's rows of the block vector X          -> one TMA bulk copy per RUN of consecutive halo rows when a
//                                                             tile spans whole rows of X (runs are listed here), else per row
//        * the chunk's packed entries {value, row-in-chunk} -> ONE contiguous TMA bulk copy
//        * the chunk's per-row entry ranges (crp)            -> ONE contiguous TMA bulk copy
// Accumulators stay in registers across the chunks of a panel; chunks are processed last-to-first so that the panel's own
// X rows (chunk 0) are resident when the epilogue needs them.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

namespace pgb {

struct StreamPanelsHost {
    int n_rows = 0, n_panels = 0, n_chunks = 0, crp_stride = 0, max_rows = 0, max_chunk_halo = 0, max_chunk_ent = 0;
    long long nnz = 0;
    std::vector<int> panel_row_ptr;     // [n_panels + 1]
    std::vector<int> panel_chunk_ptr;   // [n_panels + 1]
    std::vector<int> chunk_halo_ptr;    // [n_chunks + 1] into halo_cols
    std::vector<int> halo_cols;         // column (= row of X) of every halo entry
    std::vector<int> chunk_ent_ptr;     // [n_chunks + 1] into the re-ordered entry arrays
    std::vector<int> ent_src;           // [nnz] CSR slot of the re-ordered entry
    std::vector<unsigned> ent_idx;      // [nnz] row of the entry's column inside its chunk's staged X tile
    std::vector<int> crp;               // [n_chunks * crp_stride] entry range starts per row, relative to the chunk
    std::vector<int> chunk_run_ptr;     // [n_chunks + 1] runs of consecutive columns inside a chunk's halo list ...
    std::vector<int> run_start;         // ... first halo entry of the run, relative to the chunk
    std::vector<int> run_col;           // ... its column
    std::vector<int> run_len;           // ... number of consecutive columns
    int max_chunk_runs = 0;
    // ---- dense 8-row-group form for the FP64 tensor-core kernel (k_spmm_mma; built when mma_groups > 0) ----
    // Rows of a panel are cut into GROUPS of 8 consecutive rows (one per consumer warp).  Per (chunk, group) the distinct
    // halo entries the group's rows touch inside that chunk (the "union columns") are cut into K-STEPS of 4: one k-step is
    // one DMMA m8n8k4 per 8 source columns, A = the group's 8 x 4 block of matrix values (zeros where a row has no entry).
    int mma_groups = 0, meta_gstride = 0, max_chunk_ks = 0, max_chunk_meta = 0;
    long long n_ks = 0;                 // k-steps in total
    std::vector<int> chunk_ks_ptr;      // [n_chunks + 1] first k-step of a chunk
    std::vector<int> a_src;             // [n_ks * 32] CSR slot of A-fragment element (lane = 4 * row-in-group + column-in-step), -1 = zero
    std::vector<int> chunk_meta_ptr;    // [n_chunks + 1] offset (in 32-bit words, multiple of 4) of the chunk's meta block
    std::vector<unsigned> meta;         // per chunk: [k-step range start of every group (meta_gstride words) | one word per k-step:
                                        //             the 4 staged-row indices of its columns, one byte each]
};

// rmax: rows per panel (<= hc), hc: halo entries per chunk, max_chunks: chunks per panel.  Returns "" or an error text.
// mma_groups > 0: additionally build the 8-row-group form (rmax <= 8 * mma_groups, hc <= 255); rowb_hint = bytes of one
// staged X row in the most common launch (orders the columns of a k-step so that the four 64-byte B-fragment spans fall
// into different shared-memory banks).
inline std::string build_stream_panels(int n, const int *rowptr, const int *colidx, int rmax, int hc, int max_chunks,
                                       StreamPanelsHost &S, int mma_groups = 0, int rowb_hint = 800) {
    // 8-row-group form: the panel's own rows may span chunks 0 AND 1 (the kernel keeps both resident for the epilogue), so
    // chunks can be half a panel: a deeper ring of smaller stages
    if (n < 0 || rmax < 1 || max_chunks < 1 || (mma_groups > 0 ? 2 * hc < rmax : hc < rmax)) return "invalid stream-panel limits";
    if (mma_groups > 0 && (rmax > 8 * mma_groups || hc > 255 || max_chunks < 2)) return "invalid stream-panel limits (8-row groups)";
    S = StreamPanelsHost();
    S.mma_groups = mma_groups;
    S.meta_gstride = mma_groups > 0 ? (mma_groups + 1 + 3) / 4 * 4 : 0;
    if (mma_groups > 0) { S.chunk_ks_ptr.push_back(0); S.chunk_meta_ptr.push_back(0); }
    std::vector<int> upos((size_t)std::max(hc, 1), -1), ulist, usrc, order;
    S.n_rows = n; S.nnz = n > 0 ? rowptr[n] : 0;
    S.crp_stride = (rmax + 1 + 3) / 4 * 4;                       // 16-byte multiple: one bulk copy
    const int hmax = hc * max_chunks;
    std::vector<int> stamp((size_t)std::max(n, 1), -1), slot((size_t)std::max(n, 1), 0);
    S.ent_src.resize((size_t)S.nnz); S.ent_idx.resize((size_t)S.nnz);
    S.panel_row_ptr.push_back(0); S.panel_chunk_ptr.push_back(0); S.chunk_halo_ptr.push_back(0); S.chunk_ent_ptr.push_back(0);
    S.chunk_run_ptr.push_back(0);
    std::vector<int> others;
    long long ent_pos = 0;
    int row = 0, np = 0;
    while (row < n) {
        const int start = row;
        // 1. extent: add rows while the union of {own rows} and {touched columns} stays within hmax
        int h = 0;
        others.clear();
        while (row < n && row - start < rmax) {
            int add = 0;
            if (stamp[row] != np) add++;
            for (int p = rowptr[row]; p < rowptr[row + 1]; p++) { const int c = colidx[p]; if (c != row && stamp[c] != np) add++; }
            if (h + add > hmax) {
                if (row == start) return "a single matrix row exceeds the halo limit of the streamed SpMM";
                break;
            }
            if (stamp[row] != np) { stamp[row] = np; h++; }
            for (int p = rowptr[row]; p < rowptr[row + 1]; p++) {
                const int c = colidx[p];
                if (c < 0 || c >= n) return "column index out of range";
                if (stamp[c] != np) { stamp[c] = np; h++; }
            }
            row++;
        }
        const int nrows = row - start;
        // 2. halo order: own rows first, then the other columns ascending (neighbouring patches of the space-filling
        //    curve appear as runs of consecutive rows of X -> few, large bulk copies)
        for (int r = start; r < row; r++) slot[r] = r - start;
        for (int r = start; r < row; r++)
            for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                const int c = colidx[p];
                if ((c < start || c >= row) && stamp[c] == np) { stamp[c] = -2 - np; others.push_back(c); }
            }
        std::sort(others.begin(), others.end());
        int hn = nrows;
        for (int c : others) slot[c] = hn++;
        const int nch = (hn + hc - 1) / hc;
        // chunk boundaries: balanced, chunk 0 holds at least the own rows
        std::vector<int> cb((size_t)nch + 1, hn);
        cb[0] = 0;
        if (nch > 1 && mma_groups > 0 && nrows > hc) {
            // more own rows than a chunk holds: chunks 0 and 1 share them (both at least half the panel), the rest is balanced
            const int c01 = std::max((hn + nch - 1) / nch, (nrows + 1) / 2);
            cb[1] = std::min(hn, c01);
            if (nch > 2) {
                cb[2] = std::min(hn, 2 * c01);
                const int rem = hn - cb[2], per = (rem + nch - 3) / (nch - 2);
                for (int c = 3; c < nch; c++) cb[c] = std::min(hn, cb[2] + (c - 2) * per);
            }
        } else if (nch > 1) {
            const int c0size = std::max((hn + nch - 1) / nch, nrows);
            const int rem = hn - c0size, per = (rem + nch - 2) / (nch - 1);
            for (int c = 1; c < nch; c++) cb[c] = std::min(hn, c0size + (c - 1) * per);
        }
        for (int c = 0; c < nch; c++) {
            if (cb[c + 1] - cb[c] > hc) return "internal error: chunk exceeds the halo limit";
            for (int i = cb[c]; i < cb[c + 1]; i++) S.halo_cols.push_back(i < nrows ? start + i : others[(size_t)(i - nrows)]);
            S.chunk_halo_ptr.push_back((int)S.halo_cols.size());
            S.max_chunk_halo = std::max(S.max_chunk_halo, cb[c + 1] - cb[c]);
            {   // runs of consecutive columns
                const int *hc0 = S.halo_cols.data() + (S.halo_cols.size() - (size_t)(cb[c + 1] - cb[c]));
                const int len = cb[c + 1] - cb[c];
                int nr = 0;
                for (int i = 0; i < len;) {
                    int e = i + 1;
                    while (e < len && hc0[e] == hc0[e - 1] + 1) e++;
                    S.run_start.push_back(i); S.run_col.push_back(hc0[i]); S.run_len.push_back(e - i);
                    i = e; nr++;
                }
                S.chunk_run_ptr.push_back((int)S.run_start.size());
                S.max_chunk_runs = std::max(S.max_chunk_runs, nr);
            }
            // entries of this chunk, row by row
            const size_t crp0 = S.crp.size();
            S.crp.resize(crp0 + (size_t)S.crp_stride, 0);
            int cnt = 0;
            for (int r = start; r < row; r++) {
                S.crp[crp0 + (size_t)(r - start)] = cnt;
                for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                    const int s = slot[colidx[p]];
                    if (s >= cb[c] && s < cb[c + 1]) { S.ent_src[(size_t)ent_pos] = p; S.ent_idx[(size_t)ent_pos] = (unsigned)(s - cb[c]); ent_pos++; cnt++; }
                }
            }
            for (int i = nrows; i < S.crp_stride; i++) S.crp[crp0 + (size_t)i] = cnt;
            S.chunk_ent_ptr.push_back((int)ent_pos);
            S.max_chunk_ent = std::max(S.max_chunk_ent, cnt);
            if (mma_groups > 0) {
                const size_t m0 = S.meta.size();
                S.meta.resize(m0 + (size_t)S.meta_gstride, 0u);
                int ks_chunk = 0;
                for (int g = 0; g < mma_groups; g++) {
                    S.meta[m0 + (size_t)g] = (unsigned)ks_chunk;
                    const int ra = start + 8 * g, rb = std::min(row, ra + 8);
                    if (ra >= row) continue;
                    // union of the group's columns inside this chunk, in first-touch order; usrc[u * 8 + r] = CSR slot or -1
                    ulist.clear(); usrc.clear();
                    for (int r = ra; r < rb; r++)
                        for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                            const int s = slot[colidx[p]];
                            if (s < cb[c] || s >= cb[c + 1]) continue;
                            const int i = s - cb[c];
                            if (upos[(size_t)i] < 0) { upos[(size_t)i] = (int)ulist.size(); ulist.push_back(i); usrc.insert(usrc.end(), 8, -1); }
                            usrc[(size_t)upos[(size_t)i] * 8 + (size_t)(r - ra)] = p;
                        }
                    const int nu = (int)ulist.size();
                    for (int i : ulist) upos[(size_t)i] = -1;
                    if (nu == 0) continue;
                    // order the columns: a 64-bit LDS is served per half-warp, i.e. per k-step four 32-byte segments (one per
                    // column, start = idx * rowb mod 128, in units of 16 bytes: classes 0..7, a segment covers 2 consecutive
                    // classes): every k-step takes 4 columns whose segments overlap as little as possible
                    std::vector<int> bucket[8];
                    for (int u = nu - 1; u >= 0; u--) bucket[((long long)ulist[(size_t)u] * rowb_hint % 128) / 16].push_back(u);
                    const int nks = (nu + 3) / 4;
                    order.assign((size_t)nks * 4, -1);
                    int left = nu;
                    for (int k = 0; k < nks; k++) {
                        int cover[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int j = 0; j < 4 && left > 0; j++) {
                            int best = -1, best_max = 1 << 30; size_t best_size = 0;
                            for (int cl = 0; cl < 8; cl++) {
                                if (bucket[cl].empty()) continue;
                                int mx = 0;
                                for (int q = 0; q < 8; q++) { const int add = (((q - cl) & 7) < 2) ? 1 : 0; mx = std::max(mx, cover[q] + add); }
                                if (mx < best_max || (mx == best_max && bucket[cl].size() > best_size)) { best = cl; best_max = mx; best_size = bucket[cl].size(); }
                            }
                            for (int q = 0; q < 8; q++) if (((q - best) & 7) < 2) cover[q]++;
                            order[(size_t)k * 4 + (size_t)j] = bucket[best].back(); bucket[best].pop_back(); left--;
                        }
                    }
                    for (int k = 0; k < nks; k++) {
                        unsigned word = 0;
                        const size_t a0 = S.a_src.size();
                        S.a_src.resize(a0 + 32, -1);
                        for (int j = 0; j < 4; j++) {
                            const int u = order[(size_t)k * 4 + (size_t)j];
                            const int idx = u >= 0 ? ulist[(size_t)u] : ulist[(size_t)order[(size_t)k * 4]];   // padding: any staged row, zero values
                            word |= (unsigned)idx << (8 * j);
                            if (u >= 0) for (int r = 0; r < 8; r++) S.a_src[a0 + (size_t)(4 * r + j)] = usrc[(size_t)u * 8 + (size_t)r];
                        }
                        S.meta.push_back(word);
                    }
                    ks_chunk += nks;
                }
                for (int g = mma_groups; g < S.meta_gstride; g++) S.meta[m0 + (size_t)g] = (unsigned)ks_chunk;
                // groups without rows point at the end too (written as 'continue' above left them at their running value)
                for (int g = 0; g < mma_groups; g++) if (start + 8 * g >= row) S.meta[m0 + (size_t)g] = (unsigned)ks_chunk;
                while (S.meta.size() % 4) S.meta.push_back(0u);
                S.n_ks += ks_chunk;
                S.chunk_ks_ptr.push_back((int)S.n_ks);
                S.chunk_meta_ptr.push_back((int)S.meta.size());
                S.max_chunk_ks = std::max(S.max_chunk_ks, ks_chunk);
                S.max_chunk_meta = std::max(S.max_chunk_meta, (int)(S.meta.size() - m0));
            }
        }
        S.n_chunks += nch;
        S.max_rows = std::max(S.max_rows, nrows);
        np++;
        S.panel_row_ptr.push_back(row);
        S.panel_chunk_ptr.push_back(S.n_chunks);
    }
    S.n_panels = np;
    if (ent_pos != S.nnz) return "internal error: entries lost while building stream panels";
    return "";
}

// ---- gather form (k_spmm_gather): narrow column windows (multi-GPU source shards) -------------------------------------
// No shared-memory staging: the block vectors of a narrow window stay in L2, so every 8-row group simply lists its k-steps --
// 32 A-fragment slots and the four GLOBAL column indices of the k-step -- over the union of its rows' columns (ascending).
struct GatherGroupsHost {
    int n_groups = 0; long long n_ks = 0;
    std::vector<int> ks_ptr;     // [n_groups + 1]
    std::vector<int> cols;       // [4 * n_ks] column of X of every k-step slot (padding: the group's first row)
    std::vector<int> a_src;      // [32 * n_ks] CSR slot of fragment element (4 * row-in-group + slot) or -1
};
inline void build_gather_groups(int n, const int *rowptr, const int *colidx, GatherGroupsHost &G) {
    G = GatherGroupsHost();
    G.n_groups = (n + 7) / 8;
    G.ks_ptr.assign(1, 0);
    std::vector<int> ucols, pos((size_t)std::max(n, 1), -1);
    for (int g = 0; g < G.n_groups; g++) {
        const int ra = 8 * g, rb = std::min(n, ra + 8);
        ucols.clear();
        for (int r = ra; r < rb; r++) for (int p = rowptr[r]; p < rowptr[r + 1]; p++) if (pos[(size_t)colidx[p]] < 0) { pos[(size_t)colidx[p]] = 0; ucols.push_back(colidx[p]); }
        std::sort(ucols.begin(), ucols.end());
        for (size_t u = 0; u < ucols.size(); u++) pos[(size_t)ucols[u]] = (int)u;
        const int nks = ((int)ucols.size() + 3) / 4;
        const size_t c0 = G.cols.size(), a0 = G.a_src.size();
        G.cols.resize(c0 + (size_t)4 * nks, ra); G.a_src.resize(a0 + (size_t)32 * nks, -1);
        for (size_t u = 0; u < ucols.size(); u++) G.cols[c0 + u] = ucols[u];
        for (int r = ra; r < rb; r++)
            for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                const int u = pos[(size_t)colidx[p]];
                G.a_src[a0 + (size_t)(u / 4) * 32 + (size_t)(4 * (r - ra) + (u & 3))] = p;
            }
        for (int c : ucols) pos[(size_t)c] = -1;
        G.n_ks += nks;
        G.ks_ptr.push_back((int)G.n_ks);
    }
}

} // namespace pgb
