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
};

// rmax: rows per panel (<= hc), hc: halo entries per chunk, max_chunks: chunks per panel.  Returns "" or an error text.
inline std::string build_stream_panels(int n, const int *rowptr, const int *colidx, int rmax, int hc, int max_chunks,
                                       StreamPanelsHost &S) {
    if (n < 0 || rmax < 1 || hc < rmax || max_chunks < 1) return "invalid stream-panel limits";
    S = StreamPanelsHost();
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
        if (nch > 1) {
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

} // namespace pgb
