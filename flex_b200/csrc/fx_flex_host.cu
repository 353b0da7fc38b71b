// fx_flex_host.cu -- host part of the pillar format: round 1 of Mat::csr2_DiagTiling (mat.cu:680-759).
//
// Round 1 grows each diagonal block row by row from where the previous block ended, so the blocks form one serial chain
// along the diagonal: it stays on the host (flat arrays instead of the reference's per-row hash maps).  Rounds 2 and 3
// (mat.cu:771-903) are independent per block / per row panel and run on the GPU: fx_flex_build.cu.
#include <algorithm>
#include <cmath>

#include "fx_common.cuh"
#include "fx_flex.cuh"

namespace fx {

namespace {
bool has_col(const uint32_t* rowptr, const uint32_t* col, int r, unsigned c) {
  const uint32_t *b = col + rowptr[r], *e = col + rowptr[r + 1];
  const uint32_t* it = std::lower_bound(b, e, c);
  return it != e && *it == c;
}

}  // namespace

// Round 1 of csr2_DiagTiling (mat.cu:706-759) on the host: every diagonal block starts where the previous one ended and grows
// row by row until it holds its share of nz -- one serial chain along the diagonal.  Outputs what rounds 2 and 3 (on the GPU,
// fx_flex_build.cu) need: the block widths, the number of non-empty blocks and one "listed" byte per column.
//
// Why one byte per column is enough.  The reference keeps nnz_claimed (per row, the nz inside a block's square) and
// alpha_columns_per_sm (per SM of 64 blocks, the columns seen inside its squares), and round 2 takes a nz (j, l) of a row of SM s
// when l lies in the SM's column window and (the nz is claimed or s lists l).  A column is listed only by the block whose rows
// contain it (mat.cu:722,737: mat_r_start <= l <= j), i.e. by the SM whose window contains it, and a claimed nz always has its
// column listed by its own block -- so inside the window "claimed or listed by s" is simply listed[l], and the reference's
// assert at mat.cu:818 cannot fire.
int diag_round1_host(int M, int nnz, const uint32_t* rowptr, const uint32_t* col, int n_sm, std::vector<int>& tile_width,
                     int& warps_with_weights, std::vector<uint8_t>& listed) {
  const int warps_per_sm = 64;   // mat.cu:688
  const float alpha = 0.3f;      // mat.cu:690
  const int nnz_diagonal_tiles = (int)(alpha * rowptr[M]);
  const int partitions_node = warps_per_sm * n_sm;
  const int nnz_p_diagonal_tile = std::max(32, nnz_diagonal_tiles / partitions_node);
  const int thr = (int)(0.85 * nnz_p_diagonal_tile);
  tile_width.assign(partitions_node, 0);
  listed.assign(std::max(M, 1), 0);
  int mat_r_start = 0;
  warps_with_weights = 0;
  for (int i = 0; i < partitions_node; ++i) {
    mat_r_start += i ? tile_width[i - 1] : 0;
    int cnt = 0, j = mat_r_start;
    while (j < M && cnt <= thr) {
      for (unsigned kk = rowptr[j];; ++kk) {
        // the reference walks until it meets column j and does not stop at the end of the row
        // (mat.cu:718-727); running off the array is undefined there and refused here
        FX_REQUIRE(kk < (unsigned)nnz, FX_ERR_FORMAT,
                   "row %d has no diagonal entry and the reference's diagonal walk (mat.cu:718-727) leaves the matrix", j);
        if (!(col[kk] <= (unsigned)j)) break;
        if ((int)col[kk] >= mat_r_start) { ++cnt; listed[col[kk]] = 1; }
        if (col[kk] == (unsigned)j) break;
      }
      for (int kk = mat_r_start; kk < j; ++kk)
        if (has_col(rowptr, col, kk, (unsigned)j)) { ++cnt; listed[j] = 1; }
      ++j;
    }
    warps_with_weights += cnt > 0;
    FX_REQUIRE(j >= M || cnt, FX_ERR_FORMAT, "empty diagonal block (assert mat.cu:751)");
    tile_width[i] = j - mat_r_start;
  }
  long verify_m = 0;
  for (int w : tile_width) verify_m += w;
  FX_REQUIRE(verify_m == M, FX_ERR_FORMAT, "alpha is too small: diagonal blocks cover %ld of %d rows (assert mat.cu:759)", verify_m, M);
  FX_REQUIRE(partitions_node != warps_with_weights, FX_ERR_FORMAT, "no idle warp left for the balance queue (division by zero at mat.cu:862)");
  return FX_OK;
}

}  // namespace fx
