/*
 * fx_oracle_flex.c -- TEST INFRASTRUCTURE ONLY (see fx_oracle.h).
 * C restatement of the reference's Flex tile-format builders (mat.cu):
 *   F1 csr2flex_Rmajor / csr2flex_Cmajor  mat.cu:1345-1518   (tile format, v4-v9)
 *   F2 csr2seg_Cmajor                     mat.cu:1192-1269   (row-panel segmentation)
 *   F3 the tile-segment arrays v10-v35 read, derived from the same segmentation
 *   F4 SM bucketing of Mat::csr2tile      mat.cu:1097-1162   (row_based_split)
 *   F5 csr2_DiagTiling                    mat.cu:680-903     (pillar / "alpha" format, HEAD)
 * and of the SpMM through the pillar format (alpha_w_atomic_spmm_v36, flex.cu:4010-4124).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fx_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* F1                                                                                         */
/* ------------------------------------------------------------------------------------------ */
int orc_flex_tile_build(int m, const uint32_t *rowptr, const uint32_t *col, const float *val, int tm,
                        int tn, int cmajor, orc_flextile *o) {
  memset(o, 0, sizeof(*o));
  int n = m, nnz = (int)rowptr[m], npanels = (m + tm - 1) / tm;
  for (int r = 0; r < m; ++r) if (rowptr[r] == rowptr[r + 1]) return -1; /* "we assume there is no empty row" :1359 */
  o->m = m; o->tm = tm; o->tn = tn; o->npanels = npanels; o->nnz = nnz;
  o->tileRowPtr = calloc((size_t)npanels + 1, sizeof(uint32_t));
  size_t cap = (size_t)nnz + 1; /* a tile holds at least one nz */
  o->tileNnz = calloc(cap + 1, sizeof(uint32_t));
  o->nnzTile = calloc(cap, sizeof(int));
  o->bitMap = calloc(cap, sizeof(int));
  o->tileColIdx = calloc(cap, sizeof(uint32_t));
  o->rcOffset = calloc(cap, sizeof(int));
  o->newVals = calloc(cap, sizeof(float));
  int *cOffset = malloc(sizeof(int) * (size_t)tm);
  int pos = 0, nt = 0;
  for (int ridx = 0; ridx < npanels; ++ridx) {
    int rowStart = ridx * tm, rowEnd = (ridx + 1) * tm < m ? (ridx + 1) * tm : m;
    int left = n;
    for (int i = rowStart; i < rowEnd; ++i) {
      cOffset[i - rowStart] = 0;
      if ((int)col[rowptr[i]] < left) left = (int)col[rowptr[i]];
    }
    int right = left + tn < n ? left + tn : n;
    int tiles = 0;
    while (pos < (int)rowptr[rowEnd]) {
      int nnzInTile = 0, bit_map = 0;
      tiles++;
      if (!cmajor) {
        for (int i = rowStart; i < rowEnd; ++i) {
          int c = (int)rowptr[i] + cOffset[i - rowStart];
          while (c < (int)rowptr[i + 1] && (int)col[c] < right) {
            int rc16 = ((i - rowStart) << 16) | ((int)col[c] - left);
            bit_map |= 1 << ((int)col[c] - left);
            o->newVals[pos] = val[c];
            o->rcOffset[pos] = rc16;
            ++c; pos++; cOffset[i - rowStart]++; nnzInTile++;
          }
        }
      } else {
        for (int itn = 0; itn < tn; ++itn)
          for (int i = rowStart; i < rowEnd; ++i) {
            int c = (int)rowptr[i] + cOffset[i - rowStart];
            if (c < (int)rowptr[i + 1] && (int)col[c] == left + itn && (int)col[c] < right) {
              int rc16 = ((i - rowStart) << 16) | ((int)col[c] - left);
              bit_map |= 1 << ((int)col[c] - left);
              o->newVals[pos] = val[c];
              o->rcOffset[pos] = rc16;
              pos++; cOffset[i - rowStart]++; nnzInTile++;
            }
          }
      }
      o->nnzTile[nt] = nnzInTile;
      o->bitMap[nt] = bit_map;
      o->tileNnz[nt + 1] = o->tileNnz[nt] + (uint32_t)nnzInTile;
      o->tileColIdx[nt] = (uint32_t)left;
      nt++;
      left = n;
      for (int i = rowStart; i < rowEnd; ++i) {
        int rnnz = (int)(rowptr[i + 1] - rowptr[i]);
        if (cOffset[i - rowStart] < rnnz) {
          int cc = (int)col[rowptr[i] + cOffset[i - rowStart]];
          if (cc < left) left = cc;
        }
      }
      right = left + tn < n ? left + tn : n;
    }
    o->tileRowPtr[ridx + 1] = o->tileRowPtr[ridx] + (uint32_t)tiles;
  }
  o->ntiles = nt;
  free(cOffset);
  return 0;
}
void orc_flextile_free(orc_flextile *t) {
  free(t->tileRowPtr); free(t->tileNnz); free(t->nnzTile); free(t->bitMap); free(t->tileColIdx);
  free(t->rcOffset); free(t->newVals);
  memset(t, 0, sizeof(*t));
}

/* SpMM through the tile format (what v4-v9 compute, flex.cu:329-1117): C[panel*tm + r, :] +=
 * val * B[tileColIdx + c, :] with (r,c) unpacked from rcOffset (RC16, flex.cu:841). */
void orc_flextile_spmm(const orc_flextile *t, const float *B, int k, float *C) {
  for (int64_t i = 0; i < (int64_t)t->m * k; ++i) C[i] = 0.f;
  for (int p = 0; p < t->npanels; ++p)
    for (uint32_t ti = t->tileRowPtr[p]; ti < t->tileRowPtr[p + 1]; ++ti)
      for (uint32_t e = t->tileNnz[ti]; e < t->tileNnz[ti + 1]; ++e) {
        int r = t->rcOffset[e] >> 16, c = t->rcOffset[e] & 0xffff;
        const float *b = B + (int64_t)(t->tileColIdx[ti] + (uint32_t)c) * k;
        float *o = C + (int64_t)(p * t->tm + r) * k;
        for (int j = 0; j < k; ++j) o[j] = fmaf(t->newVals[e], b[j], o[j]);
      }
}

/* ------------------------------------------------------------------------------------------ */
/* F2 / F3: one panel of csr2seg_Cmajor, appended to the growing arrays                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t *rowPtr; size_t nrp, caprp;   /* alpha_rowPtr */
  uint32_t *colIdx; float *vals; size_t nnz; /* alpha_colIdx / alpha_vals, capacity = total nnz */
  uint32_t *pillar; size_t npil, cappil; /* alpha_pillar_rowPtr */
  uint32_t *voMap; size_t nvm, capvm;    /* segVoMap */
} alpha_t;
static void push_u32(uint32_t **a, size_t *n, size_t *cap, uint32_t v) {
  if (*n + 1 > *cap) { *cap = *cap ? *cap * 2 : 1024; *a = realloc(*a, sizeof(uint32_t) * *cap); }
  (*a)[(*n)++] = v;
}

/* claimed[e] != 0: nz e already belongs to a diagonal tile (duplicate_sparse_mat, :1220-1223) */
static int seg_panel(int m, const uint32_t *rowptr, const uint32_t *col, const float *val,
                     const int32_t *vo_mp, int tm, int nnz_limit, int ridx, unsigned char *claimed,
                     alpha_t *a, int *nnz_rowPtr) {
  int rowStart = ridx * tm, rowEnd = (ridx + 1) * tm < m ? (ridx + 1) * tm : m, rows = rowEnd - rowStart;
  int dif = (int)(0.1 * nnz_limit); /* :1203 */
  int *cur = malloc(sizeof(int) * (size_t)tm), *prev = malloc(sizeof(int) * (size_t)tm);
  int *atom = calloc((size_t)tm, sizeof(int));
  /* per-row list of unclaimed entries (in column order) */
  int total = (int)(rowptr[rowEnd] - rowptr[rowStart]);
  int *ent = malloc(sizeof(int) * (size_t)(total ? total : 1));
  int *estart = malloc(sizeof(int) * ((size_t)tm + 1));
  int ne = 0;
  for (int i = 0; i < rows; ++i) {
    estart[i] = ne;
    for (uint32_t e = rowptr[rowStart + i]; e < rowptr[rowStart + i + 1]; ++e) if (!claimed || !claimed[e]) ent[ne++] = (int)e;
    cur[i] = prev[i] = estart[i];
  }
  estart[rows] = ne;
  int nnzInSeg = 0, tiles = 0, remaining = ne;
  while (remaining > 0 || nnzInSeg > 0) {
    if (remaining > 0) {
      /* next distinct column among the unclaimed entries */
      uint32_t j = 0xffffffffu;
      for (int i = 0; i < rows; ++i) if (cur[i] < estart[i + 1] && col[ent[cur[i]]] < j) j = col[ent[cur[i]]];
      for (int i = 0; i < rows; ++i)
        if (cur[i] < estart[i + 1] && col[ent[cur[i]]] == j) {
          if (claimed) claimed[ent[cur[i]]] = 1; /* "mark it as visited" :1225 */
          cur[i]++; atom[i]++; nnzInSeg++; remaining--;
        }
    }
    /* cut: last column with something pending, or within `dif` of the limit, or over it (:1235) */
    if ((remaining == 0 && nnzInSeg) || (nnz_limit - nnzInSeg) <= dif || nnzInSeg > nnz_limit) {
      *nnz_rowPtr += nnzInSeg;
      for (int i = 0; i < rows; ++i) {
        int cnt = cur[i] - prev[i];
        push_u32(&a->rowPtr, &a->nrp, &a->caprp, a->rowPtr[a->nrp - 1] + (uint32_t)cnt);
        for (int q = prev[i]; q < cur[i]; ++q) { a->colIdx[a->nnz] = col[ent[q]]; a->vals[a->nnz] = val[ent[q]]; a->nnz++; }
        nnzInSeg -= cnt;
        prev[i] = cur[i];
      }
      for (int i = 0; i < rows; ++i) {
        int rl = (int)(rowptr[rowStart + i + 1] - rowptr[rowStart + i]);
        uint32_t v = (uint32_t)vo_mp[rowStart + i];
        push_u32(&a->voMap, &a->nvm, &a->capvm, atom[i] < rl ? (v | 0x80000000u) : v); /* :1252-1260 */
        atom[i] = 0;
      }
      push_u32(&a->pillar, &a->npil, &a->cappil, a->pillar[a->npil - 1] + (uint32_t)rows);
      tiles++;
    }
  }
  free(cur); free(prev); free(atom); free(ent); free(estart);
  return tiles;
}

int orc_seg_build(int m, const uint32_t *rowptr, const uint32_t *col, const float *val,
                  const int32_t *vo_mp, int tm, int nnz_limit, orc_seg *o) {
  memset(o, 0, sizeof(*o));
  int nnz = (int)rowptr[m], npanels = (m + tm - 1) / tm;
  for (int r = 0; r < m; ++r) if (rowptr[r] == rowptr[r + 1]) return -1; /* assert :1207 */
  alpha_t a;
  memset(&a, 0, sizeof(a));
  a.colIdx = malloc(sizeof(uint32_t) * (size_t)(nnz ? nnz : 1));
  a.vals = malloc(sizeof(float) * (size_t)(nnz ? nnz : 1));
  push_u32(&a.rowPtr, &a.nrp, &a.caprp, 0);
  push_u32(&a.pillar, &a.npil, &a.cappil, 0);
  o->segs_per_panel = malloc(sizeof(int) * (size_t)(npanels ? npanels : 1));
  int nnz_rowPtr = 0;
  for (int p = 0; p < npanels; ++p)
    o->segs_per_panel[p] = seg_panel(m, rowptr, col, val, vo_mp, tm, nnz_limit, p, NULL, &a, &nnz_rowPtr);
  o->m = m; o->tm = tm; o->nnz = nnz; o->npanels = npanels;
  o->nsegs = (int)a.npil - 1;
  o->rows_total = (int)a.nrp - 1;
  o->alpha_rowPtr = a.rowPtr; o->alpha_colIdx = a.colIdx; o->alpha_vals = a.vals;
  o->pillar_rowPtr = a.pillar; o->segVoMap = a.voMap;
  /* F3: the arrays kernels v10-v35 read (flex.cu:1130-1139, 2473-2479, 3555-3567): per segment the
   * nz in COLUMN-major order as (rowInSeg, absCol) pairs, values alongside, a prefix segPtr, the
   * row map padded to tm entries per segment, and the CV variant (row pointers per segment +
   * interleaved ((float)col, val), row-major). */
  int S = o->nsegs;
  o->segPtr = calloc((size_t)S + 1, sizeof(uint32_t));
  o->segNzRCIdx = malloc(sizeof(uint32_t) * 2 * (size_t)(nnz ? nnz : 1));
  o->segVals = malloc(sizeof(float) * (size_t)(nnz ? nnz : 1));
  o->segVoMapPad = malloc(sizeof(uint32_t) * (size_t)(S ? S : 1) * (size_t)tm);
  o->seg_rowPtr = malloc(sizeof(int) * (size_t)(S ? S : 1) * ((size_t)tm + 1));
  o->segNzCV = malloc(sizeof(float) * 2 * (size_t)(nnz ? nnz : 1));
  int *cursor = malloc(sizeof(int) * (size_t)tm);
  size_t w = 0;
  for (int s = 0; s < S; ++s) {
    int r0 = (int)a.pillar[s], rows = (int)(a.pillar[s + 1] - a.pillar[s]);
    o->segPtr[s] = a.rowPtr[r0];
    for (int i = 0; i < tm; ++i) {
      o->segVoMapPad[(size_t)s * tm + i] = i < rows ? a.voMap[r0 + i] : 0x7fffffffu;
      o->seg_rowPtr[(size_t)s * (tm + 1) + i] = (int)(a.rowPtr[r0 + (i < rows ? i : rows)] - a.rowPtr[r0]);
      cursor[i] = i < rows ? (int)a.rowPtr[r0 + i] : 0;
    }
    o->seg_rowPtr[(size_t)s * (tm + 1) + tm] = (int)(a.rowPtr[r0 + rows] - a.rowPtr[r0]);
    for (uint32_t e = a.rowPtr[r0]; e < a.rowPtr[r0 + rows]; ++e) {
      o->segNzCV[2 * (size_t)e] = (float)a.colIdx[e];
      o->segNzCV[2 * (size_t)e + 1] = a.vals[e];
    }
    int left = (int)(a.rowPtr[r0 + rows] - a.rowPtr[r0]);
    while (left > 0) { /* column-major sweep of the segment */
      uint32_t j = 0xffffffffu;
      for (int i = 0; i < rows; ++i) if (cursor[i] < (int)a.rowPtr[r0 + i + 1] && a.colIdx[cursor[i]] < j) j = a.colIdx[cursor[i]];
      for (int i = 0; i < rows; ++i)
        if (cursor[i] < (int)a.rowPtr[r0 + i + 1] && a.colIdx[cursor[i]] == j) {
          o->segNzRCIdx[2 * w] = (uint32_t)i; o->segNzRCIdx[2 * w + 1] = j; o->segVals[w] = a.vals[cursor[i]];
          w++; cursor[i]++; left--;
        }
    }
  }
  o->segPtr[S] = (uint32_t)nnz;
  free(cursor);
  return 0;
}
void orc_seg_free(orc_seg *s) {
  free(s->alpha_rowPtr); free(s->alpha_colIdx); free(s->alpha_vals); free(s->pillar_rowPtr); free(s->segVoMap);
  free(s->segs_per_panel); free(s->segPtr); free(s->segNzRCIdx); free(s->segVals); free(s->segVoMapPad);
  free(s->seg_rowPtr); free(s->segNzCV);
  memset(s, 0, sizeof(*s));
}

/* ------------------------------------------------------------------------------------------ */
/* F4: contiguous SM buckets + shared tail (mat.cu:1118-1162, row_based_split=true)             */
/* ------------------------------------------------------------------------------------------ */
void orc_sm_buckets(int n_sm, int nsegs, int npanels, const int *segs_per_panel, int *next_seg,
                    int *grouped_tailSeg) {
  int segload = nsegs / n_sm, head = 0, tail = 0, panel = 0;
  for (int i = 0; i < n_sm; ++i) {
    next_seg[i] = head;
    if (panel < npanels) {
      int cur = segs_per_panel[panel];
      tail = head + cur;
      while (++panel < npanels) {
        if (segs_per_panel[panel] + cur > segload) break;
        cur += segs_per_panel[panel];
        tail += segs_per_panel[panel];
      }
    }
    grouped_tailSeg[i] = nsegs < tail ? nsegs : tail;
    head = nsegs < tail ? nsegs : tail;
  }
  next_seg[n_sm] = head;
  grouped_tailSeg[n_sm] = nsegs;
}

/* ------------------------------------------------------------------------------------------ */
/* F5: csr2_DiagTiling (mat.cu:680-903)                                                         */
/* ------------------------------------------------------------------------------------------ */
static int has_col(const uint32_t *rowptr, const uint32_t *col, int r, uint32_t c) { /* position or -1 */
  uint32_t lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (col[mid] < c) lo = mid + 1; else hi = mid; }
  return (lo < rowptr[r + 1] && col[lo] == c) ? (int)lo : -1;
}

int orc_diag_tiling(int m, const uint32_t *rowptr, const uint32_t *col, const float *val,
                    const int32_t *vo_mp, int tm, int n_sm, orc_pillar *o) {
  memset(o, 0, sizeof(*o));
  const int warps_per_sm = 64;
  const float alpha = 0.3f;
  const int nnz = (int)rowptr[m];
  for (int r = 0; r < m; ++r) if (rowptr[r] == rowptr[r + 1]) return -1; /* assert :1207 in round 3 */
  const int nnz_diagonal_tiles = (int)(alpha * rowptr[m]);        /* :694 (float arithmetic) */
  const int partitions_node = warps_per_sm * n_sm;
  int nnz_p_diagonal_tile = nnz_diagonal_tiles / partitions_node;
  if (nnz_p_diagonal_tile < 32) nnz_p_diagonal_tile = 32;
  int *tile_width = calloc((size_t)partitions_node, sizeof(int));
  unsigned char *claimed = calloc((size_t)(nnz ? nnz : 1), 1);    /* duplicate_sparse_mat */
  /* alpha_columns_per_sm as a per-SM stamp: sm_of_col[c] holds every SM that listed c -> a column can
   * be listed by more than one SM only at block borders; keep a small open list per column */
  int *colsm_head = malloc(sizeof(int) * (size_t)m);
  for (int i = 0; i < m; ++i) colsm_head[i] = -1;
  typedef struct { int sm, next; } node_t;
  node_t *nodes = malloc(sizeof(node_t) * ((size_t)2 * m + 16));
  size_t nnodes = 0, capnodes = (size_t)2 * m + 16;
#define COL_IN_SM(c, s, res) do { res = 0; for (int q_ = colsm_head[c]; q_ >= 0; q_ = nodes[q_].next) if (nodes[q_].sm == (s)) { res = 1; break; } } while (0)
#define COL_ADD_SM(c, s) do { int f_; COL_IN_SM(c, s, f_); if (!f_) { if (nnodes == capnodes) { capnodes *= 2; nodes = realloc(nodes, sizeof(node_t) * capnodes); } nodes[nnodes].sm = (s); nodes[nnodes].next = colsm_head[c]; colsm_head[c] = (int)nnodes++; } } while (0)
  /* round 1 :706-759 */
  int mat_r_start = 0, warps_with_weights = 0;
  const int thr = (int)(0.85 * nnz_p_diagonal_tile);
  for (int i = 0; i < partitions_node; ++i) {
    mat_r_start += i ? tile_width[i - 1] : 0;
    int cnt = 0, j = mat_r_start, sm = i / warps_per_sm;
    while (j < m && cnt <= thr) {
      /* row panel (:718-727).  The reference walks colIdx from rowPtr[j] until it meets column j;
       * in a row WITHOUT its diagonal the walk runs on into the following rows' entries (it never
       * checks rowPtr[j+1]) and claims (j, colIdx[kk]) for whatever it meets.  Restated as is; only
       * running off the end of the array (undefined there) is refused. */
      for (uint32_t kk = rowptr[j];; ++kk) {
        if (kk >= (uint32_t)nnz) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -2; }
        if (!(col[kk] <= (uint32_t)j)) break;
        if ((int)col[kk] >= mat_r_start) {
          cnt++; COL_ADD_SM(col[kk], sm);
          int e_ = kk < rowptr[j + 1] ? (int)kk : has_col(rowptr, col, j, col[kk]);
          if (e_ >= 0) claimed[e_] = 1;
        }
        if (col[kk] == (uint32_t)j) break;
      }
      for (int kk = mat_r_start; kk < j; ++kk) { /* column panel */
        int l = has_col(rowptr, col, kk, (uint32_t)j);
        if (l >= 0) { cnt++; claimed[l] = 1; COL_ADD_SM(j, sm); }
      }
      j++;
    }
    warps_with_weights += cnt > 0;
    if (!(j >= m || cnt)) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -3; }
    tile_width[i] = j - mat_r_start;
  }
  long verify_m = 0;
  for (int i = 0; i < partitions_node; ++i) verify_m += tile_width[i];
  if (verify_m != m) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -4; } /* assert :759 */
  /* round 2 :771-835 */
  alpha_t a;
  memset(&a, 0, sizeof(a));
  a.colIdx = malloc(sizeof(uint32_t) * (size_t)(nnz ? nnz : 1));
  a.vals = malloc(sizeof(float) * (size_t)(nnz ? nnz : 1));
  push_u32(&a.pillar, &a.npil, &a.cappil, 0);
  size_t npi = 0;
  o->alpha_pillarIdx = malloc(sizeof(uint32_t) * ((size_t)n_sm + 2));
  int nnz_rowPtr = 0, row_end = 0, col_start = 0, col_end = 0, row_start = 0;
  for (int i = 0; i < warps_with_weights; ++i) {
    row_start = row_end;
    row_end += tile_width[i];
    if (i % warps_per_sm == 0) {
      col_start = col_end;
      for (int idx = 0; idx < warps_per_sm && (i + idx) < warps_with_weights; ++idx) col_end += tile_width[i + idx];
    }
    int sm = i / warps_per_sm, nnz_warp = 0;
    for (int j = row_start; j < row_end; ++j) {
      int entries = 0;
      push_u32(&a.rowPtr, &a.nrp, &a.caprp, (uint32_t)nnz_rowPtr);
      for (uint32_t kk = rowptr[j]; kk < rowptr[j + 1]; ++kk) {
        int l = (int)col[kk];
        if (l < col_start) continue;
        if (l >= col_end) break;
        int insm; COL_IN_SM(l, sm, insm);
        if (claimed[kk] || insm) {
          claimed[kk] = 1;
          if (!insm) { /* reference asserts alpha_columns_per_sm contains l (:818) */
            free(tile_width); free(claimed); free(colsm_head); free(nodes); return -5;
          }
          a.colIdx[a.nnz] = (uint32_t)l; a.vals[a.nnz] = val[kk]; a.nnz++;
          entries++; nnz_warp++; nnz_rowPtr++;
        }
      }
      uint32_t v = (uint32_t)vo_mp[j];
      push_u32(&a.voMap, &a.nvm, &a.capvm, entries < (int)(rowptr[j + 1] - rowptr[j]) ? (v | 0x80000000u) : v);
    }
    if (!nnz_warp) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -6; } /* assert :837 */
    push_u32(&a.pillar, &a.npil, &a.cappil, a.pillar[a.npil - 1] + (uint32_t)tile_width[i]);
    if (i % warps_per_sm == 0) o->alpha_pillarIdx[npi++] = (uint32_t)i;
  }
  while (npi <= (size_t)n_sm) o->alpha_pillarIdx[npi++] = (uint32_t)warps_with_weights;
  if ((int)a.nrp != m) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -7; } /* assert :853 */
  o->empty_wp_p = (1 - (float)warps_with_weights / partitions_node) * 100;
  o->band_nz_p = (float)a.nnz / rowptr[m] * 100;
  if (partitions_node - warps_with_weights == 0) { free(tile_width); free(claimed); free(colsm_head); free(nodes); return -8; } /* /0 at :862 */
  /* round 3 :871-878: leftovers through csr2seg_Cmajor into the shared balance queue */
  push_u32(&a.rowPtr, &a.nrp, &a.caprp, (uint32_t)nnz_rowPtr);
  int tiles_in_total = warps_with_weights;
  int tileRows = (m + tm - 1) / tm;
  for (int p = 0; p < tileRows; ++p)
    tiles_in_total += seg_panel(m, rowptr, col, val, vo_mp, tm, 128, p, claimed, &a, &nnz_rowPtr);
  o->alpha_pillarIdx[npi++] = (uint32_t)tiles_in_total;
  o->m = m; o->nnz = nnz; o->n_sm = n_sm; o->n_segs = tiles_in_total; o->warps_with_weights = warps_with_weights;
  o->rows_total = (int)a.nrp - 1;
  o->alpha_rowPtr = a.rowPtr; o->alpha_colIdx = a.colIdx; o->alpha_vals = a.vals;
  o->alpha_pillar_rowPtr = a.pillar; o->segVoMap = a.voMap;
  free(tile_width); free(claimed); free(colsm_head); free(nodes);
  return (nnz_rowPtr == nnz && (int)a.nnz == nnz) ? 0 : -9;
}
void orc_pillar_free(orc_pillar *p) {
  free(p->alpha_rowPtr); free(p->alpha_colIdx); free(p->alpha_vals); free(p->alpha_pillar_rowPtr);
  free(p->alpha_pillarIdx); free(p->segVoMap);
  memset(p, 0, sizeof(*p));
}

/* alpha_w_atomic_spmm_v36 (flex.cu:4010-4124), serial: every row of the alpha CSR adds its partial
 * dot products into C[segVoMap & 0x7fffffff] (the MSB only selects atomicAdd vs store there).
 * The same routine evaluates the F2 "alpha" layout of orc_seg. */
void orc_alpha_spmm(int rows_total, const uint32_t *alpha_rowPtr, const uint32_t *alpha_colIdx,
                    const float *alpha_vals, const uint32_t *segVoMap, int64_t m, const float *shadowB,
                    int k, float *C) {
  for (int64_t i = 0; i < m * k; ++i) C[i] = 0.f;
  for (int r = 0; r < rows_total; ++r) {
    float *o = C + (int64_t)(segVoMap[r] & 0x7fffffffu) * k;
    for (uint32_t e = alpha_rowPtr[r]; e < alpha_rowPtr[r + 1]; ++e) {
      const float *b = shadowB + (int64_t)alpha_colIdx[e] * k;
      for (int j = 0; j < k; ++j) o[j] = fmaf(alpha_vals[e], b[j], o[j]);
    }
  }
}
