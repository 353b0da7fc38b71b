/*
 * fx_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's algorithms on the SpMM hot path
 * of guohaoqiang/Flex.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (flex_b200/csrc, libflexb200.so) never links, loads or calls it.
 *
 * Every function cites the reference file:line it restates.  Pinning status is
 * listed in oracle/README.md and DESIGN.md section 3.
 */
#ifndef FX_ORACLE_H
#define FX_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- L1: CSV loader (DataLoader.cu:9-124, aspt/sspmm_128.cu:102-160) ---- */
typedef struct {
  int64_t n, nnz;
  uint32_t *rowptr; /* n+1 */
  uint32_t *col;    /* nnz */
  float *val;       /* nnz */
  int64_t uni_nb;   /* rows with exactly one nz (DataLoader.cu:24-27) */
  int c;            /* class count by file name (DataLoader.cu:62-84) */
} orc_csr;
int orc_csv_load(const char *path, orc_csr *out);
void orc_csr_free(orc_csr *m);

typedef struct { /* DataLoader.cu:86-115 */
  int is_directed;
  int64_t n_edges_one_way, n_edges_asymmetric;
  int n_nodes_z_out, n_nodes_z_in, n_nodes_z_deg;
} orc_census_t;
int orc_census(const orc_csr *m, orc_census_t *out);

/* ---- L2: random dense B (DataLoader.cu:198-209 / aspt/sspmm_128.cu:1148-1154) */
void orc_rand_B_flex(int64_t n, int k, float *out); /* srand(1); 2*rand()/RAND_MAX-1 */
void orc_rand_B_aspt(int64_t n, int k, float *out); /* srand(1); (rand()%1048576)/1048576 */

/* ---- SpMM (aspt/sspmm_128.cu:1415-1422): fp32, CSR order, row-major B/C ---- */
void orc_spmm_ref(int64_t n, const uint32_t *rowptr, const uint32_t *col, const float *val,
                  const float *B, int k, float *C);
/* same per-row summation order, rows in parallel (OpenMP); returns threads used */
int orc_spmm_omp(int64_t n, const uint32_t *rowptr, const uint32_t *col, const float *val,
                 const float *B, int k, float *C, int threads);
/* fp64 accumulate; Cabs (optional) = sum |a*b| per element, for tolerance analysis */
void orc_spmm_f64(int64_t n, const uint32_t *rowptr, const uint32_t *col, const float *val,
                  const float *B, int k, double *C, double *Cabs);
/* reference order on a subset of rows (sampled checks at full size) */
void orc_spmm_rows(const int64_t *rows, int64_t nrows, const uint32_t *rowptr, const uint32_t *col,
                   const float *val, const float *B, int k, float *C);
int orc_num_threads(void);

/* ---- V1: validators (flex.cu:4155-4213, aspt/sspmm_128.cu:1425-1446) ---- */
typedef struct {
  int64_t flex_count;  /* resCheck: err > FLT_EPSILON*row_nnz*4 */
  int64_t aspt_count;  /* rel diff > 1e-2 */
  int64_t tight_count; /* ours: |d| > 1e-5*max(1, ||gold[row,:]||_inf) */
  int64_t gold_zeros;
  double max_err;      /* resCheck metric */
  double max_tight;    /* max |d|/max(1, ||gold[row,:]||_inf) */
  double aspt_pct;     /* aspt_count / (n*k) * 100 */
} orc_errs;
void orc_check(const float *gold, const float *res, int64_t n, int k, const uint32_t *rowptr,
               orc_errs *out);

/* ---- L3: permutation apply (DataLoader.cu:244-321, 658-857 ctor bodies) ----
 * rank[old] = new.  Output: vo_mp[new]=old, permuted CSR with columns ascending. */
void orc_perm_apply(int64_t n, const uint32_t *rowptr, const uint32_t *col, const float *val,
                    const uint64_t *rank, int32_t *vo_mp, uint32_t *rowptr_out, uint32_t *col_out,
                    float *val_out);
/* P1 (flex.cu:276-289) */
void orc_permute_rows(int64_t n, int k, const int32_t *vo_mp, const float *B, float *shadowB);

/* ---- A1: ASpT tile builder, canonical (aspt/sspmm_128.cu:831-1087,1207-1333) ---- */
typedef struct {
  int n, nr, npanel, ne, BH, BW, num_dense;
  int any_flag;       /* d_flag[0] != 0 (:1221-1224) */
  int *mcsr_chk;      /* npanel */
  int *mcsr_cnt;      /* npanel+1 */
  int *mcsr_e;        /* BH*(num_dense+npanel)+1 */
  int *mcsr_list;     /* BW*num_dense */
  int *baddr, *saddr; /* num_dense */
  int *key2;          /* ne: tile id per ORIGINAL nz (30000 = sparse) */
  int *perm;          /* ne: permuted position -> original nz index ("val" array) */
  int *csr_e;         /* ne permuted columns */
  float *csr_ev;      /* ne permuted values */
  int64_t S1, S2;     /* sum len, sum len^2 over sparse groups (exact) */
  double avg, vari;   /* :1226/:1300, :1050-1073 + :1315 */
  int special_p;      /* only if vari>=200 */
  int *special, *special2;
  int regime;         /* 0 ssparse, 1 sparse_v2+dense, 2 v2l+v2h+dense (:1355-1397) */
} orc_aspt;
/* forced_cnt/forced_list: NULL for the canonical slot assignment; otherwise the mcsr_cnt and
 * mcsr_list of a reference run (its atomic-order choice), reproduced from there on. */
int orc_aspt_build(int n, const uint32_t *rowptr, const uint32_t *col, const float *val, int BW,
                   const int *forced_cnt, const int *forced_list, orc_aspt *out);
void orc_aspt_free(orc_aspt *t);
/* SpMM evaluated THROUGH the tile structure (dense groups then sparse group per row, one
 * fmaf per nz): the summation order of the GPU panel kernel. C is nr*k. */
void orc_aspt_spmm(const orc_aspt *t, const float *B, int k, float *C);

#ifdef __cplusplus
}
#endif
#endif
