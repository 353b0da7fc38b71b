/*
 * flexb200.h -- C ABI of libflexb200.so, the B200-native (sm_100a) SpMM engine that
 * stands behind guohaoqiang/Flex's driver surface.
 *
 * Plain pointers and sizes only: no C++/torch types.  Every entry point names the
 * reference interface (file:line in guohaoqiang/Flex) it replaces.  Pointers named
 * *_dev are CUDA device addresses on the current device; `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).
 *
 * Error model (reference: CUDA_CHECK throws std::runtime_error, common.h:53-60; data
 * violations are asserts): every call returns FX_OK or a negative fx_status and
 * records a message retrievable with fx_last_error().  Nothing falls back to the
 * CPU: with no usable sm_100 device every compute call returns FX_ERR_CUDA.
 */
#ifndef FLEXB200_H
#define FLEXB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FX_OK = 0,
  FX_ERR_IO = -1,       /* file missing / malformed CSV */
  FX_ERR_ARG = -2,      /* bad argument */
  FX_ERR_CUDA = -3,     /* CUDA runtime error or no device */
  FX_ERR_FORMAT = -4,   /* matrix violates a builder precondition (reference asserts) */
  FX_ERR_NOMEM = -5,
  FX_ERR_UNSUPPORTED = -6
} fx_status;

typedef enum { /* DataLoader::vertex_order_abbr, DataLoader.cu:14,326,457,660,725,791 */
  FX_ORDER_OVO = 0, /* original vertex order */
  FX_ORDER_DEG = 1, /* DataLoaderDeg    DataLoader.cu:658-721 */
  FX_ORDER_RCM = 2, /* DataLoaderRcm    DataLoader.cu:723-787 */
  FX_ORDER_GOR = 3, /* DataLoaderGorder DataLoader.cu:789-857 */
  FX_ORDER_DFS = 4, /* DataLoaderDFS    DataLoader.cu:324-453 */
  FX_ORDER_RBT = 5  /* DataLoaderRabbit DataLoader.cu:455-655 */
} fx_order;

typedef enum {
  FX_FMT_CSR = 0,    /* raw CSR (run_ge_spmm path, flex.cu:4285; ASpT "ssparse" regime) */
  FX_FMT_ASPT = 1,   /* ASpT dense/sparse tiles  aspt/sspmm_128.cu:831-1087,1207-1333 */
  FX_FMT_TILE = 2,   /* Flex tile format         mat.cu:1345-1518 */
  FX_FMT_SEG = 3,    /* Flex tile-segment format mat.cu:1192-1269 + SM buckets :1097-1162 */
  FX_FMT_PILLAR = 4, /* Flex diagonal tiling     mat.cu:680-942 */
  FX_FMT_TCW = 5     /* B200-native: per-panel tensor-core windows (the panel's most shared columns, multiplied
                        by tcgen05) + ASpT over the remaining nz.  Serves the same purpose as the reference's
                        on-chip tiles (mat.cu:1345-1518, aspt dense tiles); no reference counterpart layout. */
} fx_format;

typedef struct fx_matrix fx_matrix; /* replaces class DataLoader (DataLoader.cuh:21-112) */
typedef struct fx_tiles fx_tiles;   /* replaces class Mat / Mat_POD (mat.cuh:18-229) */

typedef struct { /* DataLoader public fields, DataLoader.cuh:57-71 */
  int64_t m, n, nnz, dim, c, uni_nb;
  int32_t is_directed, n_nodes_z_out, n_nodes_z_in, n_nodes_z_deg;
  int64_t n_edges_one_way, n_edges_asymmetric;
  int32_t order;          /* fx_order */
  char graph_name[64];    /* basename without extension, DataLoader.cu:11-12 */
  char order_abbr[4];     /* "OVO","DEG","RCM","GOR","DFS","RBT" */
} fx_matrix_info;

typedef struct {
  int32_t format;     /* fx_format */
  int32_t tm, tn;     /* Flex tile height/width (tileConfs, flex.cu:4146-4152); ASpT: ignored */
  int32_t bw;         /* ASpT tile width: 0 = reference rule (128 if k>=64 else 256) */
  int32_t nnz_limit;  /* NNZ_LIMIT (mat.cuh:16), 0 = 128 */
  int32_t n_sm;       /* SM count for F4/F5 bucketing, 0 = device value */
  int32_t row_begin, row_end; /* build only rows [row_begin,row_end) (row-panel shard); 0,0 = all */
  int32_t cmajor;     /* FX_FMT_TILE: 0 = csr2flex_Rmajor, 1 = csr2flex_Cmajor (COL_MAJ_TILE, DataLoader.cuh:18) */
  /* FX_FMT_TCW plan (flex_b200/csrc/fx_tcw_build.cu has the rule); 0 = default; negative tc_min_* = "no minimum" */
  int32_t tc_threshold;  /* a column is a candidate with >= this many nz in the panel (4) */
  int32_t tc_width;      /* at most this many window columns per panel, multiple of 32 (256) */
  int32_t tc_min_gain;   /* a panel keeps its window only if its net gain, in B-row fetches, reaches this (1024) */
  int32_t tc_chunk_cost; /* what one 32-column chunk costs, in B-row fetches (224) */
  int32_t tc_min_total;  /* the matrix keeps its windows only if the summed net gain reaches this (1000000) */
  int32_t reserved[2];
} fx_build_opts;

typedef struct { /* ASpT metadata export for bit-exact checks; pointers are host copies
                    owned by the fx_tiles handle, valid until fx_tiles_free */
  int32_t n, nr, npanel, ne, BH, BW, num_dense, any_flag, regime, special_p;
  int64_t S1, S2;
  double avg, vari;
  const int32_t *mcsr_chk, *mcsr_cnt, *mcsr_e, *mcsr_list, *baddr, *saddr;
  const int32_t *perm, *csr_e, *special, *special2;
  const float *csr_ev;
} fx_aspt_arrays;

typedef struct { /* Flex tile format (mat.cu:1345-1518; Mat_POD fields mat.cuh:27-34), host copies */
  int32_t m, tm, tn, ntiles, npanels, nnz, cmajor;
  const uint32_t *tileRowPtr; /* npanels+1 */
  const uint32_t *tileNnz;    /* ntiles+1 */
  const int32_t *nnzTile, *bitMap; /* ntiles */
  const uint32_t *tileColIdx; /* ntiles */
  const int32_t *rcOffset;    /* nnz: (rowInPanel<<16)|(col-tileColIdx) */
  const float *newVals;       /* nnz */
} fx_tile_arrays;

typedef struct { /* row-panel segmentation (mat.cu:1192-1269) + SM buckets (mat.cu:1118-1162), host copies */
  int32_t m, tm, nnz, nsegs, rows_total, npanels, n_sm;
  const uint32_t *alpha_rowPtr, *alpha_colIdx, *alpha_pillar_rowPtr, *segVoMap;
  const float *alpha_vals;
  const int32_t *segs_per_panel;
  const uint32_t *segPtr, *segNzRCIdx, *segVoMapPad; /* tile-segment arrays of kernels v10-v35 */
  const float *segVals, *segNzCV;
  const int32_t *seg_rowPtr;
  const int32_t *next_seg, *grouped_tailSeg; /* n_sm+1 each */
} fx_seg_arrays;

typedef struct { /* diagonal tiling / pillar format (mat.cu:680-903; Mat_POD mat.cuh:49-55), host copies */
  int32_t m, nnz, n_sm, n_segs, rows_total, warps_with_weights;
  const uint32_t *alpha_rowPtr, *alpha_colIdx, *alpha_pillar_rowPtr, *alpha_pillarIdx, *segVoMap;
  const float *alpha_vals;
  float empty_wp_p, band_nz_p;
  int32_t round1_on_gpu; /* 1: the diagonal blocks (round 1) were grown on the GPU too; 0: on the host (a row without its
                            diagonal, or an nz near the diagonal without its transpose) */
} fx_pillar_arrays;

typedef struct { /* tensor-window format (FX_FMT_TCW), host copies owned by the handle */
  int32_t n, nr, npanel, W, T, min_gain, ntc, dropped, chunk_cost, reserved0;
  int64_t win_nnz, rest_nnz, min_total, net_gain;
  const int32_t *tc_cols;     /* npanel*W : ascending column list of each panel, -1 padded (W a multiple of 32) */
  const int32_t *tc_ncol;     /* npanel */
  const int32_t *win_cptr;    /* npanel*(W/32)+1 : window nz of each (panel, 32-column chunk of its list) */
  const uint16_t *win_code;   /* win_nnz : word of (row r in panel, kk = list position & 31) in the chunk's K-major 128x32 operand
                                 tile: (r>>3)<<8 | (kk>>2)<<5 | (r&7)<<2 | (kk&3); (row, position) order in a chunk */
  const float *win_val;       /* win_nnz */
  const uint32_t *rest_rowptr; /* n+1 */
  const uint32_t *rest_col;    /* rest_nnz */
  const float *rest_val;       /* rest_nnz */
} fx_tcw_arrays;

typedef struct { /* what run()/process() print: flex.cu:5134-5631, aspt/sspmm_128.cu:1406-1446 */
  float tPre_ms, tElap_ms;
  double gflops;        /* 2*nnz*k / tElap (aspt/sspmm_128.cu:1406) */
  double tpre_over_telap;
  int64_t errs_flex;    /* resCheck count, flex.cu:4155-4213 */
  int64_t errs_tight;   /* |d| > 1e-5*max(1, ||gold[row,:]||_inf) */
  double errs_aspt_pct; /* aspt/sspmm_128.cu:1425-1446 */
  double max_err;
} fx_report;

const char *fx_last_error(void);
int fx_version(void);
/* number of kernels launched by this library in this process (all streams) */
int64_t fx_launch_count(void);
int fx_device_sm_count(int *n_sm);

/* ---- L0: CSR load + device upload ------------------------------------------------ */
/* DataLoader::DataLoader(path,k) DataLoader.cu:9-124 + cuda_alloc_cpy :167-218 (CSR part) */
int fx_csr_load(const char *path, int k, fx_matrix **out);
/* same from host arrays (no reference counterpart; the arrays DataLoader would have parsed) */
int fx_csr_from_arrays(int64_t n, int64_t nnz, const uint32_t *rowptr, const uint32_t *col,
                       const float *val, int k, const char *name, fx_matrix **out);
/* CSR already resident in HBM (arrays are copied device-to-device).  Same preconditions as the host loaders, checked by a
 * device kernel: rowPtr monotone over [0,nnz], every column < n, columns strictly ascending within a row (unique and
 * sorted, DataLoader.cu:97,272) -- FX_ERR_FORMAT otherwise.  vo_mp is the identity. */
int fx_csr_from_device(int64_t n, int64_t nnz, const uint32_t *rowptr_dev, const uint32_t *col_dev,
                       const float *val_dev, int k, const char *name, fx_matrix **out);
/* N2: Matrix Market coordinate file -> CSR, as data/SuiteSparse/mtx2csr.cc (mmio_allinone :57-222 +
 * the 6-significant-digit text round trip of writeCSR2csv :224-246) followed by DataLoader would
 * load it; rows sorted by column */
int fx_mtx_load(const char *path, int k, fx_matrix **out);
/* the 3-line CSV writer of mtx2csr.cc:224-246 */
int fx_csr_write_csv(const fx_matrix *m, const char *path);
/* binary CSR cache (no reference counterpart: replaces re-parsing gigabytes of ASCII) */
int fx_csr_save_bin(const fx_matrix *m, const char *path);
int fx_csr_load_bin(const char *path, int k /* 0 = stored k */, fx_matrix **out);
int fx_matrix_get_info(const fx_matrix *m, fx_matrix_info *info);
/* host views of rowPtr/col/vals (DataLoader.cuh:32-34); NULL if created from device arrays */
int fx_matrix_host_csr(const fx_matrix *m, const uint32_t **rowptr, const uint32_t **col,
                       const float **val);
int fx_matrix_device_csr(const fx_matrix *m, const uint32_t **rowptr_dev, const uint32_t **col_dev,
                         const float **val_dev);
void fx_matrix_free(fx_matrix *m); /* DataLoader::freeAll DataLoader.cuh:100-111 */
/* B = 2*rand()/RAND_MAX-1 from the glibc stream seeded 1 (DataLoader.cu:198-209); host buffer n*k */
int fx_rand_B(int64_t n, int k, float *B_host);

/* ---- L1: reordering hooks ---------------------------------------------------------- */
/* DataLoaderDeg/Rcm/Gorder(const DataLoader&) DataLoader.cuh:128-145: new matrix with the
 * permuted CSR (columns ascending per row) and vo_mp[new]=old. */
int fx_reorder(const fx_matrix *m, int order /* fx_order */, fx_matrix **out);
/* apply an explicit rank[old]=new (DataLoader::perm_apply DataLoader.cu:244-321) */
int fx_reorder_with_rank(const fx_matrix *m, const uint64_t *rank, int order_tag, fx_matrix **out);
int fx_permutation(const fx_matrix *m, const int32_t **vo_mp /* host, n entries */);
/* shadow_b[r,:] = B[vo_mp[r],:]  (flexspmm_v9_permuteX flex.cu:276-289) */
int fx_permute_rows(const fx_matrix *m, const float *B_dev, float *shadowB_dev, int k, void *stream);
/* C_out[vo_mp[r],:] = C[r,:] (VO_RECOVER off path, flex.cu:994 v9 writes C[voMp[row]]) */
int fx_unpermute_rows(const fx_matrix *m, const float *C_dev, float *C_out_dev, int k, void *stream);

/* ---- L2: tile-format build on the GPU ---------------------------------------------- */
/* Mat::Mat + csr2tile/csr2_DiagTiling + transfer + launch_prep (mat.cuh:74,82-83,176-182);
 * ASpT: process() pre-process section aspt/sspmm_128.cu:1207-1333.  tPre_ms = GPU time of the
 * build with the arena already allocated (events around the build kernels and the small
 * device->host reads they need). */
int fx_build(const fx_matrix *m, const fx_build_opts *opts, fx_tiles **out, float *tPre_ms);
/* repeat the build into the same arena (timing loops) */
int fx_rebuild(fx_tiles *t, float *tPre_ms);
int fx_tiles_export_aspt(fx_tiles *t, fx_aspt_arrays *out);
int fx_tiles_export_tile(fx_tiles *t, fx_tile_arrays *out);
int fx_tiles_export_seg(fx_tiles *t, fx_seg_arrays *out);
int fx_tiles_export_pillar(fx_tiles *t, fx_pillar_arrays *out);
int fx_tiles_export_tcw(fx_tiles *t, fx_tcw_arrays *out);
/* the scalars of fx_tcw_arrays without copying any array: out = {npanel, ntc, win_nnz, rest_nnz, listed columns,
 * net_gain, W, T} */
int fx_tiles_tcw_info(const fx_tiles *t, int64_t out[8]);
void fx_tiles_free(fx_tiles *t); /* Mat::freeMatGPU* mat.cuh:184-220 */

/* ---- L3: SpMM ---------------------------------------------------------------------- */
/* C[m x k] = A * B[n x k], row-major fp32, C fully overwritten (the reference pre-zeroes C and
 * accumulates with atomics: mat.cu:32-41, aspt/sspmm_128.cu:1147).  Asynchronous on `stream`
 * unless tElap_ms != NULL, in which case the call brackets the kernels with events, waits and
 * returns their elapsed time (flex.cu:5051-5068, aspt/sspmm_128.cu:1369-1380).
 * - One SpMM in flight per handle: the 512-chunk partial sums, the window products and the timing events are scratch
 *   of the fx_tiles handle (as the reference's device arrays are globals, aspt/sspmm_128.cu:76-90).  Two calls on
 *   different streams need two handles (fx_build twice) or an event between them.
 * - k may differ from the k the matrix was loaded with; a call with a LARGER k (or k % 4 != 0) on an ASpT /
 *   tensor-window handle runs the raw-CSR kernel, because that scratch is sized for the build's k.
 * - FX_FMT_TCW needs finite B: the dense window contraction multiplies the zeros of the A tile by every listed row of
 *   B, so an Inf/NaN in such a row reaches all 128 rows of the panel (0 * Inf), where the reference's sparse semantics
 *   (and every other format here) touch only the rows holding a nz in that column; tests/test_gpu_robustness.py. */
int fx_spmm(const fx_tiles *t, const float *B_dev, float *C_dev, int k, void *stream,
            float *tElap_ms);
/* The same SpMM with events between its kernels (the reference times its kernel section as a whole,
 * aspt/sspmm_128.cu:1369-1380; this is what the per-kernel NPerf counters of flex.cu:4583-4656 are for):
 * ms[0] = tensor-window kernel, ms[1] = 512-chunk kernel of long rows, ms[2] = row kernel, ms[3] = the step.
 * ASpT / tensor-window handles only; waits for the step to finish. */
int fx_spmm_kernel_times(const fx_tiles *t, const float *B_dev, float *C_dev, int k, void *stream, float ms[4]);
/* Same with HOST buffers: copies B in, runs, copies C out (DataLoader.cu:216 + flex.cu:5690).
 * ASpT / tensor-window handles with k % 32 == 0, k >= 64: pipelined over two column chunks (FLEX_HOST_CHUNKS) and, inside a
 * chunk, over up to four ranges of row panels (FLEX_HOST_GROUPS) whose rows of C are copied out as soon as they are
 * multiplied; tElap_ms is then the sum of the chunks' kernel sections.  Pass pinned buffers for full PCIe rate. */
int fx_spmm_host(const fx_tiles *t, const float *B_host, float *C_host, int k, float *total_ms,
                 float *tElap_ms);

/* ---- L3a: AXW, the GCN layer product (cusp.cu:3-208 run1 / run2, main.cu:22-79; dead code in the reference) ----------
 * X [n x k], W [k x c], C [rows x c], all row-major fp32 on the device.  order 0: C = A*(X*W) (run1: the dense factor
 * first, then the SpMM at width c); order 1: C = (A*X)*W (run2).  The dense factor runs on tcgen05 with the 3xTF32 split
 * (fp32 accuracy); the intermediate lives in a scratch of the handle.  k and c multiples of 4; with ASpT / tensor-window
 * tiles c (order 0) resp. k (order 1) must not exceed the build's k, else the SpMM takes the CSR kernel.  gemm_ms / spmm_ms
 * (optional) receive the two device times and make the call wait. */
int fx_axw(const fx_tiles *t, const float *X_dev, const float *W_dev, float *C_dev, int k, int c, int order, void *stream,
           float *gemm_ms, float *spmm_ms);

/* ---- L3b: row-panel sharded SpMM from host buffers (the reference is single-GPU: no counterpart) ------
 * One process per GPU.  Every rank builds the tiles of ITS row-panel shard (fx_build with row_begin/row_end; B is
 * replicated, no collective is on the multiply itself).  What a G-rank job must not do is push all of B through the
 * host's PCIe root G times: rank r uploads rows [r*ceil(n/G), (r+1)*ceil(n/G)) of B only, ncclAllGather assembles B on
 * every GPU over NVLink -- the one real exchange step of the path --, the rank multiplies its shard and copies its own
 * rows of C back, pipelined over column chunks like fx_spmm_host.
 * fx_comm wraps an ncclComm_t created from a unique id that the host program distributes (MPI, torch.distributed,
 * a file ...): rank 0 calls fx_comm_unique_id and hands the 128 bytes to every rank, every rank calls fx_comm_init
 * on its device.  NCCL is resolved at run time (the libnccl.so.2 already loaded in the process, else the system's);
 * without it these calls return FX_ERR_UNSUPPORTED and everything else works. */
typedef struct fx_comm fx_comm;
/* cudaSetDevice for host programs that do not include the CUDA headers (one process per GPU: call it first) */
int fx_set_device(int ordinal);
/* the row-panel shards of a G-rank job: rank r owns rows [cuts[r], cuts[r+1]) -- contiguous ranges of 128-row panels
 * holding about nnz/G nonzeros each (cuts has nranks+1 entries) */
int fx_panel_shards(const fx_matrix *m, int nranks, int64_t *cuts);
int fx_comm_unique_id(char id[128]);
int fx_comm_init(int nranks, int rank, const char id[128], fx_comm **out);
void fx_comm_free(fx_comm *c);
/* rows [*row_lo, *row_hi) of B that this rank uploads for a matrix of n rows */
int fx_comm_slice(const fx_comm *c, int64_t n, int64_t *row_lo, int64_t *row_hi);
/* B_rows_host: this rank's rows of B (row stride k); C_local_host: rows [row_begin,row_end) of C (row stride k).
 * Collective: every rank of the communicator must call it with the same k.  total_ms = first H2D to last D2H on this
 * rank's device clock; tElap_ms = kernels only.  ASpT / tensor-window tiles, k % 4 == 0, k <= the build's k. */
int fx_spmm_sharded_host(const fx_tiles *t, fx_comm *c, const float *B_rows_host, float *C_local_host, int k,
                         float *total_ms, float *tElap_ms);

/* ---- L4: validation + reporting ----------------------------------------------------- */
/* resCheck (flex.cu:4155-4213) and the ASpT validator (aspt/sspmm_128.cu:1425-1446) over host
 * buffers; rowptr may be NULL (row_nnz taken as 1). */
int fx_check(const float *gold, const float *res, int64_t n, int k, const uint32_t *rowptr,
             fx_report *rep);

#ifdef __cplusplus
}
#endif
#endif
