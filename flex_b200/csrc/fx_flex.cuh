// fx_flex.cuh -- device/host state of the Flex formats (Mat_POD fields, mat.cuh:18-63)
#pragma once
#include <stdint.h>

#include <vector>

struct fx_flex_dev {
  struct PillarHost {  // csr2_DiagTiling outputs (mat.cuh:89-100)
    std::vector<unsigned> alpha_rowPtr, alpha_colIdx, alpha_pillar_rowPtr, alpha_pillarIdx, segVoMap;
    std::vector<float> alpha_vals;
    int n_segs = 0, warps_with_weights = 0;
    float empty_wp_p = 0, band_nz_p = 0;
  };
  int m = 0, nnz = 0, tm = 4, tn = 4, cmajor = 0, nnz_limit = 128, n_sm = 148, npanels = 0;
  int *count = nullptr, *off = nullptr;  // per-panel tile/segment counts and their exclusive scan
  // tile format
  unsigned *tileRowPtr = nullptr, *tileNnz = nullptr, *tileColIdx = nullptr;
  int *nnzTile = nullptr, *bitMap = nullptr, *rcOffset = nullptr;
  float* newVals = nullptr;
  int ntiles = 0;
  // seg / pillar ("alpha") format
  unsigned *alpha_rowPtr = nullptr, *alpha_colIdx = nullptr, *pillar_rowPtr = nullptr, *segVoMap = nullptr, *pillarIdx = nullptr;
  float* alpha_vals = nullptr;
  unsigned *segPtr = nullptr, *segNzRCIdx = nullptr, *segVoMapPad = nullptr;
  float *segVals = nullptr, *segNzCV = nullptr;
  int* seg_rowPtr = nullptr;
  int *next_seg = nullptr, *grouped_tailSeg = nullptr;
  unsigned* counter = nullptr;
  int nsegs = 0, rows_total = 0, seg_cap = 0;
  // pillar builder (fx_flex_build.cu, rounds 2-3 of csr2_DiagTiling on the GPU)
  unsigned char* listed = nullptr;            // [m]   column seen inside a diagonal block in round 1
  int* pstart = nullptr;                      // [partitions+1] first row of every diagonal block
  int *nzcnt = nullptr, *nzoff = nullptr;     // per row panel: nz left for round 3 and their exclusive scan
  int* perr = nullptr;                        // error flags of the device rounds
  unsigned* rest_col = nullptr;               // the nz round 2 leaves, compacted row by row (input of round 3)
  float* rest_val = nullptr;
  int *dpos = nullptr, *ends = nullptr;       // round 1 on the GPU: position of the diagonal in every row, end of the block starting at every row
  bool round1_on_gpu = false;
  void* scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  int partitions = 0, r2_nnz = 0;
  PillarHost ph;
  std::vector<int> h_count, h_next, h_tail;
  // host copies handed out by the export calls
  std::vector<unsigned> e_u32[8];
  std::vector<int> e_i32[4];
  std::vector<float> e_f32[4];
};
