// Data layout shared by the host driver and the kernels of the weighted pairwise-MI scan (see mi_scan.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ldw {

// One record per (slot, variant): everything the epilogue needs to know about one SNP.
//   T   : fixed-point weighted marginals of the SNP's allele slots q = 0..r-1 in the epilogue's count unit
//         (floor(V / 2^sb), V = the exact fixed-point sum) PLUS r' * M, the pseudocounts of the r' cells of that row of
//         the joint table (M = 0.5 in count units, r' = variant + 2); 0 for q >= r.  The last observed allele (slot r-1)
//         is the "complement" slot whose joint counts are derived by subtraction instead of from the GEMM.
//   q   : p_q + 0.5 r' for r' = variant+2 (r of the partner SNP); rp = 1/q.
// Layout: four 16-byte vectors {T0..T3} {rp0..rp3} {q0..q3} {T4, rp4, q4, -}; kinds with at most four allele slots
// read two of the first three (the Q1-corrected form uses q, the plain form rp).
struct __align__(16) Rec {
  uint32_t T[4];
  float rp[4];
  float q[4];
  uint32_t T4;
  float rp4;
  float q4;
  uint32_t pad;
};
static_assert(sizeof(Rec) == 64, "Rec must be 64 bytes");

// Per-block, per from-slot: local row index (or -1) and r of the TO-list SNP that sits at that local index
// (quirk Q1: the reference reads the transposed rft by linear index).
struct RowDyn {
  int32_t il;
  float rtl;
};

// Per-block, per to-slot: local column index (or -1), r of the FROM-list SNP at that local index (Q1), the
// column's short-range row intervals [a0,a1) u [b0,b1) in local-row space and the column's output bases.
struct __align__(16) ColDyn {
  int32_t jl;
  float rfl;
  int32_t a0, a1, b0, b1;
  uint32_t baseU, baseL;
};
static_assert(sizeof(ColDyn) == 32, "ColDyn must be 32 bytes");

enum : uint32_t { TILE_HAS_SR = 1u, TILE_NULL = 2u };

// One unit of work of a CTA pair: 2 x 128 row SNPs (PA planes each; CTA `rank` of the pair owns rows rank*128 ..) x NJ column
// SNPs (PB planes each; CTA `rank` stages columns rank*NJ/2 ..).  The row-side fields name the first half; the second half
// follows 128 slots / operand rows later.  A half that holds no wanted pair (or does not exist: odd number of row tiles)
// is flagged TILE_NULL: its MMAs run with the pair's, its epilogue is skipped.
struct __align__(16) TileDesc {
  int32_t a_row0, a_pstride;  // operand row of plane 0 of the row tile; rows between planes
  int32_t b_row0, b_pstride;
  int32_t i_slot0, j_slot0;   // global slot ids (index into Rec)
  int32_t i_dyn0, j_dyn0;     // offsets into RowDyn / ColDyn
  uint8_t PA, PB, njlog2, flags;   // flags: rows 0..127 (CTA 0)
  uint8_t flags1, pad8[3];         // flags1: rows 128..255 (CTA 1)
  int32_t pad[2];
};
static_assert(sizeof(TileDesc) == 48, "TileDesc must be 48 bytes");

struct Cand {
  int32_t il, jl;
  float mi;
};

constexpr int MI_HIST_BINS = 4096;  // positive-float bits >> 19

struct ScanParams {
  const TileDesc* tiles;
  int32_t n_tiles;
  const Rec* rec;
  int64_t rec_vstride;
  const RowDyn* rowdyn;
  const ColDyn* coldyn;
  int32_t nkb;  // K blocks of 128 sequences
  int32_t nf, nt;
  int32_t diag, ragged, qcorr, sr_only, dense, emit_all;
  const uint8_t* rfl_arr;  // r of from-list by local index (ragged Q1 path)
  const uint8_t* rtl_arr;  // r of to-list by local index
  float kT;               // joint count = t * kT with t = (H << sa) + (L >> sb) = floor(V / 2^sb)
  uint32_t sa, sb;
  uint32_t M;             // the pseudocount 0.5 in count units (kT = 0.5 / M exactly): cells are evaluated as t + M
  // per tile kind [r_i - 2][r_j - 2], with den = neff + 0.5 r_i r_j: the epilogue's constants, formed on the host in the same
  // single-precision operations the kernel used to run per tile (IEEE: identical bits)
  float scale[4][4];        // (ln 2 / den) * kT
  float q0s[4][4];          // (0.25 r_i r_j / den) / kT
  float qod[4][4];          // (0.25 / den) / kT
  float rp_qc[4][4];        // (1 / den) * (1 / kT): row factor of the Q1 form
  float rp_plain[4][4];     // den * kT: row factor of the plain form
  float* sr_out;            // this block's SR slots
  float* dense_out;         // debug: nf x nt, column-major
  Cand* cand;
  uint32_t cand_cap;
  uint32_t* cand_count;
  uint32_t* tcand_bits;
  uint32_t* hist;
  uint32_t kprime, delta;
  uint32_t* overflow;
  const uint8_t* dig;       // [4][Kpad] per-sequence weight digits D3, D2, D1, D0
  int64_t kpad;
  unsigned long long* dbg;  // optional per-CTA cycle counters (profiling aid; NULL in production)
};

// One global operand array: the 0/1 one-hot planes X [plane rows][Kpad sequences].  Row tiles (A) and column
// tiles (B) are boxes of the same tensor; the x128 copy of A and the four digit-weighted copies of B are built
// inside shared memory by the expander warps.
struct TmapSet {
  CUtensorMap a;      // box 128 rows
  CUtensorMap b[4];   // box rows 64, 32, 16, 8: one CTA's half of a column tile of 128, 64, 32, 16 SNPs
};

}  // namespace ldw
