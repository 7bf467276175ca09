// kp_tables.h — plan-time tables shared by host and device code.
//
// IUPAC facts restated from the reference (src/kmerpapa/pattern_utils.py):
//   :5-19    letter -> nucleotide list (`code`), also the k-mer base order of a general letter
//   :48-57   two-way splits (c1,c2) of a letter in scan order (`complements`)
//   :86-100  digit order of the sub-letters of a general letter (`perm_code`)
//   :237-257 dense pattern number = sum digit_i * w_i, position 0 least significant
//
// Geometry.  "Effective" positions are the multi-letter positions of the general pattern (fixed letters
// carry no digit), numbered in string order.  A subset of them (largest radices first, product of
// radices <= KP_MAX_TILE) is kept on chip: the LOW positions.  One of the low positions, `estar` (the
// one with the largest radix r0), is the register dimension of the DP kernel; the other low positions
// enumerate the `nrows` rows of a tile; the HIGH positions enumerate the tiles.
//
// Device layout of a float table: tile t, row in schedule order ("srow"), digit d of position estar at
//   t * tile_stride + ((d >> 2) * rp + srow) * 4 + (d & 3)          (tile_stride = ng * rp * 4 floats)
// i.e. each row is ng float4 groups and the 32 lanes of a warp touch consecutive rows (coalesced).
#pragma once
#include <stdint.h>

#define KP_MAXK 32        // pattern length
#define KP_MAXPOS 16      // positions with more than one letter
#define KP_MAX_TILE 4096  // cells per tile upper bound
#define KP_MAXHS (KP_MAXPOS * 7)

struct KpTables {
    int32_t npos;             // effective positions
    int32_t nlow, nhigh;      // low (on-chip) and high (tile) positions
    int32_t estar;            // effective index of the register position (-1: the pattern has no free position)
    int32_t r0, nb0, ng;      // its radix / number of bases / float4 groups per row (1,1,1 when estar < 0)
    int32_t nrows, rp;        // rows per tile, row pitch (nrows rounded up to a multiple of 8)
    int32_t nrounds;          // schedule rounds (<= 32 rows each)
    uint32_t tile_cells, tile_stride, tile_kmers;
    uint32_t ntiles;
    uint32_t total_level;

    uint8_t pos_id[KP_MAXPOS];    // string position of an effective position
    uint8_t radix[KP_MAXPOS];     // 3, 7 or 15
    uint8_t nbase[KP_MAXPOS];     // 2, 3 or 4
    uint8_t is_low[KP_MAXPOS];
    uint8_t highpos[KP_MAXPOS];   // effective indices of the high positions, ascending
    uint32_t roww[KP_MAXPOS];     // row weight (low positions other than estar)
    uint32_t highw[KP_MAXPOS];    // tile weight (high positions)
    uint32_t kw[KP_MAXPOS];       // k-mer index weight (all effective positions, position 0 fastest)
    uint32_t lkw[KP_MAXPOS];      // weight inside the tile's low k-mer index (estar fastest) for low positions
    uint64_t extw[KP_MAXPOS];     // dense pattern-number weight

    uint8_t digit_mask[KP_MAXPOS][16];  // digit -> nucleotide subset (A=1,C=2,G=4,T=8)
    uint8_t mask_digit[KP_MAXPOS][16];  // subset -> digit (0xFF if not a sub-letter)
    // universal, indexed by subset mask
    uint8_t ms_n[16];       // number of two-way splits
    uint8_t ms_c1[16][7];   // c1 subset of split j
    uint8_t ms_c2[16][7];

    // row tables blob (device global; CTAs copy it to shared memory): byte offsets into the blob
    uint32_t rt_bytes;
    uint32_t rt_round_start;  // u16 [nrounds + 1]   first srow of each round
    uint32_t rt_row_level;    // u8  [nrows]         by srow; 0 = every row digit is a single nucleotide
    uint32_t rt_xs_off;       // u16 [nrows + 1]     by srow; cross-row splits (CSR)
    uint32_t rt_xs;           // u32 [...]           srow of c1 | srow of c2 << 16, scan order
    uint32_t rt_xs_rank;      // u8  [...]           scan rank of each cross-row split (string position * 8 + j)
    uint32_t rt_bs_off;       // u16 [nrows + 1]     by srow; base rows covered by the row (CSR)
    uint32_t rt_bs;           // u16 [...]           base-row index (low k-mer index / nb0)
    uint32_t rt_srow_of_row;  // u16 [nrows]         natural row -> srow
    uint32_t maxhs;                  // capacity of a tile's high-split list (7 per high position, rounded to 4)
    uint32_t warp_smem_bytes[2];     // per-warp shared memory of the DP kernel [wide]
};

// ---------------------------------------------------------------------------------------------------
// Fiber kernel (kp_fiber.cuh): a FOURTH position on chip.  One high position of radix 15 (an N), the "fiber position",
// joins the three low positions: a CTA owns the 15 tiles that differ only in that digit (a fiber, 15^4 = 50 625
// patterns, 202.5 KB of float32 in shared memory), so the splits of the fiber position never leave the SM.  Inside
// the CTA the 15 x 225 rows are processed level by level (level of the fiber digit + level of the row), and the rows
// of a tile are kept in LEVEL order ("frow") in shared memory; the HBM layout (schedule order, "srow") is unchanged.
// ---------------------------------------------------------------------------------------------------
#define KP_FIBER_ROWS 225
#define KP_FIBER_DIGITS 15
#define KP_FIBER_ROW_LEVELS 7
struct KpFiberTables {
    int32_t ok;                // the general pattern has the all-N tile shape and an N among its high positions
    int32_t fe;                // effective index of the fiber position
    int32_t fhi;               // its index among the high positions
    uint32_t hw;               // its tile weight
    uint32_t nfibers;
    uint32_t lvl_start[KP_FIBER_ROW_LEVELS + 1];   // rows of row-level l: frow in [lvl_start[l], lvl_start[l + 1])
    // blob (device global; CTAs copy it to shared memory): byte offsets
    uint32_t ft_bytes;
    uint32_t ft_srow_of_frow;  // u8  [225]
    uint32_t ft_frow_of_srow;  // u8  [232]  (padding rows -> 0xFF)
    uint32_t ft_xs_off;        // u16 [226]  by frow; cross-row splits (CSR)
    uint32_t ft_xs;            // u16 [...]  frow of c1 | frow of c2 << 8
    uint32_t ft_bs_off;        // u16 [226]  by frow; base rows covered by the row (CSR)
    uint32_t ft_bs;            // u8  [...]  base-row index
    uint32_t maxhs;            // capacity of a fiber's high-split list
};

