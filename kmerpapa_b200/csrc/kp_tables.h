// kp_tables.h — plan-time tables shared by host and device code.
//
// IUPAC facts restated from the reference (src/kmerpapa/pattern_utils.py):
//   :5-19    letter -> nucleotide list (`code`), also the k-mer base order of a general letter
//   :48-57   two-way splits (c1,c2) of a letter in scan order (`complements`)
//   :86-100  digit order of the sub-letters of a general letter (`perm_code`)
//   :237-257 dense pattern number = sum digit_i * w_i, position 0 least significant
#pragma once
#include <stdint.h>

#define KP_MAXK 32        // pattern length
#define KP_MAXPOS 16      // positions with more than one letter (the others carry no digit)
#define KP_MAXLOW 8       // positions kept inside a tile
#define KP_MAXML 25       // mini-levels inside a tile (3 per low position + 1)
#define KP_MAX_TILE 4096  // cells per tile upper bound
#define KP_MAXHS (KP_MAXPOS * 7)

// Everything a kernel needs to know about one general pattern.  Lives in device global memory;
// CTAs copy the hot parts to shared memory once.
struct KpTables {
    int32_t npos;    // effective (multi-letter) positions, ascending string position
    int32_t nlow;    // the first nlow of them live inside a tile
    int32_t nhigh;   // npos - nlow
    int32_t nml;     // mini-levels inside a tile
    uint32_t tile_cells, tile_stride, tile_kmers;
    uint32_t ntiles;
    uint32_t total_level;

    uint8_t pos_id[KP_MAXPOS];   // string position (rank code = pos_id * 8 + split index)
    uint8_t radix[KP_MAXPOS];    // 3, 7 or 15
    uint8_t nbase[KP_MAXPOS];    // 2, 3 or 4
    uint8_t shift[KP_MAXLOW];    // bit field of the digit inside a packed cell word
    uint8_t fmask[KP_MAXLOW];
    uint32_t loww[KP_MAXLOW];    // cell weight of a low position
    uint32_t lowkw[KP_MAXLOW];   // low k-mer weight of a low position
    uint32_t highw[KP_MAXPOS];   // tile weight of a high position (index npos-relative: [nlow..npos))
    uint32_t highkw[KP_MAXPOS];  // high k-mer weight of a high position
    uint64_t extw[KP_MAXPOS];    // dense pattern-number weight

    uint8_t digit_mask[KP_MAXPOS][16];  // digit -> nucleotide subset (A=1,C=2,G=4,T=8)
    uint8_t mask_digit[KP_MAXPOS][16];  // subset -> digit (0xFF if not a sub-letter)
    // universal, indexed by subset mask
    uint8_t ms_n[16];       // number of two-way splits
    uint8_t ms_c1[16][7];   // c1 subset of split j
    uint8_t ms_c2[16][7];
    // per low position, by digit: splits as negative cell offsets (c1 in .x, c2 in .y)
    uint8_t low_ns[KP_MAXLOW][16];
    int16_t low_d1[KP_MAXLOW][16][8];
    int16_t low_d2[KP_MAXLOW][16][8];
    uint32_t ml_off[KP_MAXML + 2];  // mini-level offsets into the cell list
};
