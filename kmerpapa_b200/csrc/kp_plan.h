// kp_plan.h — host-side construction of the per-general-pattern tables (no CUDA in here).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "kp_tables.h"

struct KpHostPlan {
    std::string gen;
    int k = 0;
    uint64_t npat = 0, nkmer = 0;
    KpTables t;                        // uploaded verbatim
    std::vector<uint8_t> rowtab;       // row tables blob (offsets in t.rt_*)
    std::vector<uint16_t> srow_of_row, row_of_srow;  // natural row <-> schedule position
    std::vector<uint32_t> tile_order;  // tile ids sorted by (high level, tile id)
    std::vector<uint64_t> hl_off;      // offsets of the high levels in tile_order (size nhl + 1)
    uint8_t gen_mask[KP_MAXK];         // nucleotide subset of every string position (fixed ones too)
    uint8_t eff_of_pos[KP_MAXK];       // string position -> effective position index, 0xFF if fixed
    bool lattice = true;               // false: no tile lattice (tile_order / hl_off empty, the DP entry points refuse)
    // fiber kernel (kp_fiber.cuh): tables, and the fibers (tile id of their digit-0 tile) sorted by (fiber wave, id)
    KpFiberTables ft;
    std::vector<uint8_t> fibertab;
    std::vector<uint32_t> fiber_order;
    std::vector<uint64_t> fl_off;      // offsets of the fiber waves in fiber_order
};

// Returns 0 on success; on failure fills err.
// lattice = false: digit / k-mer tables only (greedy estimator, counts); no tile lattice, so any k the k-mer table allows
int kp_build_host_plan(const char *gen_pat, KpHostPlan &P, std::string &err, bool lattice = true);

// dense pattern number -> (tile, row in schedule order, digit of the register position)
void kp_locate(const KpHostPlan &P, uint64_t pat, uint64_t *tile, uint32_t *srow, uint32_t *d0);

// dense pattern number -> IUPAC string / nucleotide masks (host utility, mirrors num2pattern)
void kp_num2masks(const KpHostPlan &P, uint64_t num, uint8_t *masks_out);
