// kp_plan.cpp — builds the index/split tables of one general pattern on the host.
//
// Restates, as nucleotide-subset masks (A=1, C=2, G=4, T=8), the reference's letter tables:
//   src/kmerpapa/pattern_utils.py:5-19 (`code`), :48-57 (`complements`), :86-100 (`perm_code`).
#include "kp_plan.h"

#include <string.h>

#include <algorithm>

namespace {

struct Letter {
    char ch;
    uint8_t mask;
    const char *bases;   // k-mer base order of a general letter
    const char *digits;  // digit order of its sub-letters
    const char *splits;  // two-way splits "c1c2c1c2..." in scan order
};

const Letter kLetters[15] = {
    {'A', 1, "A", "A", ""},
    {'C', 2, "C", "C", ""},
    {'G', 4, "G", "G", ""},
    {'T', 8, "T", "T", ""},
    {'R', 5, "AG", "AGR", "AG"},
    {'Y', 10, "CT", "CTY", "CT"},
    {'S', 6, "GC", "GCS", "GC"},
    {'W', 9, "AT", "ATW", "AT"},
    {'K', 12, "GT", "GTK", "GT"},
    {'M', 3, "AC", "ACM", "AC"},
    {'B', 14, "CGT", "CGTSYKB", "CKGYTS"},
    {'D', 13, "AGT", "AGTRWKD", "AKGWTR"},
    {'H', 11, "ACT", "ACTMWYH", "AYCWTM"},
    {'V', 7, "ACG", "ACGMRSV", "ASCRGM"},
    {'N', 15, "ACGT", "ACGTRYSWKMBDHVN", "SWKMRYABCDGHTV"},
};

const Letter *find_letter(char c)
{
    for (const Letter &l : kLetters)
        if (l.ch == c) return &l;
    return nullptr;
}

int popc4(unsigned m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

}  // namespace

int kp_build_host_plan(const char *gen_pat, KpHostPlan &P, std::string &err)
{
    size_t len = strlen(gen_pat);
    if (len < 1 || len > KP_MAXK) { err = "general pattern length must be 1.." + std::to_string(KP_MAXK); return 1; }
    P.gen = gen_pat;
    P.k = (int)len;
    KpTables &t = P.t;
    memset(&t, 0, sizeof t);
    memset(t.mask_digit, 0xFF, sizeof t.mask_digit);

    // universal split table by subset mask
    for (const Letter &l : kLetters) {
        int ns = (int)strlen(l.splits) / 2;
        t.ms_n[l.mask] = (uint8_t)ns;
        for (int j = 0; j < ns; j++) {
            t.ms_c1[l.mask][j] = find_letter(l.splits[2 * j])->mask;
            t.ms_c2[l.mask][j] = find_letter(l.splits[2 * j + 1])->mask;
        }
    }

    int npos = 0;
    uint64_t npat = 1, nkmer = 1;
    int total_level = 0;
    for (int i = 0; i < P.k; i++) {
        const Letter *g = find_letter(gen_pat[i]);
        if (!g) { err = std::string("not an IUPAC letter: '") + gen_pat[i] + "'"; return 2; }
        P.gen_mask[i] = g->mask;
        P.eff_of_pos[i] = 0xFF;
        int radix = (int)strlen(g->digits);
        if (radix == 1) continue;
        if (npos >= KP_MAXPOS) { err = "too many multi-letter positions (max " + std::to_string(KP_MAXPOS) + ")"; return 3; }
        int e = npos++;
        P.eff_of_pos[i] = (uint8_t)e;
        t.pos_id[e] = (uint8_t)i;
        t.radix[e] = (uint8_t)radix;
        t.nbase[e] = (uint8_t)strlen(g->bases);
        t.extw[e] = npat;
        for (int d = 0; d < radix; d++) {
            uint8_t m = find_letter(g->digits[d])->mask;
            t.digit_mask[e][d] = m;
            t.mask_digit[e][m] = (uint8_t)d;
        }
        // singleton digits must enumerate the bases in `code` order (k-mer index == digit for k-mers)
        for (int b = 0; b < t.nbase[e]; b++)
            if (g->digits[b] != g->bases[b]) { err = "internal: base/digit order mismatch"; return 4; }
        if (npat > (UINT64_MAX / 16)) { err = "pattern table too large"; return 5; }
        npat *= (uint64_t)radix;
        nkmer *= (uint64_t)t.nbase[e];
        total_level += t.nbase[e] - 1;
    }
    P.npat = npat;
    P.nkmer = nkmer;
    t.npos = npos;
    t.total_level = (uint32_t)total_level;

    // low positions: longest prefix whose cells fit a tile and whose digit fields fit 16 bits
    int nlow = 0, bits = 0;
    uint32_t cells = 1, tk = 1;
    while (nlow < npos && nlow < KP_MAXLOW) {
        int r = t.radix[nlow];
        int b = r <= 3 ? 2 : (r <= 7 ? 3 : 4);
        if ((uint64_t)cells * r > KP_MAX_TILE || bits + b > 16) break;
        t.shift[nlow] = (uint8_t)bits;
        t.fmask[nlow] = (uint8_t)((1 << b) - 1);
        t.loww[nlow] = cells;
        t.lowkw[nlow] = tk;
        bits += b;
        cells *= r;
        tk *= t.nbase[nlow];
        nlow++;
    }
    t.nlow = nlow;
    t.nhigh = npos - nlow;
    t.tile_cells = cells;
    t.tile_stride = (cells + 31u) & ~31u;
    t.tile_kmers = tk;
    uint64_t ntiles = npat / cells;
    if (ntiles >= (1ull << 31)) { err = "too many tiles"; return 6; }
    t.ntiles = (uint32_t)ntiles;
    {
        uint32_t hw = 1, hkw = 1;
        for (int e = nlow; e < npos; e++) {
            t.highw[e] = hw;
            t.highkw[e] = hkw;
            hw *= t.radix[e];
            hkw *= t.nbase[e];
        }
    }

    // per low position split offsets
    for (int e = 0; e < nlow; e++) {
        for (int d = 0; d < t.radix[e]; d++) {
            uint8_t m = t.digit_mask[e][d];
            int ns = t.ms_n[m];
            t.low_ns[e][d] = (uint8_t)ns;
            for (int j = 0; j < ns; j++) {
                int c1 = t.mask_digit[e][t.ms_c1[m][j]], c2 = t.mask_digit[e][t.ms_c2[m][j]];
                if (c1 == 0xFF || c2 == 0xFF || c1 >= d || c2 >= d) { err = "internal: split table"; return 7; }
                t.low_d1[e][d][j] = (int16_t)((c1 - d) * (int)t.loww[e]);
                t.low_d2[e][d][j] = (int16_t)((c2 - d) * (int)t.loww[e]);
            }
        }
    }

    // cells sorted by mini-level, then by the per-position subset sizes (keeps warps uniform), then id
    struct CellKey { uint32_t ml, sig, cell, packed; };
    std::vector<CellKey> keys(cells);
    int nml = 0;
    for (uint32_t c = 0; c < cells; c++) {
        uint32_t x = c, ml = 0, sig = 0, packed = 0, hns1 = 0;
        for (int e = 0; e < nlow; e++) {
            uint32_t d = x % t.radix[e];
            x /= t.radix[e];
            int sz = popc4(t.digit_mask[e][d]);
            ml += (uint32_t)(sz - 1);
            sig = sig * 4 + (uint32_t)(sz - 1);
            packed |= d << t.shift[e];
            if (sz > 1) hns1 = (uint32_t)e + 1;  // highest multi-letter low position, +1 (0: a k-mer cell)
        }
        packed |= hns1 << 28;
        keys[c] = {ml, sig, c, packed};
        nml = std::max(nml, (int)ml + 1);
    }
    if (nml > KP_MAXML) { err = "internal: too many mini-levels"; return 8; }
    std::sort(keys.begin(), keys.end(), [](const CellKey &a, const CellKey &b) {
        if (a.ml != b.ml) return a.ml < b.ml;
        if (a.sig != b.sig) return a.sig < b.sig;
        return a.cell < b.cell;
    });
    t.nml = nml;
    P.cell_list.resize(cells);
    for (uint32_t i = 0; i < cells; i++) {
        P.cell_list[i] = (keys[i].cell << 16) | keys[i].packed;  // [31:28] hns1, [27:16] cell, [15:0] digits
        t.ml_off[keys[i].ml + 1]++;
    }
    for (int l = 0; l < nml; l++) t.ml_off[l + 1] += t.ml_off[l];

    // tiles sorted by high level
    int nhl = 1;
    for (int e = nlow; e < npos; e++) nhl += t.nbase[e] - 1;
    std::vector<uint8_t> tl(ntiles);
    P.hl_off.assign((size_t)nhl + 1, 0);
    for (uint64_t tile = 0; tile < ntiles; tile++) {
        uint64_t x = tile;
        int l = 0;
        for (int e = nlow; e < npos; e++) {
            l += popc4(t.digit_mask[e][x % t.radix[e]]) - 1;
            x /= t.radix[e];
        }
        tl[tile] = (uint8_t)l;
        P.hl_off[(size_t)l + 1]++;
    }
    for (int l = 0; l < nhl; l++) P.hl_off[(size_t)l + 1] += P.hl_off[(size_t)l];
    P.tile_order.resize(ntiles);
    {
        std::vector<uint64_t> cur(P.hl_off.begin(), P.hl_off.end() - 1);
        for (uint64_t tile = 0; tile < ntiles; tile++) P.tile_order[cur[tl[tile]]++] = (uint32_t)tile;
    }
    return 0;
}

void kp_num2masks(const KpHostPlan &P, uint64_t num, uint8_t *masks_out)
{
    for (int i = 0; i < P.k; i++) {
        uint8_t e = P.eff_of_pos[i];
        if (e == 0xFF) { masks_out[i] = P.gen_mask[i]; continue; }
        masks_out[i] = P.t.digit_mask[e][num % P.t.radix[e]];
        num /= P.t.radix[e];
    }
}
