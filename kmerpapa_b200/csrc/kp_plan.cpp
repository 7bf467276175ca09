// kp_plan.cpp — builds the index/split/schedule tables of one general pattern on the host.
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

// The DP kernel hard-codes, per radix of the register position, which base digits each digit covers.
// (digit-space structure; identical for every letter of the same size — checked below)
const uint8_t kDigitBases3[3] = {1, 2, 3};
const uint8_t kDigitBases7[7] = {1, 2, 4, 3, 5, 6, 7};
const uint8_t kDigitBases15[15] = {1, 2, 4, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};

}  // namespace

int kp_build_host_plan(const char *gen_pat, KpHostPlan &P, std::string &err, bool lattice)
{
    P.lattice = lattice;
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
        if (nkmer >= (1ull << 31)) { err = "too many k-mers"; return 5; }
        t.kw[e] = (uint32_t)nkmer;
        for (int d = 0; d < radix; d++) {
            uint8_t m = find_letter(g->digits[d])->mask;
            t.digit_mask[e][d] = m;
            t.mask_digit[e][m] = (uint8_t)d;
        }
        // single-nucleotide digits must enumerate the bases in `code` order (k-mer base index == digit)
        for (int b = 0; b < t.nbase[e]; b++)
            if (g->digits[b] != g->bases[b]) { err = "internal: base/digit order mismatch"; return 4; }
        // digit-space structure the kernel hard-codes
        const uint8_t *want = radix == 3 ? kDigitBases3 : radix == 7 ? kDigitBases7 : kDigitBases15;
        for (int d = 0; d < radix; d++) {
            unsigned got = 0;
            for (int b = 0; b < 4; b++)
                if ((t.digit_mask[e][d] >> b) & 1) got |= 1u << t.mask_digit[e][1 << b];
            if (got != want[d]) { err = "internal: digit structure mismatch"; return 4; }
        }
        if (npat > (UINT64_MAX / 16)) { err = "pattern table too large"; return 5; }
        npat *= (uint64_t)radix;
        nkmer *= (uint64_t)t.nbase[e];
        total_level += t.nbase[e] - 1;
    }
    P.npat = npat;
    P.nkmer = nkmer;
    t.npos = npos;
    t.total_level = (uint32_t)total_level;

    // ---- choose the low positions: largest radices first (ties: lowest position), product <= KP_MAX_TILE ----
    std::vector<int> byradix(npos);
    for (int e = 0; e < npos; e++) byradix[e] = e;
    std::stable_sort(byradix.begin(), byradix.end(), [&](int a, int b) { return t.radix[a] > t.radix[b]; });
    uint32_t cells = 1;
    for (int e : byradix)
        if ((uint64_t)cells * t.radix[e] <= KP_MAX_TILE) { t.is_low[e] = 1; cells *= t.radix[e]; }
    t.estar = npos > 0 ? byradix[0] : -1;
    t.r0 = npos > 0 ? t.radix[t.estar] : 1;
    t.nb0 = npos > 0 ? t.nbase[t.estar] : 1;
    t.ng = (t.r0 + 3) / 4;
    std::vector<int> rowpos;  // low positions other than estar, ascending
    uint32_t roww = 1, lkw = (uint32_t)t.nb0, hw = 1, tk = (uint32_t)t.nb0;
    if (t.estar >= 0) t.lkw[t.estar] = 1;
    for (int e = 0; e < npos; e++) {
        if (t.is_low[e]) {
            t.nlow++;
            if (e == t.estar) continue;
            rowpos.push_back(e);
            t.roww[e] = roww;
            t.lkw[e] = lkw;
            roww *= t.radix[e];
            lkw *= t.nbase[e];
            tk *= t.nbase[e];
        } else {
            t.highpos[t.nhigh++] = (uint8_t)e;
            t.highw[e] = hw;
            if ((uint64_t)hw * t.radix[e] >= (1ull << 31)) {
                if (lattice) { err = "too many tiles"; return 6; }
                hw = 0;   // no lattice: the tile weights are never used
            }
            hw *= t.radix[e];
        }
    }
    t.nrows = (int32_t)roww;
    t.rp = (t.nrows + 7) & ~7;   // 8 rows = one 128-byte line per float4 group: every group starts on a line
    t.tile_cells = cells;
    t.tile_stride = (uint32_t)(t.ng * t.rp * 4);
    t.tile_kmers = tk;
    t.ntiles = hw;
    const uint64_t ntiles = hw;

    // ---- rows of a tile ----
    const int nrows = t.nrows;
    std::vector<std::vector<std::pair<int, int>>> xsplit(nrows);  // (c1 row, c2 row) in scan order
    std::vector<std::vector<uint8_t>> xrank(nrows);               // scan rank (string position * 8 + j) of each
    std::vector<std::vector<uint16_t>> bases(nrows);              // base rows covered
    std::vector<int> level(nrows, 0);
    for (int r = 0; r < nrows; r++) {
        int dig[KP_MAXPOS] = {0};
        uint32_t x = (uint32_t)r;
        for (int e : rowpos) { dig[e] = (int)(x % t.radix[e]); x /= t.radix[e]; }
        std::vector<uint32_t> acc(1, 0);
        for (int e : rowpos) {
            uint8_t m = t.digit_mask[e][dig[e]];
            level[r] += popc4(m) - 1;
            for (int j = 0; j < t.ms_n[m]; j++) {
                int c1 = t.mask_digit[e][t.ms_c1[m][j]], c2 = t.mask_digit[e][t.ms_c2[m][j]];
                if (c1 == 0xFF || c2 == 0xFF || c1 >= dig[e] || c2 >= dig[e]) { err = "internal: split table"; return 7; }
                xsplit[r].push_back({r - (dig[e] - c1) * (int)t.roww[e], r - (dig[e] - c2) * (int)t.roww[e]});
                xrank[r].push_back((uint8_t)(t.pos_id[e] * 8 + j));
            }
            std::vector<uint32_t> nxt;
            for (uint32_t a : acc)
                for (int b = 0; b < 4; b++)
                    if ((m >> b) & 1) nxt.push_back(a + (uint32_t)t.mask_digit[e][1 << b] * (t.lkw[e] / (uint32_t)t.nb0));
            acc.swap(nxt);
        }
        std::sort(acc.begin(), acc.end());
        for (uint32_t a : acc) bases[r].push_back((uint16_t)a);
    }
    // list scheduling into rounds of <= 32 rows whose children all sit in earlier rounds
    std::vector<int> height(nrows, 0), pending(nrows, 0);
    std::vector<std::vector<int>> parents(nrows);
    for (int r = 0; r < nrows; r++) {
        std::vector<int> ch;
        for (auto &pr : xsplit[r]) { ch.push_back(pr.first); ch.push_back(pr.second); }
        std::sort(ch.begin(), ch.end());
        ch.erase(std::unique(ch.begin(), ch.end()), ch.end());
        pending[r] = (int)ch.size();
        for (int c : ch) parents[c].push_back(r);
    }
    for (int r = nrows - 1; r >= 0; r--)  // parents have larger row numbers
        for (int q : parents[r]) height[r] = std::max(height[r], height[q] + 1);
    std::vector<int> ready, order;
    std::vector<uint16_t> round_start(1, 0);
    for (int r = 0; r < nrows; r++)
        if (pending[r] == 0) ready.push_back(r);
    while ((int)order.size() < nrows) {
        if (ready.empty()) { err = "internal: row schedule"; return 8; }
        std::sort(ready.begin(), ready.end(), [&](int a, int b) {
            if (height[a] != height[b]) return height[a] > height[b];
            if (xsplit[a].size() != xsplit[b].size()) return xsplit[a].size() > xsplit[b].size();
            return a < b;
        });
        int take = std::min<int>(32, (int)ready.size());
        std::vector<int> now(ready.begin(), ready.begin() + take);
        ready.erase(ready.begin(), ready.begin() + take);
        std::sort(now.begin(), now.end(), [&](int a, int b) {  // similar cost next to each other
            if (xsplit[a].size() != xsplit[b].size()) return xsplit[a].size() < xsplit[b].size();
            return a < b;
        });
        for (int r : now) order.push_back(r);
        round_start.push_back((uint16_t)order.size());
        for (int r : now)
            for (int q : parents[r])
                if (--pending[q] == 0) ready.push_back(q);
    }
    t.nrounds = (int32_t)round_start.size() - 1;
    P.row_of_srow.assign(order.begin(), order.end());
    P.srow_of_row.assign(nrows, 0);
    for (int s = 0; s < nrows; s++) P.srow_of_row[order[s]] = (uint16_t)s;

    // ---- blob ----
    {
        std::vector<uint8_t> &B = P.rowtab;
        B.clear();
        auto align = [&](size_t a) { while (B.size() % a) B.push_back(0); };
        auto put16 = [&](uint32_t v) { B.push_back((uint8_t)(v & 0xFF)); B.push_back((uint8_t)((v >> 8) & 0xFF)); };
        auto put32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) B.push_back((uint8_t)(v >> (8 * i))); };
        t.rt_round_start = (uint32_t)B.size();
        for (uint16_t v : round_start) put16(v);
        align(4);
        t.rt_row_level = (uint32_t)B.size();
        for (int s = 0; s < nrows; s++) B.push_back((uint8_t)level[order[s]]);
        align(4);
        t.rt_xs_off = (uint32_t)B.size();
        {
            uint32_t o = 0;
            for (int s = 0; s < nrows; s++) { put16(o); o += (uint32_t)xsplit[order[s]].size(); }
            put16(o);
            if (o > 65535) { err = "internal: split list too long"; return 8; }
        }
        align(4);
        t.rt_xs = (uint32_t)B.size();
        for (int s = 0; s < nrows; s++)
            for (auto &pr : xsplit[order[s]])
                put32((uint32_t)P.srow_of_row[pr.first] | ((uint32_t)P.srow_of_row[pr.second] << 16));
        align(4);
        t.rt_xs_rank = (uint32_t)B.size();
        for (int s = 0; s < nrows; s++)
            for (uint8_t rk : xrank[order[s]]) B.push_back(rk);
        align(4);
        t.rt_bs_off = (uint32_t)B.size();
        {
            uint32_t o = 0;
            for (int s = 0; s < nrows; s++) { put16(o); o += (uint32_t)bases[order[s]].size(); }
            put16(o);
            if (o > 65535) { err = "internal: base list too long"; return 8; }
        }
        align(4);
        t.rt_bs = (uint32_t)B.size();
        for (int s = 0; s < nrows; s++)
            for (uint16_t b : bases[order[s]]) put16(b);
        align(4);
        t.rt_srow_of_row = (uint32_t)B.size();
        for (int r = 0; r < nrows; r++) put16(P.srow_of_row[r]);
        align(16);
        t.rt_bytes = (uint32_t)B.size();
    }
    t.maxhs = (uint32_t)((7 * t.nhigh + 3) & ~3);
    if (t.maxhs < 4) t.maxhs = 4;
    for (int wide = 0; wide < 2; wide++) {
        size_t b = (size_t)t.ng * t.rp * 16;          // the warp's copy of its tile
        b += (size_t)tk * 2 * (wide ? 8 : 4);          // base counts of the tile
        b += (size_t)t.maxhs * 8 + 16;                 // high split list (two child tiles each) + count
        t.warp_smem_bytes[wide] = (uint32_t)((b + 15) & ~(size_t)15);
    }

    memset(&P.ft, 0, sizeof P.ft);
    if (!lattice) {
        t.ntiles = 0;
        P.hl_off.assign(1, 0);
        return 0;
    }
    // ---- fiber kernel tables: all-N tile shape (register N + two N row positions) and an N among the high positions ----
    {
        KpFiberTables &f = P.ft;
        int fhi = -1;
        for (int i = 0; i < t.nhigh && fhi < 0; i++)
            if (t.radix[t.highpos[i]] == KP_FIBER_DIGITS) fhi = i;
        if (t.r0 == 15 && nrows == KP_FIBER_ROWS && t.rp == 232 && rowpos.size() == 2 && fhi >= 0) {
            f.ok = 1;
            f.fhi = fhi;
            f.fe = t.highpos[fhi];
            f.hw = t.highw[f.fe];
            f.nfibers = (uint32_t)(ntiles / KP_FIBER_DIGITS);
            // rows in level order (stable in schedule order)
            std::vector<int> frow_srow(nrows);
            for (int s = 0; s < nrows; s++) frow_srow[s] = s;
            std::stable_sort(frow_srow.begin(), frow_srow.end(), [&](int a, int b) { return level[order[a]] < level[order[b]]; });
            std::vector<int> frow_of_srow(nrows);
            for (int fr = 0; fr < nrows; fr++) frow_of_srow[frow_srow[fr]] = fr;
            int maxlvl = 0;
            for (int r = 0; r < nrows; r++) maxlvl = std::max(maxlvl, level[r]);
            if (maxlvl != KP_FIBER_ROW_LEVELS - 1) { err = "internal: fiber row levels"; return 9; }
            for (int l = 0; l <= KP_FIBER_ROW_LEVELS; l++) f.lvl_start[l] = 0;
            for (int r = 0; r < nrows; r++) f.lvl_start[level[r] + 1]++;
            for (int l = 0; l < KP_FIBER_ROW_LEVELS; l++) f.lvl_start[l + 1] += f.lvl_start[l];
            std::vector<uint8_t> &B = P.fibertab;
            B.clear();
            auto align = [&](size_t a) { while (B.size() % a) B.push_back(0); };
            auto put16 = [&](uint32_t v) { B.push_back((uint8_t)(v & 0xFF)); B.push_back((uint8_t)((v >> 8) & 0xFF)); };
            f.ft_srow_of_frow = (uint32_t)B.size();
            for (int fr = 0; fr < nrows; fr++) B.push_back((uint8_t)frow_srow[fr]);
            align(4);
            f.ft_frow_of_srow = (uint32_t)B.size();
            for (int s = 0; s < t.rp; s++) B.push_back(s < nrows ? (uint8_t)frow_of_srow[s] : (uint8_t)0xFF);
            align(4);
            f.ft_xs_off = (uint32_t)B.size();
            {
                uint32_t o = 0;
                for (int fr = 0; fr < nrows; fr++) { put16(o); o += (uint32_t)xsplit[order[frow_srow[fr]]].size(); }
                put16(o);
            }
            align(4);
            f.ft_xs = (uint32_t)B.size();
            for (int fr = 0; fr < nrows; fr++)
                for (auto &pr : xsplit[order[frow_srow[fr]]]) {
                    const int c1 = frow_of_srow[P.srow_of_row[pr.first]], c2 = frow_of_srow[P.srow_of_row[pr.second]];
                    if (c1 >= fr || c2 >= fr) { err = "internal: fiber row order"; return 9; }
                    put16((uint32_t)c1 | ((uint32_t)c2 << 8));
                }
            align(4);
            f.ft_bs_off = (uint32_t)B.size();
            {
                uint32_t o = 0;
                for (int fr = 0; fr < nrows; fr++) { put16(o); o += (uint32_t)bases[order[frow_srow[fr]]].size(); }
                put16(o);
            }
            align(4);
            f.ft_bs = (uint32_t)B.size();
            for (int fr = 0; fr < nrows; fr++)
                for (uint16_t b : bases[order[frow_srow[fr]]]) B.push_back((uint8_t)b);
            align(16);
            f.ft_bytes = (uint32_t)B.size();
            f.maxhs = (uint32_t)((7 * (t.nhigh - 1) + 3) & ~3);
            if (f.maxhs < 4) f.maxhs = 4;
        }
    }
    // ---- tiles sorted by high level ----
    int nhl = 1;
    for (int i = 0; i < t.nhigh; i++) nhl += t.nbase[t.highpos[i]] - 1;
    std::vector<uint8_t> tl(ntiles);
    P.hl_off.assign((size_t)nhl + 1, 0);
    for (uint64_t tile = 0; tile < ntiles; tile++) {
        uint64_t x = tile;
        int l = 0;
        for (int i = 0; i < t.nhigh; i++) {
            int e = t.highpos[i];
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
    if (P.ft.ok) {   // fibers (their digit-0 tile) by wave = level over the other high positions; a digit-0 tile has that level
        const KpFiberTables &f = P.ft;
        const int nfl = nhl - 3;
        std::vector<uint64_t> cnt((size_t)nfl + 1, 0);
        for (uint64_t tile = 0; tile < ntiles; tile++)
            if ((tile / f.hw) % KP_FIBER_DIGITS == 0) cnt[(size_t)tl[tile] + 1]++;
        for (int l = 0; l < nfl; l++) cnt[(size_t)l + 1] += cnt[(size_t)l];
        P.fl_off = cnt;
        P.fiber_order.resize(cnt[(size_t)nfl]);
        std::vector<uint64_t> cur(cnt.begin(), cnt.end() - 1);
        for (uint64_t tile = 0; tile < ntiles; tile++)
            if ((tile / f.hw) % KP_FIBER_DIGITS == 0) P.fiber_order[cur[tl[tile]]++] = (uint32_t)tile;
        if (P.fiber_order.size() != f.nfibers) { err = "internal: fiber count"; return 9; }
    }
    return 0;
}

void kp_locate(const KpHostPlan &P, uint64_t pat, uint64_t *tile, uint32_t *srow, uint32_t *d0)
{
    const KpTables &t = P.t;
    uint64_t tl = 0;
    uint32_t row = 0, d = 0;
    for (int e = 0; e < t.npos; e++) {
        uint32_t dig = (uint32_t)((pat / t.extw[e]) % t.radix[e]);
        if (e == t.estar) d = dig;
        else if (t.is_low[e]) row += dig * t.roww[e];
        else tl += (uint64_t)dig * t.highw[e];
    }
    *tile = tl;
    *srow = P.srow_of_row[row];
    *d0 = d;
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
