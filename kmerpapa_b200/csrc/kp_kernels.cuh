// kp_kernels.cuh — sm_100a kernels of the pattern-partition DP.
//
// Data layout: the pattern table is cut into tiles of `tile_cells` consecutive dense
// pattern numbers (the low positions of the general pattern); a tile is identified by the digits of the
// remaining (high) positions.  One CTA owns one tile at a time:
//   phase A  list the tile's high-position splits (two child tiles each, same cell offset)
//   phase B  load the tile's low-k-mer counts and expand them to all cells in shared memory
//            (subset sums, one pass per low position: each pattern = sum of disjoint sub-patterns)
//   phase C  float64 self-score of every cell (glibc-exact log), kept in shared memory
//   phase D  stream the child tiles from HBM/L2 with 16-byte coalesced loads, keep the running
//            (min, first rank) of f32(best[c1] + best[c2]) per cell
//   phase E  wavefront over the tile's own levels in shared memory: low-position splits, merge with the
//            streamed minimum, compare with the self-score (float64 compare, float32 store)
//   phase F  write the tile's best scores and split codes once
// Tiles of one "high level" (sum of the levels of the high digits) are independent: one launch per level.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kp_math.cuh"
#include "kp_tables.h"

#define KP_NT 256          // threads per CTA of the DP kernel
#define KP_QPT 4           // 16-byte chunks per thread in the streaming phase (single DP)

struct KpDpParams {
    const KpTables *tab;
    const uint32_t *cell_list;
    const uint32_t *tile_list;  // tiles of this wave, ascending
    uint32_t ntiles_wave;
    int leaf_wave;              // wave 0: cells of mini-level 0 are k-mers (level-0 formula)
    const long long *e0, *e1, *e2, *e3;  // single: M, U.  CV: Mtot, Utot, Mtest, Utest
    double alpha, beta, penalty;
    float *best;        // single
    uint8_t *split;     // single
    float *tt;          // CV: (train, test) interleaved
};

__host__ __device__ inline size_t kp_dp_smem_bytes(bool cv, bool wide, uint32_t cells, uint32_t stride, int nlow)
{
    size_t ns = (size_t)(cv ? 2 : 1) * (wide ? 2 : 1);
    size_t b = 2048;                         // log table
    b += ns * stride * 8;                    // count slots, later self-scores
    b += (size_t)(cv ? 2 : 1) * stride * 4;  // S (and T)
    b += cv ? 0 : stride;                    // R
    b += ((size_t)cells + 3) / 4 * 16;       // cell list
    b += (size_t)nlow * 128 * 4;             // low split offsets
    b += (size_t)nlow * 16;                  // low split counts
    b += KP_MAXHS * 4 * 2 + 128;             // high split list
    b += 32 * 4;                             // mini-level offsets
    b += 16 + KP_MAXLOW * 4;                 // scalars, per-low-position meta
    return b;
}

template <bool WIDE>
__device__ __forceinline__ unsigned long long kp_ldcnt(const unsigned long long *slot, uint32_t stride, int which,
                                                       uint32_t cell)
{
    if (WIDE) return slot[(size_t)which * stride + cell];
    return ((const uint32_t *)slot)[((size_t)(which >> 1) * stride + cell) * 2 + (which & 1)];
}
template <bool WIDE>
__device__ __forceinline__ void kp_stcnt(unsigned long long *slot, uint32_t stride, int which, uint32_t cell,
                                         unsigned long long v)
{
    if (WIDE) slot[(size_t)which * stride + cell] = v;
    else ((uint32_t *)slot)[((size_t)(which >> 1) * stride + cell) * 2 + (which & 1)] = (uint32_t)v;
}

template <bool CV, bool WIDE>
__global__ void __launch_bounds__(KP_NT) kp_dp_wave_kernel(const KpDpParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const KpTables &tb = *p.tab;
    const int tid = threadIdx.x;
    const uint32_t cells = tb.tile_cells, stride = tb.tile_stride, tk = tb.tile_kmers;
    const int nlow = tb.nlow, npos = tb.npos, nml = tb.nml;
    constexpr int NS = (CV ? 2 : 1) * (WIDE ? 2 : 1);
    constexpr int NC = CV ? 4 : 2;  // logical counters per cell

    double2 *logtab = (double2 *)smem;
    unsigned long long *slot = (unsigned long long *)(smem + 2048);
    float *S = (float *)(slot + (size_t)NS * stride);
    float *T = S + stride;                                    // CV only
    uint8_t *R = (uint8_t *)(S + (size_t)(CV ? 2 : 1) * stride);  // single only
    uint32_t *cl = (uint32_t *)(R + (CV ? 0 : stride));
    int *lowd = (int *)(cl + ((cells + 3) / 4) * 4);
    uint8_t *lowns = (uint8_t *)(lowd + nlow * 128);
    uint32_t *hs1 = (uint32_t *)(lowns + nlow * 16);
    uint32_t *hs2 = hs1 + KP_MAXHS;
    uint8_t *hsr = (uint8_t *)(hs2 + KP_MAXHS);
    uint32_t *mlo = (uint32_t *)(hsr + 128);
    int *s_nhs = (int *)(mlo + 32);
    uint32_t *lowmeta = (uint32_t *)(s_nhs + 4);  // shift | field mask << 8 | rank base << 16

    // ---- once per CTA: stage the read-only tables ----
    for (int i = tid; i < 128; i += KP_NT) logtab[i] = make_double2(kpc_logTab[2 * i], kpc_logTab[2 * i + 1]);
    for (uint32_t i = tid; i < cells; i += KP_NT) cl[i] = p.cell_list[i];
    for (int i = tid; i < nlow * 128; i += KP_NT) {
        int e = i >> 7, d = (i >> 3) & 15, j = i & 7;
        lowd[i] = (int)(((uint32_t)(uint16_t)tb.low_d1[e][d][j]) | ((uint32_t)(uint16_t)tb.low_d2[e][d][j] << 16));
    }
    for (int i = tid; i < nlow * 16; i += KP_NT) lowns[i] = tb.low_ns[i >> 4][i & 15];
    for (int i = tid; i <= nml; i += KP_NT) mlo[i] = tb.ml_off[i];
    for (int i = tid; i < nlow; i += KP_NT)
        lowmeta[i] = (uint32_t)tb.shift[i] | ((uint32_t)tb.fmask[i] << 8) | ((uint32_t)tb.pos_id[i] * 8u << 16);
    __syncthreads();

    const double alpha = p.alpha, beta = p.beta, penalty = p.penalty;
    const float INF = __int_as_float(0x7f800000);

    for (uint32_t it = blockIdx.x; it < p.ntiles_wave; it += gridDim.x) {
        const uint32_t tile = p.tile_list[it];

        // ---- phase A: the tile's high-position splits, in scan order (position, split) ----
        if (tid < 32) {
            int e = nlow + tid, ns = 0, d = 0;
            uint32_t m = 0, hw = 1;
            if (e < npos) {
                hw = tb.highw[e];
                d = (int)((tile / hw) % tb.radix[e]);
                m = tb.digit_mask[e][d];
                ns = tb.ms_n[m];
            }
            int off = ns;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, off, o);
                if (tid >= o) off += v;
            }
            int total = __shfl_sync(0xffffffffu, off, 31);
            off -= ns;
            for (int j = 0; j < ns; j++) {
                int c1 = tb.mask_digit[e][tb.ms_c1[m][j]], c2 = tb.mask_digit[e][tb.ms_c2[m][j]];
                hs1[off + j] = tile - (uint32_t)(d - c1) * hw;
                hs2[off + j] = tile - (uint32_t)(d - c2) * hw;
                hsr[off + j] = (uint8_t)(tb.pos_id[e] * 8 + j);
            }
            if (tid == 0) *s_nhs = total;
        }

        // ---- phase B: counts.  k-mer cells first, then one subset-sum pass per low position ----
        for (uint32_t kl = tid; kl < tk; kl += KP_NT) {
            uint32_t cell = 0;
            for (int e = 0; e < nlow; e++) cell += ((kl / tb.lowkw[e]) % tb.nbase[e]) * tb.loww[e];
            size_t g = (size_t)tile * tk + kl;
            if (!CV) {
                kp_stcnt<WIDE>(slot, stride, 0, cell, (unsigned long long)p.e0[g]);
                kp_stcnt<WIDE>(slot, stride, 1, cell, (unsigned long long)p.e1[g]);
            } else {
                long long mt = p.e2[g], ut = p.e3[g];
                kp_stcnt<WIDE>(slot, stride, 0, cell, (unsigned long long)(p.e0[g] - mt));  // train = total - held-out
                kp_stcnt<WIDE>(slot, stride, 1, cell, (unsigned long long)(p.e1[g] - ut));
                kp_stcnt<WIDE>(slot, stride, 2, cell, (unsigned long long)mt);
                kp_stcnt<WIDE>(slot, stride, 3, cell, (unsigned long long)ut);
            }
        }
        __syncthreads();
        for (int e = 0; e < nlow; e++) {
            const uint32_t sh = tb.shift[e], fm = tb.fmask[e], lw = tb.loww[e];
            for (uint32_t i = tid; i < cells; i += KP_NT) {
                uint32_t pk = cl[i];
                if ((int)(pk >> 28) != e + 1) continue;  // highest multi-letter low position of this cell
                uint32_t cell = (pk >> 16) & 0xFFFu;
                int d = (int)((pk >> sh) & fm);
                uint32_t m = tb.digit_mask[e][d];
                unsigned long long acc[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) acc[c] = 0;
                for (int b = 0; b < 4; b++) {
                    if (!((m >> b) & 1u)) continue;
                    uint32_t src = cell - (uint32_t)(d - (int)tb.mask_digit[e][1u << b]) * lw;
#pragma unroll
                    for (int c = 0; c < NC; c++) acc[c] += kp_ldcnt<WIDE>(slot, stride, c, src);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) kp_stcnt<WIDE>(slot, stride, c, cell, acc[c]);
            }
            __syncthreads();
        }

        // ---- phase C: float64 self-score per cell; overwrites the cell's own counts ----
        {
            const uint32_t nleaf = p.leaf_wave ? mlo[1] : 0u;
            float *tsl = (float *)(slot + (size_t)(WIDE ? 2 : 1) * stride);
            for (uint32_t i = tid; i < cells; i += KP_NT) {
                uint32_t cell = (cl[i] >> 16) & 0xFFFu;
                unsigned long long M = kp_ldcnt<WIDE>(slot, stride, 0, cell), U = kp_ldcnt<WIDE>(slot, stride, 1, cell);
                double s;
                if (!CV) {
                    if (i < nleaf) {
                        s = (double)__double2float_rn(kp_leaf_score(M, U, alpha, beta, penalty, logtab));
                    } else {
                        double lp, l1;
                        s = kp_self_score(M, U, alpha, beta, penalty, logtab, lp, l1);
                    }
                } else {
                    unsigned long long Mt = kp_ldcnt<WIDE>(slot, stride, 2, cell), Ut = kp_ldcnt<WIDE>(slot, stride, 3, cell);
                    double t;
                    if (i < nleaf) {
                        kp_leaf_cv(M, U, Mt, Ut, alpha, beta, penalty, logtab, s, t);
                        s = (double)__double2float_rn(s);
                    } else {
                        double lp, l1;
                        s = kp_self_score(M, U, alpha, beta, penalty, logtab, lp, l1);
                        t = kp_test_ll(Mt, Ut, lp, l1);
                    }
                    tsl[2 * (size_t)cell] = __double2float_rn(t);
                }
                ((double *)slot)[cell] = s;
            }
        }
        __syncthreads();  // hs list (phase A) visible; also orders phase C before phase E

        // ---- phase D: stream the child tiles of the high-position splits ----
        const int nhs = *s_nhs;
        if (!CV) {
            const int nq = (int)((cells + 3) >> 2);
            float4 hv[KP_QPT];
            uint32_t hr[KP_QPT];
#pragma unroll
            for (int c = 0; c < KP_QPT; c++) { hv[c] = make_float4(INF, INF, INF, INF); hr[c] = 0xFFFFFFFFu; }
            for (int s = 0; s < nhs; s++) {
                const float4 *a = (const float4 *)(p.best + (size_t)hs1[s] * stride);
                const float4 *b = (const float4 *)(p.best + (size_t)hs2[s] * stride);
                const uint32_t rk = hsr[s];
#pragma unroll
                for (int c = 0; c < KP_QPT; c++) {
                    int q = tid + c * KP_NT;
                    if (q < nq) {
                        float4 x = __ldg(a + q), y = __ldg(b + q);
                        float v0 = __fadd_rn(x.x, y.x), v1 = __fadd_rn(x.y, y.y), v2 = __fadd_rn(x.z, y.z), v3 = __fadd_rn(x.w, y.w);
                        if (v0 < hv[c].x) { hv[c].x = v0; hr[c] = (hr[c] & 0xFFFFFF00u) | rk; }
                        if (v1 < hv[c].y) { hv[c].y = v1; hr[c] = (hr[c] & 0xFFFF00FFu) | (rk << 8); }
                        if (v2 < hv[c].z) { hv[c].z = v2; hr[c] = (hr[c] & 0xFF00FFFFu) | (rk << 16); }
                        if (v3 < hv[c].w) { hv[c].w = v3; hr[c] = (hr[c] & 0x00FFFFFFu) | (rk << 24); }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < KP_QPT; c++) {
                int q = tid + c * KP_NT;
                if (q < nq) { ((float4 *)S)[q] = hv[c]; ((uint32_t *)R)[q] = hr[c]; }
            }
        } else {
            const int nq = (int)((cells + 1) >> 1);  // 16 bytes = 2 cells x (train, test)
            for (int q0 = tid; q0 < nq; q0 += KP_NT * KP_QPT) {
                float2 hv[KP_QPT], ht[KP_QPT];
#pragma unroll
                for (int c = 0; c < KP_QPT; c++) { hv[c] = make_float2(INF, INF); ht[c] = make_float2(0.f, 0.f); }
                for (int s = 0; s < nhs; s++) {
                    const float4 *a = (const float4 *)(p.tt + (size_t)hs1[s] * stride * 2);
                    const float4 *b = (const float4 *)(p.tt + (size_t)hs2[s] * stride * 2);
#pragma unroll
                    for (int c = 0; c < KP_QPT; c++) {
                        int q = q0 + c * KP_NT;
                        if (q < nq) {
                            float4 x = __ldg(a + q), y = __ldg(b + q);
                            float v0 = __fadd_rn(x.x, y.x), v1 = __fadd_rn(x.z, y.z);
                            if (v0 < hv[c].x) { hv[c].x = v0; ht[c].x = __fadd_rn(x.y, y.y); }
                            if (v1 < hv[c].y) { hv[c].y = v1; ht[c].y = __fadd_rn(x.w, y.w); }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < KP_QPT; c++) {
                    int q = q0 + c * KP_NT;
                    if (q < nq) { ((float2 *)S)[q] = hv[c]; ((float2 *)T)[q] = ht[c]; }
                }
            }
        }
        __syncthreads();

        // ---- phase E: wavefront over the tile's own levels ----
        for (int ml = 0; ml < nml; ml++) {
            const uint32_t lo = mlo[ml], hi = mlo[ml + 1];
            for (uint32_t i = lo + tid; i < hi; i += KP_NT) {
                const uint32_t pk = cl[i];
                const uint32_t cell = (pk >> 16) & 0xFFFu;
                float lv = INF;
                uint32_t lr = 0xFFu;
                int o1 = 0, o2 = 0;
                for (int e = 0; e < nlow; e++) {
                    const uint32_t lm = lowmeta[e];
                    int d = (int)((pk >> (lm & 0xFFu)) & ((lm >> 8) & 0xFFu));
                    int ns = lowns[e * 16 + d];
                    const int *dl = lowd + e * 128 + d * 8;
                    uint32_t rb = lm >> 16;
                    for (int j = 0; j < ns; j++) {
                        int w = dl[j];
                        int a1 = (int)cell + (int)(short)(w & 0xFFFF), a2 = (int)cell + (w >> 16);
                        float v = __fadd_rn(S[a1], S[a2]);
                        if (v < lv) { lv = v; lr = rb + j; o1 = a1; o2 = a2; }
                    }
                }
                const float hv = S[cell];
                const double s = ((const double *)slot)[cell];
                if (!CV) {
                    if (hv < lv) { lv = hv; lr = R[cell]; }       // low positions scan first: they keep ties
                    if (s < (double)lv) { lv = __double2float_rn(s); lr = 0xFFu; }
                    S[cell] = lv;
                    R[cell] = (uint8_t)lr;
                } else {
                    float tv;
                    if (hv < lv) { lv = hv; tv = T[cell]; }
                    else tv = __fadd_rn(T[o1], T[o2]);
                    if (s < (double)lv) {
                        lv = __double2float_rn(s);
                        tv = ((const float *)(slot + (size_t)(WIDE ? 2 : 1) * stride))[2 * (size_t)cell];
                    }
                    S[cell] = lv;
                    T[cell] = tv;
                }
            }
            __syncthreads();
        }

        // ---- phase F: write the tile ----
        if (!CV) {
            const int nq = (int)((cells + 3) >> 2);
            float4 *ob = (float4 *)(p.best + (size_t)tile * stride);
            uint32_t *os = (uint32_t *)(p.split + (size_t)tile * stride);
            for (int q = tid; q < nq; q += KP_NT) { ob[q] = ((const float4 *)S)[q]; os[q] = ((const uint32_t *)R)[q]; }
        } else {
            const int nq = (int)((cells + 1) >> 1);
            float4 *ot = (float4 *)(p.tt + (size_t)tile * stride * 2);
            for (int q = tid; q < nq; q += KP_NT) {
                float2 a = ((const float2 *)S)[q], b = ((const float2 *)T)[q];
                ot[q] = make_float4(a.x, b.x, a.y, b.y);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// K1: scatter packed k-mers into the dense k-mer tables (duplicates add, like read_dict)
// ---------------------------------------------------------------------------------------------------
__global__ void kp_pack_kernel(const KpTables *tab, const uint8_t *gen_mask, int k, const unsigned long long *codes,
                               const long long *pos, const long long *neg, unsigned long long n, long long *kmerM,
                               long long *kmerU, int *err)
{
    const KpTables &tb = *tab;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long code = codes[i];
        unsigned long long kidx = 0, kw = 1;
        int e = 0;
        bool ok = true;
        for (int s = 0; s < k; s++) {
            unsigned m = (unsigned)((code >> (4 * s)) & 15u);
            unsigned g = gen_mask[s];
            if (m == 0 || (m & (m - 1)) || !(m & g)) { ok = false; break; }
            if (g & (g - 1)) {  // multi-letter position: carries a digit
                kidx += (unsigned long long)tb.mask_digit[e][m] * kw;
                kw *= tb.nbase[e];
                e++;
            }
        }
        if (!ok || (k < 16 && (code >> (4 * k)) != 0)) { atomicExch(err, 1); continue; }
        atomicAdd((unsigned long long *)&kmerM[kidx], (unsigned long long)pos[i]);
        atomicAdd((unsigned long long *)&kmerU[kidx], (unsigned long long)neg[i]);
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: expanded counts E[tile][low k-mer] = sum over the high k-mers covered by the tile's high digits
// ---------------------------------------------------------------------------------------------------
__global__ void kp_expand_base_kernel(const KpTables *tab, unsigned long long nkmer, const long long *kmerM,
                                      const long long *kmerU, long long *expM, long long *expU)
{
    const KpTables &tb = *tab;
    const uint32_t tk = tb.tile_kmers;
    for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < nkmer;
         x += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long kl = x % tk, kh = x / tk, tile = 0;
        for (int e = tb.nlow; e < tb.npos; e++) {
            tile += (kh % tb.nbase[e]) * tb.highw[e];
            kh /= tb.nbase[e];
        }
        expM[tile * tk + kl] = kmerM[x];
        expU[tile * tk + kl] = kmerU[x];
    }
}

// pass over high position e: tiles whose digit at e is multi-letter and whose higher digits are single
__global__ void kp_expand_pass_kernel(const KpTables *tab, int e, long long *expM, long long *expU)
{
    const KpTables &tb = *tab;
    const uint32_t tk = tb.tile_kmers;
    const unsigned long long total = (unsigned long long)tb.ntiles * tk;
    const uint32_t hw = tb.highw[e], rad = tb.radix[e], nb = tb.nbase[e];
    for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < total;
         x += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t tile = (uint32_t)(x / tk), kl = (uint32_t)(x % tk);
        uint32_t rest = tile / hw;
        uint32_t d = rest % rad;
        if (d < nb) continue;
        rest /= rad;
        bool ok = true;
        for (int f = e + 1; f < tb.npos; f++) {
            if (rest % tb.radix[f] >= tb.nbase[f]) { ok = false; break; }
            rest /= tb.radix[f];
        }
        if (!ok) continue;
        uint32_t m = tb.digit_mask[e][d];
        long long am = 0, au = 0;
        for (int b = 0; b < 4; b++) {
            if (!((m >> b) & 1u)) continue;
            size_t src = (size_t)(tile - (d - tb.mask_digit[e][1u << b]) * hw) * tk + kl;
            am += expM[src];
            au += expU[src];
        }
        expM[x] = am;
        expU[x] = au;
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: backtrack.  Breadth-first expansion from the general pattern; each leaf carries its path key
// (0 = c1 side, 1 = c2 side, most significant bit first), so sorting by key restores the reference's
// depth-first, c1-first emission order.
// ---------------------------------------------------------------------------------------------------
struct KpBtNode { unsigned long long pat, key; };

__global__ void __launch_bounds__(256) kp_backtrack_kernel(const KpTables *tab, const uint8_t *split,
                                                           unsigned long long top, KpBtNode *fa, KpBtNode *fb,
                                                           KpBtNode *leaves, unsigned long long cap,
                                                           unsigned long long *out_counts /* [0]=nleaves [1]=overflow */)
{
    const KpTables &tb = *tab;
    __shared__ unsigned long long s_ncur, s_nnext, s_nleaf;
    __shared__ int s_over;
    if (threadIdx.x == 0) { s_ncur = 1; s_nnext = 0; s_nleaf = 0; s_over = 0; fa[0].pat = top; fa[0].key = 0; }
    __syncthreads();
    KpBtNode *cur = fa, *nxt = fb;
    for (int depth = 0; depth < 64; depth++) {
        unsigned long long ncur = s_ncur;
        if (ncur == 0) break;
        for (unsigned long long i = threadIdx.x; i < ncur; i += blockDim.x) {
            KpBtNode nd = cur[i];
            unsigned long long tile = nd.pat / tb.tile_cells;
            uint32_t cell = (uint32_t)(nd.pat % tb.tile_cells);
            uint8_t code = split[tile * tb.tile_stride + cell];
            if (code == 0xFF) {
                unsigned long long li = atomicAdd(&s_nleaf, 1ULL);
                if (li < cap) leaves[li] = nd; else s_over = 1;
                continue;
            }
            int pos = code >> 3, j = code & 7, e = 0;
            for (int f = 0; f < tb.npos; f++) if (tb.pos_id[f] == pos) e = f;
            unsigned long long w = tb.extw[e];
            int d = (int)((nd.pat / w) % tb.radix[e]);
            uint32_t m = tb.digit_mask[e][d];
            int c1 = tb.mask_digit[e][tb.ms_c1[m][j]], c2 = tb.mask_digit[e][tb.ms_c2[m][j]];
            unsigned long long ni = atomicAdd(&s_nnext, 2ULL);
            if (ni + 2 > cap) { s_over = 1; continue; }
            nxt[ni].pat = nd.pat - (unsigned long long)(d - c1) * w;
            nxt[ni].key = nd.key;
            nxt[ni + 1].pat = nd.pat - (unsigned long long)(d - c2) * w;
            nxt[ni + 1].key = nd.key | (1ULL << (63 - depth));
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_ncur = s_over ? 0 : s_nnext; s_nnext = 0; }
        __syncthreads();
        KpBtNode *t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0) { out_counts[0] = s_nleaf; out_counts[1] = (unsigned long long)(s_over || s_ncur != 0); }
}

// rank sort by key (keys are distinct): out[rank] = pat
__global__ void kp_backtrack_sort_kernel(const KpBtNode *leaves, const unsigned long long *counts,
                                         unsigned long long cap, unsigned long long *out)
{
    unsigned long long n = counts[0];
    if (n > cap) n = cap;
    __shared__ unsigned long long keys[256];
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long mykey = i < n ? leaves[i].key : 0, rank = 0;
    for (unsigned long long base = 0; base < n; base += 256) {
        unsigned long long j = base + threadIdx.x;
        keys[threadIdx.x] = j < n ? leaves[j].key : ~0ULL;
        __syncthreads();
        unsigned long long lim = n - base < 256 ? n - base : 256;
        for (unsigned long long t = 0; t < lim; t++) rank += keys[t] < mykey;
        __syncthreads();
    }
    if (i < n) out[rank] = leaves[i].pat;
}

// ---------------------------------------------------------------------------------------------------
// counts of arbitrary patterns from the k-mer tables (one CTA per pattern)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kp_pattern_counts_kernel(const KpTables *tab, unsigned long long nkmer,
                                                                const long long *kmerM, const long long *kmerU,
                                                                const unsigned long long *patnums, long long *outM,
                                                                long long *outU)
{
    const KpTables &tb = *tab;
    __shared__ uint32_t masks[KP_MAXPOS];
    __shared__ long long redM[256], redU[256];
    unsigned long long pat = patnums[blockIdx.x];
    if (threadIdx.x == 0) {
        unsigned long long x = pat;
        for (int e = 0; e < tb.npos; e++) { masks[e] = tb.digit_mask[e][x % tb.radix[e]]; x /= tb.radix[e]; }
    }
    __syncthreads();
    long long am = 0, au = 0;
    for (unsigned long long x = threadIdx.x; x < nkmer; x += blockDim.x) {
        unsigned long long r = x;
        bool in = true;
        for (int e = 0; e < tb.npos; e++) {
            uint32_t b = (uint32_t)(r % tb.nbase[e]);
            r /= tb.nbase[e];
            if (!(masks[e] & tb.digit_mask[e][b])) { in = false; break; }
        }
        if (in) { am += kmerM[x]; au += kmerU[x]; }
    }
    redM[threadIdx.x] = am; redU[threadIdx.x] = au;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { redM[threadIdx.x] += redM[threadIdx.x + s]; redU[threadIdx.x] += redU[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { outM[blockIdx.x] = redM[0]; outU[blockIdx.x] = redU[0]; }
}

// ---------------------------------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------------------------------
__global__ void kp_debug_log_kernel(const double *x, double *y, unsigned long long n)
{
    __shared__ double2 tab[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(kpc_logTab[2 * i], kpc_logTab[2 * i + 1]);
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        y[i] = kp_log(x[i], tab);
}

__global__ void kp_debug_leaf_kernel(const long long *M, const long long *U, unsigned long long n, double alpha,
                                     double beta, double penalty, double *out)
{
    __shared__ double2 tab[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(kpc_logTab[2 * i], kpc_logTab[2 * i + 1]);
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = kp_leaf_score((unsigned long long)M[i], (unsigned long long)U[i], alpha, beta, penalty, tab);
}
