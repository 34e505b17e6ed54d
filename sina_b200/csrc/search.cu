// K-mer search, top-k ranking and family selection on the device.
//   find_tile_kernel   kmer_search::impl::find, counting part  (reference src/kmer_search.cpp:389-409)
//   find_merge_kernel  partial_sort by greater<pair<int16,int>> (:412): score desc, then id desc
//   family_kernel      famfinder::impl::match + gap filter + fs_req (src/famfinder.cpp:497-612, 474-491)
#include "common.cuh"

namespace sg {

// ---------------------------------------------------------------------------------------------------
// Valid k-mers of every query (one warp per query), compacted: ambiguous windows dropped, fast mode keeps
// A-prefixed ones, the last k-mer is never produced (src/kmer.h:69-78,110-125,179-201). A query k-mer occurring
// twice is listed twice (all_kmers / prefix_kmers, not the unique_ variants, src/kmer_search.cpp:391,397).
__global__ void __launch_bounds__(128) query_kmers_kernel(const uint8_t* __restrict__ qmasks,
                                                          const uint64_t* __restrict__ qoff, uint32_t nq, int k,
                                                          int nofast, uint32_t* __restrict__ kmers,
                                                          uint32_t* __restrict__ nk) {
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (q >= nq) return;
    const uint8_t* m = qmasks + qoff[q];
    const uint32_t n = (uint32_t)(qoff[q + 1] - qoff[q]);
    uint32_t* out = kmers + qoff[q];
    const uint32_t lane = lane_id();
    uint32_t cnt = 0;
    if (n > (uint32_t)k) {
        for (uint32_t i0 = (uint32_t)k - 1; i0 + 1 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t v = 0;
            bool ok = i + 1 < n && kmer_at(m, i, k, v);
            ok = ok && (nofast || (v >> (2 * (k - 1))) == 0);
            const uint32_t b = __ballot_sync(0xffffffffu, ok);
            if (ok) out[cnt + __popc(b & ((1u << lane) - 1u))] = v;
            cnt += __popc(b);
        }
    }
    if (lane == 0) nk[q] = cnt;
}

// ---------------------------------------------------------------------------------------------------
// Counting part of find(): one CTA per (query, tile); a tile is up to 24 sub-tiles and every warp OWNS the
// u16 score counters of one sub-tile in shared memory. The warp walks the query's k-mers and, for each, streams
// the k-mer's posting list for its own sub-tile: ids inside one list are distinct, so the 32 lanes of one load
// update 32 different counters with a plain LDS / add / STS and no atomic is needed (shared-memory atomics run
// at 2 cycles per lane and were the bound of the first version); lists are applied one after the other, in
// program order. The first 32 postings of FIND_G lists are requested before any of them is applied.
// Selection: the tile's top-`need` in rank order (score desc, id desc) are those above a threshold score T plus
// the highest ids among the ties at T. T is found on a histogram of the high scores only ((M/2, M], then
// (M/4, M/2], ... below the tile's maximum M), so the many low counters cost one compare each.
constexpr int FIND_G = 8;
constexpr uint32_t SEL_BINS = 1024;    // widest score window histogrammed at once
constexpr uint32_t TIE_CAP = 1024;     // ties at the threshold ranked in shared memory (more: id-ordered walk)

struct FindArgs {
    const uint32_t* kmers; const uint32_t* nk; const uint64_t* qoff;
    uint32_t N, sub_size, n_sub, tile_warps;
    const uint32_t* list_off; const uint16_t* postings;
    uint32_t max; uint64_t* cand; uint32_t* cand_n; unsigned long long* counters;
    uint16_t* scores_out;   // non-null: write the tile's score counters to scores_out[q][N] instead of selecting (full ranking)
};

__global__ void __launch_bounds__(32 * TILE_WARPS_MAX) find_tile_kernel(FindArgs A) {
    extern __shared__ uint32_t hist32[];  // tile_warps * sub_size u16 counters
    __shared__ uint32_t hist2[SEL_BINS];
    __shared__ uint32_t tie[TIE_CAP];
    __shared__ uint32_t red[33];
    __shared__ uint32_t sh_sel[8];      // 0: T, 1: count_gt, 2: need_eq, 3: emitted, 4: found, 5: ties gathered, 6: count_eq
    uint16_t* hist = reinterpret_cast<uint16_t*>(hist32);
    const uint32_t q = blockIdx.x, tile = blockIdx.y, n_tiles = gridDim.y;
    const uint32_t B = A.sub_size;
    const uint32_t tile_lo = tile * A.tile_warps * B;
    const uint32_t tile_n = min(A.tile_warps * B, A.N - tile_lo);
    const uint32_t words = (tile_n + 1) >> 1;
    const uint32_t tid = threadIdx.x, nt = blockDim.x, lane = lane_id(), w = warp_id();
    for (uint32_t i = tid; i < words; i += nt) hist32[i] = 0;
    __syncthreads();

    // ---- counting
    const uint32_t sub = tile * A.tile_warps + w;
    uint32_t lmax = 0;   // highest counter value this lane wrote: the tile maximum needs no pass of its own
    if (sub < A.n_sub) {
        uint16_t* hw = hist + (size_t)w * B;
        const uint32_t* kl = A.kmers + A.qoff[q];
        const uint32_t nk = A.nk[q];
        const uint16_t* __restrict__ post = A.postings;
        unsigned long long my_post = 0;
        for (uint32_t base = 0; base < nk; base += 32) {
            uint32_t a = 0, len = 0;
            if (base + lane < nk) {
                const uint32_t* o = A.list_off + (uint64_t)kl[base + lane] * A.n_sub + sub;
                a = __ldg(o);
                len = __ldg(o + 1) - a;
            }
            my_post += len;
            if (!__any_sync(0xffffffffu, len != 0)) continue;
            for (uint32_t i0 = 0; i0 < 32; i0 += FIND_G) {
                uint32_t al[FIND_G], ll[FIND_G], x[FIND_G];
#pragma unroll
                for (int g = 0; g < FIND_G; g++) {
                    al[g] = __shfl_sync(0xffffffffu, a, i0 + g);
                    ll[g] = __shfl_sync(0xffffffffu, len, i0 + g);
                    x[g] = lane < ll[g] ? (uint32_t)__ldg(post + al[g] + lane) : 0u;
                }
#pragma unroll
                for (int g = 0; g < FIND_G; g++) {
                    if (ll[g] == 0) continue;                       // warp-uniform
                    if (lane < ll[g]) { const uint32_t nv = hw[x[g]] + 1u; hw[x[g]] = (uint16_t)nv; lmax = max(lmax, nv); }
                    __syncwarp();
                    for (uint32_t e0 = 32; e0 < ll[g]; e0 += 128) {  // long lists: four more loads at a time
                        uint32_t y[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t e = e0 + 32 * u + lane;
                            y[u] = e < ll[g] ? (uint32_t)__ldg(post + al[g] + e) : 0xffffffffu;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (y[u] != 0xffffffffu) { const uint32_t nv = hw[y[u]] + 1u; hw[y[u]] = (uint16_t)nv; lmax = max(lmax, nv); }
                        __syncwarp();
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) my_post += __shfl_xor_sync(0xffffffffu, my_post, o);
        if (lane == 0 && my_post) atomicAdd(&A.counters[0], my_post);
    }
    __syncthreads();
    if (A.scores_out) {   // full ranking (rank_full_kernel): hand the whole score vector over
        uint16_t* dst = A.scores_out + (uint64_t)q * A.N + tile_lo;
        for (uint32_t i = tid; i < tile_n; i += nt) dst[i] = hist[i];
        return;
    }

    // ---- selection
    const uint32_t need = min(A.max, tile_n);
    uint64_t* out = A.cand + ((uint64_t)q * n_tiles + tile) * A.max;
    auto score_of = [&](uint32_t i) -> uint32_t { return hist[i]; };
    // tile maximum
    uint32_t mx = lmax;   // tracked while counting (a counter's last write is its final value)
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[w] = mx;
    if (tid == 0) { sh_sel[3] = 0; sh_sel[4] = 0; sh_sel[5] = 0; }
    __syncthreads();
    if (tid == 0) { uint32_t m2 = 0; for (uint32_t i = 0; i < (nt >> 5); i++) m2 = max(m2, red[i]); red[32] = m2; }
    __syncthreads();
    const uint32_t M = red[32];
    uint32_t T = 0, count_gt = 0, need_eq = 0, count_eq = 0;
    {
        // windows (lo, hi] of scores, highest first; entries above the current window are already counted in cum
        uint32_t hi = M, cum = 0;
        for (;;) {
            uint32_t lo = hi / 2;
            if (hi - lo > SEL_BINS) lo = hi - SEL_BINS;
            for (uint32_t i = tid; i < SEL_BINS; i += nt) hist2[i] = 0;
            __syncthreads();
            if (hi > 0) {
                for (uint32_t i = tid; i < words; i += nt) {
                    const uint32_t v = hist32[i];
                    const uint32_t s0 = v & 0xffffu, s1 = v >> 16;
                    if (s0 > lo && s0 <= hi) atomicAdd(&hist2[s0 - lo - 1], 1u);
                    if (s1 > lo && s1 <= hi && 2 * i + 1 < tile_n) atomicAdd(&hist2[s1 - lo - 1], 1u);
                }
            }
            __syncthreads();
            if (w == 0) {   // lane L owns bins [32L, 32L+32); find the highest bin where cum + suffix count >= need
                uint32_t part = 0;
                for (uint32_t b = 0; b < 32; b++) part += hist2[32 * lane + b];
                // above = entries in the bins of higher lanes, suffix = entries in the whole window
                uint32_t above = 0, suffix = 0;
                for (int l = 31; l >= 0; l--) {
                    const uint32_t pl = __shfl_sync(0xffffffffu, part, l);
                    if ((int)lane == l) above = suffix;
                    suffix += pl;
                }
                const bool mine = cum + above < need && cum + above + part >= need;
                const uint32_t who = __ballot_sync(0xffffffffu, mine);
                if (who) {
                    if (mine) {
                        uint32_t c2 = cum + above;
                        for (int b = 31; b >= 0; b--) {
                            const uint32_t hb = hist2[32 * lane + b];
                            if (c2 + hb >= need) { sh_sel[0] = lo + 1 + 32 * lane + b; sh_sel[1] = c2; sh_sel[2] = need - c2; sh_sel[6] = hb; break; }
                            c2 += hb;
                        }
                        sh_sel[4] = 1;
                    }
                } else if (lane == 0) {
                    sh_sel[1] = cum + suffix;   // everything in this window ranks above the threshold
                }
            }
            __syncthreads();
            if (sh_sel[4]) { T = sh_sel[0]; count_gt = sh_sel[1]; need_eq = sh_sel[2]; count_eq = sh_sel[6]; break; }
            cum = sh_sel[1];
            if (lo == 0) { T = 0; count_gt = cum; need_eq = need - cum; count_eq = tile_n - cum; break; }  // ties at score 0
            hi = lo;
            __syncthreads();
        }
    }
    // entries above the threshold, and the ties at it
    const bool rank_ties = T > 0 && count_eq <= TIE_CAP;
    for (uint32_t i = tid; i < words; i += nt) {
        const uint32_t v = hist32[i];
        const uint32_t s0 = v & 0xffffu, s1 = v >> 16;
        if (s0 > T) out[atomicAdd(&sh_sel[3], 1u)] = ((uint64_t)s0 << 32) | (tile_lo + 2 * i);
        if (s1 > T && 2 * i + 1 < tile_n) out[atomicAdd(&sh_sel[3], 1u)] = ((uint64_t)s1 << 32) | (tile_lo + 2 * i + 1);
        if (rank_ties) {
            if (s0 == T) tie[atomicAdd(&sh_sel[5], 1u)] = 2 * i;
            if (s1 == T && 2 * i + 1 < tile_n) tie[atomicAdd(&sh_sel[5], 1u)] = 2 * i + 1;
        }
    }
    __syncthreads();
    if (rank_ties) {   // highest ids first: rank = number of tied ids above this one
        const uint32_t ne = sh_sel[5];
        for (uint32_t i = tid; i < ne; i += nt) {
            const uint32_t id = tie[i];
            uint32_t rk = 0;
            for (uint32_t j = 0; j < ne; j++) rk += tie[j] > id ? 1u : 0u;
            if (rk < need_eq) out[count_gt + rk] = ((uint64_t)T << 32) | (tile_lo + id);
        }
    } else {
        // walk ids downwards in chunks of blockDim, thread 0 <-> highest id of the chunk
        uint32_t remaining = need_eq, emitted_eq = 0;
        for (uint32_t top = tile_n; top > 0 && remaining > 0;) {
            uint32_t chunk = min(top, nt);
            uint32_t flag = 0, i = 0;
            if (tid < chunk) { i = top - 1 - tid; flag = score_of(i) == T; }
            uint32_t tot, ex = block_exscan(flag, red, &tot);
            if (flag && ex < remaining) out[count_gt + emitted_eq + ex] = ((uint64_t)T << 32) | (tile_lo + i);
            uint32_t took = min(tot, remaining);
            emitted_eq += took; remaining -= took;
            top -= chunk;
        }
    }
    if (tid == 0) A.cand_n[q * n_tiles + tile] = need;
}

// One CTA per query: gather the tiles' candidates, bitonic-sort the 64-bit keys descending in shared
// memory, keep the first `max`.
__global__ void __launch_bounds__(1024) find_merge_kernel(const uint64_t* __restrict__ cand,
                                                           const uint32_t* __restrict__ cand_n, uint32_t n_tiles,
                                                           uint32_t max, uint32_t N, uint32_t p2,
                                                           uint64_t* __restrict__ ranked, uint32_t* __restrict__ nres) {
    extern __shared__ uint64_t keys[];
    const uint32_t q = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0;
    __syncthreads();
    // compact tile lists back to back (deterministic positions: prefix over cand_n, which is tiny)
    uint32_t base = 0;
    for (uint32_t t = 0; t < n_tiles; t++) {
        uint32_t c = cand_n[q * n_tiles + t];
        const uint64_t* src = cand + ((uint64_t)q * n_tiles + t) * max;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) keys[base + i] = src[i] + 1;  // +1: real keys > padding
        base += c;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= p2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < p2 / 2; i += blockDim.x) {
                uint32_t lo = 2 * i - (i & (stride - 1));  // index with bit `stride` cleared
                uint32_t hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    const uint32_t r = min(max, N);
    for (uint32_t i = threadIdx.x; i < r; i += blockDim.x) ranked[(uint64_t)q * max + i] = keys[i] - 1;
    if (threadIdx.x == 0) nres[q] = r;
}

int launch_find(Session* s, uint32_t max, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    if (n == 0) { q0 = 0; n = s->nq; }
    if (max == 0) SG_FAIL(SG_ERR_ARG, "find: max must be > 0");
    if (max > ix->N) max = ix->N;
    uint32_t p2 = 1;
    while (p2 < (uint64_t)max * ix->n_tiles) p2 <<= 1;
    if (p2 > FIND_MAX_SORT) SG_FAIL(SG_ERR_LIMIT, "find: max * tiles exceeds the top-k merge capacity (16384)");
    if (max > s->find_cap) {
        if (s->d_cand) cudaFree(s->d_cand);
        if (s->d_ranked) cudaFree(s->d_ranked);
        s->d_cand = nullptr; s->d_ranked = nullptr;
        SG_CUDA(cudaMalloc(&s->d_cand, (uint64_t)s->max_q * ix->n_tiles * max * sizeof(uint64_t)));
        SG_CUDA(cudaMalloc(&s->d_ranked, (uint64_t)s->max_q * max * sizeof(uint64_t)));
        s->find_cap = max;
    }
    s->find_max = max;
    const size_t smem = (size_t)ix->tile_warps * ix->sub_size * 2;
    SG_CUDA(cudaFuncSetAttribute(find_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SG_CUDA(cudaFuncSetAttribute(find_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(p2 * 8)));
    query_kmers_kernel<<<(n + 3) / 4, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff + q0, n, ix->k, ix->nofast,
                                                          s->d_kmers, s->d_nk + q0);
    FindArgs A;
    A.kmers = s->d_kmers; A.nk = s->d_nk + q0; A.qoff = s->d_qoff + q0; A.N = ix->N; A.sub_size = ix->sub_size; A.n_sub = ix->n_sub;
    A.tile_warps = ix->tile_warps; A.list_off = ix->d_list_off; A.postings = ix->d_postings; A.max = max; A.scores_out = nullptr;
    A.cand = s->d_cand + (uint64_t)q0 * ix->n_tiles * max; A.cand_n = s->d_cand_n + (uint64_t)q0 * ix->n_tiles; A.counters = s->d_counters;
    dim3 grid(n, ix->n_tiles);
    find_tile_kernel<<<grid, 32 * ix->tile_warps, smem, s->stream>>>(A);
    find_merge_kernel<<<n, p2 / 2 < 1024 ? (p2 / 2 < 32 ? 32 : p2 / 2) : 1024, p2 * 8, s->stream>>>(
        A.cand, A.cand_n, ix->n_tiles, max, ix->N, p2, s->d_ranked + (uint64_t)q0 * max, s->d_nres + q0);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 3;
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Full ranking of one query's score vector: every reference in rank order (score desc, id desc), i.e. find() with
// max = N. Used when the family walk needs a window the shared-memory merge cannot hold (the reference widens its
// window x10 until it covers the index, src/famfinder.cpp:591-608; a wider window never changes the result, so the
// walk goes straight to the whole index). One CTA per query: a stable LSD radix sort of the 16-bit scores (two 8-bit
// passes, descending digits) over the ids taken in DESCENDING order, so that equal scores keep the higher id first.
constexpr int RF_THREADS = 1024;
__global__ void __launch_bounds__(RF_THREADS) rank_full_kernel(const uint16_t* __restrict__ scores, uint32_t N,
                                                               uint64_t* __restrict__ tmp, uint64_t* __restrict__ out,
                                                               uint32_t* __restrict__ nres) {
    __shared__ uint32_t base[256];               // first output position of a digit
    __shared__ uint32_t wcnt[RF_THREADS / 32][256];   // per chunk: elements of a digit per warp -> exclusive offsets
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const uint16_t* sc = scores + (uint64_t)q * N;
    uint64_t* t0 = tmp + (uint64_t)q * N;
    uint64_t* o0 = out + (uint64_t)q * N;
    for (int pass = 0; pass < 2; pass++) {
        const uint32_t shift = 8 * pass;
        auto load = [&](uint32_t i) -> uint64_t {   // element i of the pass's input sequence
            if (pass == 0) { const uint32_t id = N - 1 - i; return ((uint64_t)sc[id] << 32) | id; }
            return t0[i];
        };
        auto digit = [&](uint64_t key) -> uint32_t { return 255u - ((uint32_t)(key >> (32 + shift)) & 255u); };
        if (tid < 256) base[tid] = 0;
        __syncthreads();
        for (uint32_t c0 = 0; c0 < N; c0 += RF_THREADS) {   // warp-aggregated: the scores crowd into a few digits
            const uint32_t i = c0 + tid;
            const uint32_t d = i < N ? digit(load(i)) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (i < N && (uint32_t)__ffs((int)peers) - 1u == lane) atomicAdd(&base[d], (uint32_t)__popc(peers));
        }
        __syncthreads();
        if (w == 0) {   // exclusive prefix over the 256 digit counts
            uint32_t carry = 0;
            for (uint32_t d0 = 0; d0 < 256; d0 += 32) {
                const uint32_t v = base[d0 + lane];
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
                base[d0 + lane] = carry + x - v;
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
        }
        __syncthreads();
        uint64_t* dst = pass == 0 ? t0 : o0;
        for (uint32_t c0 = 0; c0 < N; c0 += RF_THREADS) {
            for (uint32_t i = tid; i < (RF_THREADS / 32) * 256; i += RF_THREADS) (&wcnt[0][0])[i] = 0;
            __syncthreads();
            const uint32_t i = c0 + tid;
            const bool have = i < N;
            uint64_t key = 0;
            uint32_t d = 0xffffffffu, rank = 0;
            if (have) { key = load(i); d = digit(key); }
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (have && rank == 0) wcnt[w][d] = __popc(peers);
            __syncthreads();
            if (tid < 256) {   // per digit: exclusive offsets of the warps inside this chunk, then the chunk's total
                uint32_t run = 0;
                for (uint32_t ww = 0; ww < RF_THREADS / 32; ww++) { const uint32_t c = wcnt[ww][tid]; wcnt[ww][tid] = run; run += c; }
                const uint32_t b = base[tid];
                base[tid] = b + run;
                for (uint32_t ww = 0; ww < RF_THREADS / 32; ww++) wcnt[ww][tid] += b;
            }
            __syncthreads();
            if (have) dst[wcnt[w][d] + rank] = key;
            __syncthreads();
        }
    }
    if (tid == 0) nres[q] = N;
}

// full score vectors + full ranking of queries [q0, q0 + n) into s->d_full_keys (n <= s->full_cap)
int launch_find_full(Session* s, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    const size_t smem = (size_t)ix->tile_warps * ix->sub_size * 2;
    SG_CUDA(cudaFuncSetAttribute(find_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    query_kmers_kernel<<<(n + 3) / 4, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff + q0, n, ix->k, ix->nofast,
                                                          s->d_kmers, s->d_nk + q0);
    FindArgs A;
    A.kmers = s->d_kmers; A.nk = s->d_nk + q0; A.qoff = s->d_qoff + q0; A.N = ix->N; A.sub_size = ix->sub_size; A.n_sub = ix->n_sub;
    A.tile_warps = ix->tile_warps; A.list_off = ix->d_list_off; A.postings = ix->d_postings; A.max = 1;
    A.cand = nullptr; A.cand_n = nullptr; A.counters = s->d_counters; A.scores_out = s->d_full_scores;
    dim3 grid(n, ix->n_tiles);
    find_tile_kernel<<<grid, 32 * ix->tile_warps, smem, s->stream>>>(A);
    rank_full_kernel<<<n, RF_THREADS, 0, s->stream>>>(s->d_full_scores, ix->N, s->d_full_tmp, s->d_full_keys, s->d_nres + q0);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 3;
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Orientation check (--turn): famfinder::impl::turn_check + do_turn_check (reference src/famfinder.cpp:311-378).
// The reference runs find(max = 1) on the query, its reverse, its complement and its reverse complement and keeps
// the first orientation with the strictly largest top score. Here the query buffer is transformed in place between
// the searches (orig -> reversed -> complemented -> reverse-complemented) and finally brought to the chosen one.
// op bit 0 = reverse (cseq_base::reverse, src/cseq.cpp:284-289), bit 1 = complement (base_iupac::complement,
// src/aligned_base.h:117-124: A<->T/U, G<->C on the IUPAC bits, case kept).
__device__ __forceinline__ uint8_t mask_complement(uint8_t m) {
    return (uint8_t)(((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 1u) << 3) | ((m & 8u) >> 3) | (m & 16u));
}

// one warp per query; ops: per-query op code, or null for the uniform code `op_all`
__global__ void __launch_bounds__(128) orient_kernel(uint8_t* __restrict__ qmasks, const uint64_t* __restrict__ qoff,
                                                     uint32_t nq, const uint8_t* __restrict__ ops, uint32_t op_all) {
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (q >= nq) return;
    const uint32_t op = ops ? ops[q] : op_all;
    if (op == 0) return;
    uint8_t* m = qmasks + qoff[q];
    const uint32_t n = (uint32_t)(qoff[q + 1] - qoff[q]);
    const bool rev = op & 1u, comp = op & 2u;
    if (rev) {
        for (uint32_t i = lane_id(); i < (n + 1) / 2; i += 32) {
            const uint32_t j = n - 1 - i;
            uint8_t a = m[i], b = m[j];
            if (comp) { a = mask_complement(a); b = mask_complement(b); }
            m[i] = b;
            m[j] = a;   // i == j (middle base): written twice with the same value
        }
    } else {
        for (uint32_t i = lane_id(); i < n; i += 32) m[i] = mask_complement(m[i]);
    }
}

// top score of every query after find(max = 1) into column `which` of scores[4][nq]
__global__ void turn_score_kernel(const uint64_t* __restrict__ ranked, const uint32_t* __restrict__ nres, uint32_t max,
                                  uint32_t nq, int32_t* __restrict__ scores, uint32_t which) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    scores[(uint64_t)which * nq + q] = nres[q] ? (int32_t)(ranked[(uint64_t)q * max] >> 32) : 0;
}

// best orientation (strict '>' from 0 in the order none, reversed, complemented, both; src/famfinder.cpp:369-377) and
// the op that takes the buffer from its current state `state` to it
__global__ void turn_pick_kernel(const int32_t* __restrict__ scores, uint32_t nq, uint32_t state,
                                 int32_t* __restrict__ turn, uint8_t* __restrict__ ops) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int32_t mx = 0, best = 0;
    for (int i = 0; i < 4; i++) {
        const int32_t sc = scores[(uint64_t)i * nq + q];
        if (mx < sc) { mx = sc; best = i; }
    }
    turn[q] = best;
    ops[q] = (uint8_t)((uint32_t)best ^ state);   // orientations are the group {1, rev, comp, rev*comp}: codes xor
}

int launch_turn(Session* s, int all) {
    const uint32_t nq = s->nq, gw = (nq + 3) / 4, gt = (nq + 127) / 128;
    SG_CUDA(cudaMemsetAsync(s->d_turn_scores, 0, (uint64_t)4 * nq * 4, s->stream));
    auto search = [&](uint32_t which) -> int {
        SG_TRY(launch_find(s, 1));
        turn_score_kernel<<<gt, 128, 0, s->stream>>>(s->d_ranked, s->d_nres, s->find_max, nq, s->d_turn_scores, which);
        return SG_OK;
    };
    auto orient = [&](uint32_t op) { orient_kernel<<<gw, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff, nq, nullptr, op); };
    uint32_t state = 0;                      // orientation code of the buffer: bit 0 reversed, bit 1 complemented
    SG_TRY(search(0));
    if (all) {
        orient(1); state = 1; SG_TRY(search(1));
        orient(3); state = 2; SG_TRY(search(2));   // reversed -> complemented
        orient(1); state = 3; SG_TRY(search(3));
    } else {
        orient(3); state = 3; SG_TRY(search(3));
    }
    turn_pick_kernel<<<gt, 128, 0, s->stream>>>(s->d_turn_scores, nq, state, s->d_turn, s->d_turn_ops);
    orient_kernel<<<gw, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff, nq, s->d_turn_ops, 0);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += (all ? 4 : 2) * 1 + (all ? 3 : 1) + 2;
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------------
// One thread per query walks its ranked candidates with the reference's quota rules. The outcome only
// depends on earlier items, and once both quotas are met every later item is removed, so scanning a
// window that reaches that point equals the reference's retry loop (:591-608); if the window ends first
// and does not cover the index the query is flagged (-2) and the host re-runs with a 10x window.
__global__ void family_kernel(const uint64_t* __restrict__ ranked, const uint32_t* __restrict__ nres, uint32_t nq,
                              uint32_t window, uint32_t N, const uint64_t* __restrict__ row_off,
                              const uint32_t* __restrict__ cols, const int64_t* __restrict__ excl, sg_fam_params p,
                              uint32_t fam_cap, uint32_t* __restrict__ fam_ids, float* __restrict__ fam_scores,
                              int32_t* __restrict__ fam_n, uint32_t* __restrict__ retry, const float* __restrict__ ident) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t* r = ranked + (uint64_t)q * window;
    const uint32_t w = nres[q];
    uint32_t have = 0, have_full = 0, n = 0;
    const int64_t ex = excl ? excl[q] : -1;
    uint32_t* ids = fam_ids + (uint64_t)q * fam_cap;
    float* scs = fam_scores + (uint64_t)q * fam_cap;
    for (uint32_t i = 0; i < w; i++) {
        const uint32_t id = (uint32_t)r[i];
        const float score = (float)(int16_t)(uint16_t)(r[i] >> 32);  // result_item.score = (float)int16
        const uint32_t len = (uint32_t)(row_off[id + 1] - row_off[id]);
        const bool is_full = len >= p.fs_full_len;
        bool rm = len < p.fs_min_len;                                             // remove_short :537-539
        rm = rm || (p.leave_query_out && ex == (int64_t)id);                     // remove_query :542-544
        rm = rm || (ident && ident[(uint64_t)q * window + i] > p.fs_msc_max);     // remove_similar :553-556 (identity_kernel)
        rm = rm || (have >= p.fs_min && (have >= p.fs_max || !(score < p.fs_msc)) &&   // quota :558-586
                    !(p.fs_req_full && have_full < p.fs_req_full && is_full));
        if (rm) continue;
        have++;                                                                   // count_good :519-531
        if (p.fs_req_full && is_full) have_full++;
        if (n < fam_cap) { ids[n] = id; scs[n] = score; }
        n++;
    }
    if ((have < p.fs_max || have_full < p.fs_req_full) && w < N) {  // window too small: retry (:592,604-607)
        fam_n[q] = -2;
        atomicAdd(retry, 1u);
        return;
    }
    if (n > fam_cap) n = fam_cap;
    if (p.fs_req_gaps != 0) {  // too_few_gaps :474-480
        uint32_t mo = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t id = ids[i];
            uint32_t len = (uint32_t)(row_off[id + 1] - row_off[id]);
            bool few = len == 0 || (cols[row_off[id + 1] - 1] - len + 1 < p.fs_req_gaps);
            if (!few) { ids[mo] = id; scs[mo] = scs[i]; mo++; }
        }
        n = mo;
    }
    fam_n[q] = n < p.fs_req ? -1 : (int32_t)n;  // :486-491
}

int launch_family(Session* s, const sg_fam_params& fp, uint32_t window, uint32_t q0, uint32_t n, const uint64_t* ranked, const float* ident) {
    Index* ix = s->ix;
    if (n == 0) { q0 = 0; n = s->nq; }
    SG_CUDA(cudaMemsetAsync(s->d_retry, 0, sizeof(uint32_t), s->stream));
    family_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(ranked ? ranked : s->d_ranked + (uint64_t)q0 * window, s->d_nres + q0, n, window, ix->N,
                                                         ix->d_row_off, ix->d_cols, s->d_excl + q0, fp, s->fam_cap,
                                                         s->d_fam_ids + (uint64_t)q0 * s->fam_cap,
                                                         s->d_fam_scores + (uint64_t)q0 * s->fam_cap, s->d_fam_n + q0, s->d_retry, ident);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
