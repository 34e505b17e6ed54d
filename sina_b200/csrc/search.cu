// K-mer search, top-k ranking and family selection on the device.
//   find_tile_kernel   kmer_search::impl::find, counting part  (reference src/kmer_search.cpp:389-409)
//   find_merge_kernel  partial_sort by greater<pair<int16,int>> (:412): score desc, then id desc
//   family_kernel      famfinder::impl::match + gap filter + fs_req (src/famfinder.cpp:497-612, 474-491)
#include "common.cuh"
#include <algorithm>

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
// program order.
// A (k-mer, sub-tile) list holds ~9 postings at 500 k references, so the kernel lives or dies by the instructions
// and the latency it spends per LIST, not by bytes (ncu, profiles/r02x: issue slots 49 % busy at 28 instructions
// per list, stalls on the LDS -> add -> STS chain and at the barriers):
// (1) the CTA stages the list offsets of `kc` k-mers x (tile_warps + 1) sub-tile boundaries in shared memory with
//     coalesced loads (a k-mer's offsets for the tile are contiguous), transposed to off[boundary][k-mer] so that a
//     warp reads the starts and ends of four of its lists with two LDS.128;
// (2) a warp requests the first 32 postings of G = 8 lists before it applies the previous G (double buffered in
//     registers; 8, 12 and 16 measure the same: the loads in flight are not what bounds it); lanes without a posting load a sentinel and their update is predicated off, so the apply loop has
//     no branch; lists longer than 32 (rare) get their rest applied in a pass of their own before the loop.
// Two CTAs per SM where the counters allow it (tiles of <= 14 sub-tiles): one CTA's selection and barriers hide
// behind the other's counting.
// Selection: the tile's top-`need` in rank order (score desc, id desc) are those above a threshold score T plus
// the highest ids among the ties at T. T is found on a histogram of a window of high scores; the first window
// starts at the `need`-th largest of the threads' own maxima (a lower bound of T that is almost always within a few
// ties of it), later ones halve downwards. The passes over the counters read eight at a time (LDS.128) and reject a
// pair of low counters with one packed u16x2 maximum.
// x = postings[a + lane] if lane < len, else the sentinel 0xffff stored behind the last posting (launch_index_build):
// an unconditional load from a selected address. (A predicated load into a preset register made ptxas put the preset
// BEHIND the load under register pressure: a write to a register with a load in flight, i.e. a stall on every list.)
__device__ __forceinline__ uint32_t ldg_posting(uint64_t post, uint32_t a, uint32_t lane, uint32_t len, uint32_t sentinel_at,
                                                uint32_t& off) {
    uint32_t x;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 ad;\n\tsetp.lt.u32 p, %4, %5;\n\tadd.u32 %1, %3, %4;\n\t"
                 "selp.u32 %1, %1, %6, p;\n\tmad.wide.u32 ad, %1, 2, %2;\n\tld.global.nc.u16 %0, [ad];\n\t}"
                 : "=r"(x), "=r"(off) : "l"(post), "r"(a), "r"(lane), "r"(len), "r"(sentinel_at));
    return x;
}
// counter x of the warp's sub-tile (shared byte address hb) += 1, unless x is the sentinel: no branch. The load is
// unconditional (sentinel lanes read the never-written zero word at dz), only the store is predicated: a predicated
// load left ptxas with a partially defined register per list, which it kept alive (16 registers) and spilled.
__device__ __forceinline__ void bump_posting(uint32_t hb, uint32_t dz, uint32_t x) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t.reg .u32 ad;\n\tsetp.ne.u32 p, %2, 0xffff;\n\tmad.lo.u32 ad, %2, 2, %0;\n\t"
                 "selp.u32 ad, ad, %1, p;\n\tld.shared.u16 c, [ad];\n\tadd.u32 c, c, 1;\n\t@p st.shared.u16 [ad], c;\n\t}" ::"r"(hb), "r"(dz), "r"(x) : "memory");
}
static FindLayout find_layout(const Index* ix) { return find_layout(ix->tile_warps, ix->sub_size); }

struct FindArgs {
    const uint32_t* kmers; const uint32_t* nk; const uint64_t* qoff;
    uint32_t N, sub_size, n_sub, tile_warps;
    const uint32_t* list_off; const uint16_t* postings;
    uint32_t sentinel_at;   // postings[sentinel_at] = 0xffff
    uint32_t zero;          // 0 (see request())
    uint32_t max; uint64_t* cand; uint32_t* cand_n; unsigned long long* counters;
    uint16_t* scores_out;   // non-null: write the tile's score counters to scores_out[q][N] instead of selecting (full ranking)
    uint32_t kc, ks, scratch_words;   // FindLayout
};

template <int FIND_G, int MAX_THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_CTAS) find_tile_kernel(FindArgs A) {
    extern __shared__ __align__(16) uint32_t hist32[];  // tile_warps * sub_size u16 counters, then the scratch words
    __shared__ uint32_t red[33];
    __shared__ uint32_t sh_sel[8];      // 0: T, 1: count_gt, 2: need_eq, 3: emitted, 4: found, 5: ties gathered, 6: count_eq, 7: postings
    uint16_t* hist = reinterpret_cast<uint16_t*>(hist32);
    const uint32_t q = blockIdx.x, tile = blockIdx.y, n_tiles = gridDim.y;
    const uint32_t B = A.sub_size, tw = A.tile_warps;
    uint32_t* scratch = hist32 + ((size_t)tw * B >> 1);
    uint32_t* hist2 = scratch;             // [SEL_BINS]
    uint32_t* tie = scratch + SEL_BINS;    // [TIE_CAP]
    const uint32_t tile_lo = tile * tw * B;
    const uint32_t tile_n = min(tw * B, A.N - tile_lo);
    const uint32_t quads = (tw * B) >> 3;              // all counters of the CTA, eight per 16 bytes (B is a power of two >= 32)
    const uint32_t tid = threadIdx.x, nt = blockDim.x, lane = lane_id(), w = warp_id();
    uint4* hist128 = reinterpret_cast<uint4*>(hist32);
    for (uint32_t i = tid; i < quads; i += nt) hist128[i] = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = tid; i < A.scratch_words; i += nt) scratch[i] = 0;   // zero columns
    if (tid == 0) sh_sel[7] = 0;

    // ---- counting
    const uint32_t sub0 = tile * tw, sub = sub0 + w;
    const uint32_t ow = tw + 1;                                   // staged offsets per k-mer
    const uint32_t ow_inv = ((1u << 20) + ow - 1u) / ow;
    const uint32_t kc = A.kc, ks = A.ks;
    const uint32_t* kl = A.kmers + A.qoff[q];
    const uint32_t nk = A.nk[q];
    uint32_t my_post = 0;                                         // postings of the tile's lists (mod 2^32 per thread, summed per CTA)
    uint32_t pre[FIND_PRE];
    auto fetch = [&](uint32_t c0) {   // offsets of the k-mers [c0, c0 + kc) x this tile's sub-tile boundaries
#pragma unroll
        for (int i = 0; i < FIND_PRE; i++) {
            const uint32_t idx = tid + (uint32_t)i * nt;
            const uint32_t k = (idx * ow_inv) >> 20, j = idx - k * ow;   // idx / ow, exact for idx < 2^20 / ow
            pre[i] = 0;
            if (k < kc && c0 + k < nk)
                pre[i] = __ldg(A.list_off + (uint64_t)__ldg(kl + c0 + k) * A.n_sub + min(sub0 + j, A.n_sub));
        }
    };
    const uint32_t hb = (uint32_t)__cvta_generic_to_shared(hist + (size_t)w * B);   // this warp's counters
    const uint32_t dz = (uint32_t)__cvta_generic_to_shared(scratch + A.kc);        // a zero column of the staged offsets: never written
    const uint32_t* row_a = scratch + w * ks;                     // row_a[k], row_a[ks + k]: this warp's list of k-mer k
    uint64_t post;                                                // kept in registers (not re-read from the constant bank per list)
    asm volatile("mov.u64 %0, %1;" : "=l"(post) : "l"(A.postings));
    const uint32_t sent = A.sentinel_at;
    for (uint32_t c0 = 0; c0 < nk; c0 += kc) {
        __syncthreads();              // counters zeroed / the previous chunk's offsets are no longer read
        fetch(c0);                    // (not prefetched during the previous chunk: the registers are worth more to the
                                      // lists in flight, and the SM's other CTA fills the gap)
#pragma unroll
        for (int i = 0; i < FIND_PRE; i++) {
            const uint32_t idx = tid + (uint32_t)i * nt;
            const uint32_t k = (idx * ow_inv) >> 20, j = idx - k * ow;
            if (k < kc) {
                scratch[j * ks + k] = pre[i];
                // boundaries past the index are clamped to its end: last boundary minus first = the tile's postings
                my_post += j == tw ? pre[i] : j == 0 ? 0u - pre[i] : 0u;
            }
        }
        __syncthreads();
        if (sub >= A.n_sub) continue;
        const uint32_t cn = min(kc, nk - c0);
        // lists longer than 32: everything behind the first 32 postings, four loads at a time (adds commute: done first)
        for (uint32_t kb = 0; kb < cn; kb += 32) {
            const uint32_t kk = kb + lane;                        // columns up to ks are readable and zero past the chunk
            uint32_t m = __ballot_sync(0xffffffffu, kk < cn && row_a[ks + kk] - row_a[kk] > 32u);
            for (; m; m &= m - 1u) {                              // warp-uniform
                const uint32_t k = kb + (uint32_t)__ffs((int)m) - 1u;
                const uint32_t a = row_a[k], len = row_a[ks + k] - a;
                for (uint32_t e0 = 32; e0 < len; e0 += 128) {
                    uint32_t y[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { uint32_t o; y[u] = ldg_posting(post, a + e0 + 32 * u, lane, len - min(len, e0 + 32 * u), sent, o); }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        bump_posting(hb, dz, y[u]);
                        __syncwarp();
                    }
                }
            }
        }
        // first 32 postings of the lists [k0, k0 + FIND_G) (lanes without one: the sentinel)
        // (`chain` is always 0, which ptxas cannot know: each group of four lists takes its offsets from an address that
        // depends on the previous group's, so the groups' offset loads and address arithmetic stay in program order
        // instead of being hoisted ahead of the request's first LDG. It was put in while hunting register spills whose
        // cause turned out to be bump_posting's predicated load; it costs two instructions per four lists and stayed
        // because this is the build the parity runs and the sanitizer saw.)
        auto request = [&](uint32_t k0, uint32_t (&x)[FIND_G]) {
            uint32_t chain = 0;
#pragma unroll
            for (int g = 0; g < FIND_G; g += 4) {
                const uint32_t* ra = row_a + k0 + g + chain;
                const uint4 a4 = *reinterpret_cast<const uint4*>(ra);
                const uint4 e4 = *reinterpret_cast<const uint4*>(ra + ks);
                uint32_t o0, o1, o2, o3;
                x[g] = ldg_posting(post, a4.x, lane, e4.x - a4.x, sent, o0);
                x[g + 1] = ldg_posting(post, a4.y, lane, e4.y - a4.y, sent, o1);
                x[g + 2] = ldg_posting(post, a4.z, lane, e4.z - a4.z, sent, o2);
                x[g + 3] = ldg_posting(post, a4.w, lane, e4.w - a4.w, sent, o3);
                chain = o3 & A.zero;
            }
        };
        auto apply = [&](const uint32_t (&x)[FIND_G]) {
#pragma unroll
            for (int g = 0; g < FIND_G; g++) {
                bump_posting(hb, dz, x[g]);
                __syncwarp();
            }
        };
        uint32_t xa[FIND_G], xb[FIND_G];
        request(0, xa);
#pragma unroll 1
        for (uint32_t k0 = 0; k0 < cn; k0 += 2 * FIND_G) {         // the next lists are on their way while these are applied
            request(k0 + FIND_G, xb);
            apply(xa);
            request(k0 + 2 * FIND_G, xa);
            apply(xb);
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_post += __shfl_xor_sync(0xffffffffu, my_post, o);
    if (lane == 0 && my_post) atomicAdd(&sh_sel[7], my_post);
    __syncthreads();
    if (tid == 0 && sh_sel[7]) atomicAdd(&A.counters[0], (unsigned long long)sh_sel[7]);
    if (A.scores_out) {   // full ranking (rank_full_kernel): hand the whole score vector over
        uint16_t* dst = A.scores_out + (uint64_t)q * A.N + tile_lo;
        for (uint32_t i = tid; i < tile_n; i += nt) dst[i] = hist[i];
        return;
    }

    // ---- selection
    const uint32_t need = min(A.max, tile_n);
    uint64_t* out = A.cand + ((uint64_t)q * n_tiles + tile) * A.max;
    auto score_of = [&](uint32_t i) -> uint32_t { return hist[i]; };
    // tile maximum, and the thread's own maxima (even / odd counters): packed u16x2 maxima over the counters
    uint32_t mx2 = 0;
    for (uint32_t i = tid; i < quads; i += nt) {
        const uint4 v4 = hist128[i];
        mx2 = __vmaxu2(__vmaxu2(mx2, v4.x), __vmaxu2(v4.y, __vmaxu2(v4.z, v4.w)));
    }
    const uint32_t own0 = mx2 & 0xffffu, own1 = mx2 >> 16;
    uint32_t mx = max(own0, own1);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[w] = mx;
    if (tid == 0) { sh_sel[3] = 0; sh_sel[4] = 0; sh_sel[5] = 0; }
    __syncthreads();
    if (tid == 0) { uint32_t m2 = 0; for (uint32_t i = 0; i < (nt >> 5); i++) m2 = max(m2, red[i]); red[32] = m2; }
    __syncthreads();
    const uint32_t M = red[32];
    // warp 0: the highest bin of hist2 (scores lo + 1 ...) at which `cum` + the entries in the bins above reach `need`;
    // sets sh_sel[0] = its score, [1] = entries above it, [2] = entries still needed at it, [6] = entries in it, [4] = 1;
    // if the window does not reach that far: [1] = cum + the window's entries
    auto scan_window = [&](uint32_t lo, uint32_t cum) {
        uint32_t part = 0;
        for (uint32_t b = 0; b < 32; b++) part += hist2[32 * lane + b];   // lane L owns bins [32L, 32L+32)
        // above = entries in the bins of higher lanes, suffix = entries in the whole window
        uint32_t above = 0, suffix = 0;
        for (int l = 31; l >= 0; l--) {
            const uint32_t pl2 = __shfl_sync(0xffffffffu, part, l);
            if ((int)lane == l) above = suffix;
            suffix += pl2;
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
    };
    // seed: the need-th largest of the threads' own maxima (2 per thread, distinct counters) is a lower bound of T
    uint32_t seed_lo = 0xffffffffu;
    if (M > 0) {
        const uint32_t lo0 = M > SEL_BINS ? M - SEL_BINS : 0u;
        for (uint32_t i = tid; i < SEL_BINS; i += nt) hist2[i] = 0;
        __syncthreads();
        if (own0 > lo0) atomicAdd(&hist2[own0 - lo0 - 1], 1u);
        if (own1 > lo0) atomicAdd(&hist2[own1 - lo0 - 1], 1u);
        __syncthreads();
        if (w == 0) scan_window(lo0, 0u);
        __syncthreads();
        if (sh_sel[4]) seed_lo = sh_sel[0] - 1u;
        __syncthreads();
        if (tid == 0) sh_sel[4] = 0;
    }
    uint32_t T = 0, count_gt = 0, need_eq = 0, count_eq = 0;
    {
        // windows (lo, hi] of scores, highest first; entries above the current window are already counted in cum.
        // Counters past tile_n are zero and a window's lower bound is >= 0: they never fall into one.
        uint32_t hi = M, cum = 0;
        for (;;) {
            uint32_t lo = hi / 2;
            if (seed_lo < hi) { lo = seed_lo; seed_lo = 0xffffffffu; }   // first window only
            if (hi - lo > SEL_BINS) lo = hi - SEL_BINS;
            for (uint32_t i = tid; i < SEL_BINS; i += nt) hist2[i] = 0;
            __syncthreads();
            if (hi > 0) {
                const uint32_t lo2 = lo | (lo << 16);
                for (uint32_t i = tid; i < quads; i += nt) {
                    const uint4 v4 = hist128[i];
                    const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t v = vv[c];
                        if (__vmaxu2(v, lo2) == lo2) continue;          // both counters <= lo
                        const uint32_t s0 = v & 0xffffu, s1 = v >> 16;
                        if (s0 > lo && s0 <= hi) atomicAdd(&hist2[s0 - lo - 1], 1u);
                        if (s1 > lo && s1 <= hi) atomicAdd(&hist2[s1 - lo - 1], 1u);
                    }
                }
            }
            __syncthreads();
            if (w == 0) scan_window(lo, cum);
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
    {
        const uint32_t below = T ? (T - 1u) | ((T - 1u) << 16) : 0u;    // T > 0: a pair of counters both < T holds nothing
        for (uint32_t i = tid; i < quads; i += nt) {
            const uint4 v4 = hist128[i];
            const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t v = vv[c];
                if (T && __vmaxu2(v, below) == below) continue;
                const uint32_t i0 = 8 * i + 2 * c;
                const uint32_t s0 = v & 0xffffu, s1 = v >> 16;
                if (s0 > T) out[atomicAdd(&sh_sel[3], 1u)] = ((uint64_t)s0 << 32) | (tile_lo + i0);
                if (s1 > T) out[atomicAdd(&sh_sel[3], 1u)] = ((uint64_t)s1 << 32) | (tile_lo + i0 + 1);
                if (rank_ties) {
                    if (s0 == T) tie[atomicAdd(&sh_sel[5], 1u)] = i0;
                    if (s1 == T) tie[atomicAdd(&sh_sel[5], 1u)] = i0 + 1;
                }
            }
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

static int launch_find_tile(const Index* ix, FindArgs& A, dim3 grid, cudaStream_t st) {
    const FindLayout L = find_layout(ix);
    A.kc = L.kc; A.ks = L.ks; A.scratch_words = L.scratch_words;
#define FIND_LAUNCH(V)                                                                                              \
    {                                                                                                               \
        auto kern = find_tile_kernel<FIND_VARIANTS[V].g, 32 * FIND_VARIANTS[V].max_warps, FIND_VARIANTS[V].ctas>;   \
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));              \
        kern<<<grid, 32 * ix->tile_warps, L.smem, st>>>(A);                                                         \
    }
    if (L.variant == 0) FIND_LAUNCH(0) else if (L.variant == 1) FIND_LAUNCH(1) else FIND_LAUNCH(2)
#undef FIND_LAUNCH
    return SG_OK;
}

// One CTA per (query, group of candidate lists): gather the lists [g * gs, (g + 1) * gs) of the query (lists_per_q of
// up to `max` keys each), bitonic-sort the 64-bit keys descending in shared memory, keep the first `max`.
// One level: a single group of all the tiles' lists. Two levels (max * tiles beyond the sort capacity): groups of
// tiles first, then the groups' winners.
__global__ void __launch_bounds__(1024) find_merge_kernel(const uint64_t* __restrict__ cand,
                                                           const uint32_t* __restrict__ cand_n, uint32_t lists_per_q,
                                                           uint32_t gs, uint32_t max, uint32_t p2,
                                                           uint64_t* __restrict__ out, uint32_t* __restrict__ out_n) {
    extern __shared__ uint64_t keys[];
    const uint32_t q = blockIdx.x, g = blockIdx.y, ng = gridDim.y;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0;
    __syncthreads();
    // compact the lists back to back (deterministic positions: prefix over cand_n, which is tiny)
    uint32_t base = 0;
    for (uint32_t t = g * gs; t < min((g + 1) * gs, lists_per_q); t++) {
        uint32_t c = cand_n[q * lists_per_q + t];
        const uint64_t* src = cand + ((uint64_t)q * lists_per_q + t) * max;
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
    const uint32_t r = min(max, base);   // all tiles: sum of min(max, tile size) >= min(max, N)
    uint64_t* dst = out + ((uint64_t)q * ng + g) * max;
    for (uint32_t i = threadIdx.x; i < r; i += blockDim.x) dst[i] = keys[i] - 1;
    if (threadIdx.x == 0) out_n[q * ng + g] = r;
}

static int launch_merge(Session* s, const uint64_t* cand, const uint32_t* cand_n, uint32_t lists_per_q, uint32_t gs, uint32_t ng,
                        uint32_t max, uint32_t n, uint64_t* out, uint32_t* out_n) {
    uint32_t p2 = 1;
    while (p2 < (uint64_t)max * std::min(gs, lists_per_q)) p2 <<= 1;
    SG_CUDA(cudaFuncSetAttribute(find_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(p2 * 8)));
    find_merge_kernel<<<dim3(n, ng), p2 / 2 < 1024 ? (p2 / 2 < 32 ? 32 : p2 / 2) : 1024, p2 * 8, s->stream>>>(
        cand, cand_n, lists_per_q, gs, max, p2, out, out_n);
    s->stats.kernel_launches += 1;
    return SG_OK;
}

int launch_find(Session* s, uint32_t max, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    if (n == 0) { q0 = 0; n = s->nq; }
    if (max == 0) SG_FAIL(SG_ERR_ARG, "find: max must be > 0");
    if (max > ix->N) max = ix->N;
    uint32_t gs = 0, ng = 0;
    if (!find_merge_plan(max, ix->n_tiles, &gs, &ng)) SG_FAIL(SG_ERR_LIMIT, "find: max * tiles exceeds the two-level top-k merge capacity");
    if (max > s->find_cap) {
        if (s->d_cand) cudaFree(s->d_cand);
        if (s->d_ranked) cudaFree(s->d_ranked);
        s->d_cand = nullptr; s->d_ranked = nullptr;
        SG_CUDA(cudaMalloc(&s->d_cand, (uint64_t)s->max_q * ix->n_tiles * max * sizeof(uint64_t)));
        SG_CUDA(cudaMalloc(&s->d_ranked, (uint64_t)s->max_q * max * sizeof(uint64_t)));
        s->find_cap = max;
    }
    if (ng > 1 && (uint64_t)s->max_q * ng * max > s->cand2_cap) {
        if (s->d_cand2) cudaFree(s->d_cand2);
        if (s->d_cand2_n) cudaFree(s->d_cand2_n);
        s->d_cand2 = nullptr; s->d_cand2_n = nullptr; s->cand2_cap = 0;
        SG_CUDA(cudaMalloc(&s->d_cand2, (uint64_t)s->max_q * ng * max * sizeof(uint64_t)));
        SG_CUDA(cudaMalloc(&s->d_cand2_n, (uint64_t)s->max_q * ix->n_tiles * sizeof(uint32_t)));
        s->cand2_cap = (uint64_t)s->max_q * ng * max;
    }
    s->find_max = max;
    query_kmers_kernel<<<(n + 3) / 4, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff + q0, n, ix->k, ix->nofast,
                                                          s->d_kmers, s->d_nk + q0);
    FindArgs A;
    A.kmers = s->d_kmers; A.nk = s->d_nk + q0; A.qoff = s->d_qoff + q0; A.N = ix->N; A.sub_size = ix->sub_size; A.n_sub = ix->n_sub;
    A.tile_warps = ix->tile_warps; A.list_off = ix->d_list_off; A.postings = ix->d_postings; A.sentinel_at = (uint32_t)ix->n_postings; A.zero = 0; A.max = max; A.scores_out = nullptr;
    A.cand = s->d_cand + (uint64_t)q0 * ix->n_tiles * max; A.cand_n = s->d_cand_n + (uint64_t)q0 * ix->n_tiles; A.counters = s->d_counters;
    dim3 grid(n, ix->n_tiles);
    SG_TRY(launch_find_tile(ix, A, grid, s->stream));
    uint64_t* ranked = s->d_ranked + (uint64_t)q0 * max;
    if (ng == 1) {
        SG_TRY(launch_merge(s, A.cand, A.cand_n, ix->n_tiles, ix->n_tiles, 1, max, n, ranked, s->d_nres + q0));
    } else {   // queries [q0, q0 + n) use the first n slots of the group buffers
        SG_TRY(launch_merge(s, A.cand, A.cand_n, ix->n_tiles, gs, ng, max, n, s->d_cand2, s->d_cand2_n));
        SG_TRY(launch_merge(s, s->d_cand2, s->d_cand2_n, ng, ng, 1, max, n, ranked, s->d_nres + q0));
    }
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 2;
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
    query_kmers_kernel<<<(n + 3) / 4, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff + q0, n, ix->k, ix->nofast,
                                                          s->d_kmers, s->d_nk + q0);
    FindArgs A;
    A.kmers = s->d_kmers; A.nk = s->d_nk + q0; A.qoff = s->d_qoff + q0; A.N = ix->N; A.sub_size = ix->sub_size; A.n_sub = ix->n_sub;
    A.tile_warps = ix->tile_warps; A.list_off = ix->d_list_off; A.postings = ix->d_postings; A.sentinel_at = (uint32_t)ix->n_postings; A.zero = 0; A.max = 1;
    A.cand = nullptr; A.cand_n = nullptr; A.counters = s->d_counters; A.scores_out = s->d_full_scores;
    dim3 grid(n, ix->n_tiles);
    SG_TRY(launch_find_tile(ix, A, grid, s->stream));
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
