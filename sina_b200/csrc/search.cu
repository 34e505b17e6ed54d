// K-mer search, top-k ranking and family selection on the device.
//   find_tile_kernel   kmer_search::impl::find, counting part  (reference src/kmer_search.cpp:389-409)
//   find_merge_kernel  partial_sort by greater<pair<int16,int>> (:412): score desc, then id desc
//   family_kernel      famfinder::impl::match + gap filter + fs_req (src/famfinder.cpp:497-612, 474-491)
#include "common.cuh"

namespace sg {

// ---------------------------------------------------------------------------------------------------
// One CTA per (query, reference tile). The tile's scores live in shared memory as packed u16 counters;
// each warp takes one query k-mer at a time and streams its posting list with coalesced 4-byte loads,
// one shared-memory atomic per posting. A query k-mer occurring twice is counted twice (all_kmers /
// prefix_kmers, not the unique_ variants, :391,397). Then the tile's top-`max` (score desc, id desc) is
// selected with a two-level radix select over the 16-bit scores and emitted unordered.
__global__ void __launch_bounds__(1024) find_tile_kernel(
    const uint8_t* __restrict__ qmasks, const uint64_t* __restrict__ qoff, uint32_t N, int k, int nofast,
    uint32_t tile_size, uint64_t n_slots, const uint64_t* __restrict__ list_off,
    const uint32_t* __restrict__ postings, uint32_t max, uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_n,
    unsigned long long* __restrict__ counters) {
    extern __shared__ uint32_t hist[];  // (tile_n+1)/2 words of two u16 counters
    __shared__ uint32_t h256[256];
    __shared__ uint32_t red[33];
    __shared__ uint32_t sh_sel[4];      // 0: threshold T, 1: count_gt, 2: need_eq, 3: emitted
    const uint32_t q = blockIdx.x, tile = blockIdx.y, n_tiles = gridDim.y;
    const uint32_t tile_lo = tile * tile_size;
    const uint32_t tile_n = min(tile_size, N - tile_lo);
    const uint32_t words = (tile_n + 1) >> 1;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) hist[i] = 0;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) h256[i] = 0;
    __syncthreads();

    const uint8_t* m = qmasks + qoff[q];
    const uint32_t n = (uint32_t)(qoff[q + 1] - qoff[q]);
    const uint32_t lane = lane_id(), nw = blockDim.x >> 5;
    unsigned long long my_post = 0;
    if (n > (uint32_t)k) {
        for (uint32_t i = (uint32_t)k - 1 + warp_id(); i + 1 < n; i += nw) {  // last k-mer never produced
            uint32_t v;
            if (!kmer_at(m, i, k, v)) continue;                               // warp-uniform
            if (!nofast && (v >> (2 * (k - 1))) != 0) continue;               // fast: first base A
            const uint64_t slot = (uint64_t)tile * n_slots + v;
            const uint64_t a = list_off[slot], b = list_off[slot + 1];
            uint64_t e = a + lane;
            for (; e + 96 < b; e += 128) {                                    // 4 independent loads in flight
                uint32_t i0 = postings[e], i1 = postings[e + 32], i2 = postings[e + 64], i3 = postings[e + 96];
                i0 -= tile_lo; i1 -= tile_lo; i2 -= tile_lo; i3 -= tile_lo;
                atomicAdd(&hist[i0 >> 1], 1u << ((i0 & 1) * 16));
                atomicAdd(&hist[i1 >> 1], 1u << ((i1 & 1) * 16));
                atomicAdd(&hist[i2 >> 1], 1u << ((i2 & 1) * 16));
                atomicAdd(&hist[i3 >> 1], 1u << ((i3 & 1) * 16));
            }
            for (; e < b; e += 32) {
                uint32_t i0 = postings[e] - tile_lo;
                atomicAdd(&hist[i0 >> 1], 1u << ((i0 & 1) * 16));
            }
            if (lane == 0) my_post += b - a;
        }
    }
    if (lane == 0 && my_post) atomicAdd(&counters[0], my_post);
    __syncthreads();

    // ---- top-`need` of the tile in rank order (score desc, id desc) via radix select on the score
    const uint32_t need = min(max, tile_n);
    auto score_of = [&](uint32_t i) -> uint32_t { return (hist[i >> 1] >> ((i & 1) * 16)) & 0xffffu; };
    for (uint32_t i = threadIdx.x; i < tile_n; i += blockDim.x) atomicAdd(&h256[score_of(i) >> 8], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t cum = 0;
        int b1 = 255;
        for (; b1 > 0; b1--) { if (cum + h256[b1] >= need) break; cum += h256[b1]; }
        sh_sel[0] = (uint32_t)b1; sh_sel[1] = cum;
    }
    __syncthreads();
    const uint32_t b1 = sh_sel[0];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) h256[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += blockDim.x) {
        uint32_t sc = score_of(i);
        if ((sc >> 8) == b1) atomicAdd(&h256[sc & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t cum = sh_sel[1];
        int b2 = 255;
        for (; b2 > 0; b2--) { if (cum + h256[b2] >= need) break; cum += h256[b2]; }
        sh_sel[0] = (b1 << 8) | (uint32_t)b2;  // threshold score T
        sh_sel[1] = cum;                        // entries with score > T
        sh_sel[2] = need - cum;                 // entries to take among score == T (highest ids first)
        sh_sel[3] = 0;
    }
    __syncthreads();
    const uint32_t T = sh_sel[0], count_gt = sh_sel[1];
    uint64_t* out = cand + ((uint64_t)q * n_tiles + tile) * max;
    for (uint32_t i = threadIdx.x; i < tile_n; i += blockDim.x) {
        uint32_t sc = score_of(i);
        if (sc > T) {
            uint32_t p = atomicAdd(&sh_sel[3], 1u);
            out[p] = ((uint64_t)sc << 32) | (tile_lo + i);
        }
    }
    // ties: walk ids downwards in chunks of blockDim, thread 0 <-> highest id of the chunk
    uint32_t remaining = sh_sel[2], emitted_eq = 0;
    for (uint32_t top = tile_n; top > 0 && remaining > 0;) {
        uint32_t chunk = min(top, (uint32_t)blockDim.x);
        uint32_t flag = 0, i = 0;
        if (threadIdx.x < chunk) { i = top - 1 - threadIdx.x; flag = score_of(i) == T; }
        uint32_t tot, ex = block_exscan(flag, red, &tot);
        if (flag && ex < remaining) out[count_gt + emitted_eq + ex] = ((uint64_t)T << 32) | (tile_lo + i);
        uint32_t took = min(tot, remaining);
        emitted_eq += took; remaining -= took;
        top -= chunk;
    }
    if (threadIdx.x == 0) cand_n[q * n_tiles + tile] = need;
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

int launch_find(Session* s, uint32_t max) {
    Index* ix = s->ix;
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
    const uint32_t tile_n = ix->tile_size < ix->N ? ix->tile_size : ix->N;
    size_t smem = (size_t)((tile_n + 1) / 2) * 4;
    SG_CUDA(cudaFuncSetAttribute(find_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SG_CUDA(cudaFuncSetAttribute(find_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(p2 * 8)));
    dim3 grid(s->nq, ix->n_tiles);
    find_tile_kernel<<<grid, 1024, smem, s->stream>>>(s->d_qmasks, s->d_qoff, ix->N, ix->k, ix->nofast, ix->tile_size,
                                                     ix->n_slots, ix->d_list_off, ix->d_postings, max, s->d_cand,
                                                     s->d_cand_n, s->d_counters);
    find_merge_kernel<<<s->nq, p2 / 2 < 1024 ? (p2 / 2 < 32 ? 32 : p2 / 2) : 1024, p2 * 8, s->stream>>>(
        s->d_cand, s->d_cand_n, ix->n_tiles, max, ix->N, p2, s->d_ranked, s->d_nres);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 2;
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
                              int32_t* __restrict__ fam_n, uint32_t* __restrict__ retry) {
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

int launch_family(Session* s, const sg_fam_params& fp, uint32_t window) {
    Index* ix = s->ix;
    SG_CUDA(cudaMemsetAsync(s->d_retry, 0, sizeof(uint32_t), s->stream));
    family_kernel<<<(s->nq + 127) / 128, 128, 0, s->stream>>>(s->d_ranked, s->d_nres, s->nq, window, ix->N,
                                                             ix->d_row_off, ix->d_cols, s->d_excl, fp, s->fam_cap,
                                                             s->d_fam_ids, s->d_fam_scores, s->d_fam_n, s->d_retry);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
