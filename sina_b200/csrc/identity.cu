// sina_b200 -- pairwise identity of aligned sequences and the --search stage.
//   identity_kernel:  cseq_comparator::operator()            (reference src/cseq_comparator.cpp:57-118,209-293)
//   search_select_kernel: search_filter::operator(), k-mer branch (reference src/search_filter.cpp:244-330)
//
// The reference walks the two position-sorted base vectors in a merge loop. Every count it produces is a function
// of three facts per base -- is its column inside the overlap [lo, hi] of the two (trimmed) sequences, does the other
// sequence have a base in the same column, is either base filtered (lowercase) -- so one CTA per query puts the query
// into a column-indexed table in shared memory (1 B per alignment column) and its warps stream the candidates' rows
// through it: one coalesced pass over (column, base) of the candidate per pair, no merge, no divergence.
#include "common.cuh"

namespace sg {

namespace {

constexpr int ID_THREADS = 256;
constexpr uint32_t SCORE_REMOVED = 0xFFFFFFFFu;   // pair dropped by --search-ignore-super (a NaN pattern, never a score)

struct IdentArgs {
    const uint8_t* amasks; const uint32_t* acols; const uint64_t* aoff;      // aligned queries
    const uint64_t* ranked; const uint32_t* nres; uint32_t stride;           // candidates: keys (.. | id) per query, or
    const uint32_t* pair_ids; const uint64_t* pair_off;                      // explicit pairs (sg_identity_batch)
    const uint8_t* masks; const uint32_t* cols; const uint64_t* row_off; uint32_t W;
    int iupac, cover, filter_lc, ignore_super, use_table;
    float* scores;            // [nq][stride] or [pair_off[nq]]
};

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t v) {   // first index with a[i] >= v
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// base_iupac::comp / comp_pessimistic / comp_exact (src/aligned_base.h:153-169) on the 4-bit masks
__device__ __forceinline__ bool base_match(uint32_t a, uint32_t b, int rule) {
    a &= 15u; b &= 15u;
    if (rule == 0) return (a & b) != 0;
    if (rule == 1) return __popc(a) <= 1 && a == b;   // !is_ambig() of the QUERY base
    return a == b;
}

__global__ void __launch_bounds__(ID_THREADS) identity_kernel(IdentArgs A) {
    extern __shared__ uint8_t qtab[];     // [W] base of the query in that column, 0 = none
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5, nwarp = ID_THREADS / 32;
    const uint64_t a0 = A.aoff[q];
    const uint32_t na = (uint32_t)(A.aoff[q + 1] - a0);
    const uint8_t* am = A.amasks + a0;
    const uint32_t* ac = A.acols + a0;
    uint32_t npair;
    const uint32_t* pid = nullptr;
    float* out;
    if (A.pair_ids) { npair = (uint32_t)(A.pair_off[q + 1] - A.pair_off[q]); pid = A.pair_ids + A.pair_off[q]; out = A.scores + A.pair_off[q]; }
    else { npair = A.nres[q]; out = A.scores + (uint64_t)q * A.stride; }
    if (A.use_table) {
        for (uint32_t i = tid; i < (A.W + 15u) / 16u; i += ID_THREADS) reinterpret_cast<uint4*>(qtab)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (uint32_t i = tid; i < na; i += ID_THREADS) if (ac[i] < A.W) qtab[ac[i]] = am[i];
        __syncthreads();
    }
    // the query after traverse()'s trimming of filtered bases at both ends (src/cseq_comparator.cpp:65-79)
    uint32_t ta0 = 0, ta1 = na, a_unf_total = na;
    if (A.filter_lc) {
        __shared__ uint32_t sh[3];
        if (tid == 0) { sh[0] = 0xFFFFFFFFu; sh[1] = 0; sh[2] = 0; }
        __syncthreads();
        uint32_t mn = 0xFFFFFFFFu, mx = 0, cnt = 0;
        for (uint32_t i = tid; i < na; i += ID_THREADS)
            if (!(am[i] & 16u)) { mn = min(mn, i); mx = max(mx, i + 1); cnt++; }
        atomicMin(&sh[0], mn); atomicMax(&sh[1], mx); atomicAdd(&sh[2], cnt);
        __syncthreads();
        ta0 = sh[0]; ta1 = sh[1]; a_unf_total = sh[2];
        if (ta0 == 0xFFFFFFFFu) { ta0 = 0; ta1 = 0; }
    }
    const float qnan = __int_as_float(0x7fc00000);
    for (uint32_t pi = wid; pi < npair; pi += nwarp) {
        const uint32_t r = pid ? pid[pi] : (uint32_t)A.ranked[(uint64_t)q * A.stride + pi];
        const uint64_t b0 = A.row_off[r];
        const uint32_t nb = (uint32_t)(A.row_off[r + 1] - b0);
        const uint8_t* bm = A.masks + b0;
        const uint32_t* bc = A.cols + b0;
        // --search-ignore-super: boost::algorithm::contains(target bases, query bases, comp) (src/search_filter.cpp:263-267)
        if (A.ignore_super) {
            bool found = false;
            if (na <= nb) {
                for (uint32_t off0 = 0; off0 + na <= nb && !found; off0 += 32) {
                    const uint32_t off = off0 + lane;
                    bool ok = off + na <= nb;
                    for (uint32_t j = 0; ok && j < na; j++) ok = (am[j] & bm[off + j] & 15u) != 0;
                    found = __any_sync(0xffffffffu, ok);
                }
            }
            if (!found) { if (lane == 0) out[pi] = __uint_as_float(SCORE_REMOVED); continue; }   // the reference keeps the containing ones
        }
        uint32_t tb0 = 0, tb1 = nb;
        if (A.filter_lc) {
            uint32_t mn = 0xFFFFFFFFu, mx = 0;
            for (uint32_t j = lane; j < nb; j += 32) if (!(bm[j] & 16u)) { mn = min(mn, j); mx = max(mx, j + 1); }
            mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
            tb0 = mn == 0xFFFFFFFFu ? 0 : mn; tb1 = mn == 0xFFFFFFFFu ? 0 : mx;
        }
        if (ta0 >= ta1 || tb0 >= tb1) { if (lane == 0) out[pi] = qnan; continue; }   // the reference dereferences end() here
        const uint32_t fa = ac[ta0], la = ac[ta1 - 1], fb = bc[tb0], lb = bc[tb1 - 1];
        const uint32_t lo = max(fa, fb), hi = min(la, lb);
        uint32_t match = 0, mismatch = 0, only_b = 0, ovh_b = 0;
        for (uint32_t j = tb0 + lane; j < tb1; j += 32) {
            const uint32_t col = bc[j], b = bm[j];
            const bool bf = A.filter_lc && (b & 16u);
            if (col < lo || col > hi) { ovh_b += !bf; continue; }
            uint32_t a = 0;
            if (A.use_table) a = qtab[col];
            else { const uint32_t i = lower_bound_u32(ac, na, col); if (i < na && ac[i] == col) a = am[i]; }
            if (!a) { only_b += !bf; continue; }
            const bool af = A.filter_lc && (a & 16u);
            if (!af && !bf) { if (base_match(a, b, A.iupac)) match++; else mismatch++; }   // both() (src/cseq_comparator.cpp:190-204)
            else if (af && !bf) only_b++;                                                   // (a alone: counted on the query side below)
        }
        match = warp_sum(match); mismatch = warp_sum(mismatch); only_b = warp_sum(only_b); ovh_b = warp_sum(ovh_b);
        // query side: unfiltered bases inside the overlap are matched, mismatched or alone
        uint32_t a_unf_region;
        if (!A.filter_lc) a_unf_region = lo <= hi ? lower_bound_u32(ac, na, hi + 1u) - lower_bound_u32(ac, na, lo) : 0u;
        else {
            uint32_t c = 0;
            for (uint32_t i = ta0 + lane; i < ta1; i += 32) c += (!(am[i] & 16u) && ac[i] >= lo && ac[i] <= hi) ? 1u : 0u;
            a_unf_region = warp_sum(c);
        }
        const int only_a = (int)a_unf_region - (int)match - (int)mismatch;
        const int ovh_a = (int)a_unf_total - (int)a_unf_region;
        int base;
        const int m = (int)match, mm = (int)mismatch, ob = (int)only_b, vb = (int)ovh_b;
        switch (A.cover) {   // src/cseq_comparator.cpp:240-277
        case 0: base = 1; break;
        case 1: base = m + mm + only_a + ovh_a; break;
        case 2: base = m + mm + ob + vb; break;
        case 3: base = m + mm + only_a + ob; break;
        case 4: base = m + mm + only_a + ob + ovh_a + vb; break;
        case 5: base = m + mm + (only_a + ob + ovh_a + vb) / 2; break;
        case 6: base = m + mm + min(only_a + ovh_a, ob + vb); break;
        case 7: base = m + mm + max(only_a + ovh_a, ob + vb); break;
        default: base = m + mm; break;
        }
        if (lane == 0) out[pi] = __fdiv_rn((float)m, (float)base);   // (float)m.match / base (:279)
    }
}

// partial_sort(greater<result_item>) + the min_sim cut (src/search_filter.cpp:322-330): the max_result best pairs by
// (score, name) descending, emitted while score > min_sim. name_rank[id] = rank of the reference's name in ascending
// order (null: the id itself).
__global__ void __launch_bounds__(256) search_select_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ ranked,
                                                            const uint32_t* __restrict__ nres, uint32_t stride,
                                                            const uint32_t* __restrict__ name_rank, uint32_t max_result,
                                                            float min_sim, uint32_t* out_ids, float* out_scores, uint32_t* out_n) {
    __shared__ unsigned long long best[8];
    __shared__ uint32_t besti[8];
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t n = nres[q];
    const float* sc = scores + (uint64_t)q * stride;
    const uint64_t* rk = ranked + (uint64_t)q * stride;
    unsigned long long prev = ~0ull;   // keys are distinct (the rank / id is part of the key)
    uint32_t emitted = 0;
    for (uint32_t k = 0; k < max_result; k++) {
        unsigned long long my = 0; uint32_t myi = 0xFFFFFFFFu;
        for (uint32_t i = tid; i < n; i += 256) {
            const uint32_t bits = __float_as_uint(sc[i]);
            if (bits >= 0x7f800000u) continue;   // removed, NaN, negative: never reported
            const uint32_t id = (uint32_t)rk[i];
            const unsigned long long key = ((unsigned long long)bits << 32) | (name_rank ? name_rank[id] : id);
            if (key < prev && (myi == 0xFFFFFFFFu || key > my)) { my = key; myi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, my, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, myi, o);
            if (oi != 0xFFFFFFFFu && (myi == 0xFFFFFFFFu || ok > my)) { my = ok; myi = oi; }
        }
        if (lane == 0) { best[wid] = my; besti[wid] = myi; }
        __syncthreads();
        unsigned long long b = 0; uint32_t bi = 0xFFFFFFFFu;
        for (int w = 0; w < 8; w++) if (besti[w] != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || best[w] > b)) { b = best[w]; bi = besti[w]; }
        __syncthreads();
        if (bi == 0xFFFFFFFFu) break;
        const float s = sc[bi];
        if (!(s > min_sim)) break;
        if (tid == 0) { out_ids[(uint64_t)q * max_result + emitted] = (uint32_t)rk[bi]; out_scores[(uint64_t)q * max_result + emitted] = s; }
        emitted++;
        prev = b;
    }
    if (tid == 0) out_n[q] = emitted;
}

}  // namespace

static size_t table_bytes(uint32_t W) { return ((size_t)W + 15) / 16 * 16; }

int launch_identity(Session* s, const uint8_t* d_amasks, const uint32_t* d_acols, const uint64_t* d_aoff, uint32_t nq,
                    const uint64_t* ranked, const uint32_t* nres, uint32_t stride, const uint32_t* pair_ids,
                    const uint64_t* pair_off, int iupac, int cover, int filter_lc, int ignore_super, float* d_scores) {
    Index* ix = s->ix;
    IdentArgs A;
    A.amasks = d_amasks; A.acols = d_acols; A.aoff = d_aoff; A.ranked = ranked; A.nres = nres; A.stride = stride;
    A.pair_ids = pair_ids; A.pair_off = pair_off;
    A.masks = ix->d_masks; A.cols = ix->d_cols; A.row_off = ix->d_row_off; A.W = ix->W;
    A.iupac = iupac; A.cover = cover; A.filter_lc = filter_lc; A.ignore_super = ignore_super;
    const size_t tb = table_bytes(ix->W);
    A.use_table = tb <= 200 * 1024;
    A.scores = d_scores;
    const size_t smem = A.use_table ? tb : 0;
    SG_CUDA(cudaFuncSetAttribute(identity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    identity_kernel<<<nq, ID_THREADS, smem, s->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

int launch_search_select(Session* s, const float* d_scores, const uint64_t* ranked, const uint32_t* nres, uint32_t stride,
                         uint32_t nq, uint32_t max_result, float min_sim, uint32_t* d_ids, float* d_out, uint32_t* d_n) {
    search_select_kernel<<<nq, 256, 0, s->stream>>>(d_scores, ranked, nres, stride, s->ix->d_name_rank, max_result, min_sim,
                                                    d_ids, d_out, d_n);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
