// K-mer index build on the device: CSR posting lists per (k-mer, reference sub-tile), resident in HBM.
// A k-mer's lists for consecutive sub-tiles are consecutive, so the whole posting list of a k-mer is one
// contiguous run partitioned by sub-tile; ids are stored relative to the sub-tile (u16).
// Replaces kmer_search::impl::build + IndexBuilder (reference src/kmer_search.cpp:152-276):
//   * every reference row contributes each of its k-mers ONCE (unique_kmers / unique_prefix_kmers, :164-177)
//   * fast mode keeps only k-mers whose first base is A (prefix_kmers(...,1,BASE_A), src/kmer.h:110-125)
//   * the k-mer ending on a row's last base is never produced (iterator quirk, src/kmer.h:179-201)
//   * ambiguous bases break the window (generator::push, src/kmer.h:69-78)
// Lists longer than N/2 are stored inverted by the reference (:264-266) purely to save CPU bandwidth;
// counting totals are identical, so the CSR keeps plain lists.
#include "common.cuh"

namespace sg {

// One CTA per reference row. Pass 0 counts unique k-mers per (k-mer, sub-tile) slot, pass 1 writes ids.
__global__ void __launch_bounds__(128) idx_rows_kernel(const uint8_t* __restrict__ masks,
                                                        const uint64_t* __restrict__ row_off, uint32_t N, int k,
                                                        int nofast, uint32_t sub_size, uint32_t n_sub,
                                                        uint32_t hash_size, unsigned int* __restrict__ counts,
                                                        const uint32_t* __restrict__ list_off,
                                                        uint16_t* __restrict__ postings, int pass) {
    extern __shared__ uint32_t hset[];  // open addressing, key+1, 0 = empty
    for (uint32_t row = blockIdx.x; row < N; row += gridDim.x) {
        const uint64_t a = row_off[row];
        const uint32_t n = (uint32_t)(row_off[row + 1] - a);
        const uint8_t* m = masks + a;
        for (uint32_t i = threadIdx.x; i < hash_size; i += blockDim.x) hset[i] = 0;
        __syncthreads();
        const uint32_t sub = row / sub_size;
        const uint16_t local = (uint16_t)(row - sub * sub_size);
        if (n > (uint32_t)k) {
            for (uint32_t i = (uint32_t)k - 1 + threadIdx.x; i + 1 < n; i += blockDim.x) {
                uint32_t v;
                if (!kmer_at(m, i, k, v)) continue;
                if (!nofast && (v >> (2 * (k - 1))) != 0) continue;  // first base must be A (code 0)
                uint32_t h = (v * 2654435761u) & (hash_size - 1);
                bool fresh = false;
                for (;;) {
                    uint32_t old = atomicCAS(&hset[h], 0u, v + 1u);
                    if (old == 0u) { fresh = true; break; }
                    if (old == v + 1u) break;
                    h = (h + 1) & (hash_size - 1);
                }
                if (fresh) {
                    uint64_t slot = (uint64_t)v * n_sub + sub;
                    unsigned int c = atomicAdd(&counts[slot], 1u);
                    if (pass == 1) postings[list_off[slot] + c] = local;
                }
            }
        }
        __syncthreads();
    }
}

// ---- exclusive scan of u32 counts into u32 offsets (three small kernels; n can be hundreds of millions)
constexpr int SCAN_ITEMS = 8, SCAN_THREADS = 1024, SCAN_BLOCK = SCAN_ITEMS * SCAN_THREADS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_partial_kernel(const unsigned int* __restrict__ in, uint64_t n,
                                                                     unsigned long long* __restrict__ block_sums) {
    __shared__ uint32_t red[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) if (base + j < n) s += in[base + j];
    uint32_t tot;
    block_exscan(s, red, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned long long* __restrict__ block_sums, uint64_t nb,
                                                          unsigned long long* __restrict__ total) {
    // single CTA, sequential over chunks of 1024 block sums (values fit u32 per chunk element: <= 8192*max count)
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[33];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t c = 0; c < nb; c += 1024) {
        uint64_t i = c + threadIdx.x;
        unsigned long long v = i < nb ? block_sums[i] : 0ull, x = v;
        uint32_t lane = lane_id(), w = warp_id();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            unsigned long long s = wsum[lane], t = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= (uint32_t)o) t += y;
            }
            wsum[lane] = t - s;
            if (lane == 31) wsum[32] = t;
        }
        __syncthreads();
        if (i < nb) block_sums[i] = carry + wsum[w] + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(unsigned int* __restrict__ counts, uint64_t n,
                                                                   const unsigned long long* __restrict__ block_sums,
                                                                   uint32_t* __restrict__ off) {
    __shared__ uint32_t red[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = base + j < n ? counts[base + j] : 0; s += v[j]; }
    uint32_t ex = block_exscan(s, red, nullptr);
    uint64_t run = block_sums[blockIdx.x] + ex;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) { off[base + j] = (uint32_t)run; counts[base + j] = 0; }  // counts become the fill cursors
        run += v[j];
    }
}

int launch_index_build(Index* ix, cudaStream_t st) {
    const uint64_t n = ix->n_slots * ix->n_sub;
    uint32_t hash_size = 1024;
    while (hash_size < 2 * ix->max_row_len + 2) hash_size <<= 1;
    if (hash_size > 32768) SG_FAIL(SG_ERR_LIMIT, "reference row longer than 16383 bases: not supported by the index builder");
    size_t smem = (size_t)hash_size * 4;
    SG_CUDA(cudaFuncSetAttribute(idx_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned int* counts = nullptr;
    unsigned long long *block_sums = nullptr, *total = nullptr;
    uint64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    SG_CUDA(cudaMalloc(&counts, n * sizeof(unsigned int)));
    SG_CUDA(cudaMalloc(&block_sums, (nb + 1) * sizeof(unsigned long long)));
    SG_CUDA(cudaMalloc(&total, sizeof(unsigned long long)));
    SG_CUDA(cudaMalloc(&ix->d_list_off, (n + 1) * sizeof(uint32_t)));
    SG_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(unsigned int), st));
    uint32_t grid = ix->N < 148u * 16u ? (ix->N ? ix->N : 1) : 148u * 16u;
    idx_rows_kernel<<<grid, 128, smem, st>>>(ix->d_masks, ix->d_row_off, ix->N, ix->k, ix->nofast, ix->sub_size,
                                            ix->n_sub, hash_size, counts, nullptr, nullptr, 0);
    scan_partial_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(counts, n, block_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
    scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(counts, n, block_sums, ix->d_list_off);
    unsigned long long h_total = 0;
    SG_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(h_total), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    ix->n_postings = h_total;
    if (h_total >= (1ull << 32)) {
        cudaFree(counts); cudaFree(block_sums); cudaFree(total);
        SG_FAIL(SG_ERR_LIMIT, "more than 2^32 posting entries: not supported (32-bit list offsets)");
    }
    const uint32_t total32 = (uint32_t)h_total;
    SG_CUDA(cudaMemcpyAsync(ix->d_list_off + n, &total32, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMalloc(&ix->d_postings, (ix->n_postings + 64) * sizeof(uint16_t)));
    // postings[n_postings ...] = 0xffff: what the search kernel's lanes without a posting load (never a local id: sub-tiles
    // hold at most 32768 references)
    SG_CUDA(cudaMemsetAsync(ix->d_postings + ix->n_postings, 0xff, 64 * sizeof(uint16_t), st));
    idx_rows_kernel<<<grid, 128, smem, st>>>(ix->d_masks, ix->d_row_off, ix->N, ix->k, ix->nofast, ix->sub_size,
                                            ix->n_sub, hash_size, counts, ix->d_list_off, ix->d_postings, 1);
    SG_CUDA(cudaStreamSynchronize(st));
    SG_CUDA(cudaGetLastError());
    cudaFree(counts); cudaFree(block_sums); cudaFree(total);
    return SG_OK;
}

}  // namespace sg
