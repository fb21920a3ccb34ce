// scan.cu -- K-1: word offsets of the reads' packed bases, computed on the device.
//
// When the reads' bases are packed densely (read i starts where read i-1 ends, every read padded to whole
// 32-bit words -- what every flattener in this repo produces), seq_off is the exclusive prefix sum of
// ceil(l_seq / 8) and need not travel over PCIe: 4 of the ~91 bytes per read of the end-to-end path
// (kdl_ctx_consensus with batch->seq_off == NULL).  Three small launches: per-CTA totals, one CTA scanning the
// totals, per-CTA scan + offset.  1024 reads per CTA, 128-bit loads and stores.
#include "kdl_common.cuh"

namespace kdl {

constexpr int S_THREADS = 256;
constexpr int S_PER = 4;                       // reads per thread
constexpr int S_BLOCK = S_THREADS * S_PER;     // reads per CTA

__device__ __forceinline__ uint32_t seq_words(int l) { return (uint32_t)(((l & 0x7fffffff) + 7) >> 3); }

// the CTA's 1024 lengths -> words[4] of this thread; returns the thread's total
__device__ __forceinline__ uint32_t load_words(const int32_t* __restrict__ l_seq, long long n, long long base, int tid,
                                               uint32_t (&w)[S_PER]) {
    const long long i0 = base + (long long)S_PER * tid;
    if (i0 + S_PER <= n) {
        const int4 v = *reinterpret_cast<const int4*>(l_seq + i0);  // base and S_PER * tid are multiples of 4
        w[0] = seq_words(v.x); w[1] = seq_words(v.y); w[2] = seq_words(v.z); w[3] = seq_words(v.w);
    } else {
#pragma unroll
        for (int k = 0; k < S_PER; ++k) w[k] = i0 + k < n ? seq_words(l_seq[i0 + k]) : 0u;
    }
    return (w[0] + w[1]) + (w[2] + w[3]);
}

// exclusive prefix of `v` over the CTA's 256 threads (thread order); *total = sum over the CTA
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sum[S_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();  // warp_sum may still be read by a previous call
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < S_THREADS / 32; ++k) {
        const uint32_t s = warp_sum[k];
        if (k < warp) before += s;
        all += s;
    }
    *total = all;
    return before + incl - v;
}

__global__ void __launch_bounds__(S_THREADS)
seq_off_block_sums_kernel(const int32_t* __restrict__ l_seq, long long n, uint32_t* __restrict__ block_sums) {
    uint32_t w[S_PER], total;
    const uint32_t mine = load_words(l_seq, n, (long long)blockIdx.x * S_BLOCK, threadIdx.x, w);
    cta_exclusive_scan(mine, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// one CTA: block_sums[0 .. n_blocks) -> their exclusive prefix sums, in place
__global__ void __launch_bounds__(S_THREADS)
seq_off_scan_sums_kernel(uint32_t* __restrict__ block_sums, int n_blocks) {
    uint32_t carry = 0;
    for (int base = 0; base < n_blocks; base += S_THREADS) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = cta_exclusive_scan(v, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(S_THREADS)
seq_off_write_kernel(const int32_t* __restrict__ l_seq, long long n, const uint32_t* __restrict__ block_prefix,
                     uint32_t* __restrict__ seq_off) {
    uint32_t w[S_PER], total;
    const long long base = (long long)blockIdx.x * S_BLOCK;
    const uint32_t mine = load_words(l_seq, n, base, threadIdx.x, w);
    uint32_t off = block_prefix[blockIdx.x] + cta_exclusive_scan(mine, &total);
    const long long i0 = base + (long long)S_PER * threadIdx.x;
    uint32_t o[S_PER];
#pragma unroll
    for (int k = 0; k < S_PER; ++k) { o[k] = off; off += w[k]; }
    if (i0 + S_PER <= n) {
        *reinterpret_cast<uint4*>(seq_off + i0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int k = 0; k < S_PER; ++k)
            if (i0 + k < n) seq_off[i0 + k] = o[k];
    }
}

#ifndef KDL_HOST_EMU
long long seq_off_scan_blocks(long long n) { return (n + S_BLOCK - 1) / S_BLOCK; }

// host launcher (own translation unit: see api.cu).  block_sums: scratch of seq_off_scan_blocks(n) words.
// Returns 0, or 1 if a launch failed.
int launch_seq_off_scan(const int32_t* l_seq, long long n, uint32_t* block_sums, uint32_t* seq_off, cudaStream_t st) {
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)seq_off_scan_blocks(n);
    seq_off_block_sums_kernel<<<blocks, S_THREADS, 0, st>>>(l_seq, n, block_sums);
    seq_off_scan_sums_kernel<<<1, S_THREADS, 0, st>>>(block_sums, (int)blocks);
    seq_off_write_kernel<<<blocks, S_THREADS, 0, st>>>(l_seq, n, block_sums, seq_off);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace kdl
