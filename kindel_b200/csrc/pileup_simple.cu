// pileup_simple.cu -- K1s: plain nM reads (one M/=/X op, fully inside the contig).
//
// Restates kindel/kindel.py:49-54 for the reads whose whole CIGAR is a single match op -- the bulk
// of a short-read BAM.  A simple read needs no CIGAR fetch at all: the flatten step stored the op
// length in l_seq (bit 31 clear, see include/kindel_b200.h).
//
// v1 (this kernel): one warp per read, lanes stride over the bases, one RED per base into the
// five weight columns.  Correct for any read order; bounded by L2 atomic throughput, not HBM.
// The tile-owner kernel in pileup_tiled.cu replaces it for coordinate-sorted input.
#include "kdl_common.cuh"

namespace kdl {

__global__ void __launch_bounds__(256)
pileup_simple_atomic_kernel(kdl_batch b, int32_t* __restrict__ counts, long long n_slots,
                            int32_t* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (long long r = warp0; r < b.n_reads; r += n_warps) {
        const int32_t lraw = b.l_seq[r];
        if (lraw < 0) continue;  // complex read: K1g walks it
        const int c = find_contig(b.contig_read_off, b.n_contigs, r);
        const long long base = b.contig_slot[c] + b.ref_start[r];
        const uint32_t* __restrict__ seq = b.seq4 + (size_t)b.seq_off[r];
        for (int k = lane; k < lraw; k += 32) {
            const int col = nib2col(nibble_at(seq, k));
            if (col < 0) { bad = true; continue; }
            atomicAdd(counts + (long long)col * n_slots + base + k, 1);
        }
    }
    if (bad) atomicOr(err_flag, 1);
}

}  // namespace kdl
