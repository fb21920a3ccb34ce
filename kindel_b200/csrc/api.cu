// api.cu -- the C ABI declared in include/kindel_b200.h (unity build of the kernel files).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "kdl_common.cuh"
#include "pileup_general.cu"
#include "pileup_simple.cu"
#include "pileup_tile.cu"
#include "vote.cu"
#include "assemble.cu"

namespace {

std::atomic<long long> g_launches{0};

inline int grid_for(long long items, int per_block, int cap) {
    long long g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

// SM count of the CURRENT device (a process may drive several GPUs)
inline int sm_count() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// opt-in to > 48 KB dynamic shared memory for the tile kernel's instantiations, once per device
template <int kFlush, bool kCx>
inline int ensure_tile_smem() {
    static std::atomic<int> done[kMaxDevices];
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return KDL_OK;
    if (cudaFuncSetAttribute(kdl::pileup_tile_kernel<kFlush, kCx>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(kdl::TileSmem<kdl::TileCfg<kCx>>)) != cudaSuccess)
        return KDL_ERR_CUDA;
    done[dev].store(1, std::memory_order_release);
    return KDL_OK;
}

template <int kFlush, bool kCx>
inline int launch_tile(const kdl_batch& b, int32_t* counts, long long n_slots, long long tile_lo, long long n_tiles,
                       int split, int zero_rest, cudaStream_t st) {
    int rc = ensure_tile_smem<kFlush, kCx>();
    if (rc != KDL_OK) return rc;
    const long long units = n_tiles * split, max_grid = (long long)sm_count() * 2;  // two CTAs per SM, persistent
    const long long grid = units < max_grid ? units : max_grid;
    kdl::pileup_tile_kernel<kFlush, kCx><<<(unsigned)grid, kdl::W_THREADS, sizeof(kdl::TileSmem<kdl::TileCfg<kCx>>), st>>>(
        b, counts, n_slots, b.tile_index, tile_lo, n_tiles, split, zero_rest);
    return KDL_OK;
}

inline int check_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? KDL_OK : KDL_ERR_CUDA;
}

int validate_batch(const kdl_batch* b) {
    if (!b || b->n_reads < 0 || b->n_contigs < 0 || b->n_hard < 0 || b->n_complex < b->n_hard)
        return KDL_ERR_INVALID_ARG;
    if (b->n_reads > 0 && (!b->ref_start || !b->seq_off || !b->l_seq || !b->seq4 ||
                           !b->contig_read_off || !b->contig_len || !b->contig_slot))
        return KDL_ERR_INVALID_ARG;
    if ((b->n_hard > 0 && !b->hard_idx) || (b->n_complex > 0 && !b->complex_idx)) return KDL_ERR_INVALID_ARG;
    if (b->reach_right < b->max_simple_len || b->reach_left < 0) return KDL_ERR_INVALID_ARG;
    return KDL_OK;
}

}  // namespace

extern "C" {

int kdl_abi_version(void) { return KDL_ABI_VERSION; }

const char* kdl_status_string(int status) {
    switch (status) {
        case KDL_OK: return "ok";
        case KDL_ERR_INVALID_ARG: return "invalid argument";
        case KDL_ERR_CUDA: return "CUDA error";
        case KDL_ERR_NO_DEVICE: return "no CUDA device";
        case KDL_ERR_INDEX: return "IndexError: read walks off the contig or off its SEQ";
        case KDL_ERR_KEY: return "KeyError: base outside A,C,G,T,N in an M or S op";
        default: return "unknown status";
    }
}

int64_t kdl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int kdl_pileup(const kdl_batch* batch, int32_t* counts, int64_t n_slots, int32_t* ins_events,
               int32_t* err_flag, void* stream) {
    return kdl_pileup_range(batch, counts, n_slots, 0, n_slots, 0, ins_events, err_flag, stream);
}

int kdl_pileup_range(const kdl_batch* batch, int32_t* counts, int64_t n_slots, int64_t slot_lo,
                     int64_t slot_hi, int32_t flags, int32_t* ins_events, int32_t* err_flag, void* stream) {
    int rc = validate_batch(batch);
    if (rc != KDL_OK) return rc;
    if (!counts || !err_flag || n_slots <= 0 || slot_lo < 0 || slot_hi > n_slots || slot_lo > slot_hi)
        return KDL_ERR_INVALID_ARG;
    const bool tileable = (n_slots % KDL_TILE) == 0 && (slot_lo % KDL_TILE) == 0 && (slot_hi % KDL_TILE) == 0;
    if ((slot_lo & 3) || (slot_hi & 3)) return KDL_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = sm_count() * 8;
    const bool has_tile_reads = batch->n_reads > batch->n_hard;
    const bool tiled = has_tile_reads && tileable && batch->reads_sorted && batch->tile_index &&
                       batch->reach_right > 0 && batch->reach_right <= KDL_FAST_MAXLEN + KDL_TILE_MAXREACH;
    const bool fresh = (flags & KDL_PILEUP_FRESH_WEIGHTS) != 0;
    const long long tile_lo = slot_lo / KDL_TILE, n_tiles = (slot_hi - slot_lo) / KDL_TILE;
    // Depth split: a small reference piled deep has fewer tiles than the GPU has CTA slots (30 kb = 59 tiles for
    // 296 slots); `split` CTAs then share a tile by read range and flush with REDs into a zeroed table.
    int split = 1;
    if (tiled && n_tiles > 0) {
        const long long slots = (long long)sm_count() * 2;
        if (n_tiles < slots) {
            long long want = (slots + n_tiles - 1) / n_tiles;
            const long long deep = batch->n_reads / (n_tiles * 256);  // at least ~256 reads per unit
            if (want > deep) want = deep;
            split = (int)(want < 1 ? 1 : (want > 32 ? 32 : want));
        }
        if (const char* ev = getenv("KDL_SPLIT")) { const int v = atoi(ev); if (v >= 1 && v <= 64) split = v; }
    }
    bool cx_by_atomics = (batch->n_complex - batch->n_hard) * 16 < batch->n_reads;
    if (const char* ev = getenv("KDL_CX")) cx_by_atomics = !strcmp(ev, "atomics") ? true : (!strcmp(ev, "pieces") ? false : cx_by_atomics);
    // zeroing that the chosen kernels will not do themselves: the tile kernel overwrites the weight columns of a
    // fresh table and, on request, zeroes columns 5..18 window by window in its flush
    const int zero_in_k1 = (tiled && split == 1 && fresh && n_tiles > 0 && (flags & KDL_PILEUP_ZERO_REST)) ? 1 : 0;
    const int zero_from = (fresh && !(tiled && split == 1)) ? 0 : 5;
    const int zero_to = ((flags & KDL_PILEUP_ZERO_REST) && !zero_in_k1) ? KDL_NCOL : 5;
    if (zero_to > zero_from && slot_hi > slot_lo) {
        kdl::zero_cols_kernel<<<sm_count() * 4, 256, 0, st>>>(counts, n_slots, zero_from, zero_to, slot_lo, slot_hi);
        if ((rc = check_launch()) != KDL_OK) return rc;
    }
    if (batch->n_reads == 0) return KDL_OK;
    if (tiled) {
        if (n_tiles > 0) {
            // K0: read range per tile of the slot range (the linear index of a sorted BAM, built on device)
            kdl::tile_index_kernel<<<(unsigned)((n_tiles * 32 + 255) / 256), 256, 0, st>>>(*batch, tile_lo, n_tiles,
                                                                                          batch->tile_index);
            if ((rc = check_launch()) != KDL_OK) return rc;
            // K1: the tile-owner kernel.  Tile-eligible complex reads go through its piece machinery (kCx) when
            // they are a sizeable share of the batch; when they are rare (< 1/16 of the reads) the lean instantiation
            // runs and K1e counts their bases too, with REDs
            const bool cx = batch->n_complex > batch->n_hard && !cx_by_atomics;
            if (split > 1) {
                rc = cx ? launch_tile<kdl::F_ATOMIC, true>(*batch, counts, n_slots, tile_lo, n_tiles, split, 0, st)
                        : launch_tile<kdl::F_ATOMIC, false>(*batch, counts, n_slots, tile_lo, n_tiles, split, 0, st);
            } else if (fresh) {
                rc = cx ? launch_tile<kdl::F_STORE, true>(*batch, counts, n_slots, tile_lo, n_tiles, 1, zero_in_k1, st)
                        : launch_tile<kdl::F_STORE, false>(*batch, counts, n_slots, tile_lo, n_tiles, 1, zero_in_k1, st);
            } else {
                rc = cx ? launch_tile<kdl::F_ADD, true>(*batch, counts, n_slots, tile_lo, n_tiles, 1, 0, st)
                        : launch_tile<kdl::F_ADD, false>(*batch, counts, n_slots, tile_lo, n_tiles, 1, 0, st);
            }
            if (rc != KDL_OK) return rc;
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
        if (batch->n_complex > batch->n_hard) {  // K1e: insertions / deletions / clips of the tile-eligible complex reads
            if (cx_by_atomics)  // their bases too: 8 lanes per read
                kdl::pileup_events_kernel<8><<<(unsigned)((batch->n_complex + 31) / 32), 256, 0, st>>>(
                    *batch, counts, n_slots, ins_events, 1);
            else                // a few scattered REDs per read: one thread per read
                kdl::pileup_events_kernel<1><<<(unsigned)((batch->n_complex + 255) / 256), 256, 0, st>>>(
                    *batch, counts, n_slots, ins_events, 0);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
        if (batch->n_hard > 0) {  // K1g: the reads that may wrap or raise, atomically, after the tile stores
            const int grid = grid_for(batch->n_hard, 8, cap);
            kdl::pileup_general_kernel<<<grid, 256, 0, st>>>(*batch, batch->hard_idx, batch->n_hard, counts, n_slots,
                                                            ins_events, err_flag);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
    } else {
        // order-independent fallback (unsorted input): K1s for the simple reads, K1g for every complex read
        if (batch->n_reads > batch->n_complex) {
            const int grid = grid_for(batch->n_reads, 8, cap);  // 8 warps (reads) per 256-thread CTA
            kdl::pileup_simple_atomic_kernel<<<grid, 256, 0, st>>>(*batch, counts, n_slots, err_flag);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
        if (batch->n_complex > 0) {
            const int grid = grid_for(batch->n_reads, 8, cap);
            kdl::pileup_general_kernel<<<grid, 256, 0, st>>>(*batch, nullptr, batch->n_reads, counts, n_slots,
                                                            ins_events, err_flag);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
    }
    return KDL_OK;
}

int kdl_diagnose(const kdl_batch* batch, kdl_diag* diag_dev, void* stream) {
    int rc = validate_batch(batch);
    if (rc != KDL_OK) return rc;
    if (!diag_dev) return KDL_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    kdl::diagnose_init_kernel<<<1, 1, 0, st>>>(diag_dev);
    if ((rc = check_launch()) != KDL_OK) return rc;
    if (batch->n_hard > 0) {
        const long long grid = (batch->n_hard + 255) / 256;
        kdl::diagnose_kernel<<<(unsigned)grid, 256, 0, st>>>(*batch, diag_dev);
        if ((rc = check_launch()) != KDL_OK) return rc;
    }
    kdl::diagnose_final_kernel<<<1, 1, 0, st>>>(diag_dev);
    return check_launch();
}

int kdl_vote(const int32_t* counts, int64_t n_slots, int64_t min_depth_ceil, uint8_t* calls,
             void* stream) {
    if (!counts || !calls || n_slots <= 0 || (n_slots & 3)) return KDL_ERR_INVALID_ARG;
    kdl::Peers none;
    none.n = 0;
    const long long quads = n_slots / 4;
    const long long grid = (quads + 255) / 256;
    kdl::vote_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        counts, none, n_slots, 0, n_slots, min_depth_ceil, calls, nullptr);
    return check_launch();
}

int kdl_derive(const int32_t* counts, int64_t n_slots, int32_t* out, void* stream) {
    if (!counts || !out || n_slots <= 0) return KDL_ERR_INVALID_ARG;
    const long long grid = (n_slots + 255) / 256;
    kdl::derive_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(counts, n_slots, out);
    return check_launch();
}

int kdl_vote_peers(const int32_t* const* peer_counts, int32_t n_peers, int64_t n_slots,
                   int64_t slot_lo, int64_t slot_hi, int64_t min_depth_ceil, uint8_t* calls,
                   int32_t* reduced, void* stream) {
    return kdl_vote_peers_sparse(peer_counts, nullptr, nullptr, n_peers, n_slots, slot_lo, slot_hi,
                                 min_depth_ceil, calls, reduced, stream);
}

static int make_exchange(const kdl_exchange* x, kdl::Exchange* e, int64_t n_slots) {
    if (!x || x->n_ranks < 1 || x->n_ranks > 16 || x->rank < 0 || x->rank >= x->n_ranks || !x->counter)
        return KDL_ERR_INVALID_ARG;
    e->peers.n = x->n_ranks;
    e->rank = x->rank;
    e->counter = x->counter;
    for (int p = 0; p < x->n_ranks; ++p) {
        if (!x->tables[p] || !x->calls[p] || !x->ready[p] || !x->done[p]) return KDL_ERR_INVALID_ARG;
        e->peers.tab[p] = x->tables[p];
        e->peers.lo[p] = x->foot_lo[p];
        e->peers.hi[p] = n_slots > 0 && x->foot_hi[p] > n_slots ? n_slots : x->foot_hi[p];
        if ((e->peers.lo[p] & 3) || (e->peers.hi[p] & 3)) return KDL_ERR_INVALID_ARG;
        e->calls[p] = x->calls[p];
        e->ready[p] = x->ready[p];
        e->done[p] = x->done[p];
        e->slice_lo[p] = x->slice_lo[p];
        e->slice_hi[p] = x->slice_hi[p];
        if ((x->slice_lo[p] & 3) || (x->slice_hi[p] & 3) || x->slice_hi[p] < x->slice_lo[p]) return KDL_ERR_INVALID_ARG;
    }
    e->ready_local = x->ready[x->rank];
    e->done_local = x->done[x->rank];
    return KDL_OK;
}

int kdl_exchange_signal(const kdl_exchange* x, int32_t epoch, void* stream) {
    kdl::Exchange e;
    int rc = make_exchange(x, &e, 0);
    if (rc != KDL_OK) return rc;
    kdl::exchange_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(e, epoch);
    return check_launch();
}

int kdl_exchange_wait(const kdl_exchange* x, int32_t epoch, void* stream) {
    kdl::Exchange e;
    int rc = make_exchange(x, &e, 0);
    if (rc != KDL_OK) return rc;
    dim3 grid((unsigned)(sm_count() * 2 / x->n_ranks + 1), (unsigned)x->n_ranks);
    kdl::exchange_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e, epoch);
    return check_launch();
}

int kdl_exchange_vote(const kdl_exchange* x, int64_t n_slots, int64_t min_depth_ceil, int32_t epoch,
                      void* stream) {
    if (n_slots <= 0 || (n_slots & 3)) return KDL_ERR_INVALID_ARG;
    kdl::Exchange e;
    int rc = make_exchange(x, &e, n_slots);
    if (rc != KDL_OK) return rc;
    if (x->slice_hi[x->rank] > n_slots) return KDL_ERR_INVALID_ARG;
    const long long quads = (x->slice_hi[x->rank] - x->slice_lo[x->rank]) / 4;
    long long grid = (quads + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;  // an empty slice still has to take part in the flag protocol
    kdl::vote_exchange_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(e, n_slots, min_depth_ceil, epoch);
    return check_launch();
}

int kdl_cdr_flags(const int32_t* counts, int64_t n_slots, int64_t slot_lo, int64_t slot_hi,
                  double clip_decay_threshold, uint8_t* flags, uint8_t* bases, void* stream) {
    if (!counts || !flags || !bases || n_slots <= 0 || slot_lo < 0 || slot_hi > n_slots || slot_lo > slot_hi)
        return KDL_ERR_INVALID_ARG;
    if (slot_hi == slot_lo) return KDL_OK;
    const long long grid = (slot_hi - slot_lo + 255) / 256;
    kdl::cdr_flags_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(counts, n_slots, slot_lo, slot_hi,
                                                                            clip_decay_threshold, flags, bases);
    return check_launch();
}

int64_t kdl_assemble_scratch_words(int64_t n_slots) {
    return n_slots < 0 ? 0 : (n_slots + 1 + kdl::A_BLOCK - 1) / kdl::A_BLOCK + 1;
}

int kdl_assemble(const uint8_t* calls, int64_t n_slots, const int64_t* contig_slot, const int32_t* contig_len,
                 int32_t n_contigs, const int64_t* ins_slot, const uint32_t* ins_off, const uint8_t* ins_bytes,
                 int64_t n_ins, uint32_t* block_sums, uint32_t* offsets, uint8_t* out, void* stream) {
    if (!calls || n_slots <= 0 || !contig_slot || !contig_len || n_contigs < 0 || n_ins < 0 || !block_sums || !offsets ||
        !out || (n_ins > 0 && (!ins_slot || !ins_off || !ins_bytes)))
        return KDL_ERR_INVALID_ARG;
    kdl::AssembleArgs a;
    a.calls = calls; a.n_slots = n_slots; a.contig_slot = contig_slot; a.contig_len = contig_len; a.n_contigs = n_contigs;
    a.ins_slot = ins_slot; a.ins_off = ins_off; a.ins_bytes = ins_bytes; a.n_ins = n_ins;
    const long long n_blocks = (n_slots + 1 + kdl::A_BLOCK - 1) / kdl::A_BLOCK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    kdl::assemble_sums_kernel<<<(unsigned)n_blocks, kdl::A_THREADS, 0, st>>>(a, block_sums);
    if ((rc = check_launch()) != KDL_OK) return rc;
    kdl::assemble_scan_sums_kernel<<<1, kdl::A_THREADS, 0, st>>>(block_sums, n_blocks);
    if ((rc = check_launch()) != KDL_OK) return rc;
    kdl::assemble_scatter_kernel<<<(unsigned)n_blocks, kdl::A_THREADS, 0, st>>>(a, block_sums, offsets, out);
    return check_launch();
}

int kdl_table_alloc(int64_t bytes, void** dev_ptr) {
    if (!dev_ptr || bytes <= 0) return KDL_ERR_INVALID_ARG;
    *dev_ptr = nullptr;
    if (cudaMalloc(dev_ptr, (size_t)bytes) != cudaSuccess) return KDL_ERR_CUDA;
    if (cudaMemset(*dev_ptr, 0, (size_t)bytes) != cudaSuccess) return KDL_ERR_CUDA;
    return KDL_OK;
}
int kdl_table_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? KDL_OK : KDL_ERR_CUDA; }
int kdl_ipc_export(void* dev_ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dev_ptr || !handle) return KDL_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dev_ptr) != cudaSuccess) return KDL_ERR_CUDA;
    std::memcpy(handle, &h, 64);
    return KDL_OK;
}
int kdl_ipc_open(const uint8_t handle[64], void** dev_ptr) {
    if (!dev_ptr || !handle) return KDL_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    return cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? KDL_OK : KDL_ERR_CUDA;
}
int kdl_ipc_close(void* dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? KDL_OK : KDL_ERR_CUDA; }

int kdl_vote_peers_sparse(const int32_t* const* peer_counts, const int64_t* foot_lo, const int64_t* foot_hi,
                          int32_t n_peers, int64_t n_slots, int64_t slot_lo, int64_t slot_hi,
                          int64_t min_depth_ceil, uint8_t* calls, int32_t* reduced, void* stream) {
    if (!peer_counts || n_peers < 1 || n_peers > 16 || !calls || n_slots <= 0 || (n_slots & 3) ||
        slot_lo < 0 || slot_hi > n_slots || (slot_lo & 3) || (slot_hi & 3))
        return KDL_ERR_INVALID_ARG;
    if (slot_hi <= slot_lo) return KDL_OK;
    kdl::Peers peers;
    peers.n = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_counts[p]) return KDL_ERR_INVALID_ARG;
        peers.tab[p] = peer_counts[p];
        peers.lo[p] = foot_lo ? foot_lo[p] : 0;
        peers.hi[p] = foot_hi ? foot_hi[p] : n_slots;
        if ((peers.lo[p] & 3) || (peers.hi[p] & 3)) return KDL_ERR_INVALID_ARG;
    }
    const long long quads = (slot_hi - slot_lo) / 4;
    const long long grid = (quads + 255) / 256;
    kdl::vote_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        nullptr, peers, n_slots, slot_lo, slot_hi, min_depth_ceil, calls, reduced);
    return check_launch();
}

}  // extern "C"

#include "host_ctx.inl"
