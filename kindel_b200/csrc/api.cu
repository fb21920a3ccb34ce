// api.cu -- the C ABI declared in include/kindel_b200.h (unity build of the kernel files).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "kdl_common.cuh"
#include "pileup_general.cu"
#include "pileup_simple.cu"
#include "pileup_tiled.cu"
#include "pileup_ws.cu"
#include "pileup_wide.cu"
#include "vote.cu"

namespace kdl {
// scan.cu is its own translation unit (so that adding it leaves the code generated for the kernels above
// untouched); its launcher:
int launch_seq_off_scan(const int32_t* l_seq, long long n, uint32_t* block_sums, uint32_t* seq_off, cudaStream_t st);
long long seq_off_scan_blocks(long long n);
}  // namespace kdl

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

// opt-in to > 48 KB dynamic shared memory for K1f, once per device
inline int ensure_k1f_smem(int smem) {
    static std::atomic<int> done[kMaxDevices];
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return KDL_OK;
    if (cudaFuncSetAttribute(kdl::pileup_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess ||
        cudaFuncSetAttribute(kdl::pileup_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
        return KDL_ERR_CUDA;
    if (cudaFuncSetAttribute(kdl::pileup_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(kdl::WsSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(kdl::pileup_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(kdl::WsSmem)) != cudaSuccess)
        return KDL_ERR_CUDA;
    done[dev].store(1, std::memory_order_release);
    return KDL_OK;
}

// the same opt-in for the experimental K1x, made only when that kernel is selected so that the default
// path never depends on it
inline int ensure_k1x_smem() {
    static std::atomic<int> done[kMaxDevices];
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return KDL_OK;
    if (cudaFuncSetAttribute(kdl::pileup_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(kdl::WideSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(kdl::pileup_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(kdl::WideSmem)) != cudaSuccess)
        return KDL_ERR_CUDA;
    done[dev].store(1, std::memory_order_release);
    return KDL_OK;
}

// the same opt-in for the experimental lean instantiation of K1f (KDL_K1F=lean)
inline int ensure_k1f_lean_smem(int smem) {
    static std::atomic<int> done[kMaxDevices];
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return KDL_OK;
    if (cudaFuncSetAttribute(kdl::pileup_tiled_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess ||
        cudaFuncSetAttribute(kdl::pileup_tiled_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess)
        return KDL_ERR_CUDA;
    done[dev].store(1, std::memory_order_release);
    return KDL_OK;
}

// K1w2 = pileup_ws_kernel<.., WsCfg2>: two CTAs per SM, setmaxnreg (experimental, KDL_K1F=ws2)
inline int ensure_k1w2_smem() {
    static std::atomic<int> done[kMaxDevices];
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return KDL_OK;
    const int bytes = (int)sizeof(kdl::WsSmemT<kdl::WsCfg2>);
    if (cudaFuncSetAttribute(kdl::pileup_ws_kernel<false, kdl::WsCfg2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             bytes) != cudaSuccess ||
        cudaFuncSetAttribute(kdl::pileup_ws_kernel<true, kdl::WsCfg2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             bytes) != cudaSuccess)
        return KDL_ERR_CUDA;
    done[dev].store(1, std::memory_order_release);
    return KDL_OK;
}

inline bool use_ws2_kernel() {
    const char* ev = getenv("KDL_K1F");
    return ev && !strcmp(ev, "ws2");
}

inline bool use_lean_kernel() {  // K1f<.., kLean = true>, experimental (not yet validated on a GPU)
    const char* ev = getenv("KDL_K1F");
    return ev && !strcmp(ev, "lean");
}

inline bool use_wide_kernel() {  // K1x, experimental (not yet validated on a GPU)
    const char* ev = getenv("KDL_K1F");
    return ev && !strcmp(ev, "wide");
}

// which tile-owner kernel: "ws" = warp-specialised pipeline (K1w), "tiled" = K1f
inline bool use_ws_kernel() {
    const char* ev = getenv("KDL_K1F");
    if (ev && !strcmp(ev, "tiled")) return false;
    if (ev && !strcmp(ev, "ws")) return true;
    return false;
}

inline int check_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? KDL_OK : KDL_ERR_CUDA;
}

int validate_batch(const kdl_batch* b) {
    if (!b || b->n_reads < 0 || b->n_contigs < 0) return KDL_ERR_INVALID_ARG;
    if (b->n_reads > 0 && (!b->ref_start || !b->seq_off || !b->l_seq || !b->seq4 ||
                           !b->contig_read_off || !b->contig_len || !b->contig_slot))
        return KDL_ERR_INVALID_ARG;
    if (b->n_complex > 0 && (!b->complex_idx || !b->evt_off || !b->cig_off || (b->n_ops > 0 && !b->cigar)))
        return KDL_ERR_INVALID_ARG;
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
    const bool has_simple = batch->n_reads > batch->n_complex;
    const bool tiled = has_simple && tileable && batch->reads_sorted && batch->tile_index &&
                       batch->max_simple_len > 0 && batch->max_simple_len <= KDL_FAST_MAXLEN;
    const bool fresh = (flags & KDL_PILEUP_FRESH_WEIGHTS) != 0;
    // zeroing that the chosen kernels will not do themselves
    const int zero_from = (fresh && !tiled) ? 0 : 5;
    const int zero_to = (flags & KDL_PILEUP_ZERO_REST) ? KDL_NCOL : 5;
    if (zero_to > zero_from && slot_hi > slot_lo) {
        kdl::zero_cols_kernel<<<sm_count() * 4, 256, 0, st>>>(counts, n_slots, zero_from, zero_to, slot_lo, slot_hi);
        if ((rc = check_launch()) != KDL_OK) return rc;
    }
    if (batch->n_reads == 0) return KDL_OK;
    if (tiled) {
        // K0: read range per tile of the slot range (the linear index of a sorted BAM, built on device)
        const long long tile_lo = slot_lo / KDL_TILE, n_tiles = (slot_hi - slot_lo) / KDL_TILE;
        if (n_tiles > 0) {
            kdl::tile_index_kernel<<<(unsigned)((n_tiles * 32 + 255) / 256), 256, 0, st>>>(*batch, tile_lo, n_tiles,
                                                                                          batch->tile_index);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
        // K1f: one CTA per tile, 2 CTAs per SM (2 x ~90 KB shared memory)
        const int smem = (int)sizeof(kdl::FastSmem);
        if ((rc = ensure_k1f_smem(smem)) != KDL_OK) return rc;
        if (n_tiles > 0 && use_ws_kernel()) {
            // K1w: one persistent CTA per SM (4 producer + 8 consumer warps, ~200 KB shared memory)
            long long grid = n_tiles < (long long)sm_count() ? n_tiles : (long long)sm_count();
            const int wsmem = (int)sizeof(kdl::WsSmem);
            if (fresh)
                kdl::pileup_ws_kernel<true><<<(unsigned)grid, kdl::W_THREADS, wsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            else
                kdl::pileup_ws_kernel<false><<<(unsigned)grid, kdl::W_THREADS, wsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            if ((rc = check_launch()) != KDL_OK) return rc;
        } else if (n_tiles > 0 && use_ws2_kernel()) {
            if ((rc = ensure_k1w2_smem()) != KDL_OK) return rc;
            const long long max_grid = (long long)sm_count() * 2;
            const long long grid = n_tiles < max_grid ? n_tiles : max_grid;
            const int wsmem = (int)sizeof(kdl::WsSmemT<kdl::WsCfg2>);
            if (fresh)
                kdl::pileup_ws_kernel<true, kdl::WsCfg2><<<(unsigned)grid, kdl::W_THREADS, wsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            else
                kdl::pileup_ws_kernel<false, kdl::WsCfg2><<<(unsigned)grid, kdl::W_THREADS, wsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            if ((rc = check_launch()) != KDL_OK) return rc;
        } else if (n_tiles > 0 && use_wide_kernel()) {
            long long grid = n_tiles < (long long)sm_count() * 2 ? n_tiles : (long long)sm_count() * 2;
            const int xsmem = (int)sizeof(kdl::WideSmem);
            if ((rc = ensure_k1x_smem()) != KDL_OK) return rc;
            if (fresh)
                kdl::pileup_wide_kernel<true><<<(unsigned)grid, kdl::F_THREADS, xsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            else
                kdl::pileup_wide_kernel<false><<<(unsigned)grid, kdl::F_THREADS, xsmem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            if ((rc = check_launch()) != KDL_OK) return rc;
        } else if (n_tiles > 0) {
            // CTAs per SM-slot: 2 are resident per SM; each CTA walks its tiles with a software pipeline
            // (metadata of its next tile streams in while it counts), so a persistent grid is best
            long long mult = 1;
            if (const char* ev = getenv("KDL_K1F_GRID_MULT")) { mult = atoll(ev); if (mult < 1) mult = 1; }
            const long long max_grid = (long long)sm_count() * 2 * mult;
            long long grid = n_tiles < max_grid ? n_tiles : max_grid;
            if (use_lean_kernel()) {
                if ((rc = ensure_k1f_lean_smem(smem)) != KDL_OK) return rc;
                if (fresh)
                    kdl::pileup_tiled_kernel<true, true><<<(unsigned)grid, kdl::F_THREADS, smem, st>>>(
                        *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
                else
                    kdl::pileup_tiled_kernel<false, true><<<(unsigned)grid, kdl::F_THREADS, smem, st>>>(
                        *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            } else if (fresh)
                kdl::pileup_tiled_kernel<true><<<(unsigned)grid, kdl::F_THREADS, smem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            else
                kdl::pileup_tiled_kernel<false><<<(unsigned)grid, kdl::F_THREADS, smem, st>>>(
                    *batch, counts, n_slots, batch->tile_index, tile_lo, n_tiles);
            if ((rc = check_launch()) != KDL_OK) return rc;
        }
    } else if (has_simple) {
        const int grid = grid_for(batch->n_reads, 8, cap);  // 8 warps (reads) per 256-thread CTA
        kdl::pileup_simple_atomic_kernel<<<grid, 256, 0, st>>>(*batch, counts, n_slots, err_flag);
        if ((rc = check_launch()) != KDL_OK) return rc;
    }
    if (batch->n_complex > 0) {
        const int grid = grid_for(batch->n_complex, 8, cap);
        kdl::pileup_general_kernel<<<grid, 256, 0, st>>>(*batch, counts, n_slots, ins_events, err_flag);
        if ((rc = check_launch()) != KDL_OK) return rc;
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
    if (batch->n_complex > 0) {
        const long long grid = (batch->n_complex + 255) / 256;
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
