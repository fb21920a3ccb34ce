// emu_pileup.cpp -- TEST INFRASTRUCTURE: K0 + the tile-owner kernels (K1f, K1x) compiled for the host and run
// under tests/emu/cuda_emu.h.  The kernel sources are included as they are (KDL_HOST_EMU only swaps the inline
// PTX for functional stand-ins); nothing here is part of the product.
#define KDL_HOST_EMU 1
#include "cuda_emu.h"

#include "../../kindel_b200/csrc/kdl_common.cuh"
#include "../../kindel_b200/csrc/pileup_tiled.cu"
#include "../../kindel_b200/csrc/pileup_wide.cu"
#include "../../kindel_b200/csrc/pileup_ws.cu"

static char g_error[512];

extern "C" {

const char* emu_last_error() { return g_error; }

// variant 0 = K1f (pileup_tiled_kernel), 1 = K1x (pileup_wide_kernel), 2 = K1f with kLean,
// 3 = K1w (pileup_ws_kernel, WsCfg1), 4 = K1w2 (WsCfg2).  All pointers are HOST pointers;
// `counts` is int32 [KDL_NCOL][n_slots]; tile_index is scratch of 8 words per tile of the whole slot space.
// Returns 0, or 1 with emu_last_error() set.
int emu_pileup(const kdl_batch* batch, int32_t* counts, long long n_slots, uint32_t* tile_index, long long tile_lo,
               long long n_tiles, int variant, int fresh, int grid) {
    g_error[0] = 0;
    if (n_tiles <= 0) return 0;
    const kdl_batch b = *batch;
    const unsigned idx_grid = (unsigned)((n_tiles * 32 + 255) / 256);
    const char* err = emu::launch(idx_grid, 256, [&] { kdl::tile_index_kernel(b, tile_lo, n_tiles, tile_index); });
    if (!err) {
        const unsigned threads = variant >= 3 ? (unsigned)kdl::W_THREADS : (unsigned)kdl::F_THREADS;
        err = emu::launch((unsigned)grid, threads, [&] {
            if (variant == 3) {
                if (fresh) kdl::pileup_ws_kernel<true, kdl::WsCfg1>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
                else kdl::pileup_ws_kernel<false, kdl::WsCfg1>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
            } else if (variant == 4) {
                if (fresh) kdl::pileup_ws_kernel<true, kdl::WsCfg2>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
                else kdl::pileup_ws_kernel<false, kdl::WsCfg2>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
            } else if (variant == 0) {
                if (fresh) kdl::pileup_tiled_kernel<true>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
                else kdl::pileup_tiled_kernel<false>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
            } else if (variant == 2) {
                if (fresh) kdl::pileup_tiled_kernel<true, true>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
                else kdl::pileup_tiled_kernel<false, true>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
            } else {
                if (fresh) kdl::pileup_wide_kernel<true>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
                else kdl::pileup_wide_kernel<false>(b, counts, n_slots, tile_index, tile_lo, n_tiles);
            }
        });
    }
    if (err) {
        snprintf(g_error, sizeof g_error, "%s", err);
        return 1;
    }
    return 0;
}

// ---- self-test of the emulator: a bulk copy must not be visible before its mbarrier completes -----------
namespace {
struct alignas(16) SelfSmem {
    uint32_t data[64];
    uint64_t bar;
};
int selftest(bool wait) {
    static uint32_t src[64] __attribute__((aligned(16)));
    static int mismatches;
    for (int i = 0; i < 64; ++i) src[i] = 0x1000u + (uint32_t)i;
    mismatches = 0;
    const char* err = emu::launch(1, 64, [&] {
        SelfSmem& sm = *reinterpret_cast<SelfSmem*>(kdl::smem_raw);
        const int tid = (int)threadIdx.x;
        if (tid == 0) kdl::mbar_init(&sm.bar, 1);
        __syncthreads();
        if (tid == 0) {
            kdl::mbar_expect_tx(&sm.bar, sizeof sm.data);
            kdl::bulk_g2s(sm.data, src, sizeof sm.data, &sm.bar);
        }
        if (wait) kdl::mbar_wait(&sm.bar, 0);
        if (sm.data[tid] != src[tid]) atomicAdd(&mismatches, 1);
        __syncthreads();
    });
    if (err) return -1;
    return mismatches ? 1 : 0;
}
}  // namespace

int emu_selftest_missing_wait() { return selftest(false); }
int emu_selftest_missing_wait_fixed() { return selftest(true); }

}  // extern "C"
