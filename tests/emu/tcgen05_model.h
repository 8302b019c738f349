// tcgen05_model.h -- TEST INFRASTRUCTURE: a FUNCTIONAL host model of the Blackwell pieces that csrc/ncc_tc.cu reaches
// through its PTX wrappers (the section "PTX wrappers" of that file is left out when the kernels are compiled for the
// host and these definitions take its place, name for name):
//   * shared-memory addresses   offsets into the block's dynamic shared memory (emu_dyn_smem);
//   * mbarriers                 arrival count + transaction bytes + phase parity; a waiting fiber is parked until the phase flips;
//   * cp.async.bulk             memcpy, then complete_tx on the barrier;
//   * tensor memory             128 lanes x 512 columns of 32 bits per CTA; tcgen05.ld.32x32b.x16;
//   * tcgen05.mma kind::i8      D[m][n] (+)= sum_k A[m][k] * B[n][k], u8 x u8 -> 32 bit, K = 32, operands fetched through
//                               K-major no-swizzle descriptors (8 x 16-byte core matrices; LBO between K blocks, SBO
//                               between 8-row groups) exactly as addressed -- aliased layouts included;
//   * tcgen05.commit            arrives at once: the model executes an MMA when it is issued, which is one of the orders
//                               the hardware may take.
// It checks the LOGIC of the kernels (tile geometry, Toeplitz slabs, descriptor arithmetic, barrier protocol, epilogue
// index maps), not timing, and not whether the real hardware reads a descriptor the way this file does: that is what the
// -m gpu parity tests are for.
#pragma once
#include <cstdio>
#include <map>

static inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const uint8_t*>(p) - emu_dyn_smem); }
static inline uint8_t* emu_smem_ptr(uint32_t a) { return emu_dyn_smem + a; }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
struct EmuMbar { uint32_t expected = 0; int64_t pending = 0, tx = 0; volatile uint32_t phase = 0; };
static std::map<uint32_t, EmuMbar> emu_mbars;
static long long emu_mma_count = 0, emu_bulk_bytes = 0;

static inline void emu_mbar_settle(EmuMbar& b)
{
    if (b.pending == 0 && b.tx == 0) { b.phase = b.phase ^ 1u; b.pending = b.expected; }
}
static inline void mbar_init(uint64_t* bar, uint32_t count)
{
    EmuMbar b; b.expected = count; b.pending = count;
    emu_mbars[smem_u32(bar)] = b;
}
static inline EmuMbar& emu_mbar_at(uint64_t* bar)
{
    auto it = emu_mbars.find(smem_u32(bar));
    if (it == emu_mbars.end()) { fprintf(stderr, "mbarrier at %u used before mbarrier.init\n", smem_u32(bar)); abort(); }
    return it->second;
}
static inline void mbar_arrive(uint64_t* bar)
{
    emu_preempt_point();
    EmuMbar& b = emu_mbar_at(bar);
    if (b.pending <= 0) { fprintf(stderr, "mbarrier at %u: more arrivals than its count\n", smem_u32(bar)); abort(); }
    b.pending--; emu_mbar_settle(b);
}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    EmuMbar& b = emu_mbar_at(bar);
    b.tx += bytes; b.pending--; emu_mbar_settle(b);
}
static inline void emu_mbar_complete_tx(uint64_t* bar, uint32_t bytes)
{
    EmuMbar& b = emu_mbar_at(bar);
    b.tx -= bytes; emu_mbar_settle(b);
}
// parks the fiber until the phase with this parity has completed (a wait that can never end is reported by the scheduler)
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    EmuMbar& b = emu_mbar_at(bar);
    emu_wait_change(&b.phase, parity & 1u);
    return b.phase != (parity & 1u);
}
// mbarrier.test_wait: one poll, never parks
static inline bool mbar_try_wait_once(uint64_t* bar, uint32_t parity)
{
    emu_preempt_point();
    return emu_mbar_at(bar).phase != (parity & 1u);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// elect.sync: the lanes of the warp meet here before one of them is chosen.  The kernels' warp-uniform loops (every lane
// polls the same barriers, one elected lane issues) rely on that: a lane may not fall a whole barrier phase behind.
static inline bool elect_one()
{
    if (emu_coop) emu_gate_arrive_and_wait(emu_block->warps[emu_warp].gate);
    return emu_lane == 0;
}
static inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    if ((bytes & 15u) || (smem_u32(dst_smem) & 15u) || (reinterpret_cast<uintptr_t>(src_gmem) & 15u)) {
        fprintf(stderr, "cp.async.bulk: size / addresses must be multiples of 16\n"); abort();
    }
    emu_preempt_point();
    if (getenv("EMU_TC_TRACE")) fprintf(stderr, "[block %u] bulk copy %u bytes -> smem %u, barrier %u\n", blockIdx.x, bytes, smem_u32(dst_smem), smem_u32(bar));
    memcpy(dst_smem, src_gmem, bytes);
    emu_bulk_bytes += bytes;
    emu_mbar_complete_tx(bar, bytes);
}

// ---- TMA tensor copies ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled / cp.async.bulk.tensor.2d for the one shape the library uses: a u8 image of `H` rows of `pitch`
// bytes, box = 16 bytes x box_rows rows, no swizzle, out-of-bounds elements read as zero; the barrier receives the bytes
// of the WHOLE box.  Destinations must be 128-byte aligned, coordinates are in elements (bytes, rows).
struct CUtensorMap { const uint8_t* base; int64_t pitch; int H; int box_rows; };
static long long emu_tma_boxes = 0;
static bool encode_tile_map(CUtensorMap* out, const uint8_t* base, int64_t pitch, int H, int box_rows)
{
    if (getenv("EMU_NO_TMA")) return false;                  // lets a test drive the register-staging path
    if (box_rows < 1 || box_rows > 256 || (pitch & 15) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    out->base = base; out->pitch = pitch; out->H = H; out->box_rows = box_rows;
    return true;
}
static inline void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int x, int y, uint64_t* bar)
{
    if (smem_u32(dst_smem) & 127u) { fprintf(stderr, "cp.async.bulk.tensor: shared-memory destination %u is not 128-byte aligned\n", smem_u32(dst_smem)); abort(); }
    emu_preempt_point();
    uint8_t* d = static_cast<uint8_t*>(dst_smem);
    for (int r = 0; r < tmap->box_rows; ++r)
        for (int b = 0; b < 16; ++b) {
            const int64_t gx = (int64_t)x + b, gy = (int64_t)y + r;
            d[r * 16 + b] = (gx >= 0 && gx < tmap->pitch && gy >= 0 && gy < tmap->H) ? tmap->base[gy * tmap->pitch + gx] : (uint8_t)0;
        }
    emu_tma_boxes++;
    emu_mbar_complete_tx(bar, (uint32_t)tmap->box_rows * 16u);
}

// setmaxnreg (warpgroup register reallocation): nothing to model
template <uint32_t N> static inline void reg_release() {}
template <uint32_t N> static inline void reg_acquire() {}

// ---- tensor memory ------------------------------------------------------------------------------------------------
static uint32_t emu_tmem[128][512];
static uint32_t emu_tmem_cols = 0;
static inline void tmem_alloc(uint32_t* dst_smem, uint32_t ncols)
{
    if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) { fprintf(stderr, "tcgen05.alloc: %u columns\n", ncols); abort(); }
    if (emu_lane == 0) { emu_tmem_cols = ncols; memset(emu_tmem, 0xA5, sizeof(emu_tmem)); *dst_smem = 0u; }
}
static inline void tmem_dealloc(uint32_t, uint32_t ncols) { if (emu_lane == 0 && ncols != emu_tmem_cols) { fprintf(stderr, "tcgen05.dealloc: column count differs from the allocation\n"); abort(); } }
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void fence_async_smem() {}
static inline void prefetch_l1(const void*) {}
static inline long long clock64() { return 0; }
static inline void __trap() { fprintf(stderr, "__trap()\n"); abort(); }

// ncc_tc.cu's n1_to_float (PTX: mul.wide.u32 x 2, sub.s64, cvt.rn.f32.s64): same integers, round to nearest even
static inline float n1_to_float(uint32_t area, uint32_t cc, uint32_t s, uint32_t sumT)
{
    return (float)((long long)((unsigned long long)area * cc) - (long long)((unsigned long long)s * sumT));
}

static inline uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// byte (row, k) of a K-major no-swizzle operand: 8 rows x 16 bytes per core matrix
static inline uint8_t emu_operand_byte(uint64_t desc, int row, int k)
{
    const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFFu) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
    return *emu_smem_ptr(start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k >> 4) * lbo + (uint32_t)(k & 15));
}
static inline const uint8_t* emu_operand_row16(uint64_t desc, int row, int k)
{
    const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFFu) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
    return emu_smem_ptr(start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k >> 4) * lbo);
}
static inline void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    emu_preempt_point();
    const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
    const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xFFFFu;
    if (M != 128 || N < 16 || N > 256 || (N & 15) || lane0 != 0 || col0 + (uint32_t)N > emu_tmem_cols || ((idesc >> 4) & 3u) != 2u) {
        fprintf(stderr, "tcgen05.mma: bad instruction descriptor / accumulator (M %d N %d lane %u col %u of %u)\n", M, N, lane0, col0, emu_tmem_cols); abort();
    }
    alignas(32) uint8_t A[128][32], B[256][32];
    for (int m = 0; m < M; ++m) for (int k = 0; k < 32; k += 16) memcpy(&A[m][k], emu_operand_row16(a_desc, m, k), 16);   // a core-matrix row is 16 contiguous bytes
    for (int n = 0; n < N; ++n) for (int k = 0; k < 32; k += 16) memcpy(&B[n][k], emu_operand_row16(b_desc, n, k), 16);
    for (int m = 0; m < M; ++m) {
        uint16_t a16[32];
        for (int k = 0; k < 32; ++k) a16[k] = A[m][k];
        for (int n = 0; n < N; ++n) {
            uint32_t s = 0;
            for (int k = 0; k < 32; ++k) s += (uint32_t)a16[k] * (uint32_t)B[n][k];
            emu_tmem[m][col0 + n] = (accumulate ? emu_tmem[m][col0 + n] : 0u) + s;
        }
    }
    emu_mma_count++;
}
static inline void umma_commit(uint64_t* bar)
{
    if (getenv("EMU_TC_TRACE")) fprintf(stderr, "[block %u] commit -> barrier %u after %lld MMAs\n", blockIdx.x, smem_u32(bar), emu_mma_count);
    mbar_arrive(bar);
}        // every MMA issued so far has already executed
static inline void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    emu_preempt_point();
    const uint32_t lane = (taddr >> 16) + (uint32_t)emu_lane, col = taddr & 0xFFFFu;
    if (lane >= 128 || col + 16 > 512) { fprintf(stderr, "tcgen05.ld outside tensor memory\n"); abort(); }
    if ((taddr >> 16) != 32u * (uint32_t)(emu_warp & 3)) { fprintf(stderr, "tcgen05.ld: warp %d may only read lanes %d..%d\n", emu_warp, 32 * (emu_warp & 3), 32 * (emu_warp & 3) + 31); abort(); }
    for (int j = 0; j < 16; ++j) v[j] = emu_tmem[lane][col + j];
}
static void emu_tc_reset() { emu_mbars.clear(); emu_mma_count = 0; emu_bulk_bytes = 0; }
