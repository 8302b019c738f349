"""CPU tests of device code: every kernel of csrc/ is compiled for the host -- source text taken from the .cu files as it is, apart
from `extern __shared__` declarations (a pointer handed over by the launcher) and two `fence.mbarrier_init` lines -- behind
tests/emu/cuda_runtime.h (one fiber per CUDA thread of a block: barriers, shuffles, votes, shared memory, atomics) and, for the
tcgen05 kernels of ncc_tc.cu, tests/emu/tcgen05_model.h (a functional model of mbarriers, bulk copies, tensor memory and tcgen05.mma
kind::i8 that takes the place of the file's PTX wrappers).  Without a GPU this checks, mostly bit for bit:
  * transform.cu, window_stats.cu, ncc_direct.cu, ncc_points.cu, peaks.cu, nms.cu against numpy / oracle/ (the golden vectors of
    the unmodified reference are checked through the whole-library host build, tests/test_library_emulation.py);
  * ncc_tc.cu (persistent and one-tile kernels, modes A / B, both epilogues, the 16-bit accumulate flavour), ncc_float.cu (float32 and
    masked matching) against the oracle;
  * the experiment-knob kernels (row-walking and box-sum window moments) against the default moment kernel.
TEST INFRASTRUCTURE: the emulation is a checker of kernel LOGIC (one legal order of execution, no timing), not a CPU path of the
product; whether sm_100a hardware agrees with the model is what the -m gpu tests establish."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multitemplatematching-python_b200", "csrc")


def _function(src, start):
    """Source text of the function whose definition starts with ``start`` (up to its closing brace)."""
    i = src.index(start)
    depth = 0
    for k in range(src.index("{", i), len(src)):
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            return src[i:k + 1]
    raise ValueError(start)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    xf = open(os.path.join(CSRC, "transform.cu")).read()
    tc = open(os.path.join(CSRC, "ncc_tc.cu")).read()
    ws = open(os.path.join(CSRC, "window_stats.cu")).read()
    pt = open(os.path.join(CSRC, "ncc_points.cu")).read()
    pk = open(os.path.join(CSRC, "peaks.cu")).read()
    dr = open(os.path.join(CSRC, "ncc_direct.cu")).read()
    dyn = "extern __shared__ uint32_t smem[];"                                # the one edit: dynamic shared memory comes from the launcher
    assert dr.count(dyn) == 1
    dr = dr.replace(dyn, "uint32_t* smem = reinterpret_cast<uint32_t*>(emu_dyn_smem);")
    nm = open(os.path.join(CSRC, "nms.cu")).read()
    bm = open(os.path.join(CSRC, "box_moments.cu")).read()
    tcn = tc[tc.index("namespace {"):tc.index("}  // namespace") + 1]         # all device code of ncc_tc.cu ...
    a, b = tcn.index("// ---------------------------------------------------------------- PTX wrappers"), tcn.index("// Normalise 16 consecutive")
    tcn = tcn[:a] + tcn[b:]                                                   # ... minus its PTX wrappers: tests/emu/tcgen05_model.h stands in
    for text, count, repl in (('__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }', 1, ""),
                              ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', 2, ""),
                              ("extern __shared__ __align__(1024) uint8_t smem[];", 3, "uint8_t* smem = emu_dyn_smem;")):
        assert tcn.count(text) == count, text
        tcn = tcn.replace(text, repl)
    assert "asm" not in tcn
    fl = open(os.path.join(CSRC, "ncc_float.cu")).read()
    dynf = "extern __shared__ float smemf[];"
    assert fl.count(dynf) == 1
    fl = fl.replace(dynf, "float* smemf = reinterpret_cast<float*>(emu_dyn_smem);")
    a, b = fl.index("template <int C, int TT>\nint launch_f("), fl.index("// ------------------------------------------------------------------ masked matching")
    fl_dev = fl[fl.index("namespace {"):a] + fl[b:fl.index("}  // namespace") + 1]            # device code of ncc_float.cu without its launchers
    a = fl.index("namespace {", fl.index("// ------------------------------------------------------------------ 16-bit images"))
    fl_dev += "\n" + fl[a:fl.index("}  // namespace", a) + 1]                                 # ... and the 16-bit kernels
    tc_host = tc[tc.index("struct TcEnv {"):tc.index("int launch_toeplitz_prep(")]       # experiment knobs + tc_plan_group (plain host code)
    parts = ['#include "cuda_runtime.h"', '#include "tcgen05_model.h"', '#include "mtm_internal.cuh"', '#include "ncc_epilogue.cuh"',
             xf[xf.index("namespace {"):xf.index("}  // namespace") + 1],          # all device code of transform.cu
             ws[ws.index("namespace {"):ws.index("}  // namespace") + 1],          # ... of window_stats.cu
             pt[pt.index("namespace {"):pt.index("}  // namespace") + 1],          # ... of ncc_points.cu
             pk[pk.index("namespace {"):pk.index("}  // namespace") + 1],          # ... of peaks.cu
             dr[dr.index("namespace {"):dr.index("template <int C, int TT>\nint launch_one(")] + "}",   # ... of ncc_direct.cu
             nm[nm.index("namespace {"):nm.index("}  // namespace") + 1],          # ... of nms.cu
             bm[bm.index("namespace {"):bm.index("}  // namespace") + 1],          # ... of box_moments.cu
             tcn, tc_host, fl_dev,
             r'''
extern "C" int emu_sizeof_tmplmeta() { return (int)sizeof(TmplMeta); }
extern "C" void emu_transform(const uint8_t* src, uint8_t* dst, const XformDesc* descs, int n_out, int C, int dtype, int f, int grid_x)
{
    const float scale = 1.f / (float)(f * f);
    dim3 g(grid_x, n_out), b(256);
    if (dtype == MTM_U8) emu_launch(g, b, [&] { transform_kernel<uint8_t>(src, dst, descs, C, f, scale); });
    else if (dtype == MTM_U16) emu_launch(g, b, [&] { transform_kernel<uint16_t>(src, dst, descs, C, f, scale); });
    else emu_launch(g, b, [&] { transform_kernel<float>(src, dst, descs, C, f, scale); });
}
extern "C" void emu_moments(int rows_form, const uint32_t* sat_s, const uint32_t* sat_q32, int64_t pitch, int64_t plane, const SizeDesc* sizes,
                            int n_sizes, int C, uint32_t* S, float* rsD, int64_t mom_plane, int gx, int gy)
{
    SatView sv{sat_s, nullptr, pitch, plane};
    dim3 b(256);
    if (!rows_form) {
        emu_launch(dim3(gx, n_sizes), b, [&] { window_moments_kernel<false>(sv, sat_q32, sizes, S, rsD, C, mom_plane); });
        return;
    }
    dim3 g(gx, gy, n_sizes);
    if (C == 1) emu_launch(g, b, [&] { window_moments_rows_kernel<1>(sv, sat_q32, sizes, S, rsD, mom_plane); });
    else if (C == 3) emu_launch(g, b, [&] { window_moments_rows_kernel<3>(sv, sat_q32, sizes, S, rsD, mom_plane); });
    else emu_launch(g, b, [&] { window_moments_rows_kernel<4>(sv, sat_q32, sizes, S, rsD, mom_plane); });
}
// launch_build_sat (window_stats.cu) with the same grids
extern "C" void emu_build_sat(const uint8_t* img, int64_t pitch, int H, int W, int C, uint32_t* scratch, uint32_t* sat_s,
                              unsigned long long* sat_q, uint32_t* sat_q32, int64_t sat_pitch)
{
    const int SP = (W + 3) / 4 * 4;
    dim3 g1(H), b1(256);
    switch (C) {
        case 1: emu_launch_coop(g1, b1, [&] { sat_rows_c1_kernel(img, pitch, H, W, SP, scratch); }); break;
        case 2: emu_launch_coop(g1, b1, [&] { sat_rows_kernel<2>(img, pitch, H, W, SP, scratch); }); break;
        case 3: emu_launch_coop(g1, b1, [&] { sat_rows_kernel<3>(img, pitch, H, W, SP, scratch); }); break;
        default: emu_launch_coop(g1, b1, [&] { sat_rows_kernel<4>(img, pitch, H, W, SP, scratch); }); break;
    }
    emu_launch_coop(dim3((W + 1 + 31) / 32, C + 1), dim3(32, 32), [&] { sat_cols_kernel(scratch, H, W, C, SP, sat_s, sat_q, sat_q32, sat_pitch); });
}
extern "C" void emu_tmpl_stats(const uint8_t* tmpl, TmplMeta* meta, int n, int C)
{
    emu_launch_coop(dim3(n), dim3(1024), [&] { tmpl_stats_kernel(tmpl, meta, C); });
}
extern "C" void emu_ncc_points(const uint8_t* img, int64_t pitch, int H, int C, const uint32_t* sat_s, const unsigned long long* sat_q,
                               int64_t sat_pitch, const uint8_t* tmpl, const TmplMeta* meta, const int32_t* order, int count,
                               int npos, float* maps, int method)
{
    PointsParams p{};
    p.img = img; p.pitch = pitch;
    p.sat.s = sat_s; p.sat.q = sat_q; p.sat.pitch = sat_pitch; p.sat.plane = (int64_t)(H + 1) * sat_pitch;
    p.tmpl = tmpl; p.meta = meta; p.order = order; p.maps = maps; p.method = method;
    dim3 g(npos, count), b(PT_THREADS);
    if (C == 1) emu_launch_coop(g, b, [&] { ncc_points_kernel<1>(p); });
    else if (C == 3) emu_launch_coop(g, b, [&] { ncc_points_kernel<3>(p); });
    else emu_launch_coop(g, b, [&] { ncc_points_kernel<4>(p); });
}
// The post-map half of mtm_find_matches / mtm_match_templates (mtm_api.cu) with launch_peaks' grids: raw peaks -> block A,
// then the one-launch route (finalize_small_kernel) or, when it declines (> 1024 raw hits) or `force_general`, the
// multi-launch route (sort_hits_kernel x2 + nms_kernel).  n_cand >= 0: the candidate-list route (verify_candidates_kernel).
// Returns the hit count; *where = 0 -> hits in block A, 1 -> block B; *route = 0 one launch, 1 general.
extern "C" int emu_postprocess(const TmplMeta* meta, int nt, const float* maps, int method, long long n_object, double thr,
                               double max_overlap, int do_nms, int force_general, const DevHit* cand, int n_cand, int cand_cap,
                               int cap, uint8_t* blockA, uint8_t* blockB, int32_t* nontrivial, unsigned long long* best,
                               int32_t* keep, uint8_t* mirror, int* where, int* route)
{
    const int minimize = method_is_min(method) ? 1 : 0;
    const int ascending = (method == MTM_TM_SQDIFF_NORMED) ? 1 : 0;
    const float thr32 = (float)thr;
    const float thr_nms = ascending ? (float)(1.0 - thr) : thr32;
    int32_t* countA = reinterpret_cast<int32_t*>(blockA); DevHit* hitsA = reinterpret_cast<DevHit*>(blockA + MTM_HIT_HEADER);
    int32_t* countB = reinterpret_cast<int32_t*>(blockB); DevHit* hitsB = reinterpret_cast<DevHit*>(blockB + MTM_HIT_HEADER);
    memset(countA, 0, MTM_HIT_HEADER);
    int64_t max_px = 0;
    bool any2d = false, any1d = false;
    for (int t = 0; t < nt; ++t) {
        max_px = std::max<int64_t>(max_px, (int64_t)meta[t].mh * meta[t].mw);
        if (meta[t].mh == 1 || meta[t].mw == 1) any1d = true; else any2d = true;
    }
    int bx = (int)((max_px + 4095) / 4096);
    if (bx < 1) bx = 1;
    if (n_object == 1) {
        memset(best, 0, nt * sizeof(unsigned long long));
        emu_launch_coop(dim3(bx, nt), dim3(256), [&] { argbest_kernel(meta, maps, minimize, best); });
        emu_launch(dim3((nt + 127) / 128), dim3(128), [&] { emit_best_kernel(meta, nt, maps, best, hitsA, countA, cap); });
    } else if (n_cand >= 0) {
        int32_t cc = n_cand;
        emu_launch(dim3(4), dim3(64), [&] { verify_candidates_kernel(meta, nt, maps, cand, &cc, cand_cap, hitsA, cap, countA, nontrivial); });
    } else {
        memset(nontrivial, 0, nt * sizeof(int32_t));
        const float t32 = minimize ? -thr32 : thr32;
        const double t64 = minimize ? -thr : thr;
        if (any2d) emu_launch_coop(dim3(bx, nt), dim3(256), [&] { peaks2d_kernel(meta, maps, t32, minimize, hitsA, cap, countA, nontrivial, 0); });
        if (any1d) emu_launch(dim3((nt + 63) / 64), dim3(64), [&] { peaks1d_kernel(meta, nt, maps, t32, t64, minimize, hitsA, cap, countA, nontrivial); });
    }
    *route = 0;
    int declined = force_general;
    if (!force_general) {
        const int check_trivial = n_object != 1, presorted = n_object == 1;
        if (do_nms) {
            memset(countB, 0, MTM_HIT_HEADER);
            const long long no = n_object; const float mo = (float)max_overlap;
            if (presorted) emu_launch_coop(dim3(1), dim3(256), [&] { finalize_small_kernel<true, true>(hitsA, cap, countA, meta, nontrivial, minimize, check_trivial, hitsB, countB, thr_nms, ascending, no, mo, mirror, 0); });
            else emu_launch_coop(dim3(1), dim3(256), [&] { finalize_small_kernel<false, true>(hitsA, cap, countA, meta, nontrivial, minimize, check_trivial, hitsB, countB, thr_nms, ascending, no, mo, mirror, 0); });
            declined = countB[2];
        } else if (n_object != 1) {
            emu_launch_coop(dim3(1), dim3(256), [&] { finalize_small_kernel<false, false>(hitsA, cap, countA, meta, nontrivial, minimize, 1, hitsB, countB, 0.f, 0, -1ll, 0.f, mirror, 0); });
            declined = countA[2];
        }
    }
    if (declined == 2) return -2;                               // candidate list overflow: the caller streams the maps instead
    if (declined) {
        *route = 1;
        if (countA[0] > cap) return -1;                         // the library grows the hit blocks and starts over
        if (n_object != 1) {
            emu_launch_coop(dim3(1), dim3(1024), [&] { sort_hits_kernel(hitsA, cap, countA, meta, nontrivial, 0, minimize, 0, 1, 0); });
            if (do_nms) emu_launch_coop(dim3(1), dim3(1024), [&] { sort_hits_kernel(hitsA, cap, countA, meta, nontrivial, 1, minimize, ascending, 0, 0); });
        }
        if (do_nms) emu_launch_coop(dim3(1), dim3(1024), [&] { nms_kernel(hitsA, cap, countA, hitsB, countB, keep, thr_nms, ascending, (long long)n_object, (float)max_overlap); });
    }
    *where = do_nms ? 1 : 0;
    return do_nms ? countB[0] : countA[0];
}
// launch_ncc_direct (ncc_direct.cu) restated: same template tile / chunk / shared-memory pitch choices and grid.
extern "C" int emu_ncc_direct(const uint8_t* img, int64_t pitch, int H, int W, int C, const uint32_t* sat_s, const unsigned long long* sat_q,
                              int64_t sat_pitch, const uint8_t* tmpl, const TmplMeta* meta, const int32_t* order, int count,
                              float* maps, int method, int force_wide)
{
    const TmplMeta& m0 = meta[order[0]];
    DirectParams p{};
    p.img = img; p.pitch = pitch; p.H = H; p.W = W;
    p.sat.s = sat_s; p.sat.q = sat_q; p.sat.pitch = sat_pitch; p.sat.plane = (int64_t)(H + 1) * sat_pitch;
    p.tmpl = tmpl; p.meta = meta; p.order = order; p.maps = maps;
    p.count = count; p.h = m0.h; p.w = m0.w; p.wp = m0.wp; p.mh = m0.mh; p.mw = m0.mw; p.method = method;
    int TT = count >= 8 ? 8 : count >= 4 ? 4 : count >= 2 ? 2 : 1;
    if (C > 1 && TT > 4) TT = 4;
    int TW = (BX * C + p.wp) / 4 + 4;
    TW = ((TW + 15) / 16) * 16 + 8;
    const int CH = p.h < 16 ? p.h : 16;
    p.CH = CH; p.TW = TW;
    p.wide = (force_wide || (double)p.h * p.w * C * 65025.0 >= 4294967296.0) ? 1 : 0;
    const size_t smem = (size_t)(BY + CH) * TW * 4 + (size_t)TT * CH * p.wp;
    std::vector<uint8_t> dyn(smem + 64, 0xCD);
    emu_dyn_smem = dyn.data();
    dim3 grid((p.mw + BX - 1) / BX, (p.mh + BY - 1) / BY, (count + TT - 1) / TT), block(NTHREADS);
#define EMU_DIRECT(CC) \
    switch (TT) { \
        case 1: emu_launch_coop(grid, block, [&] { ncc_direct_u8_kernel<CC, 1>(p); }); break; \
        case 2: emu_launch_coop(grid, block, [&] { ncc_direct_u8_kernel<CC, 2>(p); }); break; \
        case 4: emu_launch_coop(grid, block, [&] { ncc_direct_u8_kernel<CC, 4>(p); }); break; \
        default: emu_launch_coop(grid, block, [&] { ncc_direct_u8_kernel<CC, 8>(p); }); break; \
    }
    if (C == 1) { EMU_DIRECT(1) } else if (C == 3) { EMU_DIRECT(3) } else if (C == 4) { EMU_DIRECT(4) } else return -1;
#undef EMU_DIRECT
    emu_dyn_smem = nullptr;
    return TT;
}
// launch_box_moments (box_moments.cu): grid = (strips of the widest map, bands, sizes); the number of bands is the caller's (the
// library picks it from the SM count).
extern "C" int emu_box_moments(const uint8_t* img, int64_t pitch, int C, const SizeDesc* sizes, int n_sizes, uint32_t* S, float* rsD,
                               int64_t mom_plane, int bands, int y_begin, int rows)
{
    BoxParams p{};
    p.img = img; p.pitch = pitch; p.sizes = sizes; p.S = S; p.rsD = rsD; p.mom_plane = mom_plane;
    p.y_begin = y_begin; p.rows = rows;
    int strips = 1;
    for (int k = 0; k < n_sizes; ++k) {
        const int strip_out = (BM_COLS - (sizes[k].w - 1)) & ~3;
        strips = std::max(strips, (sizes[k].mw + strip_out - 1) / strip_out);
    }
    if (C == 1 && bands < 0) {                              // the throughput form of the single-channel kernel (strips of 2048 columns)
        int strips1 = 1;
        for (int k = 0; k < n_sizes; ++k) {
            const int strip_out = (B1_COLS - (sizes[k].w - 1)) & ~15;
            strips1 = std::max(strips1, (sizes[k].mw + strip_out - 1) / strip_out);
        }
        bool lean = getenv("EMU_BOX_GENERAL_LOOP") == nullptr;      // as launch_box_moments: the lean output loop when every window is <= 256 px wide
        for (int k = 0; k < n_sizes; ++k) lean = lean && sizes[k].w <= B1_THREADS;
        if (lean && getenv("EMU_BOX_UNROLLED") == nullptr) emu_launch_coop(dim3(strips1, -bands, n_sizes), dim3(B1_THREADS), [&] { box_moments_c1_kernel<true, true>(p); });
        else if (lean) emu_launch_coop(dim3(strips1, -bands, n_sizes), dim3(B1_THREADS), [&] { box_moments_c1_kernel<true>(p); });
        else emu_launch_coop(dim3(strips1, -bands, n_sizes), dim3(B1_THREADS), [&] { box_moments_c1_kernel<false>(p); });
        return strips1;
    }
    const dim3 grid(strips, bands, n_sizes), block(BM_THREADS);
    if (C == 1) emu_launch_coop(grid, block, [&] { box_moments_kernel<1>(p); });
    else if (C == 3) emu_launch_coop(grid, block, [&] { box_moments_kernel<3>(p); });
    else if (C == 4) emu_launch_coop(grid, block, [&] { box_moments_kernel<4>(p); });
    else return -1;
    return strips;
}
// One tcgen05 launch of a template group, as launch_toeplitz_prep + launch_ncc_tc_impl (ncc_tc.cu) set it up; the tile height N,
// the ring (ds, stages), the number of epilogue warps and the number of persistent CTAs are the caller's (the library derives
// them from a clock model; every valid choice must give the same maps).  persist = 0: the one-tile-per-CTA kernel.
// method 5 -> MODE 0 (integer / fp32 epilogue on the window moments), else MODE 1 (float64 epilogue on the tables).
extern "C" long long emu_ncc_tc(const uint8_t* img, int64_t pitch, int H, int W, int C, const uint8_t* tmpl, const TmplMeta* meta,
                                const int32_t* order, int count, int mode, int h, int w, int h_min, int w_min,
                                const uint32_t* S, const float* rsD, int64_t mom_plane, const uint32_t* sat_s,
                                const unsigned long long* sat_q, int64_t sat_pitch, float* maps, int method, int N, int stages, int ds,
                                int EW, int persist, int ctas, DevHit* cand, int32_t* cand_count, int cand_cap, float cand_thr,
                                int y_base, int rows, int band_rows)
{
    TcGroup g{};
    if (!tc_plan_group(mode, h, w, C, g)) return -1;
    std::vector<uint8_t> slab_store((size_t)h * g.slab_bytes + 64);
    uint8_t* slabs = slab_store.data() + ((64 - (reinterpret_cast<uintptr_t>(slab_store.data()) & 63)) & 63);
    const int64_t pieces = (int64_t)h * (g.slab_bytes / 16);
    emu_launch(dim3((unsigned)std::min<int64_t>((pieces + 255) / 256, 4096)), dim3(256),
               [&] { toeplitz_prep_kernel(tmpl, meta, order, count, mode, h, w, g.nk, g.slab_bytes, C, slabs, nullptr); });
    TcParams p{};
    p.method = method;
    p.sat = SatView{sat_s, sat_q, sat_pitch, (int64_t)(H + 1) * sat_pitch};
    p.img = img; p.pitch = pitch; p.H = H; p.W = W;
    p.slabs = slabs; p.slab_bytes = g.slab_bytes; p.a_kblk = g.a_kblk; p.nk = g.nk;
    p.mode = mode; p.h = h; p.w = w; p.mh = H - h_min + 1; p.mw = W - w_min + 1;
    p.meta = meta; p.order = order; p.count = count; p.S = S; p.rsD = rsD; p.maps = maps; p.C = C; p.mom_plane = mom_plane;
    if (cand) { p.cand = cand; p.cand_count = cand_count; p.cand_cap = cand_cap; p.cand_thr = cand_thr; }
    const int xw = mode == 0 ? 16 : 128, gx = (p.mw + xw - 1) / xw;
    p.y_base = y_base; p.rows = std::min(rows, p.mh - y_base); p.band_rows = band_rows;      // one band of output rows per launch
    p.N = N; p.R = tc_tile_rows(N, h);
    CUtensorMap tmap{};                                     // image tiles through the (modelled) TMA unit unless EMU_NO_TMA is set
    p.tma_chunks = (p.R + 255) / 256;
    p.tma_rc = ((p.R + p.tma_chunks - 1) / p.tma_chunks + 7) & ~7;
    p.tma = encode_tile_map(&tmap, img, pitch, H, p.tma_rc) ? 1 : 0;
    static uint8_t smem_store[227 * 1024 + 1024];
    emu_dyn_smem = smem_store + ((1024 - (reinterpret_cast<uintptr_t>(smem_store) & 1023)) & 1023);
    memset(emu_dyn_smem, 0xCD, 227 * 1024);
    emu_tc_reset();
    const size_t tile_b = ((size_t)2 * g.nk * p.R * 16 + 127) & ~(size_t)127;
    const int kmode = method != MTM_TM_CCOEFF_NORMED ? 1 : 0;
    if (persist) {
        p.ds = std::max(1, std::min(h, ds)); p.stages = stages;
        if (stages < 2 || stages > TCP_MAX_STAGES || 256 + 2 * tile_b + (size_t)stages * p.ds * g.slab_bytes > 227 * 1024) return -2;
        p.tiles_x = gx; p.tiles_total = gx * ((p.rows + N - 1) / N);
        const dim3 grid((unsigned)std::min(p.tiles_total, ctas)), block(32 * (EW + 3));
        if (EW == 12) { if (kmode) emu_launch_coop(grid, block, [&] { ncc_tc_persist_kernel<false, 12, 1>(p, tmap); }); else emu_launch_coop(grid, block, [&] { ncc_tc_persist_kernel<false, 12, 0>(p, tmap); }); }
        else { if (kmode) emu_launch_coop(grid, block, [&] { ncc_tc_persist_kernel<false, 8, 1>(p, tmap); }); else emu_launch_coop(grid, block, [&] { ncc_tc_persist_kernel<false, 8, 0>(p, tmap); }); }
    } else {
        p.ds = g.ds;
        if (tile_b + (size_t)TC_STAGES * g.ds * g.slab_bytes + 256 > 227 * 1024) return -2;
        const dim3 grid(gx, (p.rows + N - 1) / N), block(TC_THREADS);
        if (kmode) emu_launch_coop(grid, block, [&] { ncc_tc_kernel<1>(p, tmap); }); else emu_launch_coop(grid, block, [&] { ncc_tc_kernel<0>(p, tmap); });
    }
    emu_dyn_smem = nullptr;
    return emu_mma_count;
}
// float32 branch (ncc_float.cu): launch_build_sat_f32 + launch_tmpl_stats_f32 + launch_ncc_direct_f32 restated for one group of
// equal-sized templates.  `arena` holds the packed float32 templates, `centred` receives their mean-centred copies.
static int emu_f32_launch(const FloatParams& p0, int C, int count, int we4)
{
    FloatParams p = p0;
    const int BXf = (C == 1) ? FBX : 16;
    int TT = count >= 4 ? 4 : count >= 2 ? 2 : 1;
    int TW = BXf * C + we4 + 8;
    TW = ((TW + 31) / 32) * 32 + 8;
    const int CH = p.h < 8 ? p.h : 8;
    p.CH = CH; p.TW = TW;
    const size_t smem = ((size_t)(FBY + CH) * TW + (size_t)TT * CH * we4) * sizeof(float);
    if (smem > 200 * 1024) return -1;
    std::vector<uint8_t> dyn(smem + 64, 0xCD);
    emu_dyn_smem = dyn.data();
    dim3 grid((p.mw + BXf - 1) / BXf, (p.mh + FBY - 1) / FBY, (count + TT - 1) / TT), block(FTHREADS);
#define EMU_F32(CC) \
    switch (TT) { \
        case 1: emu_launch_coop(grid, block, [&] { ncc_direct_f32_kernel<CC, 1>(p); }); break; \
        case 2: emu_launch_coop(grid, block, [&] { ncc_direct_f32_kernel<CC, 2>(p); }); break; \
        default: emu_launch_coop(grid, block, [&] { ncc_direct_f32_kernel<CC, 4>(p); }); break; \
    }
    if (C == 1) { EMU_F32(1) } else if (C == 3) { EMU_F32(3) } else if (C == 4) { EMU_F32(4) } else return -1;
#undef EMU_F32
    emu_dyn_smem = nullptr;
    return TT;
}
extern "C" void emu_satf(const float* img, int64_t pitch_e, int H, int W, int C, double* scratch, double* sat_s, double* sat_q, int64_t sat_pitch)
{
    if (C == 1) emu_launch_coop(dim3(H), dim3(256), [&] { satf_rows_kernel<1>(img, pitch_e, H, W, scratch); });
    else if (C == 3) emu_launch_coop(dim3(H), dim3(256), [&] { satf_rows_kernel<3>(img, pitch_e, H, W, scratch); });
    else emu_launch_coop(dim3(H), dim3(256), [&] { satf_rows_kernel<4>(img, pitch_e, H, W, scratch); });
    emu_launch_coop(dim3((W + 1 + 31) / 32, C + 1), dim3(32, 32), [&] { satf_cols_kernel(scratch, H, W, C, sat_s, sat_q, sat_pitch); });
}
extern "C" int emu_ncc_f32(const float* img, int64_t pitch_e, int H, int W, int C, const double* sat_s, const double* sat_q, int64_t sat_pitch,
                           const float* arena, float* centred, TmplMeta* meta, int n_tmpl, int stats, const int32_t* order, int count,
                           float* maps, int method)
{
    if (stats) emu_launch_coop(dim3(n_tmpl), dim3(256), [&] { tmplf_stats_kernel(arena, centred, meta, C); });
    const TmplMeta& m0 = meta[order[0]];
    FloatParams p{};
    p.img = img; p.pitch_e = pitch_e; p.H = H; p.W = W;
    p.sat_s = sat_s; p.sat_q = sat_q; p.sat_pitch = sat_pitch; p.sat_plane = (int64_t)(H + 1) * sat_pitch;
    p.centred = (method == MTM_TM_CCOEFF || method == MTM_TM_CCOEFF_NORMED) ? 1 : 0;
    p.tmpl = p.centred ? centred : arena;
    p.meta = meta; p.order = order; p.maps = maps;
    p.count = count; p.h = m0.h; p.w = m0.w; p.mh = m0.mh; p.mw = m0.mw; p.method = method;
    return emu_f32_launch(p, C, count, (m0.w * C + 3) & ~3);
}
// masked matching, methods 0 / 3 (compute_maps_masked in mtm_api.cu): T*M^2 and M^2 arenas, image and image^2 in float32, two plain
// correlations, the combine kernel.  `img8` != nullptr: uint8 image (pitch in bytes), else the float image already in `imgf`.
extern "C" int emu_masked(const uint8_t* img8, int64_t pitch8, float* imgf, float* imgf2, int64_t pitch_e, int H, int W, int C,
                          const uint8_t* raw_t, const uint8_t* raw_m, int is_f32, float* tm2, float* m2, TmplMeta* meta, int n_tmpl,
                          const int32_t* order, int count, float* mapsA, float* mapsB, int method)
{
    emu_launch_coop(dim3(n_tmpl), dim3(256), [&] { masked_prep_kernel(raw_t, raw_m, is_f32, tm2, m2, meta, C); });
    emu_launch(dim3(4), dim3(256), [&] { u8_to_f32_sq_kernel(img8, pitch8, imgf, imgf, imgf2, pitch_e, H, W * C); });
    const TmplMeta& m0 = meta[order[0]];
    FloatParams p{};
    p.pitch_e = pitch_e; p.H = H; p.W = W; p.centred = 0;
    p.meta = meta; p.order = order; p.count = count; p.h = m0.h; p.w = m0.w; p.mh = m0.mh; p.mw = m0.mw; p.method = MTM_TM_CCORR;
    const int we4 = (m0.w * C + 3) & ~3;
    p.img = imgf; p.tmpl = tm2; p.maps = mapsA;
    if (emu_f32_launch(p, C, count, we4) < 0) return -1;
    p.img = imgf2; p.tmpl = m2; p.maps = mapsB;
    if (emu_f32_launch(p, C, count, we4) < 0) return -1;
    emu_launch(dim3(2, n_tmpl), dim3(256), [&] { masked_combine_kernel(mapsA, mapsB, meta, method); });
    return 0;
}
// The 16-bit grayscale route (compute_maps with tensor16 in mtm_api.cu): u16_split_image_kernel -> float32 image + high / low byte
// planes; float64 tables; float32 template statistics; Toeplitz slabs of both template byte planes; four accumulate launches
// (MODE = 2) CC = 65536 hh + 256 (hl + lh) + ll into the float64 map; cc16_epilogue_kernel.  One template group.
extern "C" long long emu_ncc_tc16(const uint16_t* src, int H, int W, float* pixf, int64_t pitch_e, uint8_t* hi, uint8_t* lo, int64_t pitch,
                                  double* scratch, double* sat_s, double* sat_q, int64_t sat_pitch, const float* arena, float* centred,
                                  TmplMeta* meta, const uint8_t* tmpl8, int64_t tmpl8_plane, const TmplPix8* pix8, const int32_t* order,
                                  int count, int mode, int h, int w, int h_min, int w_min, double* acc, float* maps, int method,
                                  int N, int stages, int ds, int ctas, int epilogue_only)
{
    int64_t n_px = 0;
    for (int t = 0; t < count; ++t) n_px = std::max<int64_t>(n_px, (int64_t)meta[t].mh * meta[t].mw);
    if (epilogue_only) {                                    // the numerator map does not depend on the method
        emu_launch(dim3((unsigned)std::max<int64_t>(1, (n_px + 255) / 256), (unsigned)count), dim3(256),
                   [&] { cc16_epilogue_kernel(acc, maps, meta, 0, method, sat_s, sat_q, sat_pitch); });
        return 1;
    }
    emu_launch(dim3((unsigned)std::min(8, (W + 255) / 256), (unsigned)H), dim3(256),
               [&] { u16_split_image_kernel(src, (int64_t)W, H, W, pixf, pitch_e, hi, lo, pitch); });
    emu_satf(pixf, pitch_e, H, W, 1, scratch, sat_s, sat_q, sat_pitch);
    emu_launch_coop(dim3(count), dim3(256), [&] { tmplf_stats_kernel(arena, centred, meta, 1); });
    TcGroup g{};
    if (!tc_plan_group(mode, h, w, 1, g)) return -1;
    std::vector<uint8_t> slab_store((size_t)2 * h * g.slab_bytes + 64);
    uint8_t* slabs = slab_store.data() + ((64 - (reinterpret_cast<uintptr_t>(slab_store.data()) & 63)) & 63);
    const int64_t slab_plane = (int64_t)h * g.slab_bytes;
    const int64_t pieces = (int64_t)h * (g.slab_bytes / 16);
    for (int plane = 0; plane < 2; ++plane)
        emu_launch(dim3((unsigned)std::min<int64_t>((pieces + 255) / 256, 4096)), dim3(256),
                   [&] { toeplitz_prep_kernel(tmpl8 + plane * tmpl8_plane, meta, order, count, mode, h, w, g.nk, g.slab_bytes, 1, slabs + plane * slab_plane, pix8); });
    static uint8_t smem_store[227 * 1024 + 1024];
    long long mmas = 0;
    const int planes[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    const double weights[4] = {65536.0, 256.0, 256.0, 1.0};
    for (int k = 0; k < 4; ++k) {
        TcParams p{};
        p.method = MTM_TM_CCORR;
        p.img = planes[k][0] ? lo : hi; p.pitch = pitch; p.H = H; p.W = W;
        p.slabs = slabs + planes[k][1] * slab_plane; p.slab_bytes = g.slab_bytes; p.a_kblk = g.a_kblk; p.nk = g.nk;
        p.mode = mode; p.h = h; p.w = w; p.mh = H - h_min + 1; p.mw = W - w_min + 1;
        p.meta = meta; p.order = order; p.count = count; p.maps = maps; p.C = 1;
        p.acc = acc; p.acc_weight = weights[k]; p.acc_first = k == 0 ? 1 : 0;
        const int xw = mode == 0 ? 16 : 128, gx = (p.mw + xw - 1) / xw;
        p.y_base = 0; p.rows = p.mh; p.band_rows = p.mh;
        p.N = N; p.R = tc_tile_rows(N, h); p.ds = std::max(1, std::min(h, ds)); p.stages = stages;
        CUtensorMap tmap{};
        p.tma_chunks = (p.R + 255) / 256;
        p.tma_rc = ((p.R + p.tma_chunks - 1) / p.tma_chunks + 7) & ~7;
        p.tma = encode_tile_map(&tmap, p.img, pitch, H, p.tma_rc) ? 1 : 0;
        p.tiles_x = gx; p.tiles_total = gx * ((p.mh + N - 1) / N);
        const size_t tile_b = ((size_t)2 * g.nk * p.R * 16 + 127) & ~(size_t)127;
        if (256 + 2 * tile_b + (size_t)stages * p.ds * g.slab_bytes > 227 * 1024) return -2;
        emu_dyn_smem = smem_store + ((1024 - (reinterpret_cast<uintptr_t>(smem_store) & 1023)) & 1023);
        memset(emu_dyn_smem, 0xCD, 227 * 1024);
        emu_tc_reset();
        emu_launch_coop(dim3((unsigned)std::min(p.tiles_total, ctas)), dim3(32 * 11), [&] { ncc_tc_persist_kernel<false, 8, 2>(p, tmap); });
        mmas += emu_mma_count;
    }
    emu_dyn_smem = nullptr;
    emu_launch(dim3((unsigned)std::max<int64_t>(1, (n_px + 255) / 256), (unsigned)count), dim3(256),
               [&] { cc16_epilogue_kernel(acc, maps, meta, 0, method, sat_s, sat_q, sat_pitch); });
    return mmas;
}
''']
    d = tmp_path_factory.mktemp("emu")
    (d / "emu.cpp").write_text("\n".join(parts))
    lib = d / "libemu.so"
    r = subprocess.run([gxx, "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "tests", "emu"), "-I", CSRC,
                        "-I", os.path.join(ROOT, "include"), str(d / "emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    return ctypes.CDLL(str(lib))


XFORM_DTYPE = np.dtype([("src_off", "<i8"), ("src_pitch", "<i8"), ("dst_off", "<i8"), ("dst_pitch", "<i8"),
                        ("dh", "<i4"), ("dw", "<i4"), ("oh", "<i4"), ("ow", "<i4"), ("op", "<i4"), ("pad", "<i4")])
SIZE_DTYPE = np.dtype([("h", "<i4"), ("w", "<i4"), ("mh", "<i4"), ("mw", "<i4"), ("off", "<i8"), ("band", "<i4"), ("pad", "<i4")])


def _mom_segment(mw, band):
    """Entries of a size's moment segment (csrc/mtm_internal.cuh, SizeDesc): tile-major, `band` rows."""
    return (mw + 15) // 16 * 16 * band


def _mom_index(x, r, band):
    return ((x >> 4) * band + r) * 16 + (x & 15)
OPS = ["identity", "rot90", "rot180", "rot270", "fliplr", "flipud", "transpose", "antitranspose"]


@pytest.mark.parametrize("dtype,code", [(np.uint8, 0), (np.uint16, 2), (np.float32, 1)])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_transform_kernel_source_on_the_host(emu, dtype, code, channels):
    from oracle import augment_port as ap
    rng = np.random.default_rng(3)
    for f in (1, 2, 3, 4, 7, 16):
        h, w = 5 * f + (f - 1), 6 * f + 1
        shape = (h, w) + ((channels,) if channels > 1 else ())
        src = (rng.random(shape) * 255).astype(np.float32) if dtype == np.float32 else \
            rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
        small = ap.area_downscale(src, f)
        want = [np.ascontiguousarray(ap.HOST_TRANSFORMS[t](small)) for t in OPS]
        item = src.dtype.itemsize
        descs = np.zeros(len(OPS), XFORM_DTYPE)
        off = 0
        for k, wnt in enumerate(want):
            pitch = (wnt.shape[1] * channels * item + 3) // 4 * 4
            descs[k] = (0, w * channels * item, off, pitch, small.shape[0], small.shape[1], wnt.shape[0], wnt.shape[1], k, 0)
            off += (pitch * wnt.shape[0] + 15) // 16 * 16
        dst = np.full(off, 0xAB, np.uint8)
        src_c = np.ascontiguousarray(src)
        emu.emu_transform(ctypes.c_void_p(src_c.ctypes.data), ctypes.c_void_p(dst.ctypes.data), ctypes.c_void_p(descs.ctypes.data),
                          len(OPS), channels, code, f, 2)
        for k, wnt in enumerate(want):
            d = descs[k]
            rows = [dst[int(d["dst_off"]) + r * int(d["dst_pitch"]): int(d["dst_off"]) + r * int(d["dst_pitch"]) + wnt.shape[1] * channels * item]
                    .view(dtype).reshape(wnt.shape[1:]) for r in range(wnt.shape[0])]
            got = np.stack(rows)
            if dtype == np.float32:
                assert np.max(np.abs(got - wnt)) <= 1e-6 * 255, (f, OPS[k])
            else:
                assert np.array_equal(got, wnt), (f, OPS[k])


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_row_walking_moment_kernel_equals_the_default_one(emu, channels):
    rng = np.random.default_rng(11)
    H, W = 70, 93
    pitch = (W + 1 + 3) // 4 * 4
    plane = (H + 1) * pitch
    img = rng.integers(0, 256, (H, W, channels)).astype(np.int64)
    img[10:40, 20:70] = 77                                                    # flat windows: the rsD = 0 rule
    sat_s = np.zeros((channels, H + 1, pitch), np.uint32)
    for c in range(channels):
        sat_s[c, 1:, 1:W + 1] = np.cumsum(np.cumsum(img[:, :, c], axis=0), axis=1).astype(np.uint32)
    sat_q = np.zeros((H + 1, pitch), np.uint32)
    sat_q[1:, 1:W + 1] = (np.cumsum(np.cumsum((img ** 2).sum(axis=2), axis=0), axis=1) & 0xFFFFFFFF).astype(np.uint32)
    sizes = np.zeros(3, SIZE_DTYPE)
    off = 0
    for k, (h, w) in enumerate([(5, 9), (17, 16), (32, 40)]):
        sizes[k] = (h, w, H - h + 1, W - w + 1, off, H - h + 1, 0)              # whole maps: one band
        off += _mom_segment(W - w + 1, H - h + 1)
    outs = []
    for rows_form in (0, 1):
        S = np.full(off * max(2, channels), 0xDEADBEEF, np.uint32)
        R = np.full(off, -1.0, np.float32)
        emu.emu_moments(rows_form, ctypes.c_void_p(sat_s.ctypes.data), ctypes.c_void_p(sat_q.ctypes.data), ctypes.c_int64(pitch),
                        ctypes.c_int64(plane), ctypes.c_void_p(sizes.ctypes.data), 3, channels, ctypes.c_void_p(S.ctypes.data),
                        ctypes.c_void_p(R.ctypes.data), ctypes.c_int64(off), 3 if not rows_form else 1, 7)
        outs.append((S, R))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
    # and the numbers mean what they should: window sum of channel 0 of the first size, and rsD == 0 on a flat window
    h, w = 5, 9
    win = img[:h, :w, 0].sum()
    first = outs[0][0][0] if channels > 1 else outs[0][0].view(np.uint32)[0]
    assert int(first) == int(win)
    y, x = 12, 25                                                             # a flat 5 x 9 window inside the constant patch
    idx = _mom_index(x, y, H - h + 1)
    rs = outs[0][1][idx] if channels > 1 else outs[0][0].view(np.float32)[2 * idx + 1]
    assert rs == 0.0


TMPL_META_DTYPE = np.dtype([("map_off", "<i8"), ("mh", "<i4"), ("mw", "<i4"), ("mom_off", "<i8"), ("pix_off", "<i8"),
                            ("h", "<i4"), ("w", "<i4"), ("wp", "<i4"), ("is_const", "<i4"), ("mean", "<f8", 4), ("sum2", "<f8"),
                            ("norm_ccoeff", "<f8"), ("norm_plain", "<f8"), ("inv_area", "<f8"), ("isum", "<i8", 4),
                            ("inv_sqrt_d2", "<f4"), ("pad_f", "<f4")])


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _host_sat(emu, img):
    """launch_build_sat's two kernels on the host: (sat_s [C][H+1][pitch] u32, sat_q [H+1][pitch] u64, sat_q32, pitch)."""
    H, W = img.shape[:2]
    C = 1 if img.ndim == 2 else img.shape[2]
    ipitch = (W * C + 3) // 4 * 4
    buf = np.full(H * ipitch + 16, 0xEE, np.uint8)                           # padding bytes are garbage on purpose
    rows = buf[:H * ipitch].reshape(H, ipitch)
    rows[:, :W * C] = img.reshape(H, W * C)
    SP = (W + 3) // 4 * 4
    pitch = (W + 1 + 3) // 4 * 4
    scratch = np.full((C + 1) * H * SP, 0xDEADBEEF, np.uint32)
    sat_s = np.full((C, H + 1, pitch), 0xDEADBEEF, np.uint32)
    sat_q = np.full((H + 1, pitch), 0xDEADBEEFDEADBEEF, np.uint64)
    sat_q32 = np.full((H + 1, pitch), 0xDEADBEEF, np.uint32)
    emu.emu_build_sat(_ptr(buf), ctypes.c_int64(ipitch), H, W, C, _ptr(scratch), _ptr(sat_s), _ptr(sat_q), _ptr(sat_q32),
                      ctypes.c_int64(pitch))
    return buf, ipitch, sat_s, sat_q, sat_q32, pitch


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(37, 45), (5, 301), (70, 3)])
def test_summed_area_kernels_on_the_host(emu, channels, shape):
    """sat_rows_*_kernel + sat_cols_kernel (window_stats.cu) == numpy cumulative sums: u32 tables modulo 2^32, u64 table of
    squares, zero first row and column, the low words in sat_q32."""
    rng = np.random.default_rng(5)
    H, W = shape
    img = rng.integers(0, 256, (H, W, channels)).astype(np.uint8)
    img[:H // 2] = 255                                                       # large sums early
    _, _, sat_s, sat_q, sat_q32, _ = _host_sat(emu, img if channels > 1 else img[:, :, 0])
    wide = img.astype(np.uint64)
    for c in range(channels):
        want = np.zeros((H + 1, W + 1), np.uint64)
        want[1:, 1:] = wide[:, :, c].cumsum(0).cumsum(1)
        assert np.array_equal(sat_s[c, :, :W + 1], (want & 0xFFFFFFFF).astype(np.uint32)), c
    wq = np.zeros((H + 1, W + 1), np.uint64)
    wq[1:, 1:] = (wide * wide).sum(2).cumsum(0).cumsum(1)
    assert np.array_equal(sat_q[:, :W + 1], wq)
    assert np.array_equal(sat_q32[:, :W + 1], (wq & 0xFFFFFFFF).astype(np.uint32))


def _pack_templates(templates, channels):
    """The template arena of mtm_set_templates: rows padded with zeros to a multiple of 4 bytes, 16-byte aligned starts."""
    meta = np.zeros(len(templates), TMPL_META_DTYPE)
    chunks, off = [], 0
    for k, t in enumerate(templates):
        h, w = t.shape[:2]
        wp = (w * channels + 3) // 4 * 4
        rows = np.zeros((h, wp), np.uint8)
        rows[:, :w * channels] = t.reshape(h, w * channels)
        meta[k]["pix_off"], meta[k]["h"], meta[k]["w"], meta[k]["wp"] = off, h, w, wp
        size = (h * wp + 15) // 16 * 16
        chunks.append(np.concatenate([rows.ravel(), np.zeros(size - h * wp, np.uint8)]))
        off += size
    return np.concatenate(chunks + [np.zeros(16, np.uint8)]), meta


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_template_statistics_kernel_on_the_host(emu, channels):
    """tmpl_stats_kernel (window_stats.cu): the word/byte-mask sums give OpenCV's meanStdDev-derived constants exactly as
    the oracle's epilogue derives them, and flag constant templates."""
    assert emu.emu_sizeof_tmplmeta() == TMPL_META_DTYPE.itemsize
    rng = np.random.default_rng(9)
    tmpls = [rng.integers(0, 256, (h, w, channels)).astype(np.uint8) for h, w in [(7, 5), (33, 41), (1, 1), (64, 130)]]
    tmpls.append(np.full((9, 11, channels), 200, np.uint8))                  # constant: is_const, inv_sqrt_d2 = 0
    arena, meta = _pack_templates(tmpls, channels)
    emu.emu_tmpl_stats(_ptr(arena), _ptr(meta), len(tmpls), channels)
    for k, t in enumerate(tmpls):
        m = meta[k]
        T = t.reshape(-1, channels).astype(np.int64)
        area = t.shape[0] * t.shape[1]
        s, q = T.sum(0), (T * T).sum(0)
        assert list(m["isum"][:channels]) == list(s) and not m["isum"][channels:].any()
        inv_area = 1.0 / float(area)
        mean = s.astype(np.float64) * inv_area
        var = np.maximum(q.astype(np.float64) * inv_area - mean * mean, 0.0)
        norm = 0.0
        mean2 = 0.0
        for c in range(channels):                                            # same summation order as the kernel
            norm += var[c]
            mean2 += mean[c] * mean[c]
        sum2 = norm + mean2
        assert m["inv_area"] == inv_area and np.array_equal(m["mean"][:channels], mean)
        assert m["is_const"] == (1 if norm < np.finfo(np.float64).eps else 0)
        assert m["sum2"] == sum2 / inv_area
        assert m["norm_ccoeff"] == np.sqrt(norm) / np.sqrt(inv_area)
        assert m["norm_plain"] == np.sqrt(sum2) / np.sqrt(inv_area)
        d2 = int((area * q - s * s).sum())
        assert m["inv_sqrt_d2"] == (np.float32(1.0 / np.sqrt(float(d2))) if d2 > 0 else np.float32(0))
    assert meta[-1]["is_const"] == 1 and meta[-1]["inv_sqrt_d2"] == 0 and meta[0]["is_const"] == 0


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_small_map_kernel_on_the_host(emu, channels):
    """ncc_points_kernel (ncc_points.cu) end to end on the host -- summed-area tables and template statistics from their
    own kernels, then one CTA per output pixel -- against the oracle's exact score maps.  Covers the
    funnel-shift rebuild of unaligned image words (every x offset modulo 4) and rows that are not a whole number of
    words (zero-padded template words against live image bytes)."""
    from oracle import ncc_exact
    rng = np.random.default_rng(21)
    H, W = 23, 30
    img = rng.integers(0, 256, (H, W, channels)).astype(np.uint8)
    image = img if channels > 1 else img[:, :, 0]
    shapes = [(21, 25), (20, 27), (23, 30)]                                  # maps of 3x6, 4x4 and 1x1 positions
    tmpls = []
    for h, w in shapes:
        y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        t = img[y0:y0 + h, x0:x0 + w].astype(np.int64) + rng.integers(-20, 21, (h, w, channels))   # a noisy crop: scores near 1
        tmpls.append(np.clip(t, 0, 255).astype(np.uint8))
    buf, ipitch, sat_s, sat_q, _, pitch = _host_sat(emu, image)
    arena, meta = _pack_templates(tmpls, channels)
    off = 0
    for k, (h, w) in enumerate(shapes):
        meta[k]["mh"], meta[k]["mw"], meta[k]["map_off"] = H - h + 1, W - w + 1, off
        off += ((H - h + 1) * (W - w + 1) + 31) // 32 * 32
    emu.emu_tmpl_stats(_ptr(arena), _ptr(meta), len(tmpls), channels)
    order = np.array([2, 0, 1], np.int32)
    for method in ((0, 1, 2, 3, 4, 5) if channels == 1 else (1, 5)):          # the epilogue is shared with the dp4a kernel (all six there)
        maps = np.full(off, np.nan, np.float32)
        emu.emu_ncc_points(_ptr(buf), ctypes.c_int64(ipitch), H, channels, _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(pitch),
                           _ptr(arena), _ptr(meta), _ptr(order), 3, 7, _ptr(maps), method)      # 7 CTAs stride over 18 positions
        for k, t in enumerate(tmpls):
            mh, mw = int(meta[k]["mh"]), int(meta[k]["mw"])
            got = maps[int(meta[k]["map_off"]):int(meta[k]["map_off"]) + mh * mw].reshape(mh, mw)
            want = ncc_exact.match_template_exact(image, t if channels > 1 else t[:, :, 0], method)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (method, k, np.abs(got - want).max())


DEVHIT_DTYPE = np.dtype([("tmpl", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("score", "<f4"), ("seq", "<i4"),
                         ("key", "<f4")])


def _host_postprocess(emu, maps, sizes, method, n_object, thr, max_overlap, do_nms, force_general=False, candidates=None,
                      cand_cap=32768, cap=4096):
    """Peak extraction [+ sort + NMS] kernels on the host for a list of score maps; returns ([(tmpl, (x, y, w, h), score)], route)."""
    meta = np.zeros(len(maps), TMPL_META_DTYPE)
    off = 0
    for k, (m, (h, w)) in enumerate(zip(maps, sizes)):
        meta[k]["mh"], meta[k]["mw"], meta[k]["map_off"], meta[k]["h"], meta[k]["w"] = m.shape[0], m.shape[1], off, h, w
        off += (m.size + 31) // 32 * 32
    flat = np.full(off + 32, np.nan, np.float32)
    for k, m in enumerate(maps):
        flat[int(meta[k]["map_off"]):int(meta[k]["map_off"]) + m.size] = m.ravel()
    block_a = np.zeros(32 + 32 * cap, np.uint8)
    block_b = np.zeros(32 + 32 * cap, np.uint8)
    nontrivial = np.zeros(len(maps), np.int32)
    best = np.zeros(len(maps), np.uint64)
    keep = np.zeros(cap, np.int32)
    mirror = np.zeros(32 + 32 * 256, np.uint8)
    where, route = ctypes.c_int(0), ctypes.c_int(0)
    cand = np.zeros(1, DEVHIT_DTYPE) if candidates is None else candidates
    n = emu.emu_postprocess(_ptr(meta), len(maps), _ptr(flat), method, ctypes.c_longlong(-1 if n_object == float("inf") else n_object),
                            ctypes.c_double(thr), ctypes.c_double(max_overlap), int(do_nms), int(force_general), _ptr(cand),
                            -1 if candidates is None else len(candidates), cand_cap, cap, _ptr(block_a), _ptr(block_b),
                            _ptr(nontrivial), _ptr(best), _ptr(keep), _ptr(mirror), ctypes.byref(where), ctypes.byref(route))
    assert n >= 0, n
    hits = (block_b if where.value else block_a)[32:32 + 32 * n].view(DEVHIT_DTYPE)
    out = [(int(r["tmpl"]), (int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])), float(r["score"])) for r in hits]
    if route.value == 0 and (do_nms or n_object != 1):                         # the mapped mirror carries the same header + first 256 hits
        assert int(mirror[:4].view(np.int32)[0]) == n
        m = min(n, 256)
        assert np.array_equal(mirror[32:32 + 32 * m], (block_b if where.value else block_a)[32:32 + 32 * m])
    return out, route.value


def _port_postprocess(maps, sizes, method, n_object, thr, max_overlap, do_nms):
    """The same list through oracle/mtm_port.py (MTM/__init__.py:22-53, 222-241 and MTM/NMS.py on live cv2.dnn.NMSBoxes)."""
    import cv2
    from oracle import mtm_port
    hits = []
    for t, (m, (h, w)) in enumerate(zip(maps, sizes)):
        if n_object == 1:
            _, _, lo, hi = cv2.minMaxLoc(m)
            peaks = [(lo if method in (0, 1) else hi)[::-1]]
        elif method in (0, 1):
            peaks = mtm_port.find_local_min(m, thr)
        else:
            peaks = mtm_port.find_local_max(m, thr)
        hits += [(t, (int(p[1]), int(p[0]), w, h), float(m[tuple(p)])) for p in peaks]
    return mtm_port.nms(hits, thr, method == 1, n_object, max_overlap) if do_nms else hits


def _score_maps(rng, levels):
    """A mix of every shape class of MTM._findLocalMax_: 2-D maps with ties and plateaus (few score levels), a constant 2-D map
    above the threshold (peak_local_max: no peaks), a 1 x n row, an n x 1 column and a 1 x 1 map."""
    q = lambda a: (np.round(a * levels) / levels).astype(np.float32)
    maps = [q(rng.random((20, 30))), np.full((6, 7), 0.9, np.float32), q(rng.random((1, 40))), q(rng.random((33, 1))),
            np.array([[0.75]], np.float32), q(rng.random((17, 9))), np.array([[0.25]], np.float32)]
    maps[0][3:6, 10:13] = 1.0                                                 # a 3 x 3 plateau of the maximum
    sizes = [(12, 9), (30, 30), (8, 8), (5, 20), (64, 64), (10, 10), (3, 3)]
    return maps, sizes


@pytest.mark.parametrize("method", [1, 3, 5])
@pytest.mark.parametrize("n_object", [float("inf"), 1, 4])
def test_peak_sort_nms_kernels_on_the_host(emu, method, n_object):
    """peaks.cu + nms.cu from their source on the CPU: the findMatches list (order included) and the post-NMS list of
    MTM.matchTemplates equal the port's on maps full of ties, for the one-launch route and the general multi-launch route."""
    rng = np.random.default_rng(100 * method + (0 if n_object == float("inf") else n_object))
    for levels in ((8, 1000) if n_object == float("inf") else (8,)):
        maps, sizes = _score_maps(rng, levels)
        if method == 1:
            maps = [(1.0 - m).astype(np.float32) for m in maps]
        thr = 0.5
        want_list = _port_postprocess(maps, sizes, method, n_object, thr, 0.3, False)
        got_list, _ = _host_postprocess(emu, maps, sizes, method, n_object, thr, 0.3, False)
        assert got_list == want_list
        want = _port_postprocess(maps, sizes, method, n_object, thr, 0.3, True)
        for force_general in (False, True):
            got, route = _host_postprocess(emu, maps, sizes, method, n_object, thr, 0.3, True, force_general)
            assert route == int(force_general)
            assert got == want, (levels, force_general)
        if n_object != 1:
            got_list_g, _ = _host_postprocess(emu, maps, sizes, method, n_object, thr, 0.3, False, True)
            assert got_list_g == want_list


def test_near_tied_minimising_scores_keep_the_peak_finders_order_on_the_host(emu):
    """TM_SQDIFF_NORMED: two overlapping hits of one template whose scores differ by one ulp share the float32 NMS key 1 - score.
    The reference's stable sort keeps the peak finder's order there (ascending score first), so the later pixel with the smaller
    score wins the overlap: the one-launch route (single sort by key, template, then score) must agree with the general route."""
    s_hi = np.float32(0.25)
    s_lo = np.nextafter(s_hi, np.float32(0))
    assert np.float32(1) - s_hi == np.float32(1) - s_lo and s_lo < s_hi
    m = np.full((24, 40), 0.9, np.float32)
    m[2, 20] = s_hi                                  # first in row-major order, worse score
    m[6, 18] = s_lo                                  # later, better score: overlaps the first box
    m[15, 5] = 0.125
    maps, sizes = [m, np.full((9, 9), 0.95, np.float32)], [(12, 9), (5, 5)]
    for n_object in (float("inf"), 2):
        want = _port_postprocess(maps, sizes, 1, n_object, 0.5, 0.1, True)
        assert [h[1][:2] for h in want[:2]] == [(5, 15), (18, 6)], want
        for force_general in (False, True):
            got, route = _host_postprocess(emu, maps, sizes, 1, n_object, 0.5, 0.1, True, force_general)
            assert route == int(force_general) and got == want, (n_object, force_general, got, want)


def test_candidate_list_route_on_the_host(emu):
    """verify_candidates_kernel: the pixels above the threshold that the tcgen05 epilogue lists (any order) give the same hit lists as
    the streaming pass; a list that overflowed (a constant map above the threshold always does: the route needs maps larger than the
    list) is reported with header[2] = 2 so that the library streams the maps instead."""
    rng = np.random.default_rng(77)
    maps = [(np.round(rng.random((24, 30)) * 16) / 16).astype(np.float32), rng.random((26, 31)).astype(np.float32)]
    sizes = [(9, 9), (14, 6)]
    thr = 0.6
    cand = []
    for t, m in enumerate(maps):
        ys, xs = np.nonzero(m > np.float32(thr))
        cand += [(t, int(x), int(y), sizes[t][1], sizes[t][0], float(m[y, x]), 0, 0.0) for y, x in zip(ys, xs)]
    cand = np.array([cand[i] for i in rng.permutation(len(cand))], DEVHIT_DTYPE)          # epilogue warps append in any order
    assert 0 < len(cand) < 700
    for do_nms in (False, True):
        want = _port_postprocess(maps, sizes, 5, float("inf"), thr, 0.25, do_nms)
        got, _ = _host_postprocess(emu, maps, sizes, 5, float("inf"), thr, 0.25, do_nms, candidates=cand, cand_cap=700)
        assert got == want and len(want) > 5
    flat = [np.full((24, 30), 0.9, np.float32)]
    every = np.array([(0, x, y, 9, 9, 0.9, 0, 0.0) for y in range(24) for x in range(30)], DEVHIT_DTYPE)
    with pytest.raises(AssertionError, match="-2"):
        _host_postprocess(emu, flat, [(9, 9)], 5, float("inf"), thr, 0.25, True, candidates=every, cand_cap=700)
    assert _port_postprocess(flat, [(9, 9)], 5, float("inf"), thr, 0.25, True) == []
    assert _host_postprocess(emu, flat, [(9, 9)], 5, float("inf"), thr, 0.25, True)[0] == []          # ... and the streaming pass agrees


def test_more_than_1024_raw_hits_take_the_general_route_on_the_host(emu):
    """finalize_small_kernel declines lists above 1024 raw hits; sort_hits_kernel (global-memory bitonic sorts, padding to a power of
    two) + nms_kernel then give the port's lists, including the [:N_object] cut."""
    rng = np.random.default_rng(5)
    maps = [rng.random((110, 120)).astype(np.float32), (np.round(rng.random((40, 50)) * 32) / 32).astype(np.float32)]
    sizes = [(4, 5), (7, 3)]
    want_list = _port_postprocess(maps, sizes, 5, float("inf"), 0.2, 0.0, False)
    assert len(want_list) > 1024
    got_list, route = _host_postprocess(emu, maps, sizes, 5, float("inf"), 0.2, 0.0, False)
    assert route == 1 and got_list == want_list
    for n_object in (60, 1):
        want = _port_postprocess(maps, sizes, 5, n_object, 0.2, 0.1, True)
        got, route = _host_postprocess(emu, maps, sizes, 5, n_object, 0.2, 0.1, True)
        assert got == want and route == (1 if n_object != 1 else 0)


def _host_direct_maps(emu, image, tmpls, methods, force_wide=False):
    """K1 + K2 of the library on the host: summed-area tables, template statistics (once) and ncc_direct_u8_kernel per method for a
    list of equal-sized uint8 templates; returns {method: score maps} and the template-tile width the launch chose."""
    H, W = image.shape[:2]
    C = 1 if image.ndim == 2 else image.shape[2]
    buf, ipitch, sat_s, sat_q, _, pitch = _host_sat(emu, image)
    arena, meta = _pack_templates([t.reshape(t.shape[0], t.shape[1], C) for t in tmpls], C)
    h, w = tmpls[0].shape[:2]
    mh, mw = H - h + 1, W - w + 1
    per = (mh * mw + 31) // 32 * 32
    for k in range(len(tmpls)):
        meta[k]["mh"], meta[k]["mw"], meta[k]["map_off"] = mh, mw, k * per
    emu.emu_tmpl_stats(_ptr(arena), _ptr(meta), len(tmpls), C)
    order = np.arange(len(tmpls), dtype=np.int32)
    out = {}
    for method in methods:
        maps = np.full(per * len(tmpls), np.nan, np.float32)
        tt = emu.emu_ncc_direct(_ptr(buf), ctypes.c_int64(ipitch), H, W, C, _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(pitch), _ptr(arena),
                                _ptr(meta), _ptr(order), len(tmpls), _ptr(maps), method, int(force_wide))
        assert tt > 0
        out[method] = [maps[k * per:k * per + mh * mw].reshape(mh, mw) for k in range(len(tmpls))]
    return out, tt


@pytest.mark.parametrize("channels,count", [(1, 1), (1, 3), (1, 9), (3, 2), (3, 5), (4, 4)])
def test_direct_kernel_on_the_host(emu, channels, count):
    """ncc_direct_u8_kernel (register tile of dp4a accumulators over byte-shifted image words, chunked template rows, fused float64
    epilogue) for every template-tile width TT and channel count: score maps bit-identical to the oracle's exact maps, all six
    methods; the 64-bit accumulation route gives the same bits."""
    from oracle import ncc_exact
    rng = np.random.default_rng(40 + channels * 10 + count)
    H, W, h, w = 45, 83, 19, 13                                              # 27 x 71 maps: two tiles in x, a second chunk of 3 rows
    img = rng.integers(0, 256, (H, W, channels)).astype(np.uint8)
    img[5:30, 40:70] = 128                                                   # flat windows (the epilogue's zero-variance rule)
    image = img if channels > 1 else img[:, :, 0]
    tmpls = []
    for k in range(count):
        y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        t = np.clip(img[y0:y0 + h, x0:x0 + w].astype(np.int64) + rng.integers(-30, 31, (h, w, channels)), 0, 255).astype(np.uint8)
        tmpls.append(t if channels > 1 else t[:, :, 0])
    if count >= 3:
        tmpls[1] = np.full_like(tmpls[1], 99)                                # a constant template (TM_CCOEFF_NORMED map := 1)
    got, tt = _host_direct_maps(emu, image, tmpls, range(6))
    assert tt == (1 if count == 1 else 2 if count < 4 else 4 if (count < 8 or channels > 1) else 8)
    for method in range(6):
        for k, t in enumerate(tmpls):
            want = ncc_exact.match_template_exact(image, t, method)
            assert np.array_equal(got[method][k].view(np.uint32), want.view(np.uint32)), (method, k, np.abs(got[method][k] - want).max())
    wide, _ = _host_direct_maps(emu, image, tmpls, [5], force_wide=True)
    for a, b in zip(wide[5], got[5]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("channels", [1, 3, 4])
@pytest.mark.parametrize("shape,windows,bands", [((41, 1100), [(17, 40)], 3), ((30, 333), [(30, 5)], 1), ((64, 70), [(8, 64)], 57),
                                                 ((25, 2100), [(3, 512)], 6), ((48, 1150), [(5, 9), (17, 16), (32, 200), (48, 1)], 4)])
def test_box_sum_moment_kernel_equals_the_summed_area_one(emu, channels, shape, windows, bands):
    """box_moments_kernel (experiment knob MTM_B200_MOM_BOX: window moments from running box sums, no summed-area tables) writes
    exactly what window_moments_kernel writes from the tables: several strips and bands, a one-row map, a window as wide as half a
    strip, four sizes in one launch (grid sized for the largest map), flat windows (rsD = 0), garbage in the row padding."""
    rng = np.random.default_rng(13)
    H, W = shape
    img = rng.integers(0, 256, (H, W, channels)).astype(np.uint8)
    img[H // 3:, W // 4:W // 4 + 3 * windows[0][1]] = 201                     # flat windows
    ipitch = (W * channels + 64 * channels + 64 + 127) // 128 * 128          # the library's row pitch
    buf = rng.integers(0, 256, H * ipitch + 256).astype(np.uint8)            # (the library pads with zeros; any bytes do)
    buf[:H * ipitch].reshape(H, ipitch)[:, :W * channels] = img.reshape(H, W * channels)
    wide = img.astype(np.int64)
    pitch = (W + 1 + 3) // 4 * 4
    plane = (H + 1) * pitch
    sat_s = np.zeros((channels, H + 1, pitch), np.uint32)
    for c in range(channels):
        sat_s[c, 1:, 1:W + 1] = np.cumsum(np.cumsum(wide[:, :, c], axis=0), axis=1).astype(np.uint32)
    sat_q = np.zeros((H + 1, pitch), np.uint32)
    sat_q[1:, 1:W + 1] = (np.cumsum(np.cumsum((wide ** 2).sum(axis=2), axis=0), axis=1) & 0xFFFFFFFF).astype(np.uint32)
    sizes = np.zeros(len(windows), SIZE_DTYPE)
    total = 64                                                                # the maps do not start at element 0
    for k, (h, w) in enumerate(windows):
        sizes[k] = (h, w, H - h + 1, W - w + 1, total, H - h + 1, 0)          # whole maps: one band each
        total += _mom_segment(W - w + 1, H - h + 1)
    outs = []
    for box in (0, 1, 2):                                                      # summed-area kernel, generic box kernel, throughput form (C == 1)
        if box == 2 and channels != 1:
            continue
        S = np.full(total * max(2, channels), 0xDEADBEEF, np.uint32)
        R = np.full(total, -1.0, np.float32)
        if box:
            assert emu.emu_box_moments(_ptr(buf), ctypes.c_int64(ipitch), channels, _ptr(sizes), len(sizes), _ptr(S), _ptr(R),
                                       ctypes.c_int64(total), bands if box == 1 else -bands, 0, H) >= 1
        else:
            emu.emu_moments(0, _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(pitch), ctypes.c_int64(plane), _ptr(sizes), len(sizes), channels,
                            _ptr(S), _ptr(R), ctypes.c_int64(total), 3, 1)
        outs.append((S, R))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0])
        assert np.array_equal(outs[0][1].view(np.uint32), other[1].view(np.uint32))
    h, w = windows[0]
    first = outs[1][0][64] if channels > 1 else outs[1][0][2 * 64]
    assert int(first) == int(wide[:h, :w, 0].sum())
    # the same rows band by band into a RING: every band overwrites the previous one (what the library does when the maps of a
    # template group exceed the ring budget); entry (x, r) of a band = entry (x, y_begin + r) of the whole map
    B = 16
    ring = np.zeros(len(windows), SIZE_DTYPE)
    rtotal = 32
    for k, (h, w) in enumerate(windows):
        ring[k] = (h, w, H - h + 1, W - w + 1, rtotal, B, 0)
        rtotal += _mom_segment(W - w + 1, B)
    full_S, full_R = outs[0]
    for box in (1, 2):
        if box == 2 and channels != 1:
            continue
        S = np.full(rtotal * max(2, channels), 0xDEADBEEF, np.uint32)
        R = np.full(rtotal, -1.0, np.float32)
        for y_begin in range(0, max(int(sz["mh"]) for sz in ring), B):
            emu.emu_box_moments(_ptr(buf), ctypes.c_int64(ipitch), channels, _ptr(ring), len(ring), _ptr(S), _ptr(R),
                                ctypes.c_int64(rtotal), 2 if box == 1 else -2, y_begin, B)
            for k, (h, w) in enumerate(windows):
                mh_, mw_ = H - h + 1, W - w + 1
                for y in range(y_begin, min(mh_, y_begin + B)):
                    xs = np.arange(mw_)
                    src = int(sizes[k]["off"]) + _mom_index(xs, y, mh_)
                    dst = int(ring[k]["off"]) + _mom_index(xs, y - y_begin, B)
                    if channels == 1:
                        assert np.array_equal(S.view(np.uint64)[dst], full_S.view(np.uint64)[src]), (box, k, y)
                    else:
                        for c in range(channels):
                            assert np.array_equal(S[c * rtotal + dst], full_S[c * total + src]), (box, k, y, c)
                        assert np.array_equal(R.view(np.uint32)[dst], full_R.view(np.uint32)[src]), (box, k, y)


# ---- the tcgen05 kernels on the functional model of tests/emu/tcgen05_model.h -----------------------------------------------

def _host_tensor_maps(emu, image, tmpls, method, mode, N, stages=3, ds=4, EW=8, persist=True, ctas=2, cand_thr=None, box=False, band=None):
    """ncc_tc.cu on the host for ONE template group (mode A: up to 8 templates of mixed sizes, zero padded to the largest; mode B:
    one grayscale template x 128 x-offsets): summed-area tables, template statistics, window moments, Toeplitz slabs, then the
    persistent (or the one-tile-per-CTA) kernel.  Returns (maps, number of MMAs issued, candidate records or None)."""
    H, W = image.shape[:2]
    C = 1 if image.ndim == 2 else image.shape[2]
    ipitch = (W * C + 64 * C + 64 + 127) // 128 * 128                        # mtm_set_image's tile layout: zero padded rows
    buf = np.zeros(H * ipitch + 256, np.uint8)
    buf[:H * ipitch].reshape(H, ipitch)[:, :W * C] = image.reshape(H, W * C)
    _, _, sat_s, sat_q, sat_q32, spitch = _host_sat(emu, image)
    arena, meta = _pack_templates([t.reshape(t.shape[0], t.shape[1], C) for t in tmpls], C)
    order = np.asarray(sorted(range(len(tmpls)), key=lambda k: tmpls[k].shape[:2]), np.int32)      # (h, w) order of finish_templates
    sizes, moff, off = [], 0, 0
    for k in range(len(tmpls)):
        h, w = tmpls[k].shape[:2]
        meta[k]["mh"], meta[k]["mw"], meta[k]["map_off"] = H - h + 1, W - w + 1, off
        off += ((H - h + 1) * (W - w + 1) + 31) // 32 * 32
    mh_max = H - min(t.shape[0] for t in tmpls) + 1
    band = mh_max if band is None else band                                  # band < the largest map: moments + numerator band by band (ring)
    assert band == mh_max or box, "only the box-sum kernels write bands"
    for k in order:
        h, w = tmpls[k].shape[:2]
        if not sizes or (sizes[-1][0], sizes[-1][1]) != (h, w):
            sizes.append((h, w, H - h + 1, W - w + 1, moff, band, 0))
            moff += _mom_segment(W - w + 1, band)
        meta[k]["mom_off"] = sizes[-1][4]
    emu.emu_tmpl_stats(_ptr(arena), _ptr(meta), len(tmpls), C)
    sd = np.zeros(len(sizes), SIZE_DTYPE)
    for k, v in enumerate(sizes):
        sd[k] = v
    S = np.full(moff * max(2, C), 0xDEADBEEF, np.uint32)
    R = np.full(moff, np.nan, np.float32)
    maps = np.full(off + 32, np.nan, np.float32)
    hs, ws = [t.shape[0] for t in tmpls], [t.shape[1] for t in tmpls]
    cand = np.zeros(4096, DEVHIT_DTYPE)
    cand_count = np.zeros(1, np.int32)
    emu.emu_ncc_tc.restype = ctypes.c_longlong
    n_mma = 0
    for y_base in range(0, mh_max, band):
        rows = min(band, mh_max - y_base)
        if box:
            emu.emu_box_moments(_ptr(buf), ctypes.c_int64(ipitch), C, _ptr(sd), len(sizes), _ptr(S), _ptr(R), ctypes.c_int64(moff),
                                5 if C > 1 else -3, y_base, rows)
        else:
            emu.emu_moments(0, _ptr(sat_s), _ptr(sat_q32), ctypes.c_int64(spitch), ctypes.c_int64((H + 1) * spitch), _ptr(sd), len(sizes), C,
                            _ptr(S), _ptr(R), ctypes.c_int64(moff), 2, 1)
        got = emu.emu_ncc_tc(_ptr(buf), ctypes.c_int64(ipitch), H, W, C, _ptr(arena), _ptr(meta), _ptr(order), len(tmpls), mode,
                             max(hs), max(ws), min(hs), min(ws), _ptr(S), _ptr(R), ctypes.c_int64(moff), _ptr(sat_s), _ptr(sat_q),
                             ctypes.c_int64(spitch), _ptr(maps), method, N, stages, ds, EW, int(persist), ctas,
                             _ptr(cand) if cand_thr is not None else None, _ptr(cand_count), len(cand),
                             ctypes.c_float(cand_thr if cand_thr is not None else 0.0), y_base, rows, band)
        assert got > 0, got
        n_mma += got
    out = [maps[int(m["map_off"]):int(m["map_off"]) + int(m["mh"]) * int(m["mw"])].reshape(int(m["mh"]), int(m["mw"])) for m in meta]
    return out, n_mma, (cand[:int(cand_count[0])] if cand_thr is not None else None)


def _planted(rng, H, W, C, shapes):
    img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
    img[H // 2:H // 2 + 14, W // 3:W // 3 + 40] = 90                          # a flat patch: zero-variance windows for the small templates
    tmpls = []
    for h, w in shapes:
        y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        t = np.clip(img[y0:y0 + h, x0:x0 + w].astype(np.int64) + rng.integers(-25, 26, (h, w, C)), 0, 255).astype(np.uint8)
        tmpls.append(t)
    if C == 1:
        return img[:, :, 0], [t[:, :, 0] for t in tmpls]
    return img, tmpls


@pytest.mark.parametrize("channels,shapes,N,opts", [
    (1, [(17, 20), (12, 9), (17, 20)], 32, dict()),                            # mixed sizes in one mode-A group, 12 tiles over 2 CTAs
    (1, [(17, 20), (12, 9), (17, 20)], 48, dict(EW=12, stages=2, ds=1, ctas=3)),
    (1, [(17, 20), (12, 9), (17, 20)], 32, dict(persist=False)),              # the one-tile-per-CTA kernel
    (3, [(10, 13), (10, 13)], 32, dict(stages=4, ds=3)),                      # RGB: C-byte x-step of the band, per-channel moments
    (4, [(9, 6)], 16, dict(ctas=5)),
    (1, [(8, 8)] * 8, 64, dict(ctas=1)),                                      # a full group of eight, one CTA walks every tile
])
def test_tensor_core_kernel_on_the_functional_model(emu, channels, shapes, N, opts):
    """ncc_tc_persist_kernel / ncc_tc_kernel, mode A, from their source on the CPU with tests/emu/tcgen05_model.h in place of the PTX
    wrappers: Toeplitz slabs through the bulk-copy ring, descriptor arithmetic of the row shift, double-buffered image tiles and
    accumulators, epilogue index maps.  TM_CCOEFF_NORMED maps within 2e-6 of the oracle (fp32 epilogue on exact integers); the
    float64-epilogue instantiation gives the oracle's bits for TM_CCORR (the raw numerator) and TM_CCORR_NORMED."""
    from oracle import ncc_exact
    rng = np.random.default_rng(17 + channels)
    image, tmpls = _planted(rng, 70, 90, channels, shapes)
    if len(tmpls) >= 3:
        tmpls[2] = np.full_like(tmpls[2], 140)                               # a constant template: the map := 1 rule
    got, n_mma, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, N, **opts)
    for k, t in enumerate(tmpls):
        want = ncc_exact.match_template_exact(image, t, 5)
        assert got[k].shape == want.shape and np.max(np.abs(got[k] - want)) <= 2e-6, (k, float(np.nanmax(np.abs(got[k] - want))))
    for method in (2, 3):
        got, _, _ = _host_tensor_maps(emu, image, tmpls, method, 0, N, **opts)
        for k, t in enumerate(tmpls):
            want = ncc_exact.match_template_exact(image, t, method)
            assert np.array_equal(got[k].view(np.uint32), want.view(np.uint32)), (method, k)


@pytest.mark.parametrize("opts", [dict(), dict(persist=False)])
def test_tensor_core_kernel_register_staging_when_the_tensor_map_is_refused(emu, monkeypatch, opts):
    """Image tiles normally arrive through the TMA unit (cp.async.bulk.tensor boxes of 16 bytes x up to 256 rows, modelled in
    tests/emu/tcgen05_model.h: 128-byte aligned destinations, zero fill outside the image, whole-box transaction bytes).  When
    cuTensorMapEncodeTiled is unavailable the stager warps copy the tile through registers: same maps."""
    from oracle import ncc_exact
    rng = np.random.default_rng(18)
    image, tmpls = _planted(rng, 70, 90, 1, [(17, 20), (12, 9), (17, 20)])
    with_tma, n_mma, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, 32, **opts)
    monkeypatch.setenv("EMU_NO_TMA", "1")
    without, n_mma2, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, 32, **opts)
    assert n_mma == n_mma2
    for k, t in enumerate(tmpls):
        assert np.array_equal(with_tma[k].view(np.uint32), without[k].view(np.uint32))
        assert np.max(np.abs(without[k] - ncc_exact.match_template_exact(image, t, 5))) <= 2e-6


@pytest.mark.parametrize("N,opts", [(48, dict(ctas=2)), (32, dict(persist=False)), (64, dict(EW=12, stages=5, ds=2, ctas=4))])
def test_tensor_core_kernel_mode_b_on_the_functional_model(emu, N, opts):
    """Mode B (one grayscale template x 128 x-offsets): the aliased Toeplitz slab (K blocks 256 bytes apart over 128-byte row groups,
    so consecutive K blocks re-read the next row groups) and the reversed x-offset order of the accumulator lanes."""
    from oracle import ncc_exact
    rng = np.random.default_rng(23)
    image, tmpls = _planted(rng, 60, 300, 1, [(14, 33)])                     # 268 window columns: three x tiles, the last one partial
    got, n_mma, _ = _host_tensor_maps(emu, image, tmpls, 5, 1, N, **opts)
    want = ncc_exact.match_template_exact(image, tmpls[0], 5)
    assert np.max(np.abs(got[0] - want)) <= 2e-6
    nk = (33 + 127 + 31) // 32
    assert n_mma == 3 * ((47 + N - 1) // N) * 14 * nk                        # tiles x template rows x K chunks: nothing issued twice
    got, _, _ = _host_tensor_maps(emu, image, tmpls, 2, 1, N, **opts)
    assert np.array_equal(got[0].view(np.uint32), ncc_exact.match_template_exact(image, tmpls[0], 2).view(np.uint32))


def test_tensor_core_epilogue_lists_the_candidates_and_takes_box_sum_moments(emu):
    """The default epilogue appends every pixel above the threshold to the candidate list (the peak pass then visits only those);
    with the window moments coming from box_moments_kernel instead of the summed-area tables the maps do not change by a bit."""
    rng = np.random.default_rng(29)
    image, tmpls = _planted(rng, 70, 90, 1, [(16, 16), (16, 16)])
    ref, _, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, 32)
    got, _, cand = _host_tensor_maps(emu, image, tmpls, 5, 0, 32, cand_thr=0.25, box=True)
    listed = set()
    for k in range(2):
        assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32))
        ys, xs = np.nonzero(got[k] > np.float32(0.25))
        listed |= {(k, int(x), int(y), 16, 16, float(got[k][y, x])) for y, x in zip(ys, xs)}
    assert len(listed) >= 2
    assert sorted((int(c["tmpl"]), int(c["x"]), int(c["y"]), int(c["w"]), int(c["h"]), float(c["score"])) for c in cand) == sorted(listed)
    # band by band: the moments of 16 (then 48) output rows go into a ring that the next band overwrites, each band followed by its
    # numerator launch (what the library does for template groups whose moment maps exceed the ring budget) -- same maps, mixed
    # sizes (the smaller map ends earlier than the band), bands that are not a multiple of the tile height, both kernels
    image, tmpls = _planted(rng, 70, 90, 1, [(16, 16), (12, 21), (16, 16)])
    ref, _, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, 32)
    for band, opts in ((16, dict()), (48, dict(EW=12, ctas=3)), (16, dict(persist=False))):
        got, _, _ = _host_tensor_maps(emu, image, tmpls, 5, 0, 32, box=True, band=band, **opts)
        for k in range(3):
            assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32)), (band, opts, k)


# ---- float32 branch and masked matching (ncc_float.cu) --------------------------------------------------------------------

def _f32_arena(tmpls, C):
    """mtm_set_templates for float32: rows of w*C floats, 16-byte aligned template starts."""
    meta = np.zeros(len(tmpls), TMPL_META_DTYPE)
    off = 0
    for k, t in enumerate(tmpls):
        h, w = t.shape[:2]
        meta[k]["pix_off"], meta[k]["h"], meta[k]["w"], meta[k]["wp"] = off, h, w, w * C * 4
        off += (h * w * C * 4 + 15) // 16 * 16
    arena = np.zeros(off // 4 + 16, np.float32)
    for k, t in enumerate(tmpls):
        o = int(meta[k]["pix_off"]) // 4
        arena[o:o + t.size] = t.astype(np.float32).ravel()
    return arena, meta


def _geometry(meta, tmpls, H, W):
    off = 0
    for k, t in enumerate(tmpls):
        mh, mw = H - t.shape[0] + 1, W - t.shape[1] + 1
        meta[k]["mh"], meta[k]["mw"], meta[k]["map_off"] = mh, mw, off
        off += (mh * mw + 31) // 32 * 32
    return off


def _maps_of(flat, meta):
    return [flat[int(m["map_off"]):int(m["map_off"]) + int(m["mh"]) * int(m["mw"])].reshape(int(m["mh"]), int(m["mw"])) for m in meta]


def _host_f32_maps(emu, image, tmpls, methods):
    """ncc_float.cu on the host for equal-sized float32 templates: float64 tables, template statistics (+ mean-centred copies), then
    ncc_direct_f32_kernel per method.  Returns ({method: maps}, template-tile width, tables, centred arena, meta)."""
    H, W = image.shape[:2]
    C = 1 if image.ndim == 2 else image.shape[2]
    count = len(tmpls)
    pitch_e = (W * C + 3) // 4 * 4
    img = np.zeros(H * pitch_e + 64, np.float32)
    img[:H * pitch_e].reshape(H, pitch_e)[:, :W * C] = image.reshape(H, W * C)
    spitch = (W + 1 + 3) // 4 * 4
    scratch = np.zeros(2 * (C + 1) * H * W + 16, np.float64)
    sat_s = np.full((C, H + 1, spitch), np.nan)
    sat_q = np.full((H + 1, spitch), np.nan)
    emu.emu_satf(_ptr(img), ctypes.c_int64(pitch_e), H, W, C, _ptr(scratch), _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(spitch))
    arena, meta = _f32_arena(tmpls, C)
    total = _geometry(meta, tmpls, H, W)
    centred = np.full_like(arena, np.nan)
    order = np.arange(count, dtype=np.int32)
    out, tt = {}, 0
    for n, method in enumerate(methods):
        maps = np.full(total + 32, np.nan, np.float32)
        tt = emu.emu_ncc_f32(_ptr(img), ctypes.c_int64(pitch_e), H, W, C, _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(spitch), _ptr(arena),
                             _ptr(centred), _ptr(meta), count, int(n == 0), _ptr(order), count, _ptr(maps), method)
        assert tt > 0, tt
        out[method] = _maps_of(maps, meta)
    return out, tt, sat_s, centred, meta


@pytest.mark.parametrize("channels,count", [(1, 1), (1, 5), (3, 2), (4, 1)])
def test_float32_kernels_on_the_host(emu, channels, count):
    """satf_rows / satf_cols (float64 tables), tmplf_stats_kernel (OpenCV constants + mean-centred copy) and ncc_direct_f32_kernel for
    equal-sized float32 templates: all six methods within the GPU parity bar (1e-4 of max(1, max |exact|)) of the oracle's
    exact maps; 16-bit-valued data as the reference's uint16 -> float32 cast produces it."""
    from oracle import ncc_exact
    rng = np.random.default_rng(60 + channels)
    H, W, h, w = 50, 77, 11, 14
    image8, tmpls8 = _planted(rng, H, W, channels, [(h, w)] * count)
    scale = np.float32(97.0)                                                  # 16-bit range
    image = image8.astype(np.float32) * scale
    tmpls = [t.astype(np.float32) * scale for t in tmpls8]
    if count >= 3:
        tmpls[1] = np.full_like(tmpls[1], 1234.0)                            # constant template
    got, tt, sat_s, centred, meta = _host_f32_maps(emu, image, tmpls, range(6))
    assert tt == (4 if count >= 4 else 2 if count >= 2 else 1)
    wide = image.reshape(H, W, channels).astype(np.float64)
    for c in range(channels):                                                 # the tables themselves: float64 prefix sums (exact here: integers)
        want = np.zeros((H + 1, W + 1))
        want[1:, 1:] = wide[:, :, c].cumsum(0).cumsum(1)
        assert np.array_equal(sat_s[c, :, :W + 1], want)
    for method in range(6):
        for k, t in enumerate(tmpls):
            want = ncc_exact.match_template_exact(image, t, method)
            scale_m = max(1.0, float(np.abs(want).max()))                     # the bar of tests/test_gpu_float.py
            assert np.max(np.abs(got[method][k].astype(np.float64) - want)) <= 1e-4 * scale_m, (method, k)
    for k, t in enumerate(tmpls):                                             # the mean-centred copies
        o = int(meta[k]["pix_off"]) // 4
        mean = t.reshape(-1, channels).astype(np.float64).mean(0)
        assert np.allclose(centred[o:o + t.size].reshape(-1, channels), t.reshape(-1, channels) - mean, rtol=0, atol=1e-2)


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("channels", [1, 3])
def test_masked_kernels_on_the_host(emu, dtype, channels):
    """Masked matching (methods 0 / 3; the optional third element of MTM's template tuples): masked_prep_kernel, the float image and
    its square, two plain correlations and masked_combine_kernel against the oracle's restatement of OpenCV's matchTemplateMask --
    binary uint8 masks and weighted float32 masks."""
    from oracle import ncc_exact
    rng = np.random.default_rng(80 + channels)
    H, W, h, w = 40, 61, 9, 12
    image8, tmpls8 = _planted(rng, H, W, channels, [(h, w), (h, w)])
    if dtype == np.uint8:
        image, tmpls = image8, tmpls8
        masks = [(rng.random(t.shape) > 0.3).astype(np.uint8) * rng.integers(1, 255, t.shape).astype(np.uint8) for t in tmpls]   # non-zero -> 1
    else:
        image, tmpls = image8.astype(np.float32), [t.astype(np.float32) for t in tmpls8]
        masks = [rng.random(t.shape).astype(np.float32) for t in tmpls]
    arena, meta = _f32_arena(tmpls, channels)                                  # geometry of the float arenas (T*M^2, M^2)
    total = _geometry(meta, tmpls, H, W)
    esz = np.dtype(dtype).itemsize
    raw_t = np.zeros(arena.size * esz + 64, np.uint8)
    raw_m = np.zeros(arena.size * esz + 64, np.uint8)
    for k, (t, m) in enumerate(zip(tmpls, masks)):
        o = int(meta[k]["pix_off"]) // 4 * esz
        raw_t[o:o + t.nbytes] = np.ascontiguousarray(t).view(np.uint8).ravel()
        raw_m[o:o + m.nbytes] = np.ascontiguousarray(m).view(np.uint8).ravel()
    pitch_e = (W * channels + 3) // 4 * 4
    imgf = np.zeros(H * pitch_e + 64, np.float32)
    imgf2 = np.zeros(H * pitch_e + 64, np.float32)
    if dtype == np.uint8:
        pitch8 = (W * channels + 3) // 4 * 4
        img8 = np.zeros(H * pitch8 + 16, np.uint8)
        img8[:H * pitch8].reshape(H, pitch8)[:, :W * channels] = image.reshape(H, W * channels)
    else:
        pitch8, img8 = 0, None
        imgf[:H * pitch_e].reshape(H, pitch_e)[:, :W * channels] = image.reshape(H, W * channels)
    order = np.arange(2, dtype=np.int32)
    for method in (0, 3):
        tm2, m2 = np.full_like(arena, np.nan), np.full_like(arena, np.nan)
        maps_a, maps_b = np.full(total + 32, np.nan, np.float32), np.full(total + 32, np.nan, np.float32)
        rc = emu.emu_masked(_ptr(img8) if img8 is not None else None, ctypes.c_int64(pitch8), _ptr(imgf), _ptr(imgf2), ctypes.c_int64(pitch_e),
                            H, W, channels, _ptr(raw_t), _ptr(raw_m), int(dtype == np.float32), _ptr(tm2), _ptr(m2), _ptr(meta), 2,
                            _ptr(order), 2, _ptr(maps_a), _ptr(maps_b), method)
        assert rc == 0
        for k, got in enumerate(_maps_of(maps_a, meta)):
            want = ncc_exact.match_template_masked_exact(image, tmpls[k], masks[k], method)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.max(np.abs(got.astype(np.float64) - want)) <= 1e-4 * scale, (method, k, float(np.max(np.abs(got - want))), scale)


@pytest.mark.parametrize("mode,shapes", [(0, [(13, 18), (10, 9)]), (1, [(12, 30)])])
def test_sixteen_bit_route_on_the_functional_model(emu, mode, shapes):
    """uint16 grayscale images and templates: the integers are split into byte planes on the device, the numerator is assembled
    EXACTLY from four u8 x u8 tensor-core correlations (accumulate flavour of the persistent kernel), then OpenCV's float64 epilogue --
    all six methods within the float32 bar of the oracle (the reference casts uint16 to float32, MTM/__init__.py:71-74), and the
    raw numerator (TM_CCORR) equal to the exact integer correlation rounded once to float32."""
    from oracle import ncc_exact
    rng = np.random.default_rng(91 + mode)
    H, W = 52, 170 if mode else 64
    image = rng.integers(0, 65536, (H, W)).astype(np.uint16)
    tmpls = []
    for h, w in shapes:
        y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        tmpls.append(np.clip(image[y0:y0 + h, x0:x0 + w].astype(np.int64) + rng.integers(-3000, 3001, (h, w)), 0, 65535).astype(np.uint16))
    count = len(tmpls)
    arena, meta = _f32_arena(tmpls, 1)
    total = _geometry(meta, tmpls, H, W)
    order = np.asarray(sorted(range(count), key=lambda k: tmpls[k].shape), np.int32)
    pix8 = np.zeros(count, np.dtype([("off", "<i8"), ("wp", "<i4"), ("pad", "<i4")]))
    total8 = 0
    for k, t in enumerate(tmpls):
        pix8[k] = (total8, (t.shape[1] + 3) // 4 * 4, 0)
        total8 += (int(pix8[k]["wp"]) * t.shape[0] + 15) // 16 * 16
    planes = np.zeros(2 * total8 + 64, np.uint8)
    for k, t in enumerate(tmpls):
        for y in range(t.shape[0]):
            o = int(pix8[k]["off"]) + y * int(pix8[k]["wp"])
            planes[o:o + t.shape[1]] = t[y] >> 8
            planes[total8 + o:total8 + o + t.shape[1]] = t[y] & 255
    pitch_e = (W + 3) // 4 * 4
    pitch = (W + 64 + 64 + 127) // 128 * 128
    pixf = np.zeros(H * pitch_e + 64, np.float32)
    hi, lo = np.zeros(H * pitch + 256, np.uint8), np.zeros(H * pitch + 256, np.uint8)
    spitch = (W + 1 + 3) // 4 * 4
    scratch = np.zeros(4 * H * W + 16, np.float64)
    sat_s, sat_q = np.full((H + 1, spitch), np.nan), np.full((H + 1, spitch), np.nan)
    centred = np.full_like(arena, np.nan)
    src = np.ascontiguousarray(image)
    hs, ws = [t.shape[0] for t in tmpls], [t.shape[1] for t in tmpls]
    emu.emu_ncc_tc16.restype = ctypes.c_longlong
    imagef = image.astype(np.float32)
    acc = np.full(total + 32, np.nan, np.float64)
    for method in range(6):
        maps = np.full(total + 32, np.nan, np.float32)
        n_mma = emu.emu_ncc_tc16(_ptr(src), H, W, _ptr(pixf), ctypes.c_int64(pitch_e), _ptr(hi), _ptr(lo), ctypes.c_int64(pitch), _ptr(scratch),
                                 _ptr(sat_s), _ptr(sat_q), ctypes.c_int64(spitch), _ptr(arena), _ptr(centred), _ptr(meta), _ptr(planes),
                                 ctypes.c_int64(total8), _ptr(pix8), _ptr(order), count, mode, max(hs), max(ws), min(hs), min(ws), _ptr(acc),
                                 _ptr(maps), method, 32, 3, 4, 2, int(method > 0))
        assert n_mma > 0, n_mma
        for k, got in enumerate(_maps_of(maps, meta)):
            want = ncc_exact.match_template_exact(imagef, tmpls[k].astype(np.float32), method)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.max(np.abs(got.astype(np.float64) - want)) <= 1e-4 * scale, (method, k)
            if method == 2:                                                   # the exact integer numerator, rounded once
                exact = ncc_exact.cc_direct(image.astype(np.float32), tmpls[k].astype(np.float32))     # float64 sums of integers: exact
                assert np.array_equal(got, exact.astype(np.float32))
    assert np.array_equal(hi[:H * pitch].reshape(H, pitch)[:, :W], image >> 8) and np.array_equal(lo[:H * pitch].reshape(H, pitch)[:, :W], image & 255)


# ---- seeded random sweeps (a few cases each; the same generators ran hundreds of cases while the emulation was written) ---------

@pytest.mark.parametrize("seed", [111, 137, 166, 203])
def test_tensor_core_kernel_random_geometries(emu, seed):
    """Random image / template-group geometry, tile height, ring shape, epilogue-warp count, kernel flavour and CTA count on the
    tcgen05 model: the raw numerator and TM_CCORR_NORMED bit-identical to the oracle, TM_CCOEFF_NORMED within 2e-6."""
    from oracle import ncc_exact
    rng = np.random.default_rng(seed)
    C = int(rng.choice([1, 1, 3, 4]))
    mode = int(rng.integers(0, 2)) if C == 1 else 0
    H, W = int(rng.integers(12, 90)), int(rng.integers(20, 200 if mode == 0 else 400))
    hmax, wmax = int(rng.integers(4, min(H, 40) + 1)), int(rng.integers(4, min(W, 60) + 1))
    shapes = [(hmax, wmax)]
    for _ in range(0 if mode == 1 else int(rng.integers(0, 8))):
        shapes.append((hmax, wmax) if rng.random() < 0.5 else
                      (int(rng.integers(max(1, hmax // 2), hmax + 1)), int(rng.integers(max(1, wmax // 2), wmax + 1))))
    shapes = [(h, w) for h, w in shapes if h * w >= 16]
    N = int(rng.choice([16, 32, 48, 64, 96]))
    opts = dict(stages=int(rng.integers(2, 6)), ds=int(rng.integers(1, 3)), EW=int(rng.choice([8, 12])), persist=bool(rng.random() < 0.8),
                ctas=int(rng.integers(1, 5)))
    image, tmpls = _planted(rng, H, W, C, shapes)
    for method in (2, 3, 5):
        got, _, _ = _host_tensor_maps(emu, image, tmpls, method, mode, N, **opts)
        for k, t in enumerate(tmpls):
            want = ncc_exact.match_template_exact(image, t, method)
            if method == 5:
                assert np.max(np.abs(got[k] - want)) <= 2e-6, (k, shapes, N, opts)
            else:
                assert np.array_equal(got[k].view(np.uint32), want.view(np.uint32)), (method, k, shapes, N, opts)


@pytest.mark.parametrize("seed", [1001, 1002, 1003, 1004, 1005, 1006, 1008, 1009])
def test_peak_sort_nms_kernels_random_maps(emu, seed):
    """Random score maps (few levels: ties and plateaus; constant maps; 1 x n, n x 1 and 1 x 1 shapes), method, N_object, thresholds and
    route: the device lists equal the port's, order included."""
    rng = np.random.default_rng(seed)
    levels = int(rng.choice([2, 4, 16, 1000]))
    maps, sizes = [], []
    for _ in range(int(rng.integers(1, 7))):
        kind = rng.integers(0, 6)
        shp = (1, int(rng.integers(1, 50))) if kind == 0 else (int(rng.integers(1, 50)), 1) if kind == 1 else \
            (int(rng.integers(2, 40)), int(rng.integers(2, 60)))
        m = (np.round(rng.random(shp) * levels) / levels).astype(np.float32)
        if rng.random() < 0.15:
            m[:] = m.flat[0]
        maps.append(m)
        sizes.append((int(rng.integers(1, 40)), int(rng.integers(1, 40))))
    method = int(rng.choice([1, 3, 5]))
    n_object = [float("inf"), 1, int(rng.integers(2, 30))][int(rng.integers(0, 3))]
    thr, overlap = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9])), float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0]))
    for do_nms in (False, True):
        want = _port_postprocess(maps, sizes, method, n_object, thr, overlap, do_nms)
        for force_general in (False, True):
            if force_general and not do_nms and n_object == 1:
                continue
            got, _ = _host_postprocess(emu, maps, sizes, method, n_object, thr, overlap, do_nms, force_general)
            assert got == want, (do_nms, force_general, method, n_object, thr, overlap)
