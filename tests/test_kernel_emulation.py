"""CPU tests of device code: the SIMPLE kernels of csrc/ (no shared memory, no warp intrinsics, no PTX) are compiled
for the host behind tests/emu/cuda_shim.h -- their source text is taken verbatim from the .cu files -- and run thread by
thread.  Checks, bit for bit and without a GPU:
  * transform.cu: the 8 symmetries and the integer-factor INTER_AREA rounding against numpy / the restated OpenCV rule
    (which tests/test_oracle.py pins to the live cv2.resize);
  * ncc_tc.cu: the row-walking window-moment kernel (experiment knob MTM_B200_MOM_ROWS) writes exactly what the default
    grid-stride kernel writes.
TEST INFRASTRUCTURE: the emulation is a checker of kernel logic, not a CPU path of the product."""
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
    epi = open(os.path.join(CSRC, "ncc_epilogue.cuh")).read()
    internal = open(os.path.join(CSRC, "mtm_internal.cuh")).read()
    parts = ['#include "cuda_shim.h"', '#include "mtm_b200.h"',
             "struct SatView { const uint32_t* s; const unsigned long long* q; int64_t pitch; int64_t plane; };",
             _function(internal, "struct XformDesc {") + ";", _function(internal, "struct SizeDesc {") + ";",
             _function(epi, "__device__ __forceinline__ uint32_t sat_window_s("),
             xf[xf.index("namespace {"):xf.index("}  // namespace") + 1],          # all device code of transform.cu
             "namespace {",
             _function(tc, "template <bool STREAM>\n__global__ void window_moments_kernel("),
             _function(tc, "template <int C>\n__global__ void __launch_bounds__(256)\nwindow_moments_rows_kernel("),
             "}",
             r'''
extern "C" void emu_transform(const uint8_t* src, uint8_t* dst, const XformDesc* descs, int n_out, int C, int dtype, int f, int grid_x)
{
    const float scale = 1.f / (float)(f * f);
    dim3 g; g.x = grid_x; g.y = n_out; dim3 b; b.x = 256;
    if (dtype == MTM_U8) emu_launch(g, b, [&] { transform_kernel<uint8_t>(src, dst, descs, C, f, scale); });
    else if (dtype == MTM_U16) emu_launch(g, b, [&] { transform_kernel<uint16_t>(src, dst, descs, C, f, scale); });
    else emu_launch(g, b, [&] { transform_kernel<float>(src, dst, descs, C, f, scale); });
}
extern "C" void emu_moments(int rows_form, const uint32_t* sat_s, const uint32_t* sat_q32, int64_t pitch, int64_t plane, const SizeDesc* sizes,
                            int n_sizes, int C, uint32_t* S, float* rsD, int64_t mom_plane, int gx, int gy)
{
    SatView sv{sat_s, nullptr, pitch, plane};
    dim3 b; b.x = 256;
    if (!rows_form) {
        dim3 g; g.x = gx; g.y = n_sizes;
        emu_launch(g, b, [&] { window_moments_kernel<false>(sv, sat_q32, sizes, S, rsD, C, mom_plane); });
        return;
    }
    dim3 g; g.x = gx; g.y = gy; g.z = n_sizes;
    if (C == 1) emu_launch(g, b, [&] { window_moments_rows_kernel<1>(sv, sat_q32, sizes, S, rsD, mom_plane); });
    else if (C == 3) emu_launch(g, b, [&] { window_moments_rows_kernel<3>(sv, sat_q32, sizes, S, rsD, mom_plane); });
    else emu_launch(g, b, [&] { window_moments_rows_kernel<4>(sv, sat_q32, sizes, S, rsD, mom_plane); });
}
''']
    d = tmp_path_factory.mktemp("emu")
    (d / "emu.cpp").write_text("\n".join(parts))
    lib = d / "libemu.so"
    r = subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"), "-I", os.path.join(ROOT, "include"),
                        str(d / "emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return ctypes.CDLL(str(lib))


XFORM_DTYPE = np.dtype([("src_off", "<i8"), ("src_pitch", "<i8"), ("dst_off", "<i8"), ("dst_pitch", "<i8"),
                        ("dh", "<i4"), ("dw", "<i4"), ("oh", "<i4"), ("ow", "<i4"), ("op", "<i4"), ("pad", "<i4")])
SIZE_DTYPE = np.dtype([("h", "<i4"), ("w", "<i4"), ("mh", "<i4"), ("mw", "<i4"), ("off", "<i8")])
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
        sizes[k] = (h, w, H - h + 1, W - w + 1, off)
        off += ((H - h + 1) * (W - w + 1) + 31) // 32 * 32
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
    idx = y * (W - w + 1) + x
    rs = outs[0][1][idx] if channels > 1 else outs[0][0].view(np.float32)[2 * idx + 1]
    assert rs == 0.0
