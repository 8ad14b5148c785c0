"""slr_b200 — Python host-side plumbing over the C ABI of libslr_b200.so (include/slr_b200.h).

The product is the CUDA library; this module only (a) loads it with ctypes, (b) wraps the entry
points so tests and bench.py can hand it torch device tensors / numpy host arrays, and (c) offers
small helpers (synthetic calibration, numpy scene synthesis).  There is no CPU fallback: if the
shared library is missing, importing `capi()` raises.

Import as `import slr_b200` (the repo-root shim) — the directory name carries a hyphen.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLR_B200_LIB: load another build of the same library (the phase-clock debug build of csrc/Makefile PHASE=1)
LIB_PATH = os.environ.get("SLR_B200_LIB") or os.path.join(_HERE, "libslr_b200.so")

MODE_STRICT = 0
MODE_CORRECTED = 1


class SlrError(RuntimeError):
    pass


class CCamera(C.Structure):
    """slr_camera (include/slr_b200.h) == the VirtualCamera fields the path reads."""
    _fields_ = [("fc", C.c_float * 2), ("cc", C.c_float * 2), ("dist", C.c_float * 5),
                ("R", C.c_float * 9), ("t", C.c_float * 3)]


@dataclass
class Camera:
    fc: tuple = (2400.0, 2400.0)
    cc: tuple = (640.0, 512.0)
    dist: tuple = (0.0, 0.0, 0.0, 0.0, 0.0)
    R: tuple = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    t: tuple = (0.0, 0.0, 0.0)

    def to_c(self) -> CCamera:
        c = CCamera()
        c.fc[:] = [np.float32(v) for v in self.fc]
        c.cc[:] = [np.float32(v) for v in self.cc]
        c.dist[:] = [np.float32(v) for v in self.dist]
        c.R[:] = [np.float32(v) for v in self.R]
        c.t[:] = [np.float32(v) for v in self.t]
        return c


_lib = None

# every symbol include/slr_b200.h declares (tests check the library exports each of them)
EXPORTS = [
    "slr_create", "slr_destroy", "slr_set_stream", "slr_synchronize", "slr_set_calib", "slr_last_error",
    "slr_version", "slr_gray_num_bits", "slr_gray_num_imgs", "slr_generate_gray_patterns",
    "slr_generate_mf_patterns", "slr_mf_decode", "slr_gray_decode", "slr_match_triangulate_phase",
    "slr_match_triangulate_code", "slr_bucket_triangulate", "slr_run_mf", "slr_run_ge", "slr_run_mf_host",
    "slr_run_ge_host", "slr_host_alloc", "slr_host_free", "slr_synth_mf", "slr_synth_mf_fs", "slr_auto_contrast", "slr_set_auto_contrast", "slr_peer_alloc", "slr_peer_open", "slr_peer_close",
    "slr_peer_free", "slr_set_gather_targets", "slr_set_row_offset", "slr_synth_gray",
    "slr_kernel_launches", "slr_mesh_index", "slr_mesh_index_host", "slr_allgather", "slr_nccl_unique_id",
    "slr_nccl_comm_create", "slr_nccl_comm_destroy", "slr_set_rectify_maps", "slr_rectify_stack", "slr_set_host_input_raw", "slr_run_gray_host",
    "slr_run_mf_raw", "slr_ingest_begin", "slr_ingest_image", "slr_run_mf_ingested", "slr_run_ge_ingested", "slr_ingest_abort",
    "slr_horn_method", "slr_register_scan", "slr_merge_scans", "slr_png_unfilter", "slr_strict_tables_check",
]


def capi():
    """Load libslr_b200.so and declare the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SlrError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, u64p = C.c_void_p, C.c_int, C.c_uint, C.c_void_p
    lib.slr_last_error.restype = C.c_char_p
    lib.slr_version.restype = C.c_char_p
    lib.slr_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    lib.slr_destroy.argtypes = [vp]
    lib.slr_set_stream.argtypes = [vp, vp]
    lib.slr_synchronize.argtypes = [vp]
    lib.slr_set_calib.argtypes = [vp, C.POINTER(CCamera), C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.slr_strict_tables_check.argtypes = [vp]
    lib.slr_gray_num_bits.argtypes = [i32]
    lib.slr_gray_num_imgs.argtypes = [i32, i32, i32]
    lib.slr_generate_gray_patterns.argtypes = [vp, i32, i32, i32]
    lib.slr_generate_mf_patterns.argtypes = [vp, i32, i32]
    lib.slr_mf_decode.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.slr_gray_decode.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.slr_match_triangulate_phase.argtypes = [vp, vp, vp, i32, vp, vp, vp, u64p]
    lib.slr_match_triangulate_code.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, u64p]
    lib.slr_bucket_triangulate.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, u64p]
    lib.slr_run_mf.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, u64p]
    lib.slr_run_mf_raw.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, u64p]
    lib.slr_horn_method.argtypes = [vp, i32, i32, vp, vp]
    lib.slr_register_scan.argtypes = [vp, i32, vp, vp]
    lib.slr_merge_scans.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.slr_png_unfilter.argtypes = [vp, vp, vp, i32]
    lib.slr_ingest_begin.argtypes = [vp, i32]
    lib.slr_ingest_image.argtypes = [vp, i32, vp, i32, i32]
    lib.slr_run_mf_ingested.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(C.c_ulonglong)]
    lib.slr_ingest_abort.argtypes = [vp]
    lib.slr_run_ge_ingested.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_ulonglong)]
    lib.slr_run_ge.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, u64p]
    lib.slr_run_mf_host.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, C.POINTER(C.c_ulonglong)]
    lib.slr_run_ge_host.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp,
                                    C.POINTER(C.c_ulonglong)]
    lib.slr_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.slr_host_free.argtypes = [vp]
    lib.slr_synth_mf.argtypes = [vp, vp, i32, i32, u32, i32, C.c_float]
    lib.slr_synth_mf_fs.argtypes = [vp, vp, i32, i32, i32, i32, u32, i32, C.c_float]
    lib.slr_synth_gray.argtypes = [vp, vp, i32, i32, u32, i32, C.c_float]
    lib.slr_auto_contrast.argtypes = [vp, vp, i32]
    lib.slr_set_auto_contrast.argtypes = [vp, i32]
    lib.slr_set_rectify_maps.argtypes = [vp, vp, vp]
    lib.slr_rectify_stack.argtypes = [vp, vp, i32, i32, vp]
    lib.slr_set_host_input_raw.argtypes = [vp, i32]
    lib.slr_run_gray_host.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, C.POINTER(C.c_ulonglong)]
    lib.slr_mesh_index.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.slr_mesh_index_host.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.slr_allgather.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.slr_nccl_unique_id.argtypes = [vp]
    lib.slr_nccl_comm_create.argtypes = [vp, C.POINTER(vp), i32, i32, vp]
    lib.slr_nccl_comm_destroy.argtypes = [vp]
    lib.slr_peer_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), vp]
    lib.slr_peer_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.slr_peer_close.argtypes = [vp, vp]
    lib.slr_peer_free.argtypes = [vp, vp]
    lib.slr_set_gather_targets.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.c_longlong]
    lib.slr_set_row_offset.argtypes = [vp, i32]
    lib.slr_kernel_launches.argtypes = [vp]
    lib.slr_kernel_launches.restype = C.c_ulonglong
    _lib = lib
    return lib


def _check(status: int, what: str):
    if status != 0:
        msg = capi().slr_last_error().decode("utf-8", "replace")
        raise SlrError(f"{what} failed (status {status}): {msg}")


def gray_num_bits(n: int) -> int:
    return capi().slr_gray_num_bits(n)


def generate_gray_patterns(W: int, H: int, use_epi: bool) -> np.ndarray:
    """GrayCodes::generateGrays (Duke/graycodes.cpp:55-114) -> uint8 [nimgs, H, W]."""
    lib = capi()
    n = lib.slr_gray_num_imgs(W, H, int(use_epi))
    out = np.empty((n, H, W), np.uint8)
    _check(lib.slr_generate_gray_patterns(out.ctypes.data, W, H, int(use_epi)), "slr_generate_gray_patterns")
    return out


def strict_tables() -> np.ndarray:
    """The MF kernels' strict-mode lookup (slr_device.cuh: wrapped_strict_fx) for every a = G4-G2, b = G1-G3 in
    [-255, 255] -> int32 [511 (b + 255), 511 (a + 255)] in units of 2^-24, INT32_MIN where the pixel is dropped
    (Duke/mfreconstruct.cpp:246-261).  Host-only: raises if the library's own check against the branch form fails."""
    out = np.empty((511, 511), np.int32)
    _check(capi().slr_strict_tables_check(out.ctypes.data), "slr_strict_tables_check")
    return out


def generate_mf_patterns(projW: int, projH: int) -> np.ndarray:
    """MultiFrequency::generateMutiFreq (Duke/multifrequency.cpp:14-33) -> uint8 [14, projH, projW]."""
    out = np.empty((14, projH, projW), np.uint8)
    _check(capi().slr_generate_mf_patterns(out.ctypes.data, projW, projH), "slr_generate_mf_patterns")
    return out


def horn_method(base, moving, force_unit_scale=False):
    """slr_horn_method: (7-vector {tx ty tz qr qx qy qz}, scale) of the transform that moves `moving` [n,3] onto `base`."""
    pairs = np.ascontiguousarray(np.concatenate([np.asarray(base, np.float64), np.asarray(moving, np.float64)], 1))
    out, scale = np.zeros(7), C.c_double(0)
    _check(capi().slr_horn_method(C.c_void_p(pairs.ctypes.data), len(pairs), int(force_unit_scale), C.c_void_p(out.ctypes.data),
                                  C.byref(scale)), "slr_horn_method")
    return out, scale.value


def register_scan(base, moving, prev=None):
    """slr_register_scan: DotMatch::calMatrix's transfer matrix [3,4] (float64) for a scan whose markers `moving` were
    matched to the previous scan's `base`; prev = the previous scan's accumulated matrix (None for the first)."""
    pairs = np.ascontiguousarray(np.concatenate([np.asarray(base, np.float64), np.asarray(moving, np.float64)], 1))
    out = np.zeros(12)
    pv = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(12)) if prev is not None else None
    _check(capi().slr_register_scan(C.c_void_p(pairs.ctypes.data), len(pairs), C.c_void_p(pv.ctypes.data) if pv is not None else None,
                                    C.c_void_p(out.ctypes.data)), "slr_register_scan")
    return out.reshape(3, 4)


def synthetic_rig(W: int, H: int, f: float = 2400.0, tx: float = -200.0, distort: bool = True):
    """A synthetic rectified rig: two cameras, Q as cv::stereoRectify lays it out.

    Q = [[1,0,0,-cx],[0,1,0,-cy],[0,0,0,f],[0,0,-1/Tx,(cx-cx')/Tx]].
    """
    cx, cy = W / 2.0, H / 2.0
    distL = (-0.12, 0.08, 0.0007, -0.0004, 0.0) if distort else (0.0,) * 5
    distR = (-0.10, 0.05, -0.0005, 0.0006, 0.0) if distort else (0.0,) * 5
    camL = Camera(fc=(f, f * 1.001), cc=(cx + 3.25, cy - 2.5), dist=distL)
    camR = Camera(fc=(f * 0.999, f), cc=(cx - 1.75, cy + 1.25), dist=distR,
                  R=(0.9998, 0.0, 0.02, 0.0, 1.0, 0.0, -0.02, 0.0, 0.9998), t=(tx, 0.0, 0.0))
    Q = np.array([[1, 0, 0, -cx], [0, 1, 0, -cy], [0, 0, 0, f], [0, 0, -1.0 / tx, 0.0]], np.float64)
    return [camL, camR], Q


class Engine:
    """One engine per GPU; thin torch-tensor wrapper over the C ABI.  Work is issued on torch's current
    CUDA stream so torch.cuda.Event timing brackets the library's kernels."""

    def __init__(self, width: int, height: int, max_batch: int = 1, device: int = 0):
        import torch
        self._torch = torch
        self.lib = capi()
        self.W, self.H, self.max_batch, self.device = width, height, max_batch, device
        h = C.c_void_p()
        _check(self.lib.slr_create(C.byref(h), device, width, height, max_batch), "slr_create")
        self.h = h
        self.dev = torch.device("cuda", device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.slr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------------
    def _bind_stream(self):
        s = self._torch.cuda.current_stream(self.dev).cuda_stream
        _check(self.lib.slr_set_stream(self.h, C.c_void_p(s)), "slr_set_stream")

    def _empty(self, shape, dtype):
        return self._torch.empty(shape, dtype=dtype, device=self.dev)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def launches(self) -> int:
        return int(self.lib.slr_kernel_launches(self.h))

    def synchronize(self):
        _check(self.lib.slr_synchronize(self.h), "slr_synchronize")

    def set_calib(self, cams, Q, rigid=None):
        arr = (CCamera * 2)(cams[0].to_c(), cams[1].to_c())
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = None
        if rigid is not None:
            r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12))
        self._bind_stream()
        _check(self.lib.slr_set_calib(self.h, arr, q.ctypes.data_as(C.POINTER(C.c_double)),
                                      r.ctypes.data_as(C.POINTER(C.c_float)) if r is not None else None),
               "slr_set_calib")
        self.synchronize()

    def set_rectify_maps(self, map1, map2):
        """map1: int16 [2,H,W,2], map2: uint16 [2,H,W] (cv::initUndistortRectifyMap CV_16SC2 layout)"""
        m1 = np.ascontiguousarray(map1, np.int16)
        m2 = np.ascontiguousarray(map2, np.uint16)
        assert m1.shape == (2, self.H, self.W, 2) and m2.shape == (2, self.H, self.W)
        self._bind_stream()
        _check(self.lib.slr_set_rectify_maps(self.h, C.c_void_p(m1.ctypes.data), C.c_void_p(m2.ctypes.data)),
               "slr_set_rectify_maps")

    def rectify_stack(self, raw, out=None):
        out = self._torch.empty_like(raw) if out is None else out
        B, _, N = raw.shape[:3]
        self._bind_stream()
        _check(self.lib.slr_rectify_stack(self.h, self._p(raw), B, N, self._p(out)), "slr_rectify_stack")
        return out

    def auto_contrast(self, images):
        """Utilities::autoContrast on every [H, W] image of a uint8 device tensor [..., H, W], in place."""
        assert images.is_contiguous() and tuple(images.shape[-2:]) == (self.H, self.W)
        self._bind_stream()
        _check(self.lib.slr_auto_contrast(self.h, self._p(images), images.numel() // (self.H * self.W)), "slr_auto_contrast")
        return images

    def set_auto_contrast(self, on: bool):
        _check(self.lib.slr_set_auto_contrast(self.h, int(on)), "slr_set_auto_contrast")

    def set_host_input_raw(self, raw: bool):
        _check(self.lib.slr_set_host_input_raw(self.h, int(raw)), "slr_set_host_input_raw")

    # -- kernels ----------------------------------------------------------------------------
    def mf_decode(self, stack, F=3, S=4, black_thr=40, mode=MODE_STRICT):
        t = self._torch
        B = stack.shape[0]
        assert stack.dtype == t.uint8 and stack.is_contiguous() and tuple(stack.shape) == (B, 2, 2 + F * S, self.H, self.W)
        phase = self._empty((B, 2, self.H, self.W), t.float32)
        mask = self._empty((B, 2, self.H, self.W), t.uint8)
        self._bind_stream()
        _check(self.lib.slr_mf_decode(self.h, self._p(stack), B, F, S, black_thr, mode, self._p(phase), self._p(mask)),
               "slr_mf_decode")
        return phase, mask

    def gray_decode(self, stack, nbits_col, nbits_row=0, black_thr=40, white_thr=0, scan_w=None, scan_h=None):
        t = self._torch
        B = stack.shape[0]
        N = 2 + 2 * nbits_col + 2 * nbits_row
        assert stack.dtype == t.uint8 and stack.is_contiguous() and tuple(stack.shape) == (B, 2, N, self.H, self.W)
        col = self._empty((B, 2, self.H, self.W), t.int32)
        row = self._empty((B, 2, self.H, self.W), t.int32) if nbits_row > 0 else None
        mask = self._empty((B, 2, self.H, self.W), t.uint8)
        self._bind_stream()
        _check(self.lib.slr_gray_decode(self.h, self._p(stack), B, nbits_col, nbits_row, black_thr, white_thr,
                                        scan_w if scan_w is not None else self.W,
                                        scan_h if scan_h is not None else self.H,
                                        self._p(col), self._p(row), self._p(mask)), "slr_gray_decode")
        return col, row, mask

    def _outputs(self, B, want_k=True, want_color=False):
        t = self._torch
        xyz = self._empty((B, self.H, self.W, 3), t.float32)
        valid = self._empty((B, self.H, self.W), t.uint8)
        k = self._empty((B, self.H, self.W), t.int32) if want_k else None
        color = self._empty((B, self.H, self.W), t.uint8) if want_color else None
        n = t.zeros(1, dtype=t.int64, device=self.dev)
        return xyz, valid, k, color, n

    def match_triangulate_phase(self, phase, mask, want_k=True):
        B = phase.shape[0]
        xyz, valid, k, _, n = self._outputs(B, want_k)
        self._bind_stream()
        _check(self.lib.slr_match_triangulate_phase(self.h, self._p(phase), self._p(mask), B, self._p(xyz),
                                                    self._p(valid), self._p(k), self._p(n)),
               "slr_match_triangulate_phase")
        return xyz, valid, k, n

    def match_triangulate_code(self, col, mask, white=None, want_k=True):
        B = col.shape[0]
        xyz, valid, k, color, n = self._outputs(B, want_k, white is not None)
        self._bind_stream()
        _check(self.lib.slr_match_triangulate_code(self.h, self._p(col), self._p(mask), B, self._p(white),
                                                   self._p(xyz), self._p(valid), self._p(k), self._p(color),
                                                   self._p(n)), "slr_match_triangulate_code")
        return xyz, valid, k, color, n

    def bucket_triangulate(self, col, row, mask, scan_w, scan_h):
        t = self._torch
        B = col.shape[0]
        ssum = self._empty((B, scan_w * scan_h, 3), t.float32)
        cnt = self._empty((B, scan_w * scan_h), t.uint8)
        n = t.zeros(1, dtype=t.int64, device=self.dev)
        self._bind_stream()
        _check(self.lib.slr_bucket_triangulate(self.h, self._p(col), self._p(row), self._p(mask), B, scan_w, scan_h,
                                               self._p(ssum), self._p(cnt), self._p(n)), "slr_bucket_triangulate")
        return ssum, cnt, n

    def run_mf(self, stack, F=3, S=4, black_thr=40, mode=MODE_STRICT, want_k=True, out=None):
        B = stack.shape[0]
        xyz, valid, k, _, n = out if out is not None else self._outputs(B, want_k)
        self._bind_stream()
        _check(self.lib.slr_run_mf(self.h, self._p(stack), B, F, S, black_thr, mode, self._p(xyz), self._p(valid),
                                   self._p(k), self._p(n)), "slr_run_mf")
        return xyz, valid, k, n

    def run_mf_raw(self, raw_stack, F=3, S=4, black_thr=40, mode=MODE_STRICT, want_k=True, out=None):
        """slr_run_mf_raw: RAW camera stacks in, rectification inside the fused kernel's stage fill."""
        B = raw_stack.shape[0]
        xyz, valid, k, _, n = out if out is not None else self._outputs(B, want_k)
        self._bind_stream()
        _check(self.lib.slr_run_mf_raw(self.h, self._p(raw_stack), B, F, S, black_thr, mode, self._p(xyz),
                                       self._p(valid), self._p(k), self._p(n)), "slr_run_mf_raw")
        return xyz, valid, k, n

    def run_ge(self, stack, nbits_col, black_thr=40, white_thr=0, scan_w=None, have_color=False, want_k=True,
               out=None):
        B = stack.shape[0]
        xyz, valid, k, color, n = out if out is not None else self._outputs(B, want_k, have_color)
        self._bind_stream()
        _check(self.lib.slr_run_ge(self.h, self._p(stack), B, nbits_col, black_thr, white_thr,
                                   scan_w if scan_w is not None else self.W, int(have_color), self._p(xyz),
                                   self._p(valid), self._p(k), self._p(color), self._p(n)), "slr_run_ge")
        return xyz, valid, k, color, n

    # -- host-buffer entry points (numpy or pinned torch CPU tensors) ---------------------------
    @staticmethod
    def _hp(a):
        if a is None:
            return C.c_void_p(0)
        if isinstance(a, np.ndarray):
            return C.c_void_p(a.ctypes.data)
        return C.c_void_p(a.data_ptr())

    def run_mf_host(self, h_stack, h_xyz, h_valid, h_k=None, F=3, S=4, black_thr=40, mode=MODE_STRICT) -> int:
        B = h_stack.shape[0]
        n = C.c_ulonglong(0)
        self._bind_stream()
        _check(self.lib.slr_run_mf_host(self.h, self._hp(h_stack), B, F, S, black_thr, mode, self._hp(h_xyz),
                                        self._hp(h_valid), self._hp(h_k), C.byref(n)), "slr_run_mf_host")
        return int(n.value)

    def run_mf_ingested(self, images, filtered, h_sum=None, h_cnt=None, h_xyz=None, h_valid=None, scan_w=0, scan_h=0,
                        F=3, S=4, black_thr=40, mode=MODE_STRICT) -> int:
        """slr_ingest_begin + slr_ingest_image per image + slr_run_mf_ingested.  images[k] = numpy uint8 host array of
        image k (cam * N + i): [H, 1 + W] PNG scanlines when filtered[k], else [H, W] pixels."""
        self._bind_stream()
        _check(self.lib.slr_ingest_begin(self.h, len(images)), "slr_ingest_begin")
        keep = []
        for k, (img, f) in enumerate(zip(images, filtered)):
            a = np.ascontiguousarray(img, np.uint8)
            keep.append(a)
            up = int(bool(f) and bool((a[:, 0] == 2).any()))
            _check(self.lib.slr_ingest_image(self.h, k, C.c_void_p(a.ctypes.data), int(bool(f)), up), "slr_ingest_image")
        n = C.c_ulonglong(0)
        _check(self.lib.slr_run_mf_ingested(self.h, F, S, black_thr, mode, scan_w, scan_h, self._hp(h_sum), self._hp(h_cnt),
                                            self._hp(h_xyz), self._hp(h_valid), C.byref(n)), "slr_run_mf_ingested")
        return n.value

    def run_ge_host(self, h_stack, h_xyz, h_valid, h_k=None, h_color=None, nbits_col=11, black_thr=40, white_thr=0,
                    scan_w=None) -> int:
        B = h_stack.shape[0]
        n = C.c_ulonglong(0)
        self._bind_stream()
        _check(self.lib.slr_run_ge_host(self.h, self._hp(h_stack), B, nbits_col, black_thr, white_thr,
                                        scan_w if scan_w is not None else self.W, int(h_color is not None),
                                        self._hp(h_xyz), self._hp(h_valid), self._hp(h_k), self._hp(h_color),
                                        C.byref(n)), "slr_run_ge_host")
        return int(n.value)

    # -- synthetic inputs ---------------------------------------------------------------------
    # ---- multi-GPU assembly through the C ABI (NCCL bound at run time) ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(capi().slr_nccl_unique_id(buf), "slr_nccl_unique_id")
        return buf.raw

    def nccl_comm_create(self, world: int, rank: int, unique_id: bytes):
        comm = C.c_void_p(0)
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.lib.slr_nccl_comm_create(self.h, C.byref(comm), world, rank, buf), "slr_nccl_comm_create")
        return comm

    def nccl_comm_destroy(self, comm):
        _check(self.lib.slr_nccl_comm_destroy(comm), "slr_nccl_comm_destroy")

    def allgather(self, comm, world: int, rank: int, scans_per_rank: int, xyz_all, valid_all):
        """slr_allgather: in-place NCCL all-gather of the assembled cloud tensors (block `rank` already filled)."""
        self._bind_stream()
        _check(self.lib.slr_allgather(self.h, comm, world, rank, scans_per_rank, self._p(xyz_all), self._p(valid_all)),
               "slr_allgather")

    # ---- multi-GPU assembly by peer stores (CUDA IPC mappings over NVLink; see slr_b200.parallel.PeerAssembly) ----
    def peer_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle) of a buffer other ranks can map"""
        ptr, handle = C.c_void_p(0), C.create_string_buffer(64)
        _check(self.lib.slr_peer_alloc(self.h, nbytes, C.byref(ptr), handle), "slr_peer_alloc")
        return ptr.value, handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p(0)
        _check(self.lib.slr_peer_open(self.h, C.create_string_buffer(handle, 64), C.byref(ptr)), "slr_peer_open")
        return ptr.value

    def peer_close(self, ptr: int):
        _check(self.lib.slr_peer_close(self.h, C.c_void_p(ptr)), "slr_peer_close")

    def peer_free(self, ptr: int):
        _check(self.lib.slr_peer_free(self.h, C.c_void_p(ptr)), "slr_peer_free")

    def set_gather_targets(self, xyz_ptrs, valid_ptrs, first_scan: int = 0):
        """Device pointers of every rank's assembled cloud, this rank's own first; [] switches the targets off."""
        n = len(xyz_ptrs)
        assert n == len(valid_ptrs)
        xa = (C.c_void_p * max(n, 1))(*xyz_ptrs)
        va = (C.c_void_p * max(n, 1))(*valid_ptrs)
        _check(self.lib.slr_set_gather_targets(self.h, n, xa, va, first_scan), "slr_set_gather_targets")

    def set_row_offset(self, first_row: int):
        self._bind_stream()
        _check(self.lib.slr_set_row_offset(self.h, first_row), "slr_set_row_offset")

    def merge_scans(self, xyz_all, valid_all, rigid=None, has_rigid=None):
        """slr_merge_scans -> (points [count,3] device tensor, source [count] int64 device tensor)"""
        n = xyz_all.shape[0]
        cells = n * self.H * self.W
        pts = self._empty((cells, 3), self._torch.float32)
        src = self._empty((cells,), self._torch.int64)
        cnt = self._torch.zeros(1, dtype=self._torch.int64, device=xyz_all.device)
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(n, 12)) if rigid is not None else None
        hr = np.ascontiguousarray(np.asarray(has_rigid, np.uint8)) if has_rigid is not None else None
        self._bind_stream()
        _check(self.lib.slr_merge_scans(self.h, self._p(xyz_all), self._p(valid_all), n, self._hp(r), self._hp(hr), self._p(pts),
                                        self._p(src), self._p(cnt)), "slr_merge_scans")
        c = int(cnt.item())
        return pts[:c], src[:c]

    def mesh_index(self, sums, counts, first_vertex=0):
        """slr_mesh_index on device tensors sums [h,w,3] f32 / counts [h,w] u8 -> (vertices [nv,3], vertex_src [nv],
        faces [nf,3]) device tensors (MeshCreator's vertex numbering + faces, Duke/meshcreator.cpp:16-166)."""
        t = self._torch
        h, w = counts.shape
        px = w * h
        vert = self._empty((px, 3), t.float32)
        src = self._empty((px,), t.int32)
        faces = self._empty((2 * px, 3), t.int32)
        cnts = t.zeros(2, dtype=t.int64, device=vert.device)
        self._bind_stream()
        _check(self.lib.slr_mesh_index(self.h, self._p(sums), self._p(counts), w, h, first_vertex, self._p(vert),
                                       self._p(src), self._p(faces), self._p(cnts)), "slr_mesh_index")
        nv, nf = (int(v) for v in cnts.tolist())
        return vert[:nv], src[:nv], faces[:nf]

    def synth_mf(self, batch, proj_w=None, seed=0, integer_disparity=True, noise_dn=0.0, F=3, S=4):
        t = self._torch
        stack = self._empty((batch, 2, 2 + F * S, self.H, self.W), t.uint8)
        self._bind_stream()
        _check(self.lib.slr_synth_mf_fs(self.h, self._p(stack), batch, F, S, proj_w or self.W, seed,
                                        int(integer_disparity), float(noise_dn)), "slr_synth_mf_fs")
        return stack

    def synth_gray(self, batch, scan_w=None, seed=0, integer_disparity=True, noise_dn=0.0):
        t = self._torch
        scan_w = scan_w or self.W
        nb = gray_num_bits(scan_w)
        stack = self._empty((batch, 2, 2 + 2 * nb, self.H, self.W), t.uint8)
        self._bind_stream()
        _check(self.lib.slr_synth_gray(self.h, self._p(stack), batch, scan_w, seed, int(integer_disparity),
                                       float(noise_dn)), "slr_synth_gray")
        return stack
