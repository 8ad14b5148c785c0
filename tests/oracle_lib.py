"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU legs — never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


class OrcCamera(C.Structure):
    _fields_ = [("fc", C.c_float * 2), ("cc", C.c_float * 2), ("dist", C.c_float * 5),
                ("R", C.c_float * 9), ("t", C.c_float * 3)]


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _cam(cam) -> OrcCamera:
    c = OrcCamera()
    c.fc[:] = [np.float32(v) for v in cam.fc]
    c.cc[:] = [np.float32(v) for v in cam.cc]
    c.dist[:] = [np.float32(v) for v in cam.dist]
    c.R[:] = [np.float32(v) for v in cam.R]
    c.t[:] = [np.float32(v) for v in cam.t]
    return c


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp, i32 = C.c_void_p, C.c_int
        lib.orc_gray_num_bits.argtypes = [i32]
        lib.orc_gray_num_imgs.argtypes = [i32, i32, i32]
        lib.orc_generate_gray.argtypes = [vp, i32, i32, i32]
        lib.orc_gray_to_dec.argtypes = [vp, i32]
        lib.orc_generate_mf.argtypes = [vp, i32, i32]
        lib.orc_shadow_mask.argtypes = [vp, vp, i32, i32, vp]
        lib.orc_get_phase_strict.argtypes = [vp, vp]
        lib.orc_wrapped_phase_strict.argtypes = [i32, i32, vp]
        lib.orc_mf_decode.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
        lib.orc_mf_decode_mt.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32]
        lib.orc_gray_decode.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
        lib.orc_undistort_point.argtypes = [C.c_float, C.c_float, C.POINTER(OrcCamera), vp, vp]
        lib.orc_cam2world.argtypes = [C.POINTER(OrcCamera), vp]
        lib.orc_normalize.argtypes = [vp]
        lib.orc_line_line_intersection.argtypes = [vp, vp, vp, vp, vp]
        lib.orc_mf_triangulate.argtypes = [vp, vp, vp, vp, i32, i32, C.POINTER(OrcCamera), C.POINTER(OrcCamera),
                                           vp, vp, vp, vp, vp, i32]
        lib.orc_mf_triangulate.restype = C.c_int64
        lib.orc_ge_triangulate.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32]
        lib.orc_ge_triangulate.restype = C.c_int64
        lib.orc_gray_triangulate.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(OrcCamera),
                                             C.POINTER(OrcCamera), vp, vp, vp]
        lib.orc_gray_triangulate.restype = C.c_int64
        lib.orc_pointcloud_from_dense.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        lib.orc_run_mf.argtypes = [vp, i32, i32, i32, i32, i32, i32, C.POINTER(OrcCamera), vp, vp, vp, vp, vp, i32]
        lib.orc_run_mf.restype = C.c_int64
        lib.orc_mesh_index.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.orc_export_mesh.argtypes = [vp, vp, vp, i32, i32, i32, C.c_char_p]
        lib.orc_max_threads.restype = i32

    # ---- patterns ----
    def gray_num_bits(self, n):
        return self.lib.orc_gray_num_bits(n)

    def generate_gray(self, W, H, use_epi):
        n = self.lib.orc_gray_num_imgs(W, H, int(use_epi))
        out = np.empty((n, H, W), np.uint8)
        self.lib.orc_generate_gray(out.ctypes.data, W, H, int(use_epi))
        return out

    def gray_to_dec(self, bits):
        b = np.ascontiguousarray(bits, np.uint8)
        return self.lib.orc_gray_to_dec(b.ctypes.data, len(b))

    def generate_mf(self, W, H):
        out = np.empty((14, H, W), np.uint8)
        self.lib.orc_generate_mf(out.ctypes.data, W, H)
        return out

    # ---- decode ----
    def get_phase_strict(self, G):
        g = np.ascontiguousarray(G, np.int32)
        ph = C.c_float(0)
        ok = self.lib.orc_get_phase_strict(g.ctypes.data, C.byref(ph))
        return bool(ok), np.float32(ph.value)

    def wrapped_phase_strict(self, a, b):
        p = C.c_float(0)
        ok = self.lib.orc_wrapped_phase_strict(a, b, C.byref(p))
        return bool(ok), np.float32(p.value)

    def mf_decode(self, stack, F=3, S=4, black_thr=40, mode=0, nthreads=1):
        """stack: uint8 [N, H, W] (one camera) -> (phase f32 [H,W], mask u8 [H,W])"""
        stack = np.ascontiguousarray(stack, np.uint8)
        N, H, W = stack.shape
        assert N == 2 + F * S
        ph = np.empty((H, W), np.float32)
        mk = np.empty((H, W), np.uint8)
        rc = self.lib.orc_mf_decode_mt(stack.ctypes.data, W, H, F, S, black_thr, mode, ph.ctypes.data, mk.ctypes.data,
                                       nthreads)
        assert rc == 0
        return ph, mk

    def gray_decode(self, stack, nbits_col, nbits_row=0, black_thr=40, white_thr=0, scan_w=None, scan_h=None):
        stack = np.ascontiguousarray(stack, np.uint8)
        N, H, W = stack.shape
        assert N == 2 + 2 * nbits_col + 2 * nbits_row
        col = np.empty((H, W), np.int32)
        row = np.empty((H, W), np.int32)
        mk = np.empty((H, W), np.uint8)
        self.lib.orc_gray_decode(stack.ctypes.data, W, H, nbits_col, nbits_row, black_thr, white_thr,
                                 scan_w if scan_w is not None else W, scan_h if scan_h is not None else H,
                                 col.ctypes.data, row.ctypes.data, mk.ctypes.data)
        return col, row, mk

    # ---- geometry ----
    def undistort_point(self, x, y, cam):
        c = _cam(cam)
        ox, oy = C.c_float(0), C.c_float(0)
        self.lib.orc_undistort_point(np.float32(x), np.float32(y), C.byref(c), C.byref(ox), C.byref(oy))
        return np.float32(ox.value), np.float32(oy.value)

    def line_line_intersection(self, p1, v1, p2, v2):
        a = [np.ascontiguousarray(v, np.float32) for v in (p1, v1, p2, v2)]
        out = np.zeros(3, np.float32)
        ok = self.lib.orc_line_line_intersection(*[x.ctypes.data for x in a], out.ctypes.data)
        return bool(ok), out

    # ---- match + triangulate ----
    def mf_triangulate(self, phL, mkL, phR, mkR, cams, Q, rigid=None, nthreads=1):
        phL, phR = np.ascontiguousarray(phL, np.float32), np.ascontiguousarray(phR, np.float32)
        mkL, mkR = np.ascontiguousarray(mkL, np.uint8), np.ascontiguousarray(mkR, np.uint8)
        H, W = phL.shape
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        xyz = np.empty((H, W, 3), np.float32)
        valid = np.empty((H, W), np.uint8)
        mk = np.empty((H, W), np.int32)
        cl, cr = _cam(cams[0]), _cam(cams[1])
        n = self.lib.orc_mf_triangulate(phL.ctypes.data, mkL.ctypes.data, phR.ctypes.data, mkR.ctypes.data, W, H,
                                        C.byref(cl), C.byref(cr), q.ctypes.data,
                                        r.ctypes.data if r is not None else None,
                                        xyz.ctypes.data, valid.ctypes.data, mk.ctypes.data, nthreads)
        return xyz, valid, mk, int(n)

    def ge_triangulate(self, colL, mkL, colR, mkR, Q, rigid=None, whiteL=None, whiteR=None, nthreads=1):
        colL, colR = np.ascontiguousarray(colL, np.int32), np.ascontiguousarray(colR, np.int32)
        mkL, mkR = np.ascontiguousarray(mkL, np.uint8), np.ascontiguousarray(mkR, np.uint8)
        H, W = colL.shape
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        xyz = np.empty((H, W, 3), np.float32)
        valid = np.empty((H, W), np.uint8)
        mk = np.empty((H, W), np.int32)
        color = np.zeros((H, W), np.uint8) if whiteL is not None else None
        wl = np.ascontiguousarray(whiteL, np.uint8) if whiteL is not None else None
        wr = np.ascontiguousarray(whiteR, np.uint8) if whiteR is not None else None
        n = self.lib.orc_ge_triangulate(colL.ctypes.data, mkL.ctypes.data, colR.ctypes.data, mkR.ctypes.data, W, H,
                                        q.ctypes.data, r.ctypes.data if r is not None else None,
                                        wl.ctypes.data if wl is not None else None,
                                        wr.ctypes.data if wr is not None else None,
                                        xyz.ctypes.data, valid.ctypes.data, mk.ctypes.data,
                                        color.ctypes.data if color is not None else None, nthreads)
        return xyz, valid, mk, color, int(n)

    def gray_triangulate(self, colL, rowL, mkL, colR, rowR, mkR, scan_w, scan_h, cams, rigid=None):
        arrs = [np.ascontiguousarray(a, t) for a, t in ((colL, np.int32), (rowL, np.int32), (mkL, np.uint8),
                                                        (colR, np.int32), (rowR, np.int32), (mkR, np.uint8))]
        H, W = arrs[0].shape
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        ssum = np.empty((scan_w * scan_h, 3), np.float32)
        cnt = np.empty((scan_w * scan_h,), np.uint8)
        cl, cr = _cam(cams[0]), _cam(cams[1])
        n = self.lib.orc_gray_triangulate(*[a.ctypes.data for a in arrs], W, H, scan_w, scan_h, C.byref(cl),
                                          C.byref(cr), r.ctypes.data if r is not None else None,
                                          ssum.ctypes.data, cnt.ctypes.data)
        return ssum, cnt, int(n)

    def pointcloud_from_dense(self, xyz, valid, scan_w, scan_h):
        xyz = np.ascontiguousarray(xyz, np.float32)
        valid = np.ascontiguousarray(valid, np.uint8)
        H, W = valid.shape
        pts = np.empty((scan_h, scan_w, 3), np.float32)
        cnt = np.empty((scan_h, scan_w), np.uint8)
        self.lib.orc_pointcloud_from_dense(xyz.ctypes.data, valid.ctypes.data, W, H, scan_w, scan_h,
                                           pts.ctypes.data, cnt.ctypes.data)
        return pts, cnt

    def run_mf(self, stacks, cams, Q, F=3, S=4, black_thr=40, mode=0, rigid=None, nthreads=1):
        """stacks: uint8 [2, N, H, W] -> (xyz, valid, match_k, n)"""
        stacks = np.ascontiguousarray(stacks, np.uint8)
        _, N, H, W = stacks.shape
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        xyz = np.empty((H, W, 3), np.float32)
        valid = np.empty((H, W), np.uint8)
        mk = np.empty((H, W), np.int32)
        carr = (OrcCamera * 2)(_cam(cams[0]), _cam(cams[1]))
        n = self.lib.orc_run_mf(stacks.ctypes.data, W, H, F, S, black_thr, mode, carr, q.ctypes.data,
                                r.ctypes.data if r is not None else None, xyz.ctypes.data, valid.ctypes.data,
                                mk.ctypes.data, nthreads)
        return xyz, valid, mk, int(n)

    def remap_linear(self, src, map1, map2):
        src = np.ascontiguousarray(src, np.uint8)
        H, W = src.shape
        m1 = np.ascontiguousarray(map1, np.int16)
        m2 = np.ascontiguousarray(map2, np.uint16)
        dst = np.empty((H, W), np.uint8)
        self.lib.orc_remap_linear(C.c_void_p(src.ctypes.data), W, H, C.c_void_p(m1.ctypes.data), C.c_void_p(m2.ctypes.data),
                                  C.c_void_p(dst.ctypes.data))
        return dst

    def auto_contrast(self, img):
        """Utilities::autoContrast (channel 0) on a copy of a uint8 [H, W] image"""
        out = np.ascontiguousarray(img, np.uint8).copy()
        H, W = out.shape
        self.lib.orc_auto_contrast(C.c_void_p(out.ctypes.data), W, H)
        return out

    def max_threads(self):
        return self.lib.orc_max_threads()

    # ---- N3: mesh ----
    def mesh_index(self, points, count, w, h, first_vertex):
        """-> (vertices [nv,3] f32, vertex_src [nv] i32, faces [nf,3] i32)"""
        points = np.ascontiguousarray(points, np.float32)
        count = np.ascontiguousarray(count, np.uint8)
        px = w * h
        pn = np.empty(px, np.int32)
        vert = np.empty((px, 3), np.float32)
        src = np.empty(px, np.int32)
        faces = np.empty((2 * px, 3), np.int32)
        nv, nf = np.zeros(1, np.int64), np.zeros(1, np.int64)
        self.lib.orc_mesh_index(points.ctypes.data, count.ctypes.data, w, h, first_vertex, pn.ctypes.data,
                                vert.ctypes.data, src.ctypes.data, faces.ctypes.data, nv.ctypes.data, nf.ctypes.data)
        return vert[:nv[0]].copy(), src[:nv[0]].copy(), faces[:nf[0]].copy()

    def export_mesh(self, points, count, w, h, path, obj=False, color=None):
        points = np.ascontiguousarray(points, np.float32)
        count = np.ascontiguousarray(count, np.uint8)
        cptr = None
        if color is not None:
            color = np.ascontiguousarray(color, np.int32)
            cptr = color.ctypes.data
        rc = self.lib.orc_export_mesh(points.ctypes.data, count.ctypes.data, cptr, w, h, int(obj), str(path).encode())
        assert rc == 0

_oracle = None



def load() -> Oracle:
    global _oracle
    if _oracle is None:
        if not os.path.exists(LIB):
            build()
        _oracle = Oracle(C.CDLL(LIB))
    return _oracle
