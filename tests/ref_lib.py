"""ctypes binding of oracle/_ref/libref.so — the reference's OWN sources (Duke/*.cpp) compiled unmodified
against oracle/ref_shim/.  Only exists where /root/reference was mounted at build time; tests fall back to the
committed fixtures in tests/golden/ (made by tests/golden/make_ref_fixtures.py) elsewhere."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref.so")


class RefCamera(C.Structure):
    _fields_ = [("fc", C.c_float * 2), ("cc", C.c_float * 2), ("dist", C.c_float * 5),
                ("R", C.c_float * 9), ("t", C.c_float * 3)]


def _cam(cam):
    c = RefCamera()
    c.fc[:] = [np.float32(v) for v in cam.fc]
    c.cc[:] = [np.float32(v) for v in cam.cc]
    c.dist[:] = [np.float32(v) for v in cam.dist]
    c.R[:] = [np.float32(v) for v in cam.R]
    c.t[:] = [np.float32(v) for v in cam.t]
    return c


def available():
    return os.path.exists(LIB)


class Ref:
    def __init__(self):
        self.lib = C.CDLL(LIB)
        self.lib.ref_undistort.argtypes = [C.c_float, C.c_float, C.POINTER(RefCamera), C.c_void_p, C.c_void_p]

    def gray_layout(self, W, H, epi):
        nc, nr = C.c_int(0), C.c_int(0)
        n = self.lib.ref_gray_layout(W, H, int(epi), C.byref(nc), C.byref(nr))
        return n, nc.value, nr.value

    def generate_gray(self, W, H, epi):
        n, _, _ = self.gray_layout(W, H, epi)
        out = np.empty((n, H, W), np.uint8)
        self.lib.ref_generate_gray(W, H, int(epi), C.c_void_p(out.ctypes.data))
        return out

    def gray_to_dec(self, bits):
        b = np.ascontiguousarray(bits, np.uint8)
        return self.lib.ref_gray_to_dec(C.c_void_p(b.ctypes.data), len(b))

    def generate_mf(self, W, H):
        out = np.empty((14, H, W), np.uint8)
        self.lib.ref_generate_mf(W, H, C.c_void_p(out.ctypes.data))
        return out

    def mf_decode(self, stack, black_thr=40):
        stack = np.ascontiguousarray(stack, np.uint8)
        _, H, W = stack.shape
        ph = np.empty((H, W), np.float32)
        has = np.empty((H, W), np.uint8)
        mk = np.empty((H, W), np.uint8)
        self.lib.ref_mf_decode(C.c_void_p(stack.ctypes.data), W, H, black_thr, C.c_void_p(ph.ctypes.data),
                               C.c_void_p(has.ctypes.data), C.c_void_p(mk.ctypes.data))
        return ph, has, mk

    def mf_triangulate(self, phL, hasL, phR, hasR, cams, Q, rigid=None, scan=None):
        phL, phR = np.ascontiguousarray(phL, np.float32), np.ascontiguousarray(phR, np.float32)
        hasL, hasR = np.ascontiguousarray(hasL, np.uint8), np.ascontiguousarray(hasR, np.uint8)
        H, W = phL.shape
        sw, sh = scan if scan else (max(W, H), max(W, H))
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        pts = np.empty((sh, sw, 3), np.float32)
        cnt = np.empty((sh, sw), np.uint8)
        cl, cr = _cam(cams[0]), _cam(cams[1])
        self.lib.ref_mf_triangulate(C.c_void_p(phL.ctypes.data), C.c_void_p(hasL.ctypes.data), C.c_void_p(phR.ctypes.data),
                                    C.c_void_p(hasR.ctypes.data), W, H, C.byref(cl), C.byref(cr), C.c_void_p(q.ctypes.data),
                                    C.c_void_p(r.ctypes.data) if r is not None else None, sw, sh,
                                    C.c_void_p(pts.ctypes.data), C.c_void_p(cnt.ctypes.data))
        return pts, cnt

    def gray_decode(self, stack, nbits_col, nbits_row, black_thr, white_thr, scan_w, scan_h):
        stack = np.ascontiguousarray(stack, np.uint8)
        _, H, W = stack.shape
        col = np.empty((H, W), np.int32)
        row = np.empty((H, W), np.int32)
        mk = np.empty((H, W), np.uint8)
        self.lib.ref_gray_decode(C.c_void_p(stack.ctypes.data), W, H, nbits_col, nbits_row, black_thr, white_thr, scan_w,
                                 scan_h, C.c_void_p(col.ctypes.data), C.c_void_p(row.ctypes.data), C.c_void_p(mk.ctypes.data))
        return col, row, mk

    def ge_triangulate(self, colL, hasL, colR, hasR, Q, rigid=None, whiteL=None, whiteR=None, scan=None):
        colL, colR = np.ascontiguousarray(colL, np.int32), np.ascontiguousarray(colR, np.int32)
        hasL, hasR = np.ascontiguousarray(hasL, np.uint8), np.ascontiguousarray(hasR, np.uint8)
        H, W = colL.shape
        sw, sh = scan if scan else (max(W, H), max(W, H))
        q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(16))
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        pts = np.empty((sh, sw, 3), np.float32)
        cnt = np.empty((sh, sw), np.uint8)
        color = np.zeros((sh, sw), np.uint8)
        wl = np.ascontiguousarray(whiteL, np.uint8) if whiteL is not None else None
        wr = np.ascontiguousarray(whiteR, np.uint8) if whiteR is not None else None
        self.lib.ref_ge_triangulate(C.c_void_p(colL.ctypes.data), C.c_void_p(hasL.ctypes.data), C.c_void_p(colR.ctypes.data),
                                    C.c_void_p(hasR.ctypes.data), W, H, C.c_void_p(q.ctypes.data),
                                    C.c_void_p(r.ctypes.data) if r is not None else None,
                                    C.c_void_p(wl.ctypes.data) if wl is not None else None,
                                    C.c_void_p(wr.ctypes.data) if wr is not None else None, sw, sh,
                                    C.c_void_p(pts.ctypes.data), C.c_void_p(cnt.ctypes.data), C.c_void_p(color.ctypes.data))
        return pts, cnt, color

    def gray_triangulate(self, colL, rowL, hasL, colR, rowR, hasR, scan_w, scan_h, cams, rigid=None):
        arrs = [np.ascontiguousarray(a, t) for a, t in ((colL, np.int32), (rowL, np.int32), (hasL, np.uint8),
                                                        (colR, np.int32), (rowR, np.int32), (hasR, np.uint8))]
        H, W = arrs[0].shape
        r = np.ascontiguousarray(np.asarray(rigid, np.float32).reshape(12)) if rigid is not None else None
        ssum = np.empty((scan_w * scan_h, 3), np.float32)
        cnt = np.empty((scan_w * scan_h,), np.uint8)
        cl, cr = _cam(cams[0]), _cam(cams[1])
        self.lib.ref_gray_triangulate(*[C.c_void_p(a.ctypes.data) for a in arrs], W, H, scan_w, scan_h, C.byref(cl),
                                      C.byref(cr), C.c_void_p(r.ctypes.data) if r is not None else None,
                                      C.c_void_p(ssum.ctypes.data), C.c_void_p(cnt.ctypes.data))
        return ssum, cnt

    def undistort(self, x, y, cam):
        c = _cam(cam)
        ox, oy = C.c_float(0), C.c_float(0)
        self.lib.ref_undistort(np.float32(x), np.float32(y), C.byref(c), C.byref(ox), C.byref(oy))
        return np.float32(ox.value), np.float32(oy.value)

    def line_line(self, p1, v1, p2, v2):
        a = [np.ascontiguousarray(v, np.float32) for v in (p1, v1, p2, v2)]
        out = np.zeros(3, np.float32)
        ok = self.lib.ref_line_line(*[C.c_void_p(x.ctypes.data) for x in a], C.c_void_p(out.ctypes.data))
        return bool(ok), out

    def cam2world(self, cam, p):
        c = _cam(cam)
        q = np.ascontiguousarray(p, np.float32).copy()
        self.lib.ref_cam2world(C.byref(c), C.c_void_p(q.ctypes.data))
        return q

    def normalize(self, v):
        q = np.ascontiguousarray(v, np.float32).copy()
        self.lib.ref_normalize(C.c_void_p(q.ctypes.data))
        return q

    def pointcloud_add(self, w, h, iw, jh, pts):
        iw, jh = np.ascontiguousarray(iw, np.int32), np.ascontiguousarray(jh, np.int32)
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.empty((h, w, 3), np.float32)
        cnt = np.empty((h, w), np.uint8)
        self.lib.ref_pointcloud_add(w, h, C.c_void_p(iw.ctypes.data), C.c_void_p(jh.ctypes.data), C.c_void_p(pts.ctypes.data),
                                    len(iw), C.c_void_p(out.ctypes.data), C.c_void_p(cnt.ctypes.data))
        return out, cnt


    def export_mesh(self, points, count, w, h, path, obj=False, color=None):
        """MeshCreator::exportPlyMesh / exportObjMesh on a cloud given as sums [h,w,3] + counts [h,w]."""
        points = np.ascontiguousarray(points, np.float32)
        count = np.ascontiguousarray(count, np.uint8)
        cptr = None
        if color is not None:
            color = np.ascontiguousarray(color, np.uint8)
            cptr = C.c_void_p(color.ctypes.data)
        self.lib.ref_export_mesh(C.c_void_p(points.ctypes.data), C.c_void_p(count.ctypes.data), cptr, int(w), int(h),
                                 int(obj), str(path).encode())

    def export_xyz(self, points, count, w, h, path, export_off=True, color_flag=True, color=None):
        points = np.ascontiguousarray(points, np.float32)
        count = np.ascontiguousarray(count, np.uint8)
        cptr = None
        if color is not None:
            color = np.ascontiguousarray(color, np.uint8)
            cptr = C.c_void_p(color.ctypes.data)
        self.lib.ref_export_xyz(C.c_void_p(points.ctypes.data), C.c_void_p(count.ctypes.data), cptr, int(w), int(h),
                                int(export_off), int(color_flag), str(path).encode())


_ref = None


def load():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref
