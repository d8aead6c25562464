"""Pass backend for jpegqs_b200.multigpu.run_slab built on the ORACLE's per-block functions
(CPU tensors).  Test infrastructure: lets the world_size>1 gloo tests exercise the product's
sharding / halo-exchange / stop logic without a GPU."""
import ctypes as C

import numpy as np

import oracle_lib as ol
from jpegqs_b200 import multigpu as mg


class OraclePasses:
    def __init__(self, flags):
        self.lib = ol.oraclelib()
        size = 272 if flags & 1 else 160
        self.tables = np.zeros((64, size), dtype=np.float32)
        self.lib.qso_tables(flags, self.tables.ctypes.data)

    @staticmethod
    def _views(c):
        coef = c.coef.numpy()
        plane = c.plane.numpy()
        return coef, plane

    def idct(self, comps, mode, top_edge, bottom_edge, want_bad):
        bad = 0
        for idx, c in enumerate(comps):
            coef, plane = self._views(c)
            stride = plane.shape[1]
            rows, W = c.rows, c.wblk
            if mode & mg.PASS_DEQUANT:
                t = coef.astype(np.int32) * np.asarray(c.quant, dtype=np.int32)[None, None, :]
                if np.any((t + 0x800) >> 12):
                    bad |= 1 << min(idx, 31)
                coef[...] = t.astype(np.int16)
            for by in range(rows):
                for bx in range(W):
                    self.lib.qso_idct_islow(
                        C.c_void_p(coef[by, bx].ctypes.data),
                        C.c_void_p(plane.ctypes.data + (by * 8 + 1) * stride + mg.PLANE_PAD + bx * 8), stride)
            h, w = rows * 8, W * 8
            plane[1:h + 1, mg.PLANE_PAD - 1] = plane[1:h + 1, mg.PLANE_PAD]
            plane[1:h + 1, mg.PLANE_PAD + w] = plane[1:h + 1, mg.PLANE_PAD + w - 1]
            if top_edge:
                plane[0] = plane[1]
            if bottom_edge:
                plane[h + 1] = plane[h]
            if mode & mg.PASS_CLAMP:
                np.clip(coef, -1023, 1023, out=coef)
        return bad

    def smooth(self, comps, flags, clamp_out, top_edge, bottom_edge):
        for c in comps:
            coef, plane = self._views(c)
            stride = plane.shape[1]
            q = np.asarray(c.quant, dtype=np.uint16).copy()
            q[q == 0] = 1
            p2 = c.plane2.numpy() if c.plane2 is not None else None
            for by in range(c.rows):
                for bx in range(c.wblk):
                    off = (by * 8 + 1) * stride + mg.PLANE_PAD + bx * 8
                    self.lib.qso_smooth_block(
                        C.c_void_p(coef[by, bx].ctypes.data), C.c_void_p(q.ctypes.data),
                        C.c_void_p(plane.ctypes.data + off),
                        C.c_void_p(p2.ctypes.data + off) if p2 is not None else None, stride, flags, C.c_void_p(self.tables.ctypes.data), int(c.luma))
            if clamp_out:
                np.clip(coef, -1023, 1023, out=coef)

    def clamp(self, comps):
        for c in comps:
            c.coef.clamp_(-1023, 1023)

    def dequantize(self, c):
        coef = c.coef.numpy()
        coef[...] = (coef.astype(np.int32) * np.asarray(c.quant, dtype=np.int32)[None, None, :]).astype(np.int16)

    def new_plane(self, like, rows, wblk):
        import torch
        return torch.zeros((rows * 8 + 2, mg.plane_stride(wblk)), dtype=torch.uint8)

    @staticmethod
    def _p00(t):
        """address of pixel (0,0) of a plane tensor, stride"""
        a = t.numpy()
        return a.ctypes.data + a.shape[1] + mg.PLANE_PAD, a.shape[1]

    def downsample(self, y, c, plane2, ws, hs, top_edge, bottom_edge):
        y00, ys = self._p00(y.plane); d00, ds = self._p00(plane2)
        h = (y.hblk_total or y.rows) * 8; h2 = (c.hblk_total or c.rows) * 8
        first = -1 if top_edge else c.row0 * 8
        last = h2 if bottom_edge else (c.row0 + c.rows) * 8 - 1
        if last < first:
            return
        self.lib.qso_downsample_rows(C.c_void_p(y00), ys, y.wblk * 8, h, y.row0 * 8, C.c_void_p(d00), ds,
                                     c.wblk * 8, ws, hs, c.row0 * 8, first, last - first + 1, (h + hs - 1) // hs)

    def upsample(self, c, y, ws, hs, geom):
        import torch
        c00, cs = self._p00(c.plane); d00, _ = self._p00(c.plane2); y00, ys = self._p00(y.plane)
        ww, hh = y.wblk * 8, y.rows * 8
        out = np.zeros((max(hh, 1), ww), dtype=np.uint8)
        c.coef_up = torch.zeros((y.rows, y.wblk, 64), dtype=torch.int16)
        if not hh:
            return
        w1 = (geom.image_width + ws - 1) // ws; h1 = (geom.image_height + hs - 1) // hs
        self.lib.qso_upsample_rows(C.c_void_p(c00), C.c_void_p(d00), cs, C.c_void_p(y00), ys,
                                   C.c_void_p(out.ctypes.data), ww, w1, h1, ws, hs, ww, hh, y.row0 * 8)
        self.lib.qso_fdct_plane(C.c_void_p(out.ctypes.data), ww, C.c_void_p(c.coef_up.numpy().ctypes.data),
                                y.wblk, y.rows)
