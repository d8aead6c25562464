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
            for by in range(c.rows):
                for bx in range(c.wblk):
                    self.lib.qso_smooth_block(
                        C.c_void_p(coef[by, bx].ctypes.data), C.c_void_p(q.ctypes.data),
                        C.c_void_p(plane.ctypes.data + (by * 8 + 1) * stride + mg.PLANE_PAD + bx * 8),
                        None, stride, flags, C.c_void_p(self.tables.ctypes.data), int(c.luma))
            if clamp_out:
                np.clip(coef, -1023, 1023, out=coef)

    def clamp(self, comps):
        for c in comps:
            c.coef.clamp_(-1023, 1023)
