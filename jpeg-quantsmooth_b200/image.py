"""Host-side containers mirroring what libjpeg hands to do_quantsmooth.

A `CoefImage` is the flat equivalent of (j_decompress_ptr, jvirt_barray_ptr[])
as read by the reference driver (reference quantsmooth.h:2404-2453):
per component the quantized JCOEF blocks in natural (row-major) order, the raw
quantval table, the sampling factors and the libjpeg block geometry.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# libjpeg J_COLOR_SPACE values (jpeglib.h)
JCS_UNKNOWN, JCS_GRAYSCALE, JCS_RGB, JCS_YCbCr, JCS_CMYK, JCS_YCCK = range(6)


@dataclass
class Component:
    coef: np.ndarray                    # int16 [hblk, wblk, 64]
    quant: Optional[np.ndarray]         # uint16 [64] raw quantval, None = no table
    h_samp: int = 1
    v_samp: int = 1
    quant_tbl_no: int = 0

    @property
    def wblk(self) -> int:
        return int(self.coef.shape[1])

    @property
    def hblk(self) -> int:
        return int(self.coef.shape[0])


@dataclass
class CoefImage:
    width: int
    height: int
    colorspace: int
    comps: List[Component] = field(default_factory=list)

    def clone(self) -> "CoefImage":
        return copy.deepcopy(self)

    @property
    def num_blocks(self) -> int:
        return sum(c.wblk * c.hblk for c in self.comps)

    @property
    def mpixels(self) -> float:
        return self.width * self.height / 1e6


def blocks_for(image_dim: int, samp: int, max_samp: int) -> int:
    """libjpeg's width_in_blocks / height_in_blocks (jdmaster / jdinput):
    ceil(image_dim * samp / (max_samp * 8)) - NOT padded to the MCU."""
    return -(-image_dim * samp // (max_samp * 8))
