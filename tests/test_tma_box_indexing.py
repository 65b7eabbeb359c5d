"""Index algebra of the remap kernel's TMA form (csrc/align.cu, kTma), simulated on the CPU.

A TMA tensor copy faults when its innermost start coordinate is not a multiple of 16 bytes (profiles/r02_tma_min2_xy.log),
so the kernel loads a 132 x 34 box from a 4-pixel boundary and addresses it with the tile's own offset, behind a guard band
for the tiles whose halo starts left of / above the image.  This test replays that addressing with numpy for every tile
of several image shapes: every read of the remap and median phases must see the clamp-sampled pixel the default form
loads, no two halo positions may share a word, and nothing may leave the buffer.  Constants are parsed from the source."""

import re
from pathlib import Path

import numpy as np
import pytest

SRC = (Path(__file__).resolve().parent.parent / "depthdensifier_b200" / "csrc" / "align.cu").read_text()


def const(name):
    m = re.search(rf"constexpr int {name} = (\d+)", SRC)
    assert m, name
    return int(m.group(1))


TILE_W, TILE_H = (int(x) for x in re.search(r"constexpr int kTileW = (\d+), kTileH = (\d+)", SRC).groups())
HALO_W, HALO_H = TILE_W + 2, TILE_H + 2
BOX_W, GUARD = const("kTmaBoxW"), const("kTmaGuard")


def test_constants_keep_the_copy_legal():
    assert (BOX_W * 4) % 16 == 0 and BOX_W <= 256 and HALO_H <= 256  # box rows are 16-byte multiples, box dims <= 256
    assert BOX_W >= HALO_W + 3  # room for the 0..3 columns of alignment slack
    assert (GUARD * 4) % 128 == 0 and GUARD >= BOX_W + 1  # destination stays 128-byte aligned; row -1 / column -1 fit
    assert "max(tx0 - 1, 0) & ~3" in SRC and "max(ty0 - 1, 0)" in SRC  # the start coordinates this test replays


@pytest.mark.parametrize("W,H", [(160, 120), (200, 152), (132, 34), (256, 64), (504, 100), (380, 97), (1920 // 4, 70)])
def test_box_addressing_equals_clamp_sampling(W, H):
    img = np.arange(1, W * H + 1, dtype=np.float64).reshape(H, W)
    for ty in range((H + TILE_H - 1) // TILE_H):
        for tx in range((W + TILE_W - 1) // TILE_W):
            tx0, ty0 = tx * TILE_W, ty * TILE_H
            bx0, by0 = max(tx0 - 1, 0) & ~3, max(ty0 - 1, 0)
            assert bx0 % 4 == 0 and bx0 >= 0 and by0 >= 0
            buf = np.full(GUARD + HALO_H * BOX_W, np.nan)
            ys, xs = by0 + np.arange(HALO_H)[:, None], bx0 + np.arange(BOX_W)[None, :]
            inside = (ys < H) & (xs < W)  # elements past the far edges arrive as zeros
            buf[GUARD:] = np.where(inside, img[np.minimum(ys, H - 1), np.minimum(xs, W - 1)], 0.0).ravel()
            base = GUARD + (ty0 - 1 - by0) * BOX_W + (tx0 - 1 - bx0)
            hy, hx = np.mgrid[0:HALO_H, 0:HALO_W]
            at = base + hy * BOX_W + hx
            assert at.min() >= 0 and at.max() < buf.size
            assert np.unique(at).size == at.size
            x, y = np.clip(tx0 + hx - 1, 0, W - 1), np.clip(ty0 + hy - 1, 0, H - 1)
            in_img = (x == tx0 + hx - 1) & (y == ty0 + hy - 1)
            # phase 1: positions inside the image hold the pixel the default form loads; "remap" them in place
            assert np.array_equal(buf[at[in_img]], img[y[in_img], x[in_img]])
            buf[at[in_img]] += 0.5
            # replicate pass: positions outside the image take the remapped clamped pixel, which lies inside the tile
            cx, cy = x - (tx0 - 1), y - (ty0 - 1)
            assert cx.min() >= 0 and cx.max() < HALO_W and cy.min() >= 0 and cy.max() < HALO_H
            assert in_img[cy, cx].all()
            buf[at[~in_img]] = buf[at[cy[~in_img], cx[~in_img]]]
            # phase 2: the 3x3 window of every output pixel of the tile
            ly, lx = np.mgrid[0:min(TILE_H, H - ty0), 0:min(TILE_W, W - tx0)]
            for dy in range(3):
                for dx in range(3):
                    want = img[np.clip(ty0 + ly - 1 + dy, 0, H - 1), np.clip(tx0 + lx - 1 + dx, 0, W - 1)] + 0.5
                    assert np.array_equal(buf[at[ly + dy, lx + dx]], want), (tx, ty, dy, dx)
