"""Region means of the reference's only published render (examples/cornell-10k-50-importance-sampling.png, README.md:4).

Run in the build container (needs /root/reference); writes tests/golden/cornell_example_regions.json, which travels.
The PNG is 400x400 RGBA, 8 bit, rendered by the REFERENCE ITSELF (10k spp, depth 50, importance sampling per its file name)
from an older variant of the scene: its tall box is diffuse white, today's is metal (src/scene.zig:348,370).  So only the
cells whose radiance the tall box does not dominate are comparable: the two coloured walls, the floor strip and the part of
the back wall right of the boxes.  SURVEY.md §8c-5 / VERDICT r1 item 1(c)."""
import json
from pathlib import Path

import numpy as np
from PIL import Image

SRC = Path("/root/reference/examples/cornell-10k-50-importance-sampling.png")
OUT = Path(__file__).resolve().parent / "cornell_example_regions.json"
CELL = 50  # 8 x 8 grid of 50 x 50 pixel cells

# (row, col) cells compared: left wall, right wall, floor strip, back wall right of the boxes
# (the cells bordering the tall box — (1,6), (2,5), (7,4), (7,5) — are 4-7 levels darker today: a mirror box sends less
# diffuse light to its neighbourhood than the white one of the published render did; they are left out)
CELLS = ([(r, c) for r in range(2, 7) for c in (0, 1)] + [(r, c) for r in range(2, 7) for c in (6, 7)] + [(1, 7)]
         + [(7, c) for c in (0, 1, 2, 3, 6, 7)] + [(3, 5)])

im = np.array(Image.open(SRC))
assert im.shape == (400, 400, 4) and im.dtype == np.uint8
rgb = im[..., :3].astype(np.float64)
grid = [[rgb[r * CELL:(r + 1) * CELL, c * CELL:(c + 1) * CELL].reshape(-1, 3).mean(0).round(3).tolist() for c in range(8)]
        for r in range(8)]
OUT.write_text(json.dumps({"source": "examples/cornell-10k-50-importance-sampling.png (reference repository)",
                           "width": 400, "height": 400, "cell": CELL, "grid_mean_rgb8": grid,
                           "cells": sorted(set(CELLS))}, indent=1))
print("wrote", OUT)
