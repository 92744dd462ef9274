"""``rgb_array`` rendering without matplotlib (render hook).

The reference draws the scene with matplotlib patches (``rendering.py:32-341``); matplotlib is
not part of this stack, and drawing is outside the device hot path.  This module rasterises the
same scene with numpy — the same regions in the same colours (tram area, waiting area, destination
rows, walls, door, boarding agents red, exiting agents blue) — from plain position arrays, so it
serves the single-env facade (``CollectiveCrossingEnv.render``) and any env of a batched device
rollout (``render_state`` on rows copied from the device) alike.  It is a pixel rasteriser, not a
reproduction of matplotlib's anti-aliased output.
"""

from __future__ import annotations

from typing import Any

import numpy as np

# palette of the reference's drawer (rendering.py:44-55)
COLORS = {
    "background": (0xF8, 0xF9, 0xFA), "tram_area": (0xE3, 0xF2, 0xFD), "waiting_area": (0xFF, 0xF3, 0xE0),
    "exiting_destination_area": (0xF4, 0x43, 0x36), "boarding_destination_area": (0x21, 0x96, 0xF3),
    "tram_wall": (0x42, 0x42, 0x42), "door": (0x90, 0xCA, 0xF9), "boarding_agent": (0xF4, 0x43, 0x36),
    "exiting_agent": (0x21, 0x96, 0xF3), "inactive_agent": (0x9E, 0x9E, 0x9E),
}


def _blend(img, y0, y1, x0, x1, color, alpha):
    y0, y1, x0, x1 = (max(0, int(round(v))) for v in (y0, y1, x0, x1))
    if y1 <= y0 or x1 <= x0:
        return
    region = img[y0:y1, x0:x1].astype(np.float32)
    img[y0:y1, x0:x1] = (region * (1.0 - alpha) + np.asarray(color, np.float32) * alpha).astype(np.uint8)


def render_state(config: Any, boundaries: Any, xs, ys, active, num_boarding: int, cell: int = 32) -> np.ndarray:
    """RGB image (uint8, [rows, cols, 3]) of one env.  ``xs, ys, active``: per-agent arrays in agent
    order (boarding agents first); lattice point (x, y) is drawn in the cell [x, x+1) x [y, y+1),
    y growing upwards as in the reference's plot."""
    W, H, D = config.width, config.height, config.division_y
    TL, TR, DL, DR = boundaries.tram_left, boundaries.tram_right, boundaries.tram_door_left, boundaries.tram_door_right
    cols, rows = W + 2, H + 2                      # lattice points 0..W, 0..H plus a margin cell
    img = np.empty((rows * cell, cols * cell, 3), np.uint8)
    img[:] = COLORS["background"]

    def rect(x, y, w, h, color, alpha):            # in lattice units, origin bottom-left
        _blend(img, (rows - (y + h)) * cell, (rows - y) * cell, x * cell, (x + w) * cell, color, alpha)

    rect(TL, D, TR - TL + 1, H - D, COLORS["tram_area"], 0.7)                      # rendering.py:60-69
    rect(0, 0, W, D, COLORS["waiting_area"], 0.7)                                  # :72-81
    if config.exiting_destination_area_y < D:                                      # :84-94
        rect(0, config.exiting_destination_area_y, W, 1, COLORS["exiting_destination_area"], 0.8)
    if config.boarding_destination_area_y >= D:                                    # :97-123
        yb = H - 1 if config.boarding_destination_area_y == H else config.boarding_destination_area_y
        rect(TL, yb, TR - TL + 1, 1, COLORS["boarding_destination_area"], 0.8)
    t = 0.1                                                                        # wall thickness, :125
    if DL > TL:
        rect(TL, D - t / 2, DL - TL + 0.5, t, COLORS["tram_wall"], 0.9)            # :127-139
    if DR < TR:
        rect(DR - 0.5, D - t / 2, TR - DR + 1.5, t, COLORS["tram_wall"], 0.9)      # :141-156
    rect(TL - t / 2, D, t, H - D, COLORS["tram_wall"], 0.9)                        # :158-181
    rect(TR + 1 - t / 2, D, t, H - D, COLORS["tram_wall"], 0.9)
    if DR - DL - 1 > 0:
        rect(DL + 0.5, D, DR - DL - 1, 1, COLORS["door"], 0.8)                     # :183-197

    yy, xx = np.mgrid[0:cell, 0:cell]
    disc = (yy - (cell - 1) / 2) ** 2 + (xx - (cell - 1) / 2) ** 2 <= (0.35 * cell) ** 2
    for k, (x, y) in enumerate(zip(np.asarray(xs).tolist(), np.asarray(ys).tolist())):
        if x is None or y is None:
            continue
        x, y = int(x), int(y)
        if not (0 <= x < cols and 0 <= y < rows):
            continue
        color = COLORS["boarding_agent" if k < num_boarding else "exiting_agent"] if bool(np.asarray(active)[k]) else COLORS["inactive_agent"]
        r0, c0 = (rows - 1 - y) * cell, x * cell
        img[r0:r0 + cell, c0:c0 + cell][disc] = color
    return img


def render_env(env: Any, cell: int = 32) -> np.ndarray:
    """``rgb_array`` of a single-env facade (its host records)."""
    ids = env.possible_agents
    agents = [env._agents[a] for a in ids]
    return render_state(env.config, env.tram_boundaries, [a.position[0] for a in agents], [a.position[1] for a in agents],
                        [a.active for a in agents], env.config.num_boarding_agents, cell)


def render_batched(env: Any, env_index: int, cell: int = 32) -> np.ndarray:
    """``rgb_array`` of env ``env_index`` of a :class:`BatchedCollectiveCrossing` (copies that env's rows)."""
    from . import _abi
    from .utils.geometry import calculate_tram_boundaries

    k = int(env_index)
    fl = env.flags[k].cpu().numpy()
    return render_state(env.config, calculate_tram_boundaries(env.config), env.x[k].cpu().numpy(), env.y[k].cpu().numpy(),
                        (fl & _abi.F_ACTIVE) != 0, env.config.num_boarding_agents, cell)
