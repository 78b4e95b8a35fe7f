"""Warp cases whose 96x96 ROI can be written down without any warp code (shared by the CPU test
of the oracle and the GPU test of the kernels)."""
import numpy as np

from oracle import lips as O


def known_answer_case(kind, H=224, W=224, T=5, seed=0):
    """Frames, matrices and landmarks for which the 96x96 ROI can be written down without any
    warp code.  All 20 mouth landmarks sit on one integer point, so cut_patch's centre is exact."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    gray = O.bgr2gray(frames)
    if kind == "translate":          # std[r, c] = gray[r + ty, c + tx]
        s, tx, ty, px, py = 1.0, -30, 17, 100, 120
    elif kind == "border":           # the ROI hangs over the frame's left and top edges
        s, tx, ty, px, py = 1.0, -100, -90, 20, 30
    else:                            # "scale2": std[r, c] = gray[r / 2 + ty, c / 2 + tx]
        s, tx, ty, px, py = 0.5, 40, 30, 100, 110
    inv = np.array([[s, 0, tx], [0, s, ty], [0, 0, 1]], dtype=np.float64)
    fwd = np.array([[1 / s, 0, -tx / s], [0, 1 / s, -ty / s], [0, 0, 1]], dtype=np.float64)
    lm = np.zeros((T, 68, 2), dtype=np.float64)
    lm[:, :, 0], lm[:, :, 1] = px, py
    cx, cy = (px - tx) / s, (py - ty) / s            # mouth centre in the 300 x 300 frame (exact)
    cx, cy = min(max(cx, 48), 300 - 48), min(max(cy, 48), 300 - 48)
    r0, c0 = int(cy) - 48, int(cx) - 48
    expect = np.zeros((T, 96, 96), dtype=np.uint8)
    g = gray.astype(np.float64) / 255.0

    def tap(t, y, x):
        return g[t, y, x] if (0 <= y < H and 0 <= x < W) else 0.0

    for i in range(96):
        for j in range(96):
            y2, x2 = (r0 + i), (c0 + j)
            if s == 1.0:
                y, x = y2 + ty, x2 + tx
                if 0 <= y < H and 0 <= x < W:
                    expect[:, i, j] = gray[:, y, x]
            else:
                y, x, fy, fx = y2 // 2 + ty, x2 // 2 + tx, (y2 % 2) * 0.5, (x2 % 2) * 0.5
                y1, x1 = y + (y2 % 2), x + (x2 % 2)
                for t in range(T):
                    top = (1 - fx) * tap(t, y, x) + fx * tap(t, y, x1)
                    bot = (1 - fx) * tap(t, y1, x) + fx * tap(t, y1, x1)
                    expect[t, i, j] = int(((1 - fy) * top + fy * bot) * 255)
    tf = np.concatenate([fwd.reshape(-1), inv.reshape(-1)])
    return frames, gray, lm, np.tile(tf, (T, 1)), expect, (r0, c0)
