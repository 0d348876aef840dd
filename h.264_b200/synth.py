"""Deterministic synthetic YUV clips: a textured canvas under a global pan, one independently moving
patch and per-frame sensor noise, so that motion search has real work to do (non-trivial vectors, few exact ties).
Also flat and striped clips that maximise ties for the spiral-order tie-break tests.
"""
import numpy as np


def _octave(rng, h, w, cell):
    gh, gw = (h + cell - 1) // cell + 1, (w + cell - 1) // cell + 1
    g = rng.random((gh, gw))
    return np.kron(g, np.ones((cell, cell)))[:h, :w]


def luma_canvas(width, height, seed=1234, margin=128):
    rng = np.random.default_rng(seed)
    H, W = height + 2 * margin, width + 2 * margin
    c = 90 * _octave(rng, H, W, 32) + 50 * _octave(rng, H, W, 8) + 25 * _octave(rng, H, W, 2)
    c = (c - c.min()) / (c.max() - c.min())
    return 16.0 + c * (235 - 16)


def luma_frames(width, height, n_frames, seed=1234, kind="pan"):
    """list of uint8 (height, width) luma frames"""
    if kind == "flat":
        return [np.full((height, width), 128, np.uint8) for _ in range(n_frames)]
    if kind == "stripes":
        row = ((np.arange(width) // 4) % 2 * 120 + 60).astype(np.uint8)
        return [np.tile(row, (height, 1)) for _ in range(n_frames)]
    margin = 128
    canvas = luma_canvas(width, height, seed, margin)
    rng = np.random.default_rng(seed + 1)
    patch = 40.0 + 180.0 * np.random.default_rng(seed + 2).random((64, 64))
    out = []
    for t in range(n_frames):
        oy, ox = margin + 2 * t, margin + 3 * t
        f = canvas[oy:oy + height, ox:ox + width].copy()
        py = (height // 2 - 32 + 4 * t) % max(1, height - 64)
        px = (width // 2 - 32 - 5 * t) % max(1, width - 64)
        f[py:py + 64, px:px + 64] = patch[:min(64, height - py), :min(64, width - px)]
        f = f + rng.normal(0.0, 1.5, f.shape)
        if kind == "fade":          # brightness ramps down and a pedestal comes in: weighted prediction gets real weights
            f = f * (1.0 - 0.06 * t) + 3.0 * t
        out.append(np.clip(np.rint(f), 0, 255).astype(np.uint8))
    return out


def yuv420_frames(width, height, n_frames, seed=1234, kind="pan"):
    """list of (Y, U, V) uint8 planes, 4:2:0"""
    res = []
    for y in luma_frames(width, height, n_frames, seed, kind):
        y2 = y[::2, ::2].astype(np.float64)
        u = np.clip(np.rint(128 + 0.25 * (y2 - 128)), 0, 255).astype(np.uint8)
        v = np.clip(np.rint(128 - 0.25 * (y2 - 128)), 0, 255).astype(np.uint8)
        res.append((y, u, v))
    return res


def write_yuv420(path, width, height, n_frames, seed=1234, kind="pan"):
    with open(path, "wb") as f:
        for y, u, v in yuv420_frames(width, height, n_frames, seed, kind):
            f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())
    return path


def write_yuv444(path, width, height, n_frames, seed=1234, kind="pan"):
    """4:4:4: U and V from the undecimated luma"""
    with open(path, "wb") as f:
        for y in luma_frames(width, height, n_frames, seed, kind):
            yf = y.astype(np.float64)
            f.write(y.tobytes())
            f.write(np.clip(np.rint(128 + 0.25 * (yf - 128)), 0, 255).astype(np.uint8).tobytes())
            f.write(np.clip(np.rint(128 - 0.25 * (yf - 128)), 0, 255).astype(np.uint8).tobytes())
    return path


def synthetic_jobs(width, height, n_refs, lambda_factor, seed=7, rdopt=1, max_pmv=24, slice_type=0):
    """A frame's worth of batch jobs with plausible predictors: per macroblock a smooth 'true motion' field plus
    jitter per partition, as a stand-in for the predictors an encoder would supply (replayed centres).
    Returns a numpy array of JOB_DTYPE ordered (ref major, then raster macroblocks)."""
    from .me_cuda import JOB_DTYPE, NUM_PART
    rng = np.random.default_rng(seed)
    mbw, mbh = width // 16, height // 16
    jobs = np.zeros(n_refs * mbw * mbh, JOB_DTYPE)
    k = 0
    for r in range(n_refs):
        base = rng.integers(-max_pmv, max_pmv + 1, size=(mbh, mbw, 2))
        # global pan of the synthetic clip: (3, 2) pixels per frame and reference distance r+1, in quarter pels
        base[..., 0] += 4 * 3 * (r + 1)
        base[..., 1] += 4 * 2 * (r + 1)
        for my in range(mbh):
            for mx in range(mbw):
                j = jobs[k]
                j["opix_x"], j["opix_y"] = 16 * mx, 16 * my
                j["list"], j["ref"], j["slice_type"] = 0, r, slice_type
                j["pmv16"] = base[my, mx]
                j["pmv"] = base[my, mx][None, :] + rng.integers(-6, 7, size=(NUM_PART, 2))
                j["pmv"][0] = j["pmv16"]
                j["lambda"] = lambda_factor
                k += 1
    return jobs
