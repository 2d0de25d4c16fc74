"""Independent numpy transliteration of /root/reference/src/compute.wgsl (test infrastructure).

Written from the WGSL text, vectorised over agents, using numpy's own f32 sin/cos --
i.e. a *different* transcendental implementation than the oracle's, which is what any
real wgpu backend would be.  Used to cross-check the oracle's logic within the
tolerance the north star states; the trail passes use only + - * / and therefore
must agree with the oracle bit for bit.
"""
import numpy as np

F = np.float32
TAU = F(6.28318530718)
TWO_PI = F(2.0) * F(3.14159265359)
TIME_STEP = F(0.016)


def mix(a, b, t):
    return a * (F(1.0) - t) + b * t


def sample_trail_map(trail, px, py):                       # compute.wgsl:7-29
    H, W = trail.shape
    x0 = np.floor(px).astype(np.int64)
    y0 = np.floor(py).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    outside = (x0 < 0) | (x1 >= W) | (y0 < 0) | (y1 >= H)
    x0c, x1c = np.clip(x0, 0, W - 1), np.clip(x1, 0, W - 1)
    y0c, y1c = np.clip(y0, 0, H - 1), np.clip(y1, 0, H - 1)
    dx = px - x0.astype(F)
    dy = py - y0.astype(F)
    v0 = mix(trail[y0c, x0c], trail[y0c, x1c], dx)
    v1 = mix(trail[y1c, x0c], trail[y1c, x1c], dx)
    return np.where(outside, F(0.0), mix(v0, v1, dy)).astype(F)


def agents_pass(agents, trail, u, ids=None):
    """compute.wgsl:57-145, phase_split semantics.  Returns (new agents, deposit counts)."""
    H, W = trail.shape
    x, y, angle, speed = (agents[:, i].astype(F) for i in range(4))
    idx = np.arange(agents.shape[0], dtype=np.int32) if ids is None else ids.astype(np.int32)
    speed = np.minimum(np.maximum(speed, F(u.agent_speed_min)), F(u.agent_speed_max))
    sa, sd = F(u.agent_sensor_angle), F(u.agent_sensor_distance)
    aL, aR, aC = angle - sa, angle + sa, angle
    vL = sample_trail_map(trail, x + sd * np.cos(aL), y + sd * np.sin(aL))
    vR = sample_trail_map(trail, x + sd * np.cos(aR), y + sd * np.sin(aR))
    vC = sample_trail_map(trail, x + sd * np.cos(aC), y + sd * np.sin(aC))
    keep = (vC > vL) & (vC > vR)
    left = ~keep & (vL > vR)
    right = ~keep & ~left & (vR > vL)
    turn = F(u.agent_turn_speed)
    dL = (angle - TAU) - angle
    dR = (angle + TAU) - angle
    angle = np.where(left, angle + np.minimum(turn, np.abs(dL)) * np.sign(dL), angle)
    angle = np.where(right, angle + np.minimum(turn, np.abs(dR)) * np.sign(dR), angle).astype(F)
    arg = (idx.astype(F) * F(12.9898) + x * F(78.233)) + y * F(37.719)
    v = np.sin(arg).astype(F) * F(43758.5453)
    rnd = v - np.floor(v)
    angle = angle + (rnd * F(2.0) - F(1.0)) * F(u.agent_jitter)
    angle = np.fmod(angle, TWO_PI)
    angle = np.where(angle < 0, angle + TWO_PI, angle).astype(F)
    move = speed * TIME_STEP
    x = x + move * np.cos(angle)
    y = y + move * np.sin(angle)
    x = np.fmod(x, F(W)); x = np.where(x < 0, x + F(W), x).astype(F)
    y = np.fmod(y, F(H)); y = np.where(y < 0, y + F(H), y).astype(F)
    cx, cy = x.astype(np.int32), y.astype(np.int32)
    ok = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
    counts = np.zeros((H, W), np.uint32)
    np.add.at(counts, (cy[ok], cx[ok]), 1)
    return np.stack([x, y, angle, speed], axis=1).astype(F), counts


def trail_pass(trail, counts, u):
    """deposit merge + decay_trail (:148-161) + diffuse_trail (:164-195, Jacobi)."""
    t = trail.astype(F)
    if counts is not None:
        merged = np.minimum(np.maximum(t + counts.astype(F) * F(u.pheromone_deposition_amount), F(0)), F(1))
        t = np.where(counts > 0, merged, t).astype(F)
    t = np.maximum(t - F(u.decay_factor) * F(0.001), F(0)).astype(F)
    s = np.zeros_like(t)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            s = s + np.roll(np.roll(t, -dy, axis=0), -dx, axis=1)
    avg = s / F(9.0)
    rate = F(min(max(u.diffusion_rate, 0.0), 1.0))
    return mix(t, avg, rate).astype(F)
