"""CPU model of the multi-GPU strip protocol (test infrastructure).

One `StripRank` per GPU: it owns global rows [row0, row1) of the W x H trail and the agents
whose row lies there, and talks to its two ring neighbours only.  Compute uses the
oracle; communication goes through a pluggable transport (in-process queues or
torch.distributed/gloo).  The protocol is exactly the one slime_mold_b200/csrc/exchange.cu
implements with NCCL (DESIGN.md "Multi-GPU"):

  per step
    1. agents pass     sense from trail rows own +- g (ghosts), count deposits into own +- m
    2. exchange A      to each neighbour: my deposit counts that landed in ITS m boundary rows
                       + my own counts of MY boundary row next to it (1 row)
                       -> owned rows complete; one ghost row each side complete (for the 3x3 blur)
    3. trail pass      merge -> decay -> 3x3 mean on the owned rows
    4. exchange B      my new top / bottom g trail rows -> neighbours' ghost rows;
                       agents that left the strip -> the neighbour that owns their row

Arrays are kept at full map size in GLOBAL coordinates with everything a rank may not read
poisoned, so any access outside "own + ghost" shows up as a mismatch against the
single-domain oracle.
"""
from __future__ import annotations

import math

import numpy as np

POISON = np.float32(1.0e30)


def strip_bounds(rank, world, H):
    return (rank * H) // world, ((rank + 1) * H) // world


def owner_row(y, H):
    r = np.where(y >= 0, np.minimum(y, np.float32(H - 1)), np.float32(0)).astype(np.int64)   # NaN/negative -> 0
    return r


class StripRank:
    def __init__(self, so, p, rank, world, agents_global, trail_global=None):
        self.so, self.p, self.rank, self.world = so, p, rank, world
        self.W, self.H = p.width, p.height
        self.row0, self.row1 = strip_bounds(rank, world, self.H)
        self.up, self.down = (rank - 1) % world, (rank + 1) % world
        self.g = int(math.ceil(abs(p.agent_sensor_distance))) + 3          # sensing ghost depth
        self.m = int(math.ceil(abs(p.agent_speed_max) * 0.016)) + 1        # deposit / migration reach
        thin = min(strip_bounds(r, world, self.H)[1] - strip_bounds(r, world, self.H)[0] for r in range(world))
        assert self.g <= thin and self.m <= thin, "strip thinner than the halo"
        rows = owner_row(agents_global[:, 1], self.H)
        mine = (rows >= self.row0) & (rows < self.row1)
        self.ids = np.nonzero(mine)[0].astype(np.uint32)
        self.agents = np.ascontiguousarray(agents_global[mine], dtype=np.float32)
        self.trail = np.full((self.H, self.W), POISON, np.float32)
        src = np.zeros((self.H, self.W), np.float32) if trail_global is None else trail_global
        own = self.rows_mod(self.row0 - self.g, self.row1 + self.g)
        self.trail[own] = src[own]
        self.counts = np.zeros((self.H, self.W), np.uint32)

    def rows_mod(self, a, b):
        return np.arange(a, b) % self.H

    # ---- phase 1 -------------------------------------------------------------------------
    def agents_pass(self):
        self.so.agents_phase_split(self.agents, self.trail, self.counts, self.p, ids=self.ids)
        allowed = np.zeros(self.H, bool)
        allowed[self.rows_mod(self.row0 - self.m, self.row1 + self.m)] = True
        assert not self.counts[~allowed].any(), "deposit outside own +- m rows"

    # ---- exchange A ----------------------------------------------------------------------
    def counts_messages(self):
        m = self.m
        to_up = (self.counts[self.rows_mod(self.row0 - m, self.row0)].copy(), self.counts[self.row0].copy())
        to_down = (self.counts[self.rows_mod(self.row1, self.row1 + m)].copy(), self.counts[self.row1 - 1].copy())
        return {"up": to_up, "down": to_down}

    def counts_receive(self, from_up, from_down):
        m = self.m
        ghost_contrib, own_row = from_up          # up's deposits into my first m rows; up's own last row
        self.counts[self.row0:self.row0 + m] += ghost_contrib
        self.counts[(self.row0 - 1) % self.H] += own_row
        ghost_contrib, own_row = from_down        # down's deposits into my last m rows; down's own first row
        self.counts[self.row1 - m:self.row1] += ghost_contrib
        self.counts[self.row1 % self.H] += own_row

    # ---- phase 3 -------------------------------------------------------------------------
    def trail_pass(self):
        idx = self.rows_mod(self.row0 - 1, self.row1 + 1)
        t = np.ascontiguousarray(self.trail[idx])
        c = np.ascontiguousarray(self.counts[idx])
        pp = type(self.p)()
        for f, _ in self.p._fields_:
            setattr(pp, f, getattr(self.p, f))
        pp.height = t.shape[0]
        out = self.so.trail_pass(t, pp, counts=c)
        new = np.full((self.H, self.W), POISON, np.float32)
        new[self.row0:self.row1] = out[1:-1]       # the padded rows wrapped onto each other: discard them
        self.trail = new
        self.counts[:] = 0

    # ---- exchange B ----------------------------------------------------------------------
    def trail_messages(self):
        g = self.g
        return {"up": self.trail[self.row0:self.row0 + g].copy(), "down": self.trail[self.row1 - g:self.row1].copy()}

    def trail_receive(self, from_up, from_down):
        g = self.g
        self.trail[self.rows_mod(self.row0 - g, self.row0)] = from_up       # up's bottom g rows
        self.trail[self.rows_mod(self.row1, self.row1 + g)] = from_down     # down's top g rows

    def migration_messages(self):
        rows = owner_row(self.agents[:, 1], self.H)
        rel = (rows - self.row0) % self.H                       # 0..rows-1 = stay
        nrows = self.row1 - self.row0
        stay = rel < nrows
        below = rel - nrows + 1                                 # rows past my last row, going down
        above = self.H - rel                                    # rows before my first row, going up
        go_down = ~stay & (below <= above)
        go_up = ~stay & ~go_down
        assert (below[go_down] <= self.m + 1).all() and (above[go_up] <= self.m + 1).all(), \
            "agent jumped further than the migration reach"
        msgs = {"up": (self.agents[go_up].copy(), self.ids[go_up].copy()),
                "down": (self.agents[go_down].copy(), self.ids[go_down].copy())}
        self.agents, self.ids = self.agents[stay], self.ids[stay]
        return msgs

    def migration_receive(self, from_up, from_down):
        for a, i in (from_up, from_down):
            if len(i):
                self.agents = np.ascontiguousarray(np.concatenate([self.agents, a]))
                self.ids = np.concatenate([self.ids, i])
        rows = owner_row(self.agents[:, 1], self.H)
        assert ((rows >= self.row0) & (rows < self.row1)).all(), "received an agent this strip does not own"


def run_in_process(so, p, agents, trail, world, steps):
    """All ranks in one process, messages passed by reference.  Returns (agents, trail) in global order."""
    ranks = [StripRank(so, p, r, world, agents, trail) for r in range(world)]

    def deliver(kind_msgs, recv_name):
        # message sent "up" by rank r arrives at rank r.up as "from_down" and vice versa
        for r in ranks:
            from_up = kind_msgs[r.up]["down"]
            from_down = kind_msgs[r.down]["up"]
            getattr(r, recv_name)(from_up, from_down)

    for _ in range(steps):
        for r in ranks:
            r.agents_pass()
        deliver([r.counts_messages() for r in ranks], "counts_receive")
        for r in ranks:
            r.trail_pass()
        deliver([r.trail_messages() for r in ranks], "trail_receive")
        deliver([r.migration_messages() for r in ranks], "migration_receive")
    return gather(ranks, agents.shape[0])


def gather(ranks, n):
    H, W = ranks[0].H, ranks[0].W
    out_a = np.full((n, 4), np.nan, np.float32)
    out_t = np.empty((H, W), np.float32)
    seen = np.zeros(n, np.int32)
    for r in ranks:
        out_a[r.ids] = r.agents
        seen[r.ids] += 1
        out_t[r.row0:r.row1] = r.trail[r.row0:r.row1]
    assert (seen == 1).all(), "an agent is owned by zero or several strips"
    return out_a, out_t


# ---- torch.distributed (gloo) transport: one process per strip -------------------------------
def run_distributed(so, p, agents, trail, steps, dist):
    """Runs this process's strip over torch.distributed.  Message order for world == 2 (both
    neighbours are the same peer) follows exchange.cu: sends [to up, to down], receives
    [from down, from up]."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    me = StripRank(so, p, rank, world, agents, trail)

    def xchg(arr_up, arr_down, like_up, like_down):
        """send arr_up to `up`, arr_down to `down`; receive (from_up, from_down)."""
        su, sd = torch.from_numpy(np.ascontiguousarray(arr_up)), torch.from_numpy(np.ascontiguousarray(arr_down))
        ru, rd = torch.empty_like(torch.from_numpy(like_up)), torch.empty_like(torch.from_numpy(like_down))
        ops = [dist.P2POp(dist.isend, su, me.up, tag=1), dist.P2POp(dist.isend, sd, me.down, tag=2),
               dist.P2POp(dist.irecv, rd, me.down, tag=1),      # what `down` sent up
               dist.P2POp(dist.irecv, ru, me.up, tag=2)]        # what `up` sent down
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return ru.numpy(), rd.numpy()

    def xchg_var(arr_up, arr_down):
        """variable-length rows: exchange the counts first, like the engine does."""
        n_up, n_down = xchg(np.array([len(arr_up)], np.int64), np.array([len(arr_down)], np.int64),
                            np.zeros(1, np.int64), np.zeros(1, np.int64))
        shape = arr_up.shape[1:]
        return xchg(arr_up, arr_down, np.zeros((int(n_up[0]),) + shape, arr_up.dtype),
                    np.zeros((int(n_down[0]),) + shape, arr_up.dtype))

    for _ in range(steps):
        me.agents_pass()
        msg = me.counts_messages()
        pack = lambda t: np.concatenate([t[0], t[1][None, :]])          # noqa: E731  (m + 1 rows)
        ru, rd = xchg(pack(msg["up"]), pack(msg["down"]), pack(msg["down"]), pack(msg["up"]))
        me.counts_receive((ru[:-1], ru[-1]), (rd[:-1], rd[-1]))
        me.trail_pass()
        msg = me.trail_messages()
        ru, rd = xchg(msg["up"], msg["down"], msg["down"], msg["up"])
        me.trail_receive(ru, rd)
        msg = me.migration_messages()
        au, ad = xchg_var(msg["up"][0], msg["down"][0])
        iu, idn = xchg_var(msg["up"][1], msg["down"][1])
        me.migration_receive((au, iu), (ad, idn))
    return me
