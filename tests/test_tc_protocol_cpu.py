"""CPU: race check of the tensor-core kernel's synchronisation protocol on a happens-before model.

The kernel (csrc/rced_net_tc.cu) updates its activation planes in place while MMAs of neighbouring row
tiles, of the next layer and of the next batch are in flight; what keeps that safe is a handful of waits
(the three commits an epilogue awaits -- the output layer's epilogue too, because it stages the next batch's
input for its row tile in place --, the scout's cumulative act_ready waits, in_ready before the first layer of
a CTA's first batch, w_free / w_full of the weight double buffer).  This test rebuilds those waits as a graph
over the events of two consecutive batches -- M(b,s,t): the MMAs of (step, row tile); E(b,s,t): its epilogue
(for the output layer of batch 0: including the staging of batch 1's rows of tile t); S(0,g): epilogue group g
staging the first batch's input; W(b,s): the producer's copy of step s's weights -- takes the rows and planes every event reads and writes from the library's own
layout tables (rced_tc_layout), and asserts that every pair of conflicting accesses (plane write vs plane
read or write on overlapping rows; accumulator write vs read; weight buffer write vs read) is ordered by
the transitive closure.  It would have caught the race of the two-commit wait with three issuing threads
(tile t-1's MMAs not awaited), which the GPU tests passed.  Not modelled: the skip scratch (written and
read by the same thread), the output partial sums and the input double buffer (bar.sync / counters)."""
import itertools

import pytest

from fullycnnspeechenhancement_b200 import _lib
from fullycnnspeechenhancement_b200.model_utils import fold
from oracle import network

TILES, GROUPS = 8, 4


def _layout(arch):
    import tc_emulator
    return tc_emulator.layout(_lib.lib(), arch)


def _accesses(name):
    """Per step: rows/planes read by the MMAs of a tile (relative to the tile's first row), planes written
    by its epilogue."""
    lay = _layout(fold.arch_id(name))
    table = network.layer_table(name)
    P16, ns = lay["plane16"], lay["ns"]
    steps = []
    for s, st in enumerate(lay["steps"]):
        reads = set()                                         # (plane, row shift)
        for u in range(st["units"]):
            off, lbo = lay["units"][st["unit_base"] + u]
            for o in (off, off + lbo):
                plane = (o + P16 // 2) // P16
                reads.add((plane, o - plane * P16))
        if st["final"]:                                       # odd-frame copy: two planes further
            reads |= {(p + 2, sh) for p, sh in set(reads)}
            writes = set()
        else:
            cg = (table[s]["cout"] + 7) // 8
            writes = set(range(cg))
            if s == ns - 2:                                   # last conv layer: even and odd copies
                writes |= {p + 2 for p in range(cg)}
        steps.append(dict(reads=reads, writes=writes, final=st["final"]))
    return steps


def _build(steps, wait_prev_tile=True, scout_waits_final_epilogue=True, final_waits_neighbours=True,
           producer_waits_w_free=True):
    ns = len(steps)
    nodes, edges = [], {}

    def node(*k):
        if k not in edges:
            edges[k] = set()
            nodes.append(k)
        return k

    def before(u, v):
        edges[node(*u)].add(node(*v))

    for b in range(2):
        if b == 0:
            for g in range(GROUPS):
                node("S", b, g)
        for s in range(ns):
            for t in range(TILES):
                node("M", b, s, t)
                node("E", b, s, t)
    for b in range(2):
        for s in range(ns):
            for t in range(TILES):
                # epilogue waits (mbar_wait3 / the output layer's single wait)
                if steps[s]["final"]:      # stages the next batch's rows of tile t if there is a next batch (b == 0)
                    need = [t] + ([t + 1, t - 1] if b == 0 and final_waits_neighbours else [])
                else:
                    need = [t, t + 1] + ([t - 1] if wait_prev_tile else [])
                for tt in need:
                    if 0 <= tt < TILES:
                        before(("M", b, s, tt), ("E", b, s, t))
                # scout: cumulative act_ready waits of the step before (of the batch before for step 0)
                prev = (b, s - 1) if s > 0 else ((b - 1, ns - 1) if b > 0 and scout_waits_final_epilogue else None)
                if prev is not None:
                    for tt in range(0, min(t + 1, TILES - 1) + 1):
                        before(("E", prev[0], prev[1], tt), ("M", b, s, t))
                if s == 0 and b == 0:                         # in_ready (first batch only): every group has staged
                    for g in range(GROUPS):
                        before(("S", b, g), ("M", b, 0, t))
        # program order of an epilogue group: tiles g, g + 4 of every step
        for g in range(GROUPS):
            seq = [("S", 0, g)] if b == 0 else []
            for s in range(ns):
                seq.append(("E", b, s, g))
                seq.append(("E", b, s, g + GROUPS))
            if b == 1:
                seq = [("E", 0, ns - 1, g + GROUPS)] + seq
            for u, v in zip(seq[:-1], seq[1:]):
                before(u, v)
    # weight producer: W(b,s) copies step s's tiles into buffer s & 1 once the step two before has released it
    # (w_free: a commit of every issuing thread behind its last tile of that step); the scout waits for w_full
    # before it clears the step's first tile, and clears the tiles in order
    for b in range(2):
        for s in range(ns):
            node("W", b, s)
            k = b * ns + s
            if k >= 2 and producer_waits_w_free:
                pb, ps = divmod(k - 2, ns)
                for t in range(TILES):
                    before(("M", pb, ps, t), ("W", b, s))
            if k >= 1:                                        # one producer thread, in step order
                pb, ps = divmod(k - 1, ns)
                before(("W", pb, ps), ("W", b, s))
            for t in range(TILES):
                before(("W", b, s), ("M", b, s, t))
    return nodes, edges


def _closure(nodes, edges):
    index = {n: i for i, n in enumerate(nodes)}
    reach = [0] * len(nodes)
    order, seen = [], set()
    for root in nodes:                                        # iterative post-order
        if root in seen:
            continue
        stack = [(root, iter(edges[root]))]
        seen.add(root)
        while stack:
            n, it = stack[-1]
            nxt = next(it, None)
            if nxt is None:
                order.append(n)
                stack.pop()
            elif nxt not in seen:
                seen.add(nxt)
                stack.append((nxt, iter(edges[nxt])))
    for n in order:                                           # successors first
        m = 0
        for v in edges[n]:
            m |= reach[index[v]] | (1 << index[v])
        reach[index[n]] = m
    return lambda u, v: bool(reach[index[u]] >> index[v] & 1)


def _races(steps, **variant):
    nodes, edges = _build(steps, **variant)
    hb = _closure(nodes, edges)
    ns = len(steps)

    def plane_reads(n):
        if n[0] != "M":
            return set()
        _, b, s, t = n
        return {(p, r) for p, sh in steps[s]["reads"] for r in range(128 * t + sh, 128 * t + sh + 128)}

    def plane_writes(n):
        if n[0] == "S":
            return {(0, r) for r in range(128 * TILES)}       # plane 0, every row (each group: interleaved rows)
        if n[0] == "E":
            _, b, s, t = n
            planes = steps[s]["writes"]
            if steps[s]["final"] and b == 0:                  # the next batch's first-layer input, rows of this tile
                planes = {0}
            return {(p, r) for p in planes for r in range(128 * t, 128 * t + 128)}
        return set()

    acc_w = {n: n[3] for n in nodes if n[0] == "M"}
    acc_r = {n: n[3] for n in nodes if n[0] == "E"}
    pr = {n: plane_reads(n) for n in nodes}
    pw = {n: plane_writes(n) for n in nodes}
    races = []
    for u, v in itertools.combinations(nodes, 2):
        conflict = bool(pw[u] & (pr[v] | pw[v])) or bool(pw[v] & pr[u])
        if u[0] == "S" and v[0] == "S" and u[1] == v[1]:
            conflict = False                                   # the groups of one staging write disjoint rows
        if not conflict and ((u in acc_w and v in acc_r) or (u in acc_r and v in acc_w)):
            conflict = acc_w.get(u, acc_r.get(u)) == acc_w.get(v, acc_r.get(v))
        if not conflict and u in acc_w and v in acc_w:
            conflict = acc_w[u] == acc_w[v]
        if not conflict and "W" in (u[0], v[0]) and {u[0], v[0]} <= {"W", "M"}:
            conflict = (u[1] * ns + u[2]) % 2 == (v[1] * ns + v[2]) % 2   # same weight buffer (write vs read, write vs write)
        if conflict and not hb(u, v) and not hb(v, u):
            races.append((u, v))
    return races


@pytest.mark.parametrize("name", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_protocol_orders_every_conflicting_access(name):
    steps = _accesses(name)
    assert _races(steps) == []


def test_model_detects_the_known_protocol_bugs():
    """The checker is not vacuous: each weakened protocol has races."""
    steps = _accesses("FullyCNNV2")
    r = _races(steps, wait_prev_tile=False)                   # the two-commit wait (tile t-1's MMAs read tile t's halo)
    assert any(u[0] != v[0] and {u[0], v[0]} == {"M", "E"} for u, v in r)
    assert _races(steps, scout_waits_final_epilogue=False)    # next batch's first layer overwrites unread accumulators
    assert _races(steps, final_waits_neighbours=False)        # staging tile t's rows under the output layer's MMAs of tiles t-1, t+1
    assert _races(steps, producer_waits_w_free=False)         # weights of step s+2 land under the MMAs of step s
