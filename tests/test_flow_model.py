"""A model of the dataflow sweeps' protocol (rp_kernels.cuh: FlowQueue, flow_need, pos_flow) in plain Python, run under random
interleavings of the warps: the protocol must terminate (no deadlock) with EVERY body's units executed in level order, pass by
pass -- which is the reference's sequential order -- whatever the timing. This checks the design argument of DESIGN.md 3, not the
CUDA code (tests/test_gpu_flow.py does that, bit for bit against the barrier form)."""
import random

import pytest


def make_schedule(rng, n_bodies, n_units, p_fixed=0.15, p_live=0.7):
    """units in array order with the dependency-level recurrence of k_schedule: level = 1 + max(last level of either non-fixed body)"""
    fixed = [rng.random() < p_fixed for _ in range(n_bodies)]
    fixed[0] = True
    last = [0] * n_bodies
    units = []
    while len(units) < n_units:
        a, b = rng.sample(range(n_bodies), 2)
        if fixed[a] and fixed[b]:
            continue
        lvl = 1 + max(0 if fixed[a] else last[a], 0 if fixed[b] else last[b])
        if lvl > 62:
            continue
        for x in (a, b):
            if not fixed[x]:
                last[x] = lvl
        units.append(dict(a=a, b=b, level=lvl, live=rng.random() < p_live, contacts=rng.randint(1, 4)))
    return fixed, units


def run_protocol(rng, fixed, units, iters, n_warps, lanes):
    live = [u for u in units if u["live"]]
    levels = max([u["level"] for u in live], default=0)
    seq = [u for it in range(iters) for l in range(1, levels + 1) for u in live if u["level"] == l]  # level-major, pass by pass
    per_pass = len(live)
    mask = {}
    for u in live:  # k_manifold: body_live
        for x in (u["a"], u["b"]):
            if not fixed[x]:
                mask[x] = mask.get(x, 0) | (1 << u["level"])
    done = {}   # body_done
    order = {}  # what actually ran on each body, in time order
    cursor = [0]

    def need(body, level, it):
        if fixed[body]:
            return 0
        m = mask.get(body, 0)
        below = m & ((1 << level) - 1)
        if below:
            return ((it + 1) << 6) | (below.bit_length() - 1)
        if it > 0 and m:
            return (it << 6) | (m.bit_length() - 1)
        return 0

    warps = [dict(next=0, end=0, more=True, lane=[None] * lanes) for _ in range(n_warps)]
    idle_rounds = 0
    while True:
        progressed = False
        for w in rng.sample(warps, len(warps)):  # one trip of one warp at a time, in random order
            if rng.random() < 0.3:
                continue  # this warp is not scheduled now
            # take(): lanes without a unit get the next claimed items in order; a new chunk is claimed when the old one runs out
            for i in range(lanes):
                if w["lane"][i] is None:
                    if w["next"] >= w["end"] and w["more"]:
                        base = cursor[0]
                        cursor[0] += lanes
                        if base >= len(seq):
                            w["more"] = False
                        else:
                            w["next"], w["end"] = base, min(base + lanes, len(seq))
                    if w["next"] < w["end"]:
                        g = w["next"]
                        w["next"] += 1
                        u = seq[g]
                        it = g // per_pass
                        w["lane"][i] = dict(u=u, it=it, c=0, ready=False, n1=need(u["a"], u["level"], it), n2=need(u["b"], u["level"], it))
                        progressed = True
            for i in range(lanes):
                s = w["lane"][i]
                if s is None:
                    continue
                u = s["u"]
                if not s["ready"]:
                    s["ready"] = done.get(u["a"], 0) >= s["n1"] and done.get(u["b"], 0) >= s["n2"]
                    if not s["ready"]:
                        continue
                s["c"] += 1  # one contact per trip
                progressed = True
                if s["c"] == u["contacts"]:
                    for x in (u["a"], u["b"]):
                        if not fixed[x]:
                            order.setdefault(x, []).append((s["it"], u["level"]))
                            done[x] = ((s["it"] + 1) << 6) | u["level"]
                    w["lane"][i] = None
        if all(l is None for w in warps for l in w["lane"]) and cursor[0] >= len(seq) and all(w["next"] >= w["end"] for w in warps):
            return order, live
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 200, "deadlock: no lane can run"


@pytest.mark.parametrize("seed", range(12))
def test_protocol_terminates_in_reference_order(seed):
    rng = random.Random(seed)
    fixed, units = make_schedule(rng, n_bodies=rng.randint(4, 40), n_units=rng.randint(5, 150))
    iters = rng.randint(1, 3)
    order, live = run_protocol(rng, fixed, units, iters, n_warps=rng.randint(1, 6), lanes=rng.choice([2, 4, 8]))
    for body, ran in order.items():
        want = [(it, u["level"]) for it in range(iters) for u in live if body in (u["a"], u["b"])]  # array order = level order per body
        assert ran == sorted(want) == want, (seed, body)
    for x in range(len(fixed)):
        if not fixed[x] and any(x in (u["a"], u["b"]) for u in live):
            assert len(order[x]) == iters * sum(1 for u in live if x in (u["a"], u["b"]))
