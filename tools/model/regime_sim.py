"""Fluid model of one SM running the column sweep, to rank launch schedules before spending GPU time.

A tile (one warp x 32 columns) is 81 bandwidth-bound upper levels ("top") followed by 56 latency-bound cloudy
levels ("bot").  Rates from profiles/r1_level_timing.txt: a top level costs a warp 4500 cycles alone; a cloudy level
8200 / 9500 / 12500 cycles with 1 / 2 / 3 cloudy warps on its scheduler; every level moves 13 KB = 593 cycles of one
SM's share of the HBM bandwidth, and when the demand exceeds the share all warps slow down proportionally.
"""
import sys

TOP_L, BOT_L = 81, 56
T_TOP = 4500.0
T_BOT = {1: 8200.0, 2: 9500.0, 3: 12500.0}
BW = 593.0
TILES_PER_SM = 5120 / 148.0


def simulate(orders, delays=None, tiles=TILES_PER_SM, dynamic=True):
    """orders[w]: list of parts for warp w, e.g. ['tA','bA','tB','bB'] -- used when dynamic is False.
    dynamic: every warp fetches a new tile (top then bottom) when done, until `tiles` are handed out."""
    nw = 12
    delays = delays or [0.0] * nw
    t = 0.0
    handed = 0.0
    cur = [None] * nw          # (phase, remaining levels)
    queue = [list(o) for o in orders] if not dynamic else None
    done_top = [set() for _ in range(nw)]

    def fetch(w):
        nonlocal handed
        if dynamic:
            if handed >= tiles:
                return None
            frac = min(1.0, tiles - handed)
            handed += frac
            return ["top", TOP_L * frac, frac]
        if not queue[w]:
            return None
        p = queue[w].pop(0)
        return ["top" if p[0] == "t" else "bot", TOP_L if p[0] == "t" else BOT_L, 1.0]

    for w in range(nw):
        cur[w] = None
    started = [False] * nw
    while True:
        for w in range(nw):
            if cur[w] is None and t >= delays[w] and (not started[w] or True):
                started[w] = True
                cur[w] = fetch(w)
        active = [w for w in range(nw) if cur[w] is not None]
        pending = [delays[w] for w in range(nw) if cur[w] is None and t < delays[w]]
        if not active and not pending:
            break
        # rates
        lt = {}
        for w in active:
            if cur[w][0] == "top":
                lt[w] = T_TOP
            else:
                s = w % 4
                nb = sum(1 for v in active if v % 4 == s and cur[v][0] == "bot")
                lt[w] = T_BOT[nb]
        demand = sum(BW / lt[w] for w in active)
        scale = max(1.0, demand)
        dt = min([cur[w][1] * lt[w] * scale for w in active] + [d - t for d in pending]) if active else min(pending) - t
        dt = max(dt, 1e-6)
        for w in active:
            cur[w][1] -= dt / (lt[w] * scale)
        t += dt
        for w in active:
            if cur[w][1] <= 1e-9:
                if dynamic and cur[w][0] == "top":
                    cur[w] = ["bot", BOT_L * cur[w][2], cur[w][2]]
                else:
                    cur[w] = None
    return t


if __name__ == "__main__":
    ghz = 1.965e9
    base = simulate(None)
    print(f"hardware order, all start together   : {base/ghz*1e3:.3f} ms")
    for D in (100e3, 200e3, 300e3, 400e3, 600e3):
        for pat in ("slot2", "slot12"):
            d = [0.0] * 12
            for w in range(12):
                slot = w // 4
                if pat == "slot2" and slot == 2: d[w] = D
                if pat == "slot12": d[w] = D * slot
            print(f"stagger {pat} D={D/1e3:.0f}k  : {simulate(None, d)/ghz*1e3:.3f} ms")
    # static reorderings, 3 tiles per warp (36 tiles per SM instead of 34.6: scale the result)
    A = ["tA", "bA", "tB", "bB", "tC", "bC"]
    B = ["tA", "tB", "bA", "tC", "bB", "bC"]
    C = ["tA", "tB", "tC", "bA", "bB", "bC"]
    for name, o in (("AAA", [A, A, A]), ("ABC", [A, B, C]), ("AAB", [A, A, B]), ("ABB", [A, B, B]), ("AAC", [A, A, C]), ("ACC", [A, C, C])):
        orders = [o[w // 4] for w in range(12)]
        tt = simulate(orders, dynamic=False)
        print(f"static {name}: {tt/ghz*1e3*TILES_PER_SM/36:.3f} ms (scaled to 34.6 tiles)")
